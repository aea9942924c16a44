#!/bin/bash
# GPU box: ncu --set full capture of one kernel; brings back the summary CSV, the per-line table and the per-SASS-
# instruction table (gzip) so that executed-instruction counts can be attributed to loops here.
#   tools/ncu_sass_dump.sh <name> <kernel regex> <skip> <command...>
name=$1; regex=$2; skip=$3; shift 3
rep=/tmp/$name.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -f -o /tmp/$name "$@" > gpurun_out/${name}_run.log 2>&1
python tools/ncu_summary.py $rep gpurun_out/$name > /dev/null
python tools/ncu_lines.py $rep 70 > gpurun_out/${name}_lines.txt
ncu -i $rep --page source --print-source sass --csv | gzip -9 > gpurun_out/${name}_sass.csv.gz
