#!/bin/bash
# Runs on the GPU box: ncu --set full capture of one kernel, summarised there (the .ncu-rep of the big
# alignment kernels with source exceeds what gpurun_out/ carries back).
#   tools/ncu_capture.sh <name> <kernel regex> <skip> <command...>
# writes gpurun_out/<name>.csv (tools/ncu_summary.py) and gpurun_out/<name>_lines.txt (tools/ncu_lines.py)
name=$1; regex=$2; skip=$3; shift 3
rep=/tmp/$name.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -f -o /tmp/$name "$@" > gpurun_out/${name}_run.log 2>&1
python tools/ncu_summary.py $rep gpurun_out/$name > /dev/null
python tools/ncu_lines.py $rep 70 > gpurun_out/${name}_lines.txt
