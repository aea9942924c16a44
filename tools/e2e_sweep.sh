#!/bin/bash
# GPU box: end-to-end time of config 2 (1 M pairs) under the pipeline's environment switches (DESIGN.md section 4,
# "Host pipeline"): the shipped pipeline against the round-1 upload order, descriptor copies through the copy
# engine, no ramps, 8 hardware queues.   usage: tools/e2e_sweep.sh <tag> [lib]
out=gpurun_out/$1_e2e_sweep.txt; : > $out
[ -n "$2" ] && export GAMX_LIB=$2
run() { env "$@" python tools/e2e_probe.py 1000000 2>&1 | grep SUMMARY >> $out; }
for rep in 1 2; do
run GAMX_X=shipped
run GAMX_UPLOAD_LAZY=1
run GAMX_JOBS_BY_DMA=1
run GAMX_NO_PIECE_RAMP=1 GAMX_NO_CHUNK_RAMP=1
run CUDA_DEVICE_MAX_CONNECTIONS=8 GAMX_X=8_hw_queues
run GAMX_HOST_THREADS=4
done
cat $out
