#!/bin/bash
# GPU box: end-to-end time of config 2 (1 M pairs) under the pipeline's environment switches
# usage: tools/e2e_sweep.sh <tag> [lib]
out=gpurun_out/$1_e2e_sweep.txt; : > $out
[ -n "$2" ] && export GAMX_LIB=$2
run() { env "$@" python tools/e2e_probe.py 1000000 2>&1 | grep SUMMARY >> $out; }
for rep in 1 2; do
run CUDA_DEVICE_MAX_CONNECTIONS=32 GAMX_X=conn32
run CUDA_DEVICE_MAX_CONNECTIONS=32 GAMX_NO_PIECE_RAMP=1 GAMX_NO_CHUNK_RAMP=1
run CUDA_DEVICE_MAX_CONNECTIONS=32 GAMX_FILL_GRID_SCALE=1.0
run CUDA_DEVICE_MAX_CONNECTIONS=16 GAMX_X=conn16
done
cat $out
