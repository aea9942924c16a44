#!/bin/bash
# GPU box: kernel-only rates of the config-2 shape (100k pairs, 3 repetitions) in both modes + ncu summaries
tag=$1
for mode in 1 0; do python tools/probe_mode.py 200000 1000 64 $mode 4; done | tee gpurun_out/${tag}_rates.jsonl
if [ "$2" != "nocap" ]; then
tools/ncu_capture.sh ${tag}_dirs k1s_kernel 0 python tools/probe_mode.py 100000 1000 64 1 1
tools/ncu_capture.sh ${tag}_score k1s_kernel 0 python tools/probe_mode.py 100000 1000 64 0 1
fi
