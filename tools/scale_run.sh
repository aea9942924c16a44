#!/bin/bash
# GPU box with N GPUs (gpurun --gpus N): the library's own N-device context test, bench.py and the config-4 merge
# stage at the given rank counts.   tools/scale_run.sh <tag> <N> [<N> ...]
tag=$1; shift
nvidia-smi -L > gpurun_out/${tag}_gpus.txt
python -m pytest tests/test_gpu_parity.py -m gpu -q -k multi_device 2>&1 | tail -2 >> gpurun_out/${tag}_gpus.txt
for N in "$@"; do
  if [ "$N" = 1 ]; then run="python"; else run="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N))"; fi
  $run bench.py --gpus $N --steps 5 --warmup 3 --no-other-configs --no-cpu-baseline > gpurun_out/${tag}_bench_${N}gpu.json 2> gpurun_out/${tag}_bench_${N}gpu.err
  $run tools/bench_merge.py 88000000 8 > gpurun_out/${tag}_cfg4_${N}gpu.json 2> gpurun_out/${tag}_cfg4_${N}gpu.err
done
for N in "$@"; do grep -h '^{' gpurun_out/${tag}_bench_${N}gpu.json | cut -c1-140; grep -h '^{' gpurun_out/${tag}_cfg4_${N}gpu.json | cut -c1-400; done
