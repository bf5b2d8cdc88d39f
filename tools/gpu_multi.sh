#!/bin/bash
# run with: gpurun --gpus N -- bash tools/gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_$N.txt
timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -x -q > gpurun_out/pytest_dist_gpu.log 2>&1; echo "pytest dist exit=$?"; tail -5 gpurun_out/pytest_dist_gpu.log
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 50 --warmup 5 > gpurun_out/bench_reddit_${n}gpu.json 2> gpurun_out/bench_reddit_${n}gpu.err; echo "bench $n exit=$?"; cat gpurun_out/bench_reddit_${n}gpu.json | cut -c1-1500; tail -5 gpurun_out/bench_reddit_${n}gpu.err
  fi
done
