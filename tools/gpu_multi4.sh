#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -x -q > gpurun_out/pytest_dist_gpu.log 2>&1; echo "pytest dist exit=$?"; tail -3 gpurun_out/pytest_dist_gpu.log
for rw in default; do
  if [ $rw = default ]; then unset GNNA_ROW_WEIGHT; else export GNNA_ROW_WEIGHT=$rw; fi
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_reddit_${N}gpu_rw$rw.json 2> gpurun_out/bench_reddit_${N}gpu_rw$rw.err; echo "bench $N rw=$rw exit=$?"
done
