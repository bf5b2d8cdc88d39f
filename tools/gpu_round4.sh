#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python tools/sweep_dims.py reddit 0.5 > gpurun_out/sweep_default.log 2>&1; echo "sweep exit=$?"
timeout 600 python tools/sweep_dims.py ogbn-products 0.5 > gpurun_out/sweep_products_default.log 2>&1
timeout 600 python tools/epoch_breakdown.py > gpurun_out/epoch_breakdown.log 2>&1; echo "epoch exit=$?"
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_reddit.json 2> gpurun_out/bench_reddit.err; echo "bench exit=$?"; tail -3 gpurun_out/bench_reddit.err
