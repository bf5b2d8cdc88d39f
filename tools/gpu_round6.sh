#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_reddit.json 2> gpurun_out/bench_reddit.err; echo "bench exit=$?"; tail -3 gpurun_out/bench_reddit.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench ref exit=$?"; cut -c1-300 gpurun_out/bench_reference.json
