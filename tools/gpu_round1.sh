#!/bin/bash
# one gpurun call: golden vectors from the reference kernels, GPU tests, smoke, bench, ncu
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
python oracle/make_golden_refgpu.py > gpurun_out/golden.log 2>&1; echo "golden exit=$?"
cp gpurun_out/refgpu.npz tests/golden/refgpu.npz 2>/dev/null
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gnna -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 1 --no-extras > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:aggregate_kernel -s 2 -c 2 -o gpurun_out/prof_agg -f python bench.py --steps 3 --warmup 1 --no-extras > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit=$?"
ls -la gpurun_out
