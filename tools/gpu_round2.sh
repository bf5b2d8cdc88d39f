#!/bin/bash
# tests + bench (reddit, products) + ncu launch list + full capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_reddit.json 2> gpurun_out/bench_reddit.err; echo "bench exit=$?"; cat gpurun_out/bench_reddit.json; tail -3 gpurun_out/bench_reddit.err
timeout 900 python bench.py --workload ogbn-products --steps 50 --warmup 5 > gpurun_out/bench_products.json 2> gpurun_out/bench_products.err; echo "bench products exit=$?"; cat gpurun_out/bench_products.json; tail -3 gpurun_out/bench_products.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'aggregate_kernel|prescale|part_|degrees_kernel' -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 1 --no-extras > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:aggregate_kernel -s 2 -c 1 -o gpurun_out/prof_agg_reddit -f python bench.py --steps 3 --warmup 1 --no-extras > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:aggregate_kernel -s 2 -c 1 -o gpurun_out/prof_agg_products -f python bench.py --workload ogbn-products --steps 3 --warmup 1 --no-extras > gpurun_out/ncu_full_products.log 2>&1; echo "ncu full products exit=$?"
ls -la gpurun_out
