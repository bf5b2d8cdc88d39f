#!/bin/bash
# final validation of a round: full GPU suite, smoke, both bench arms, ncu launch list + full capture of the hot kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_reddit.json 2> gpurun_out/bench_reddit.err; echo "bench exit=$?"; tail -2 gpurun_out/bench_reddit.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench ref exit=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'aggregate_kernel|repack|unpack|part_|degrees_kernel' -c 30 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 1 --no-extras > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:aggregate_kernel -s 2 -c 1 -o gpurun_out/prof_agg_reddit_final -f python bench.py --steps 3 --warmup 1 --no-extras > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit=$?"
