#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/bench_fused.py ogbn-products 1.0 > gpurun_out/bench_fused_products.log 2>&1; echo "fused products exit=$?"; grep "^{'din" gpurun_out/bench_fused_products.log
timeout 600 python tools/bench_fused.py reddit 1.0 > gpurun_out/bench_fused_reddit.log 2>&1; echo "fused reddit exit=$?"; grep "^{'din" gpurun_out/bench_fused_reddit.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_aggregate_gemm -s 4 -c 1 -o gpurun_out/prof_fused_products -f python tools/bench_fused.py ogbn-products 1.0 2 > gpurun_out/ncu_fused.log 2>&1; echo "ncu fused exit=$?"
