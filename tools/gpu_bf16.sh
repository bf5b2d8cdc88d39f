#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "bf16 or fused" > gpurun_out/pytest_bf16.log 2>&1; echo "pytest bf16 exit=$?"; tail -3 gpurun_out/pytest_bf16.log
timeout 300 python tools/sweep_dims.py reddit 0.5 2>&1 | grep "^{'D'" | grep -E "'D': (16|32|64|128|256)," | sed -E "s/.*'D': ([0-9]+),.*'SAG_ms': ([0-9.]+),.*'SAG_bf16_ms': ([0-9.]+).*/D=\1 fp32 \2 ms  bf16 \3 ms/"
timeout 300 python tools/sweep_dims.py ogbn-products 0.5 2>&1 | grep "^{'D'" | grep -E "'D': (16|32|64|128|256)," | sed -E "s/.*'D': ([0-9]+),.*'SAG_ms': ([0-9.]+),.*'SAG_bf16_ms': ([0-9.]+).*/D=\1 fp32 \2 ms  bf16 \3 ms/"
