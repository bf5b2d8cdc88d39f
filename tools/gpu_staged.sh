#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "staged" > gpurun_out/pytest_staged.log 2>&1; echo "pytest staged exit=$?"; tail -3 gpurun_out/pytest_staged.log
for w in reddit ogbn-products; do for st in 0 1; do echo "$w staged=$st"; GNNA_STAGED=$st timeout 300 python tools/sweep_dims.py $w 0.5 2>&1 | grep "^{'D'" | grep -E "'D': (16|32|64|128)," | cut -c1-60; done; done > gpurun_out/staged_vs_default.txt 2>&1
cat gpurun_out/staged_vs_default.txt
