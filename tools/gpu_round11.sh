#!/bin/bash
# last check of round 1: the full GPU suite + smoke on the committed tree
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest exit=$?"; tail -2 gpurun_out/pytest_gpu_final.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
