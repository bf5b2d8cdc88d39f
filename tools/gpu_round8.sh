#!/bin/bash
# Round-1 session 3, call 2: parity incl. the run-based kernel, A/B of the run-based kernel and the bf16 occupancy variant.
mkdir -p gpurun_out
t0=$SECONDS
timeout 420 python -m pytest tests/test_parity_gpu.py -x -q > gpurun_out/pytest_parity.log 2>&1; echo "parity exit=$? t=$((SECONDS-t0))"; tail -3 gpurun_out/pytest_parity.log
timeout 400 python tools/ab_chain.py --out gpurun_out/ab_runs.json --dims 16,32,64,128 > gpurun_out/ab_runs.log 2>&1; echo "ab exit=$? t=$((SECONDS-t0))"; tail -3 gpurun_out/ab_runs.log
cap() {  # name, kernel regex, env, run_once args...
    local name=$1 rx=$2 ev=$3; shift 3
    env $ev timeout 240 ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -o gpurun_out/$name -f python tools/run_once.py "$@" > gpurun_out/$name.log 2>&1
    echo "ncu $name exit=$? t=$((SECONDS-t0))"
    ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
    [ "$(stat -c %s gpurun_out/$name.ncu-rep 2>/dev/null || echo 0)" -gt 15000000 ] && rm -f gpurun_out/$name.ncu-rep
}
cap prof_runs8_reddit_bf16 aggregate_runs GNNA_RUNS=8 reddit bf16 64
cap prof_runs8_products_bf16 aggregate_runs GNNA_RUNS=8 ogbn-products bf16 64
cap prof_agg_products_bf16 aggregate_kernel GNNA_RUNS=0 ogbn-products bf16 64
du -sh gpurun_out
