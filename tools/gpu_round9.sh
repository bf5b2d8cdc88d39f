#!/bin/bash
# Round-1 session 3, call 3: the library's own choice of the run-based kernel (default -1): full suite, A/B, bench, captures.
mkdir -p gpurun_out
t0=$SECONDS
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$? t=$((SECONDS-t0))"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python tools/ab_chain.py --out gpurun_out/ab_auto.json --runs=-1,4 --dims 16,32,48,64,128 > gpurun_out/ab_auto.log 2>&1; echo "ab exit=$? t=$((SECONDS-t0))"; tail -2 gpurun_out/ab_auto.log
timeout 420 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_reddit.json 2> gpurun_out/bench_reddit.err; echo "bench exit=$? t=$((SECONDS-t0))"; tail -2 gpurun_out/bench_reddit.err; head -c 300 gpurun_out/bench_reddit.json; echo
timeout 200 python tools/epoch_breakdown.py reddit bf16 > gpurun_out/epoch_breakdown_bf16.txt 2>&1; echo "breakdown exit=$? t=$((SECONDS-t0))"; tail -1 gpurun_out/epoch_breakdown_bf16.txt
cap() {  # name, kernel regex, env, run_once args...
    local name=$1 rx=$2 ev=$3; shift 3
    env $ev timeout 240 ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -o gpurun_out/$name -f python tools/run_once.py "$@" > gpurun_out/$name.log 2>&1
    echo "ncu $name exit=$? t=$((SECONDS-t0))"
    ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
    [ "$(stat -c %s gpurun_out/$name.ncu-rep 2>/dev/null || echo 0)" -gt 15000000 ] && rm -f gpurun_out/$name.ncu-rep
}
cap prof_auto_reddit_bf16 'aggregate_runs|aggregate_kernel' GNNA_RUNS=-1 reddit bf16 64
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -1 gpurun_out/smoke.log
du -sh gpurun_out
