#!/bin/bash
# Round-1 session 3: parity of the changed kernels, A/B of the chain variants, bench of the winner, rest of the suite,
# ncu captures.  Ordered by priority: the call may be cut when the round's GPU budget runs out.
mkdir -p gpurun_out
nvidia-smi -L
t0=$SECONDS
timeout 420 python -m pytest tests/test_parity_gpu.py -x -q > gpurun_out/pytest_parity.log 2>&1; echo "parity exit=$? t=$((SECONDS-t0))"; tail -3 gpurun_out/pytest_parity.log
timeout 300 python tools/ab_chain.py > gpurun_out/ab_chain.log 2>&1; echo "ab exit=$? t=$((SECONDS-t0))"; tail -4 gpurun_out/ab_chain.log
# headline case decides which build the bench and the captures use (another build must win by > 2 %)
WIN=$(python - <<'PY'
import json, os
try:
    rows = json.load(open("gpurun_out/ab_chain.json"))
    r = [x for x in rows["rows"] if x["workload"] == "reddit" and x["D"] == 64 and x["case"] == "gcn_f32"][0]
    t = {k[:-3]: v for k, v in r.items() if k.endswith("_ms")}
    best = min(t, key=t.get)
    print(rows["libs"][best] if t[best] < 0.98 * t["A_default"] else "")
except Exception:
    print("")
PY
)
echo "lib for bench/ncu: ${WIN:-default}"
[ -n "$WIN" ] && export GNNA_B200_LIB=$WIN
timeout 420 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_reddit.json 2> gpurun_out/bench_reddit.err; echo "bench exit=$? t=$((SECONDS-t0))"; tail -2 gpurun_out/bench_reddit.err; head -c 400 gpurun_out/bench_reddit.json; echo
timeout 400 python -m pytest tests -m gpu -x -q --deselect tests/test_parity_gpu.py > gpurun_out/pytest_rest.log 2>&1; echo "rest exit=$? t=$((SECONDS-t0))"; tail -3 gpurun_out/pytest_rest.log
cap() {  # name, kernel regex, run_once args...
    local name=$1 rx=$2; shift 2
    timeout 240 ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -o gpurun_out/$name -f python tools/run_once.py "$@" > gpurun_out/$name.log 2>&1
    echo "ncu $name exit=$? t=$((SECONDS-t0))"
    ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
    # gpurun_out is merged back only up to 64 MiB: keep the report if it is small, the csv always
    [ "$(stat -c %s gpurun_out/$name.ncu-rep 2>/dev/null || echo 0)" -gt 15000000 ] && rm -f gpurun_out/$name.ncu-rep
}
cap prof_agg_reddit_f32 aggregate_kernel reddit f32 64
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'aggregate_kernel|repack|unpack|part_|degrees_kernel|scale_rows' -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 1 --no-extras > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit=$? t=$((SECONDS-t0))"
cap prof_agg_reddit_bf16 aggregate_kernel reddit bf16 64
timeout 240 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench ref exit=$? t=$((SECONDS-t0))"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -1 gpurun_out/smoke.log
(timeout 200 python tools/gin_epoch.py ogbn-products 1.0 fp32 32; timeout 200 python tools/gin_epoch.py ogbn-products 1.0 bf16 32) > gpurun_out/gin_epoch.log 2>&1; echo "gin epoch t=$((SECONDS-t0))"; grep GIN-5 gpurun_out/gin_epoch.log
cap prof_fused_reddit_bf16 fused_aggregate reddit fused_bf16 64
du -sh gpurun_out
