#!/bin/bash
# Round-1 session 3, call 5: final state -- full suite, both bench arms, launch list + full capture of the hot kernel,
# the small BASELINE configurations (#1 Cora CPU arm, #2 citeseer epoch with rabbit reorder, #5 papers look-alike x0.05).
mkdir -p gpurun_out
t0=$SECONDS
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$? t=$((SECONDS-t0))"; tail -2 gpurun_out/pytest_gpu.log
timeout 420 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_reddit.json 2> gpurun_out/bench_reddit.err; echo "bench exit=$? t=$((SECONDS-t0))"; tail -2 gpurun_out/bench_reddit.err; head -c 300 gpurun_out/bench_reddit.json; echo
timeout 240 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench ref exit=$? t=$((SECONDS-t0))"
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'aggregate|repack|unpack|part_|degrees_kernel|scale_rows|prescale' -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 1 --no-extras > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit=$? t=$((SECONDS-t0))"
cap() {  # name, kernel regex, env, run_once args...
    local name=$1 rx=$2 ev=$3; shift 3
    env $ev timeout 240 ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -o gpurun_out/$name -f python tools/run_once.py "$@" > gpurun_out/$name.log 2>&1
    echo "ncu $name exit=$? t=$((SECONDS-t0))"
    ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
    [ "$(stat -c %s gpurun_out/$name.ncu-rep 2>/dev/null || echo 0)" -gt 15000000 ] && rm -f gpurun_out/$name.ncu-rep
}
cap prof_final_reddit_f32 aggregate_kernel GNNA_RUNS=-1 reddit f32 64
timeout 200 python -m gnnadvisor_osdi21_b200.main --synthetic citeseer --dim 3703 --hidden 16 --classes 6 --model gcn --manual_mode True --enable_rabbit True --partSize 32 --dimWorker 16 --warpPerBlock 4 --num_epoches 200 > gpurun_out/config2_citeseer.log 2>&1; echo "citeseer exit=$? t=$((SECONDS-t0))"; grep "Time (ms)" gpurun_out/config2_citeseer.log
timeout 120 python bench.py --impl reference --workload cora --dim 16 --steps 3 --warmup 1 > gpurun_out/config1_cora_cpu.json 2> gpurun_out/config1_cora_cpu.err; echo "cora cpu exit=$? t=$((SECONDS-t0))"
timeout 120 python bench.py --workload cora --dim 16 --steps 200 --warmup 20 --no-extras > gpurun_out/config1_cora_gpu.json 2> gpurun_out/config1_cora_gpu.err; echo "cora gpu exit=$? t=$((SECONDS-t0))"
timeout 300 python bench.py --workload ogbn-papers100M --scale 0.05 --dim 128 --steps 20 --warmup 3 --no-extras > gpurun_out/config5_papers_x0.05_1gpu.json 2> gpurun_out/config5_papers_1gpu.err; echo "papers exit=$? t=$((SECONDS-t0))"; head -c 250 gpurun_out/config5_papers_x0.05_1gpu.json; echo
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -1 gpurun_out/smoke.log
du -sh gpurun_out
