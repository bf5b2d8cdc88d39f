#!/bin/bash
# One GPU session = a list of named legs; everything is written to gpurun_out/ (summaries worth keeping are copied to
# profiles/ by hand afterwards).  Usage on the box:   gpurun --timeout 1500 -- bash tools/gpu_session.sh golden tests bench
# Legs:
#   golden     regenerate tests/golden/refgpu.npz candidates with the reference's own kernels -> gpurun_out/refgpu.npz
#   tests      python -m pytest tests -m gpu                                -> gpurun_out/pytest_gpu.log
#   smoke      __graft_entry__.smoke()
#   bench      python bench.py --steps 20 --warmup 5                        -> gpurun_out/bench_1gpu.json
#   benchN     torchrun bench.py --gpus N (N = number of visible GPUs)      -> gpurun_out/bench_<N>gpu.json
#   refarm     python bench.py --impl reference                             -> gpurun_out/bench_reference.json
#   launches   ncu launch list of a short bench run                         -> gpurun_out/launches.csv
#   papers     bench.py --workload ogbn-papers100M --dim 128 on all visible GPUs (scale from $PAPERS_SCALE, default 1.0)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NG=$(python -c "import torch; print(torch.cuda.device_count())")
echo "visible GPUs: $NG"
for leg in "$@"; do
  echo "=== leg: $leg ($(date +%T))"
  case "$leg" in
    golden)  timeout 600 python oracle/make_golden_refgpu.py 2>&1 | tail -3 ;;
    tests)   timeout 2400 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_gpu.log ;;
    smoke)   timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ;;
    bench)   timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "rc=$?"; tail -c 1500 gpurun_out/bench_1gpu.json; tail -5 gpurun_out/bench_1gpu.err ;;
    benchN)  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$NG" --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus "$NG" --steps 20 --warmup 5 > "gpurun_out/bench_${NG}gpu.json" 2> "gpurun_out/bench_${NG}gpu.err"; echo "rc=$?"; tail -c 2500 "gpurun_out/bench_${NG}gpu.json"; tail -8 "gpurun_out/bench_${NG}gpu.err" ;;
    refarm)  timeout 900 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "rc=$?"; tail -c 800 gpurun_out/bench_reference.json ;;
    launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/launches_bench.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/launches.csv ;;
    papers)  S=${PAPERS_SCALE:-1.0}; timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$NG" --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus "$NG" --workload ogbn-papers100M --dim 128 --scale "$S" --steps 5 --warmup 3 > "gpurun_out/papers_x${S}_${NG}gpu.json" 2> "gpurun_out/papers_x${S}_${NG}gpu.err"; echo "rc=$?"; tail -c 3000 "gpurun_out/papers_x${S}_${NG}gpu.json"; tail -12 "gpurun_out/papers_x${S}_${NG}gpu.err" ;;
    *) echo "running: $leg"; timeout 1500 bash -c "$leg" ;;
  esac
done
echo "=== done ($(date +%T))"
