#!/bin/bash
# Multi-GPU A/B of the halo exchange and the partition cost model on the visible GPUs (bench.py --no-extras lines):
#   fused exchange+aggregation kernel (gated) vs per-owner sub-kernels, copy-engine exchange, push CTAs, row weights.   gpurun --gpus 4 -- bash tools/sweep_multi.sh
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NG=$(python -c "import torch; print(torch.cuda.device_count())")
run() {   # name, env...
  local name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$NG" --master-addr 127.0.0.1 --master-port 29513 \
      bench.py --gpus "$NG" --steps 30 --warmup 5 --no-extras > "gpurun_out/sweep_${NG}gpu_${name}.json" 2> "gpurun_out/sweep_${NG}gpu_${name}.err"
  echo "== $name rc=$?"
  python - "gpurun_out/sweep_${NG}gpu_${name}.json" <<'PY'
import json, sys
try:
    l = json.load(open(sys.argv[1]))
    x = l["extras"]
    print("   step %.4f ms  e2e %.4f ms  kernel-only %.4f  exchange-only %.4f  serial %.4f  edges/shard %s" % (
        l["ms_per_step"], l["e2e"]["ms_per_step"], x["ms_kernel_only"], x["ms_exchange_only"], x["ms_step_without_overlap"],
        [s["edges"] // 1000000 for s in x["shards"]]))
    print("   " + x["halo_exchange"][:150])
except Exception as e:
    print("   no line:", e)
PY
  tail -3 "gpurun_out/sweep_${NG}gpu_${name}.err" | grep -v "OMP_NUM\|^\*\*\*\|NCCL version" | head -3
}
run push_interleaved
run push_ring GNNA_PUSH_INTERLEAVE=0
run ce_parallel GNNA_HALO_CE=1
run ce_serial GNNA_HALO_CE=1 GNNA_CE_STREAMS=0
run push_interleaved_subkernels GNNA_GATED=0
