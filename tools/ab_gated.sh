cd /root/repo
for g in 1 0 1 0; do
  GNNA_GATED=$g python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --steps 30 --warmup 5 --no-extras > gpurun_out/ab_gated_$g.json 2>/dev/null
  python - <<PY
import json
l=json.load(open("gpurun_out/ab_gated_$g.json")); x=l["extras"]
print("gated=$g step %.4f e2e %.4f kern %.4f exch %.4f serial %.4f" % (l["ms_per_step"], l["e2e"]["ms_per_step"], x["ms_kernel_only"], x["ms_exchange_only"], x["ms_step_without_overlap"]))
PY
done
