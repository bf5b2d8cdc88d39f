"""GIN forward layer: unfused fp32 operator vs the fused tcgen05 tile (GPU box only)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnnadvisor_osdi21_b200 import graph, ops
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "ogbn-products"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
dev = torch.device("cuda:0")
gr = graph.lookalike(wl, device=dev, scale=scale)
rp, ci = gr["row_ptr"], gr["col_idx"]
pp, pn = ops.build_part(32, rp)
deg = ops.degrees_from_row_ptr(rp)
N, E = gr["num_nodes"], ci.numel()
res = {"workload": wl, "N": N, "E": E}
for din, dout in ((64, 64), (128, 128)):
    X = torch.randn(N, din, device=dev)
    W = (torch.rand(din, dout, device=dev) * 2 - 1) / dout ** 0.5
    Xb = X.to(torch.bfloat16)
    a = (rp, ci, 0.5, pp, pn, 32, 32, 4)
    t_unf = bench.timed(lambda: ops.forward_gin(X, W, *a), reps, 3) / reps
    t_agg = bench.timed(lambda: ops.SAG(X, rp, ci, deg, pp, pn, 32, 32, 4), reps, 3) / reps
    t_fus = bench.timed(lambda: ops.forward_gin_fused(X, W, *a), reps, 3) / reps
    t_fusb = bench.timed(lambda: ops.forward_gin_fused(Xb, W, *a), reps, 3) / reps
    t_fusb_noagg = bench.timed(lambda: ops.aggregate_gemm_fused(2, Xb, W, rp, ci, None, 0.5, pp, pn, 32, 32, 4, want_agg=False), reps, 3) / reps
    o1 = ops.forward_gin(X, W, *a)[0]
    o2 = ops.forward_gin_fused(X, W, *a)[0]
    rel = ((o1 - o2).abs().max() / o1.abs().max()).item()
    row = {"din": din, "dout": dout, "unfused_fp32_ms": t_unf, "aggregation_only_ms": t_agg, "fused_fp32x_ms": t_fus,
           "fused_bf16x_ms": t_fusb, "fused_bf16x_no_xagg_ms": t_fusb_noagg, "max_abs_diff_over_max": rel}
    print(row, flush=True)
    res.setdefault("rows", []).append(row)
print(json.dumps(res))
