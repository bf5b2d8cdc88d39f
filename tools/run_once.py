"""A few launches of one aggregation variant on a look-alike graph -- the process ncu attaches to.
   python tools/run_once.py [reddit|ogbn-products] [f32|bf16|fused_bf16] [D] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from gnnadvisor_osdi21_b200 import graph, ops  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "reddit"
case = sys.argv[2] if len(sys.argv) > 2 else "f32"
D = int(sys.argv[3]) if len(sys.argv) > 3 else 64
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 4
dev = torch.device("cuda:0")
gr = graph.lookalike(wl, device=dev)
rp, ci = gr["row_ptr"], gr["col_idx"]
pp, pn = ops.build_part(32, rp)
deg = ops.degrees_from_row_ptr(rp)
X = torch.randn(gr["num_nodes"], D, device=dev)
a = (rp, ci, deg, pp, pn, 32, 32, 4)
for _ in range(reps):
    if case == "f32":
        ops.forward(X, torch.eye(D, device=dev), *a)            # SGEMM + prescale + aggregate (GCN)
    elif case == "bf16":
        ops.aggregate_bf16(3, ops.scale_rows_bf16(X, deg), rp, ci, deg, 1.0, pp, pn, 32, 32, 4)
    else:
        W = torch.eye(D, device=dev)
        ops.aggregate_gemm_fused(2, X.to(torch.bfloat16), W, rp, ci, None, 0.5, pp, pn, 32, 32, 4, want_agg=False)
torch.cuda.synchronize()
print("done", wl, case, D)
