"""A/B of the fused aggregate -> X*W tile (bf16 rows, D=64, Reddit look-alike): this build against the unfused bf16 pair.
Run once per library build (GNNA_B200_LIB selects a variant).   python tools/ab_fused.py [reddit|ogbn-products]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
from gnnadvisor_osdi21_b200 import _lib, graph, ops  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "reddit"
dev = torch.device("cuda:0")
gr = graph.lookalike(wl, device=dev)
rp, ci = gr["row_ptr"], gr["col_idx"]
pp, pn = ops.build_part(32, rp)
deg = ops.degrees_from_row_ptr(rp)
N = gr["num_nodes"]
X = torch.randn(N, 64, device=dev)
W = (torch.rand(64, 64, device=dev) * 2 - 1) / 8
Xb = X.to(torch.bfloat16)
a = (rp, ci, 0.5, pp, pn, 32, 32, 4)
t_mixed = bench.timed(lambda: ops.forward_gin_mixed(X, W, *a), 20, 3) / 20
t_fused = bench.timed(lambda: ops.forward_gin_fused(Xb, W, *a), 20, 3) / 20
t_fused_conv = bench.timed(lambda: ops.forward_gin_fused(ops.scale_rows_bf16(X), W, *a), 20, 3) / 20
o1, s1 = ops.forward_gin_mixed(X, W, *a)
o2, s2 = ops.forward_gin_fused(ops.scale_rows_bf16(X), W, *a)
print("%s lib=%s: unfused bf16-row GIN forward %.3f ms | fused tile %.3f ms (bf16 X given) / %.3f ms (incl. conversion) | x_agg max diff %.2e, out rel diff %.2e"
      % (wl, os.path.basename(_lib.LIB_PATH), t_mixed, t_fused, t_fused_conv, float((s1 - s2).abs().max()),
         float((o1 - o2).abs().max() / o1.abs().max())), flush=True)
