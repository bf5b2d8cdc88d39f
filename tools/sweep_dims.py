"""Time the aggregation kernel over feature widths on one graph (run on the GPU box).
   GNNA_B200_LIB=<variant .so> python tools/sweep_dims.py [workload] [scale]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnnadvisor_osdi21_b200 import graph, ops, _lib
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "reddit"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
dev = torch.device("cuda:0")
gr = graph.lookalike(wl, device=dev, scale=scale)
rp, ci = gr["row_ptr"], gr["col_idx"]
pp, pn = ops.build_part(32, rp)
deg = ops.degrees_from_row_ptr(rp)
E, N = ci.numel(), gr["num_nodes"]
res = {"lib": _lib.LIB_PATH, "workload": wl, "N": N, "E": E, "rows": []}
for D in (16, 32, 41, 47, 64, 100, 128, 172, 256, 602):
    X = torch.randn(N, D, device=dev)
    row = {"D": D}
    for name, fn in (("SAG", lambda: ops.SAG(X, rp, ci, deg, pp, pn, 32, 32, 8)),
                     ("GCNfwd_I", None)):
        if fn is None:
            continue
        ms = bench.timed(fn, 10, 3) / 10
        row[name + "_ms"] = round(ms, 4)
        row[name + "_GBs"] = round(bench.alg_bytes(E, N, D, pn.numel()) / ms / 1e6, 0)
    for wpb in (4, 16):
        ms = bench.timed(lambda: ops.SAG(X, rp, ci, deg, pp, pn, 32, 32, wpb), 10, 3) / 10
        row["SAG_wpb%d_ms" % wpb] = round(ms, 4)
    Xb = X.to(torch.bfloat16)
    ms = bench.timed(lambda: ops.aggregate_bf16(0, Xb, rp, ci, deg, 1.0, pp, pn, 32, 32, 8), 10, 3) / 10
    row["SAG_bf16_ms"] = round(ms, 4)
    res["rows"].append(row)
    print(row, flush=True)
print(json.dumps(res))
