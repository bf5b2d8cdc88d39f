"""partSize / dimWorker / warpPerBlock study (the reference's s7-4_1 / s7-4_2 sweeps) on the B200 kernel."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnnadvisor_osdi21_b200 import graph, ops
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "reddit"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
D = int(sys.argv[3]) if len(sys.argv) > 3 else 64
dev = torch.device("cuda:0")
gr = graph.lookalike(wl, device=dev, scale=scale)
rp, ci = gr["row_ptr"], gr["col_idx"]
deg = ops.degrees_from_row_ptr(rp)
N, E = gr["num_nodes"], ci.numel()
X = torch.randn(N, D, device=dev)
W = torch.eye(D, device=dev)
print("workload", wl, "N", N, "E", E, "D", D, "avg degree", E / N, flush=True)
out = {"workload": wl, "N": N, "E": E, "D": D, "partSize": {}, "dimWorker": {}, "warpPerBlock": {}}
for ps in (2, 4, 8, 16, 32, 64, 128, 256, 512):
    pp, pn = ops.build_part(ps, rp)
    ms = bench.timed(lambda: ops.SAG(X, rp, ci, deg, pp, pn, ps, 32, 4), 10, 3) / 10
    out["partSize"][ps] = round(ms, 4)
    print("partSize", ps, "groups", pn.numel(), "ms", round(ms, 4), flush=True)
pp, pn = ops.build_part(32, rp)
for dw in (1, 2, 4, 8, 16, 32):
    ms = bench.timed(lambda: ops.SAG(X, rp, ci, deg, pp, pn, 32, dw, 4), 10, 3) / 10
    out["dimWorker"][dw] = round(ms, 4)
    print("dimWorker", dw, "ms", round(ms, 4), ops.launch_info(D, pn.numel(), dw, 4), flush=True)
for wpb in (1, 2, 4, 8, 16):
    ms = bench.timed(lambda: ops.SAG(X, rp, ci, deg, pp, pn, 32, 32, wpb), 10, 3) / 10
    out["warpPerBlock"][wpb] = round(ms, 4)
    print("warpPerBlock", wpb, "ms", round(ms, 4), flush=True)
print(json.dumps(out))
