"""5-layer GIN epoch (GNNA_main.py:155-171) on a look-alike graph; config #4 of BASELINE.json.
   python tools/gin_epoch.py [workload] [scale] [fp32|bf16] [dimWorker,dimWorker,...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from gnnadvisor_osdi21_b200 import graph, ops, layers
wl = sys.argv[1] if len(sys.argv) > 1 else "ogbn-products"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
gd = sys.argv[3] if len(sys.argv) > 3 else "fp32"
dws = tuple(int(v) for v in sys.argv[4].split(",")) if len(sys.argv) > 4 else (32, 16, 8, 4)
dev = torch.device("cuda:0")
gr = graph.lookalike(wl, device=dev, scale=scale)
rp, ci = gr["row_ptr"], gr["col_idx"]
pp, pn = ops.build_part(32, rp)
deg = ops.degrees_from_row_ptr(rp)
class Info: pass
n = gr["num_nodes"]
x = torch.randn(n, gr["in_dim"], device=dev); y = torch.ones(n, dtype=torch.long, device=dev)
hid, cls = gr["hidden"], gr["classes"]
convs = ([layers.GINConv(gr["in_dim"], hid, gather_dtype=gd)] + [layers.GINConv(hid, hid, gather_dtype=gd) for _ in range(3)]
         + [layers.GINConv(hid, cls, gather_dtype=gd)])
convs = [c.to(dev) for c in convs]
opt = torch.optim.Adam([p for c in convs for p in c.parameters()], lr=0.01)
for dw in dws:
    info = Info()
    info.row_pointers, info.column_index, info.degrees, info.partPtr, info.part2Node = rp, ci, deg, pp, pn
    info.partSize, info.dimWorker, info.warpPerBlock = 32, dw, 4
    def train():
        opt.zero_grad()
        h = x
        for i, c in enumerate(convs):
            h = c(h, info)
            if i < len(convs) - 1: h = F.relu(h)
        F.nll_loss(F.log_softmax(h, dim=1), y).backward(); opt.step()
    for _ in range(3): train()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(10): train()
    torch.cuda.synchronize()
    print("GIN-5 %s N=%d E=%d dims %d-%d-%d gather=%s dimWorker=%d: epoch ms %.3f" % (wl, n, ci.numel(), gr["in_dim"], hid, cls, gd, dw, (time.perf_counter() - t) / 10 * 1e3), flush=True)
