"""Kernel-time breakdown of one GCN epoch on the Reddit look-alike (torch profiler / CUPTI), GPU box only.
   python tools/epoch_breakdown.py [workload] [fp32|bf16]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from torch.profiler import profile, ProfilerActivity
from gnnadvisor_osdi21_b200 import graph, ops, layers
dev = torch.device("cuda:0")
wl = sys.argv[1] if len(sys.argv) > 1 else "reddit"
gd = sys.argv[2] if len(sys.argv) > 2 else "fp32"
gr = graph.lookalike(wl, device=dev)
rp, ci = gr["row_ptr"], gr["col_idx"]
pp, pn = ops.build_part(32, rp)
deg = ops.degrees_from_row_ptr(rp)
class Info: pass
info = Info()
info.row_pointers, info.column_index, info.degrees, info.partPtr, info.part2Node = rp, ci, deg, pp, pn
info.partSize, info.dimWorker, info.warpPerBlock = 32, 32, 4
n = gr["num_nodes"]
x = torch.randn(n, gr["in_dim"], device=dev); y = torch.ones(n, dtype=torch.long, device=dev)
c1 = layers.GCNConv(gr["in_dim"], gr["hidden"], gather_dtype=gd).to(dev)
c2 = layers.GCNConv(gr["hidden"], gr["classes"], gather_dtype=gd).to(dev)
opt = torch.optim.Adam(list(c1.parameters()) + list(c2.parameters()), lr=0.01)
def train():
    opt.zero_grad()
    h = F.relu(c1(x, info)); o = F.log_softmax(c2(h, info), dim=1)
    F.nll_loss(o, y).backward(); opt.step()
for _ in range(5): train()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(5): train()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
import time
t=time.perf_counter()
for _ in range(20): train()
torch.cuda.synchronize()
print("gathered rows %s; epoch ms (wall, 20 epochs):" % gd, (time.perf_counter()-t)/20*1e3)
