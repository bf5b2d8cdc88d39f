"""Timeline of one overlapped sharded step (events per sub-step), run under torchrun on N GPUs."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from gnnadvisor_osdi21_b200 import graph, dist as gdist, _lib
os.dup2(2, 1)
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
gr = graph.lookalike("reddit", device=dev)
rp, ci = gr["row_ptr"], gr["col_idx"]
D = 64
sg = gdist.ShardedGraph(rp, ci, 32, device=dev, row_weight=gdist.default_row_weight(world)).build_tables()
sg.build_owner_shards()
peer = gdist.PeerHalo(sg, D)
x_src = torch.randn(sg.n_local, D, device=dev)
out = torch.empty(sg.n_local, D, device=dev)
lib = _lib.load()
p = lambda t: ctypes.c_void_p(t.data_ptr() if t.numel() else 0)
cur = torch.cuda.current_stream()
prio = int(os.environ.get("COMM_PRIO", "0"))
comm = torch.cuda.Stream(dev, priority=prio)
def step(record=None):
    ev = lambda: torch.cuda.Event(enable_timing=True)
    E = {}
    E["start"] = ev(); E["start"].record(cur)
    sg.write_local(peer, x_src, prescale=True)
    peer.begin_step()
    ready = torch.cuda.Event(); ready.record(cur); comm.wait_event(ready)
    E["push0"] = ev(); E["push0"].record(comm)
    x_ext = peer.push(comm)
    E["push1"] = ev(); E["push1"].record(comm)
    st = ctypes.c_void_p(cur.cuda_stream)
    for k, (q, (rp_, ci_, pp_, pn_)) in enumerate(zip(sg.owner_order, sg.owner_shards)):
        if q != sg.rank:
            peer.wait([q], cur)
            E["wait%d" % k] = ev(); E["wait%d" % k].record(cur)
        lib.gnna_aggregate_part_f32_ex(3, 0 if k == 0 else 1, p(x_ext), sg.n_ext, p(out), sg.n_local, p(rp_), p(ci_), p(sg.degrees_ext), 0.5,
                                       p(pp_), p(pn_), D, pn_.numel(), 32, 32, 4, st)
        E["agg%d" % k] = ev(); E["agg%d" % k].record(cur)
    peer.ack()
    pushed = torch.cuda.Event(); pushed.record(comm); cur.wait_event(pushed)
    E["end"] = ev(); E["end"].record(cur)
    return E
for _ in range(5): step()
torch.cuda.synchronize(); dist.barrier(device_ids=[rank])
import time
t=time.perf_counter()
Es = [step() for _ in range(20)]
host_ms = (time.perf_counter()-t)/20*1e3
torch.cuda.synchronize()
E = Es[-1]
line = "rank %d host %.3f ms/step | " % (rank, host_ms) + " ".join("%s=%.3f" % (k, E["start"].elapsed_time(v)) for k, v in E.items() if k != "start")
tot = Es[0]["start"].elapsed_time(Es[-1]["end"]) / 20
print(line + " | avg step %.3f" % tot, flush=True)
assert peer.error() == 0
peer.close(); dist.destroy_process_group()
