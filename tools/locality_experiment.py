"""Does vertex reordering pay on the HBM-bound aggregation?  (reference: rabbit_module, dataset.py:138-175,
README "node renumbering" study.)

A graph of ogbn-products size (2.45 M nodes, ~120 M directed edges) with PLANTED communities of 512 nodes -- 80 % of a
node's edges stay inside its community -- is aggregated (D=64 fp32, one launch of the gather, CUDA events) under three
labelings:  hidden   the communities scattered by a random permutation (what an arbitrary dataset looks like),
            rabbit   after this repository's Rabbit-Order replacement (csrc/reorder.cu) applied to the hidden labeling,
            planted  the ideal: community members adjacent.
R-MAT look-alikes have no community structure to recover, which is why this uses a planted one.

    python tools/locality_experiment.py [nodes] [--once]       (--once: one launch per labeling, for ncu --metrics dram__bytes*)
"""
import ctypes
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from gnnadvisor_osdi21_b200 import _lib, graph, ops, reorder  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 2449029
once = "--once" in sys.argv
comm, deg_in, deg_out, D = 512, 20, 5, 64
dev = torch.device("cuda:0")
lib = _lib.load()
g = torch.Generator(device=dev).manual_seed(1)
ids = torch.arange(n, device=dev)
src = torch.cat([ids.repeat_interleave(deg_in), ids.repeat_interleave(deg_out)])
dst_in = ((ids.repeat_interleave(deg_in) // comm) * comm + torch.randint(0, comm, (n * deg_in,), generator=g, device=dev)).clamp_(max=n - 1)
dst = torch.cat([dst_in, torch.randint(0, n, (n * deg_out,), generator=g, device=dev)])
keep = src != dst
src, dst = src[keep], dst[keep]
hide = torch.randperm(n, generator=g, device=dev)


def measure(name, s, d, extra=""):
    rp, ci = graph.csr_from_edges(torch.cat([s, d]), torch.cat([d, s]), n)       # symmetric, duplicates merged
    pp, pn = ops.build_part(32, rp)
    deg = ops.degrees_from_row_ptr(rp)
    X = torch.randn(n, D, device=dev, generator=torch.Generator(device=dev).manual_seed(2))
    out = torch.zeros_like(X)
    p = lambda t: ctypes.c_void_p(t.data_ptr())   # noqa: E731

    def k():
        _lib.check(lib.gnna_aggregate_part_f32_ex(0, 1, p(X), n, p(out), n, p(rp), p(ci), p(deg), 0.0, p(pp), p(pn), D, pn.numel(),
                                                  32, 32, 4, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "aggregate")
    reps = 1 if once else 20
    for _ in range(0 if once else 3):
        k()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        k()
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / reps
    rows = torch.repeat_interleave(torch.arange(n, device=dev), (rp[1:] - rp[:-1]).long())
    span = (rows - ci.long()).abs().float().mean().item()
    E = ci.numel()
    print("%-8s E=%d  avg edge span %10.0f  kernel %.3f ms  %.3e edge*dim/s  algorithmic %.0f GB/s %s"
          % (name, E, span, ms, E * D / (ms * 1e-3), (E * (D * 4 + 4) + n * (D * 4 + 8)) / (ms * 1e-3) / 1e9, extra), flush=True)


print("planted-community graph: %d nodes, communities of %d, %d in / %d out edges per node, D=%d fp32" % (n, comm, deg_in, deg_out, D), flush=True)
measure("hidden", hide[src], hide[dst])
t = time.perf_counter()
e = torch.stack([hide[src], hide[dst]]).to(torch.int32).cpu()
perm = reorder.permutation(e, n).to(dev).long()
dt = time.perf_counter() - t
measure("rabbit", perm[hide[src]], perm[hide[dst]], "(reorder of %d edges took %.1f s on %d host threads)" % (e.shape[1], dt, len(os.sched_getaffinity(0))))
measure("planted", src, dst)
