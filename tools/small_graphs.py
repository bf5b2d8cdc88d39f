"""Launch-bound configurations (BASELINE.json #1/#2): Cora / citeseer look-alikes.  Per-call time of one GCN aggregation
(CUDA events over 2000 back-to-back calls) with the single-launch path on and off, and the citeseer epoch of main.py eager
vs replayed from a CUDA graph.   python tools/small_graphs.py"""
import contextlib
import io
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from gnnadvisor_osdi21_b200 import _lib, graph, main as gmain, ops  # noqa: E402

dev = torch.device("cuda:0")
for name, D in (("cora", 16), ("citeseer", 16), ("citeseer", 6)):
    g = graph.lookalike(name, device=dev)
    rp, ci = g["row_ptr"], g["col_idx"]
    pp, pn = ops.build_part(32, rp)
    deg = ops.degrees_from_row_ptr(rp)
    X = torch.randn(g["num_nodes"], D, device=dev)
    import ctypes
    out = torch.empty_like(X)
    p = lambda t: ctypes.c_void_p(t.data_ptr())   # noqa: E731
    lib = _lib.load()

    def call():
        _lib.check(lib.gnna_gcn_aggregate_f32(p(X), p(out), p(rp), p(ci), p(deg), p(pp), p(pn), g["num_nodes"], D, pn.numel(), 32, 16, 8,
                                              ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "agg")
    for limit in (16384, 0):
        _lib.set_small_parts(limit)
        for _ in range(50):
            call()
        torch.cuda.synchronize()
        _lib.launch_count(reset=True)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(2000):
            call()
        t1.record()
        torch.cuda.synchronize()
        print("%-8s D=%-3d %-22s %.2f us per GCN aggregation, %d kernel launches per call (+ memset/alloc nodes on the general path)"
              % (name, D, "single-launch path" if limit else "general path", t0.elapsed_time(t1) / 2000 * 1e3, _lib.launch_count() // 2000), flush=True)
    _lib.set_small_parts(16384)

for extra in ([], ["--cuda_graph", "True"]):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        gmain.main(["--synthetic", "citeseer", "--dim", "3703", "--hidden", "16", "--classes", "6", "--num_epoches", "200",
                    "--enable_rabbit", "True", "--partSize", "32"] + extra)
    m = re.search(r"Time \(ms\): (\d+\.\d+)", buf.getvalue())
    print("citeseer GCN 3703-16-6, rabbit reorder on, partSize 32, %s: %s ms per epoch" % ("CUDA-graph epoch" if extra else "eager epoch", m.group(1) if m else "?"),
          "(captured)" if "captured in a CUDA graph" in buf.getvalue() else "", flush=True)
