"""Host side of the dataset path: the reference's own loader, imported unchanged (GNNAdvisor/dataset.py:
:62-72 text loop, :108-111 scipy coo -> csr from Python lists, :120-122 degree list) against libgnna_b200.so's
(csrc/dataset.cu), on the same files.  CPU only -- runs in the authoring container, where /root/reference is mounted.

    OMP_WAIT_POLICY=passive python tools/dataset_path.py [edges ...]      > profiles/r02_dataset_path_cpu.txt
"""
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnnadvisor_osdi21_b200 import graph  # noqa: E402


def reference_loader(path):
    """The reference's own loader, imported UNCHANGED from /root/reference (compat/ provides dgl and rabbit, `.cuda()` is
    made a no-op: no GPU here) and timed by its own verbose phase timers (dataset.py:74-79 loading, :109-115 CSR build);
    the degree list (:120-122) is what is left of init_edges."""
    import contextlib
    import importlib
    import io
    import re
    compat = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gnnadvisor_osdi21_b200", "compat")
    sys.path[:0] = [compat, "/root/reference/GNNAdvisor"]
    try:
        ref = importlib.import_module("dataset")
    finally:
        del sys.path[:2]
    torch.Tensor.cuda = lambda self, *a, **k: self
    obj = ref.custom_dataset.__new__(ref.custom_dataset)
    torch.nn.Module.__init__(obj)
    obj.nodes, obj.load_from_txt, obj.verbose_flag = set(), True, True
    text = io.StringIO()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(text):
        obj.init_edges(path)
    total = time.perf_counter() - t0
    load = float(re.search(r"# Loading \(txt\) ([0-9.]+)s", text.getvalue()).group(1))
    csr = float(re.search(r"# Build CSR after reordering \(s\): ([0-9.]+)", text.getvalue()).group(1))
    return (obj.row_pointers, obj.column_index, obj.degrees), (load, csr, total - load - csr)


def native_loader(path):
    t0 = time.perf_counter()
    src, dst, n = graph.load_edge_text(path)
    t1 = time.perf_counter()
    rp, ci = graph.csr_from_edges(torch.from_numpy(src), torch.from_numpy(dst), n, native=True)
    t2 = time.perf_counter()
    deg = graph.degrees_from_row_ptr_host(rp)
    t3 = time.perf_counter()
    return (rp, ci, deg), (t1 - t0, t2 - t1, t3 - t2)


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [4878874, 40000000]
    threads = len(os.sched_getaffinity(0))
    print("dataset path on the host, %d threads (authoring container, no GPU involved); OMP_WAIT_POLICY=%s"
          % (threads, os.environ.get("OMP_WAIT_POLICY", "(unset)")))
    for e in sizes:
        n = max(16, e // 12)
        s, d = graph.stream_pairs(n, 0, e, kind="rmat", seed=7)
        s, d = s.numpy(), d.numpy()
        with tempfile.TemporaryDirectory() as tmp:
            path = os.path.join(tmp, "g.txt")
            with open(path, "w") as f:
                step = 1 << 20
                for i in range(0, len(s), step):
                    f.write("".join("%d %d\n" % p for p in zip(s[i:i + step].tolist(), d[i:i + step].tolist())))
            mb = os.path.getsize(path) / 1e6
            (rp, ci, deg), tn = native_loader(path)           # first: it also warms the page cache for the reference
            (rrp, rci, rdeg), tr = reference_loader(path)
        same = torch.equal(rp, rrp) and torch.equal(ci, rci) and torch.equal(deg, rdeg)
        print("\n%d edge lines, %d nodes, %.0f MB of text -> CSR with %d edges; native result == reference result: %s"
              % (len(s), rp.numel() - 1, mb, ci.numel(), same))
        for name, a, b in zip(("parse text", "coo -> csr", "rest*"), tr, tn):
            print("  %-12s reference %8.3f s   native %7.3f s   %6.1fx" % (name, a, b, a / max(b, 1e-9)))
        print("  %-12s reference %8.3f s   native %7.3f s   %6.1fx" % ("total", sum(tr), sum(tn), sum(tr) / sum(tn)))
        print("  (* reference: everything else in init_edges -- np.stack of the two lists, edge-span statistics, the [1]*E list, the degree map;"
              " native: the degree vector)")
        assert same


if __name__ == "__main__":
    main()
