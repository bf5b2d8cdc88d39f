"""Host side of the dataset path: the reference's loader as written (GNNAdvisor/dataset.py:62-72 text
loop, :108-111 scipy coo -> csr from Python lists, :120-122 degree list) against libgnna_b200.so's (csrc/dataset.cu), on the
same files.  CPU only -- runs in the authoring container.

    OMP_WAIT_POLICY=passive python tools/dataset_path.py [edges ...]      > profiles/r02_dataset_path_cpu.txt
"""
import os
import sys
import tempfile
import time

import numpy as np
import scipy.sparse as sp
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnnadvisor_osdi21_b200 import graph  # noqa: E402


def reference_loader(path):
    """dataset.py:58-122 as written (minus dgl and .cuda())."""
    t0 = time.perf_counter()
    nodes, src_li, dst_li = set(), [], []
    with open(path, "r") as fp:
        for line in fp:
            src, dst = line.strip('\n').split()
            src, dst = int(src), int(dst)
            src_li.append(src)
            dst_li.append(dst)
            nodes.add(src)
            nodes.add(dst)
    num_edges, num_nodes = len(src_li), max(nodes) + 1
    edge_index = np.stack([src_li, dst_li])
    t1 = time.perf_counter()
    val = [1] * num_edges
    csr = sp.coo_matrix((val, edge_index), shape=(num_nodes, num_nodes)).tocsr()
    column_index, row_pointers = torch.IntTensor(csr.indices), torch.IntTensor(csr.indptr)
    t2 = time.perf_counter()
    degrees = (row_pointers[1:] - row_pointers[:-1]).tolist()
    deg = torch.sqrt(torch.FloatTensor(list(map(lambda x: x if x > 0 else 1, degrees))))
    t3 = time.perf_counter()
    return (row_pointers, column_index, deg), (t1 - t0, t2 - t1, t3 - t2)


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
        for name, a, b in zip(("parse text", "coo -> csr", "degrees"), tr, tn):
            print("  %-12s reference %8.3f s   native %7.3f s   %6.1fx" % (name, a, b, a / max(b, 1e-9)))
        print("  %-12s reference %8.3f s   native %7.3f s   %6.1fx" % ("total", sum(tr), sum(tn), sum(tr) / sum(tn)))
        assert same


if __name__ == "__main__":
    main()
