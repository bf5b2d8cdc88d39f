"""Golden vectors for build_part, produced by the REFERENCE's own build_part (GNNAdvisor.cpp:210-251)
compiled from /root/reference (oracle/build_ref.py -> oracle/_ref/GNNAdvisor_ref.so).  Runs in the
authoring container only (build_part is CPU code; /root/reference must be mounted).

    python oracle/make_golden_build_part.py        ->  tests/golden/build_part.npz

Cases: seeded degree sequences (ragged, zero-degree nodes, last node isolated = SURVEY.md F6,
empty graph, exact multiples of partSize) and the one small real graph that ships in the reference
tree (Gunrock/gunrock/dataset/small/chesapeake.mtx, 39 nodes, DIMACS10; symmetrised, stored here as
its CSR row pointer).  For every case and partSize the file holds the reference's two FLOAT32
tensors verbatim.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import build_ref  # noqa: E402

PART_SIZES = (1, 2, 3, 8, 32, 64)


def cases():
    rng = np.random.default_rng(20211)
    out = {}

    def add(name, deg):
        out[name] = np.concatenate([[0], np.cumsum(np.asarray(deg, dtype=np.int64))]).astype(np.int32)

    add("ragged200", rng.integers(0, 100, 200))
    d = rng.integers(0, 100, 200); d[-1] = 0
    add("last_isolated", d)                                  # F6
    d = rng.integers(0, 40, 300); d[::3] = 0; d[-1] = 5
    add("many_isolated", d)
    add("multiples", np.tile([0, 32, 64, 96, 128], 20))      # deg % partSize == 0 branches (:222-223)
    add("single_node", [7])
    add("single_isolated", [0])
    add("no_edges", np.zeros(10, dtype=np.int64))
    add("hub", np.concatenate([[5000], rng.integers(0, 5, 100), [3]]))
    mtx = "/root/reference/Gunrock/gunrock/dataset/small/chesapeake.mtx"
    rows = [l.split() for l in open(mtx) if not l.startswith("%")]
    n = int(rows[0][0])
    e = np.array([[int(a) - 1, int(b) - 1] for a, b in rows[1:]], dtype=np.int64)
    src = np.concatenate([e[:, 0], e[:, 1]]); dst = np.concatenate([e[:, 1], e[:, 0]])
    key = np.unique(src * n + dst)
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(indptr, key // n + 1, 1)
    out["chesapeake"] = np.cumsum(indptr).astype(np.int32)
    out["chesapeake_col_idx"] = (key % n).astype(np.int32)
    return out


def main():
    build_ref.build()
    ref = build_ref.load_ref()
    data = {}
    for name, indptr in cases().items():
        data["indptr/" + name] = indptr
        if name.endswith("_col_idx"):
            continue
        for ps in PART_SIZES:
            pp, pn = ref.build_part(ps, torch.from_numpy(indptr))
            assert pp.dtype == torch.float32 and pn.dtype == torch.float32
            data["partPtr/%s/%d" % (name, ps)] = pp.numpy()
            data["part2Node/%s/%d" % (name, ps)] = pn.numpy()
    out = os.path.join(HERE, "..", "tests", "golden", "build_part.npz")
    np.savez_compressed(out, **data)
    print("wrote", os.path.abspath(out), len(data), "arrays", os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
