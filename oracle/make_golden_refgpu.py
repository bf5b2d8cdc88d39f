"""Golden vectors for the five float operators, produced by the REFERENCE's own CUDA kernels
(GNNAdvisor_kernel.cu, compiled unmodified-but-for-the-5-line-torch-2 patch by oracle/build_ref.py)
executed on a B200.  The reference ships no golden vectors or known-answer tests for this path
(SURVEY.md 8c), so these files are what pins the oracle and the product to the reference.

Run on the GPU box (needs a GPU and the prebuilt oracle/_ref/GNNAdvisor_ref.so; does not read
/root/reference):

    gpurun -- python oracle/make_golden_refgpu.py          ->  gpurun_out/refgpu.npz
    cp gpurun_out/refgpu.npz tests/golden/refgpu.npz       (committed)

Every case stores its inputs next to the reference's outputs, so the tests do not depend on RNG
reproducibility across machines.  The reference merges neighbour-groups with float atomics in
arbitrary order, so the stored outputs are ONE valid sample; comparisons use the tolerance stated
in tests/test_golden.py (1e-4 relative, BASELINE.json north_star).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import build_ref  # noqa: E402
from gnnadvisor_osdi21_b200 import graph  # noqa: E402  (graph generator only; no product kernels run here)


def hub_graph(n, hub_degree, extra_edges, seed):
    """Node 0 adjacent to `hub_degree` others plus a uniform background: one row of hub_degree/partSize groups."""
    g = torch.Generator().manual_seed(seed)
    others = torch.randperm(n - 1, generator=g)[:hub_degree] + 1
    s = torch.randint(0, n, (extra_edges,), generator=g)
    d = torch.randint(0, n, (extra_edges,), generator=g)
    keep = s != d
    src = torch.cat([torch.zeros_like(others), others, s[keep], d[keep], torch.tensor([0, n - 1])])
    dst = torch.cat([others, torch.zeros_like(others), d[keep], s[keep], torch.tensor([n - 1, 0])])
    return graph.csr_from_edges(src, dst, n)


def case_graphs():
    g = {}
    bp = np.load(os.path.join(ROOT, "tests", "golden", "build_part.npz"))
    g["chesapeake"] = (torch.from_numpy(bp["indptr/chesapeake"]), torch.from_numpy(bp["indptr/chesapeake_col_idx"]))
    g["rmat400"] = graph.synth_graph(400, 6000, kind="rmat", seed=7)
    g["uniform500"] = graph.synth_graph(500, 3000, kind="uniform", seed=8)
    g["rmat200"] = graph.synth_graph(200, 3000, kind="rmat", seed=9)
    g["uniform64"] = graph.synth_graph(64, 400, kind="uniform", seed=10)
    g["uniform32"] = graph.synth_graph(32, 160, kind="uniform", seed=12)
    g["hub2200"] = hub_graph(2200, 2100, 4000, 11)          # a row of 1050 groups at partSize 2, 5 at 512
    return g


ALL = ("SAG", "gcn", "gin")
CASES = [
    # graph        din  dout partSize dimWorker warpPerBlock  operators
    ("chesapeake", 16, 16, 32, 16, 4, ALL),
    ("chesapeake", 5, 3, 2, 3, 1, ALL),
    ("rmat400", 32, 16, 8, 16, 4, ALL),
    ("rmat400", 64, 64, 32, 32, 8, ALL),
    ("uniform500", 41, 7, 32, 32, 2, ALL),
    # round 2: the widths of BASELINE.json's configurations, a hub row of > 1000 groups, the extreme part sizes
    ("rmat200", 100, 64, 32, 32, 2, ALL),          # ogbn-products GIN layer 1 (Din = 100)
    ("rmat200", 128, 128, 32, 32, 8, ALL),         # ogbn-papers100M hidden
    ("rmat200", 128, 172, 32, 32, 8, ALL),         # ogbn-papers100M classes
    ("uniform64", 602, 64, 32, 32, 8, ALL),        # Reddit layer 1 (602 -> 64)
    ("uniform32", 3703, 16, 32, 32, 2, ("gin",)),  # citeseer GIN layer 1: a 3703-wide gathered row, warpPerBlock 2 (README.md:183)
    ("hub2200", 8, 8, 2, 8, 4, ALL),               # hub row of 1050 groups merged by atomics, partSize 2
    ("hub2200", 16, 12, 512, 16, 4, ALL),          # partSize 512
]
X_STORE_LIMIT = 20_000       # wider inputs are regenerated from graph.stream_features(seed) by the tests instead of stored


def main():
    ref = build_ref.load_ref()
    assert ref is not None, "oracle/_ref/GNNAdvisor_ref.so missing: run oracle/build_ref.py in the authoring container"
    dev = torch.device("cuda:0")
    graphs = case_graphs()
    out = {}
    gen = torch.Generator().manual_seed(20212)
    for ci, (gname, din, dout, ps, dw, wpb, what) in enumerate(CASES):
        rp, col = graphs[gname]
        n = rp.numel() - 1
        pp_f, pn_f = ref.build_part(ps, rp)                      # the reference's own table
        pp, pn = pp_f.int().to(dev), pn_f.int().to(dev)          # GNNA_main.py:109-110
        deg = torch.sqrt(torch.clamp((rp[1:] - rp[:-1]).float(), min=1.0)).to(dev)
        rp_d, col_d = rp.to(dev), col.to(dev)
        k = "case%d/" % ci
        if n * din > X_STORE_LIMIT:
            X = graph.stream_features(torch.arange(n), din, seed=1000 + ci)
            out[k + "X_seed"] = np.array([1000 + ci], dtype=np.int64)
        else:
            X = torch.randn(n, din, generator=gen)
            out[k + "X"] = X.numpy()
        W = (torch.rand(din, dout, generator=gen) * 2 - 1) / np.sqrt(dout)
        dO = torch.randn(n, dout, generator=gen)
        Xd, Wd, dOd = X.to(dev), W.to(dev), dO.to(dev)
        eps = 0.5
        out[k + "meta"] = np.array([din, dout, ps, dw, wpb], dtype=np.int64)
        out[k + "graph"] = np.array(gname)
        out[k + "ops"] = np.array(",".join(what))
        out[k + "row_ptr"], out[k + "col_idx"] = rp.numpy(), col.numpy()
        out[k + "partPtr"], out[k + "part2Node"] = pp.cpu().numpy(), pn.cpu().numpy()
        out[k + "W"], out[k + "dO"] = W.numpy(), dO.numpy()
        if "SAG" in what:
            out[k + "SAG"] = ref.SAG(Xd, rp_d, col_d, deg, pp, pn, ps, dw, wpb).cpu().numpy()
        if "gcn" in what:
            out[k + "forward"] = ref.forward(Xd, Wd, rp_d, col_d, deg, pp, pn, ps, dw, wpb)[0].cpu().numpy()
            dX, dW = ref.backward(dOd, Xd, Wd, rp_d, col_d, deg, pp, pn, ps, dw, wpb)
            out[k + "backward_dX"], out[k + "backward_dW"] = dX.cpu().numpy(), dW.cpu().numpy()
        if "gin" in what:
            o, xagg = ref.forward_gin(Xd, Wd, rp_d, col_d, eps, pp, pn, ps, dw, wpb)
            out[k + "forward_gin"], out[k + "forward_gin_agg"] = o.cpu().numpy(), xagg.cpu().numpy()
            dXg, dWg = ref.backward_gin(dOd, xagg, Wd, rp_d, col_d, eps, pp, pn, ps, dw, wpb)
            out[k + "backward_gin_dX"], out[k + "backward_gin_dW"] = dXg.cpu().numpy(), dWg.cpu().numpy()
        torch.cuda.synchronize()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", "refgpu.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", torch.cuda.get_device_name(0))


if __name__ == "__main__":
    main()
