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


def case_graphs():
    g = {}
    bp = np.load(os.path.join(ROOT, "tests", "golden", "build_part.npz"))
    g["chesapeake"] = (torch.from_numpy(bp["indptr/chesapeake"]), torch.from_numpy(bp["indptr/chesapeake_col_idx"]))
    g["rmat400"] = graph.synth_graph(400, 6000, kind="rmat", seed=7)
    g["uniform500"] = graph.synth_graph(500, 3000, kind="uniform", seed=8)
    return g


CASES = [
    # graph        din dout partSize dimWorker warpPerBlock
    ("chesapeake", 16, 16, 32, 16, 4),
    ("chesapeake", 5, 3, 2, 3, 1),
    ("rmat400", 32, 16, 8, 16, 4),
    ("rmat400", 64, 64, 32, 32, 8),
    ("uniform500", 41, 7, 32, 32, 2),
]


def main():
    ref = build_ref.load_ref()
    assert ref is not None, "oracle/_ref/GNNAdvisor_ref.so missing: run oracle/build_ref.py in the authoring container"
    dev = torch.device("cuda:0")
    graphs = case_graphs()
    out = {}
    gen = torch.Generator().manual_seed(20212)
    for ci, (gname, din, dout, ps, dw, wpb) in enumerate(CASES):
        rp, col = graphs[gname]
        n = rp.numel() - 1
        pp_f, pn_f = ref.build_part(ps, rp)                      # the reference's own table
        pp, pn = pp_f.int().to(dev), pn_f.int().to(dev)          # GNNA_main.py:109-110
        deg = torch.sqrt(torch.clamp((rp[1:] - rp[:-1]).float(), min=1.0)).to(dev)
        rp_d, col_d = rp.to(dev), col.to(dev)
        X = torch.randn(n, din, generator=gen)
        W = (torch.rand(din, dout, generator=gen) * 2 - 1) / np.sqrt(dout)
        dO = torch.randn(n, dout, generator=gen)
        Xd, Wd, dOd = X.to(dev), W.to(dev), dO.to(dev)
        eps = 0.5
        k = "case%d/" % ci
        out[k + "meta"] = np.array([din, dout, ps, dw, wpb], dtype=np.int64)
        out[k + "graph"] = np.array(gname)
        out[k + "row_ptr"], out[k + "col_idx"] = rp.numpy(), col.numpy()
        out[k + "partPtr"], out[k + "part2Node"] = pp.cpu().numpy(), pn.cpu().numpy()
        out[k + "X"], out[k + "W"], out[k + "dO"] = X.numpy(), W.numpy(), dO.numpy()
        out[k + "SAG"] = ref.SAG(Xd, rp_d, col_d, deg, pp, pn, ps, dw, wpb).cpu().numpy()
        out[k + "forward"] = ref.forward(Xd, Wd, rp_d, col_d, deg, pp, pn, ps, dw, wpb)[0].cpu().numpy()
        dX, dW = ref.backward(dOd, Xd, Wd, rp_d, col_d, deg, pp, pn, ps, dw, wpb)
        out[k + "backward_dX"], out[k + "backward_dW"] = dX.cpu().numpy(), dW.cpu().numpy()
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
