"""Shared test inputs: seeded graphs + the parity tolerance of the path."""
import numpy as np
import torch

from gnnadvisor_osdi21_b200 import graph

# BASELINE.json north_star: "outputs match the reference CUDA path within 1e-4 relative fp32".
# Reference and product both merge neighbour-groups in arbitrary order, so element-wise relative
# error is taken against max(|ref|, 1e-3 * ||ref||_inf) (SURVEY.md 8c parity policy).
RTOL = 1e-4


def assert_close(got, ref, rtol=RTOL, what="", terms=None):
    """|got - ref| <= rtol * scale, scale = max(|ref|, 1e-3*||ref||_inf).
    `terms` (optional, same shape): the sum of the ABSOLUTE values of the terms each output element
    adds up (e.g. |G| @ |W^T| for dX = G @ W^T).  Where an element is the result of cancellation
    (|ref| << terms) the summation order -- which neither the reference's atomics nor cuBLAS fix --
    decides more than rtol*|ref|; the error is then bounded relative to the terms instead."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    if ref.size == 0:
        return
    scale = np.maximum(np.abs(ref), 1e-3 * np.abs(ref).max())
    if terms is not None:
        scale = np.maximum(scale, 0.1 * np.asarray(terms, dtype=np.float64))
    scale = np.maximum(scale, 1e-30)
    err = np.abs(got - ref) / scale
    assert np.isfinite(got).all(), what + ": non-finite output"
    assert err.max() <= rtol, "%s: max rel err %.3e at %s (got %r ref %r)" % (
        what, err.max(), np.unravel_index(err.argmax(), err.shape), got.flat[err.argmax()], ref.flat[err.argmax()])


def make_graph(kind, n, e, seed):
    rp, ci = graph.synth_graph(n, e, kind=kind, seed=seed)
    return rp.numpy(), ci.numpy()


def rand_features(n, d, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, d, generator=g).numpy()


def rand_weight(din, dout, seed):
    g = torch.Generator().manual_seed(seed)
    return ((torch.rand(din, dout, generator=g) * 2 - 1) / np.sqrt(dout)).numpy()


def golden_case(g, k):
    """Inputs and the operator list of one case of tests/golden/refgpu.npz.  Inputs too wide to store are regenerated from
    the device-independent generator they were made with (oracle/make_golden_refgpu.py: X_seed)."""
    rp = g[k + "row_ptr"]
    din = int(g[k + "meta"][0])
    if k + "X" in g.files:
        X = g[k + "X"]
    else:
        X = graph.stream_features(torch.arange(len(rp) - 1), din, seed=int(g[k + "X_seed"][0])).numpy()
    what = str(g[k + "ops"]).split(",") if k + "ops" in g.files else ["SAG", "gcn", "gin"]
    return X, g[k + "W"], g[k + "dO"], what


def golden_terms(g, k, oracle):
    """Absolute-term bounds (see assert_close) for the dense-product outputs of one golden case."""
    f64 = np.float64
    rp, ci = g[k + "row_ptr"], g[k + "col_idx"]
    X, W, dO, what = golden_case(g, k)
    X, W, dO = np.abs(X).astype(f64), np.abs(W).astype(f64), np.abs(dO).astype(f64)
    t = {}
    if "gcn" in what:
        aG = oracle.closed_form(1, dO, rp, ci)                       # sum of |terms| of G = Ahat @ dO
        t.update({"fwd": oracle.closed_form(1, X @ W, rp, ci), "dX": aG @ W.T, "dW": X.T @ aG})
    if "gin" in what:
        aS = np.abs(g[k + "forward_gin_agg"]).astype(f64)
        t.update({"gin_out": oracle.closed_form(2, X, rp, ci, 0.5) @ W,
                  "gin_dX": oracle.closed_form(2, dO @ W.T, rp, ci, 0.5), "gin_dW": aS.T @ dO})
    return t
