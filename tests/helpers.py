"""Shared test inputs: seeded graphs + the parity tolerance of the path."""
import numpy as np
import torch

from gnnadvisor_osdi21_b200 import graph

# BASELINE.json north_star: "outputs match the reference CUDA path within 1e-4 relative fp32".
# Reference and product both merge neighbour-groups in arbitrary order, so element-wise relative
# error is taken against max(|ref|, 1e-3 * ||ref||_inf) (SURVEY.md 8c parity policy).
RTOL = 1e-4


def assert_close(got, ref, rtol=RTOL, what=""):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    if ref.size == 0:
        return
    scale = np.maximum(np.abs(ref), 1e-3 * np.abs(ref).max())
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
