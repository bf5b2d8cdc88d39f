"""Vertex reordering (rabbit replacement): a valid permutation, applied consistently, that improves
locality on graphs with community structure.  CPU only (host code of the C-ABI library)."""
import os
import sys

import numpy as np
import pytest
import torch

from gnnadvisor_osdi21_b200 import reorder as R


def _span(e):
    return float((e[0].long() - e[1].long()).abs().float().mean())


def _clique_ring(nc, size, seed):
    """nc cliques of `size` vertices joined in a ring, vertex ids randomly shuffled."""
    rng = np.random.default_rng(seed)
    src, dst = [], []
    for c in range(nc):
        base = c * size
        for i in range(size):
            for j in range(i + 1, size):
                src.append(base + i); dst.append(base + j)
        src.append(base); dst.append(((c + 1) % nc) * size)
    src, dst = np.array(src), np.array(dst)
    shuffle = rng.permutation(nc * size)
    return torch.from_numpy(np.stack([shuffle[src], shuffle[dst]]).astype(np.int32))


def test_permutation_is_valid_and_consistent():
    e = _clique_ring(40, 12, 1)
    n = 480
    perm = R.permutation(e, n)
    assert perm.dtype == torch.int32 and sorted(perm.tolist()) == list(range(n))
    out = R.reorder(e)
    assert out.shape == e.shape and out.dtype == torch.int32
    assert torch.equal(out, perm[e.long()])                 # same edges, same order, relabelled endpoints
    assert torch.equal(R.reorder(e), out)                   # deterministic


def test_communities_become_contiguous():
    e = _clique_ring(60, 10, 2)
    out = R.reorder(e)
    assert _span(out) < 0.1 * _span(e)
    # every clique occupies one contiguous id range
    perm = R.permutation(e, 600)
    inv = torch.empty_like(perm); inv[perm.long()] = torch.arange(600, dtype=torch.int32)
    und = torch.cat([e, e.flip(0)], 1)
    adj = {}
    for a, b in und.t().tolist():
        adj.setdefault(a, set()).add(b)
    for c_start in range(0, 600, 10):
        members = inv[c_start:c_start + 10].tolist()
        inside = sum(len(adj[m] & set(members)) for m in members)
        assert inside >= 9 * 10 - 2          # a block of 10 new ids is (almost exactly) one clique


def test_windows_are_deterministic_for_any_thread_count():
    """The parallel aggregation evaluates a WINDOW of vertices against one state and applies their merges in list order:
    the permutation is a function of the window length only.  Same graph, 1 / 3 / all host threads, in fresh processes."""
    import subprocess
    code = ("import sys, hashlib, numpy as np, torch; sys.path.insert(0, %r); "
            "from gnnadvisor_osdi21_b200 import reorder as R; "
            "rng = np.random.default_rng(3); n = 20000; "
            "c = rng.integers(0, n // 50, 12 * n); "
            "s = c * 50 + rng.integers(0, 50, 12 * n); d = np.where(rng.random(12 * n) < 0.8, c * 50 + rng.integers(0, 50, 12 * n), rng.integers(0, n, 12 * n)); "
            "sh = rng.permutation(n); e = torch.from_numpy(np.stack([sh[s], sh[d]]).astype(np.int32)); "
            "print(' '.join(hashlib.sha1(R.permutation(e, n, window=w).numpy().tobytes()).hexdigest() for w in (0, 1, 64, 1000)))"
            % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    outs = []
    for threads in ("1", "3", ""):
        env = dict(os.environ, OMP_WAIT_POLICY="passive")
        env.pop("OMP_NUM_THREADS", None)
        if threads:
            env["OMP_NUM_THREADS"] = threads
        outs.append(subprocess.run([sys.executable, "-c", code], env=env, check=True, capture_output=True, text=True).stdout.split())
    assert outs[0] == outs[1] == outs[2] and len(outs[0]) == 4
    assert len(set(outs[0])) >= 3                               # different windows are different (but equally fixed) orders


@pytest.mark.parametrize("window", [1, 7, 64, 0])
def test_every_window_recovers_hidden_communities(window):
    """200 communities of 50 vertices, 80 % of the edges inside, ids shuffled: whatever the window, the new ids put a
    community's members next to each other (window 1 is the sequential algorithm)."""
    rng = np.random.default_rng(8)
    n, size = 10000, 50
    m = 10 * n
    c = rng.integers(0, n // size, m)
    s = c * size + rng.integers(0, size, m)
    d = np.where(rng.random(m) < 0.8, c * size + rng.integers(0, size, m), rng.integers(0, n, m))
    shuffle = rng.permutation(n)
    e = torch.from_numpy(np.stack([shuffle[s], shuffle[d]]).astype(np.int32))
    perm = R.permutation(e, n, window=window)
    assert sorted(perm.tolist()) == list(range(n))
    out = perm[e.long()]
    assert _span(out) < 0.35 * _span(e)                         # 20 % of the edges are random: their span cannot shrink
    new_of_old = perm.numpy()[shuffle]                          # new id of planted vertex i
    spread = np.array([np.ptp(new_of_old[k * size:(k + 1) * size]) for k in range(n // size)])
    assert np.median(spread) < 4 * size                         # a community occupies a short id range


def test_degenerate_inputs():
    assert R.permutation(torch.zeros(2, 0, dtype=torch.int32), 5).tolist() == [0, 1, 2, 3, 4]
    e = torch.tensor([[0, 1, 1, 2, 2], [0, 1, 2, 1, 2]], dtype=torch.int32)      # self loops + duplicates
    perm = R.permutation(e, 4)
    assert sorted(perm.tolist()) == [0, 1, 2, 3]
    assert abs(int(perm[1]) - int(perm[2])) == 1


def test_compat_modules_export_the_reference_names():
    import importlib, os, sys
    compat = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gnnadvisor_osdi21_b200", "compat")
    sys.path.insert(0, compat)
    try:
        for name, attrs in (("GNNAdvisor", ["SAG", "forward", "backward", "forward_gin", "backward_gin", "build_part"]),
                            ("rabbit", ["reorder"]), ("dgl", ["DGLGraph"]), ("torch_sparse", ["spmm"])):
            mod = importlib.import_module(name)
            for a in attrs:
                assert hasattr(mod, a), (name, a)
        import torch_sparse
        idx = torch.tensor([[0, 0, 1], [1, 1, 0]])
        got = torch_sparse.spmm(idx, torch.ones(3), 2, 2, torch.tensor([[1., 2.], [3., 4.]]))
        assert torch.equal(got, torch.tensor([[6., 8.], [1., 2.]]))
    finally:
        sys.path.remove(compat)
        for name in ("GNNAdvisor", "rabbit", "dgl", "torch_sparse"):
            sys.modules.pop(name, None)


REF_PY = "/root/reference/GNNAdvisor"


@pytest.mark.skipif(not os.path.isdir(REF_PY), reason="the reference tree is only mounted in the authoring container")
def test_reference_modules_import_unchanged_against_the_compat_modules():
    """The reference's own gnn_conv.py / param.py / dataset.py, imported UNCHANGED from the read-only reference tree with
    compat/ on the path (INTEGRATION.md option A): `import GNNAdvisor`, `import rabbit`, `import dgl` resolve to this
    repository, the layer classes construct, and a call reaches our CHECK_INPUT (same message as GNNAdvisor.cpp:71-73)
    with the reference's own positional argument order.  No GPU needed: CPU tensors are what CHECK_INPUT rejects."""
    import importlib
    import types
    compat = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gnnadvisor_osdi21_b200", "compat")
    names = ("GNNAdvisor", "rabbit", "dgl", "torch_sparse", "gnn_conv", "param", "dataset")
    saved = {n: sys.modules.pop(n) for n in names if n in sys.modules}
    sys.path[:0] = [compat, REF_PY]
    try:
        gnn_conv = importlib.import_module("gnn_conv")
        assert gnn_conv.__file__.startswith(REF_PY)
        assert os.path.dirname(gnn_conv.GNNA.__file__) == compat
        importlib.import_module("dataset")                      # needs dgl + rabbit stand-ins at import time
        info = types.SimpleNamespace(row_pointers=torch.zeros(5, dtype=torch.int32), column_index=torch.zeros(0, dtype=torch.int32),
                                     degrees=torch.ones(4), partPtr=torch.zeros(1, dtype=torch.int32),
                                     part2Node=torch.zeros(0, dtype=torch.int32), partSize=32, dimWorker=16, warpPerBlock=4)
        X = torch.randn(4, 3)
        for layer in (gnn_conv.GCNConv(3, 2), gnn_conv.GINConv(3, 2)):
            with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
                layer(X, info)
        with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
            gnn_conv.ScatterAndGather.apply(X, info)
        # build_part on the host returns what GNNA_main.py:102-110 expects: two tensors whose .int() is the table
        pp, pn = gnn_conv.GNNA.build_part(2, torch.tensor([0, 3, 3, 4], dtype=torch.int32))
        assert pp.int().tolist() == [0, 2, 3, 4] and pn.int().tolist() == [0, 0, 2]
    finally:
        del sys.path[:2]
        for n in names:
            sys.modules.pop(n, None)
        sys.modules.update(saved)
