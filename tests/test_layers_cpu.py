"""Host logic of the autograd layers (gnnadvisor_osdi21_b200/layers.py, the mirror of GNNAdvisor/gnn_conv.py) without a
GPU: the extension surface the layers call is replaced by the CPU oracle, so what is tested is the WIRING -- which operator
a layer calls with which arguments, what it saves for the backward pass, what it hands back to autograd -- against
autograd on the dense closed forms.  (The operators themselves are tested against the oracle on the GPU,
tests/test_parity_gpu.py.)"""
import numpy as np
import pytest
import torch

import oracle
from gnnadvisor_osdi21_b200 import graph, layers


class OracleSurface:
    """SAG / forward / backward / forward_gin / backward_gin with the signatures of gnnadvisor_osdi21_b200.ops (and of the
    reference's pybind module, GNNAdvisor.cpp:253-263), computed by the oracle on CPU tensors.  Records the calls."""

    def __init__(self):
        self.calls = []

    @staticmethod
    def _np(*ts):
        return [t.detach().contiguous().numpy() if isinstance(t, torch.Tensor) else t for t in ts]

    def SAG(self, X, rp, ci, deg, pp, pn, ps, dw, wpb):
        self.calls.append(("SAG", ps, dw, wpb))
        return torch.from_numpy(oracle.SAG(*self._np(X, rp, ci, deg, pp, pn)))

    def forward(self, X, W, rp, ci, deg, pp, pn, ps, dw, wpb):
        self.calls.append(("forward", ps, dw, wpb))
        return [torch.from_numpy(oracle.forward(*self._np(X, W, rp, ci, deg, pp, pn))[0])]

    def backward(self, dO, X, W, rp, ci, deg, pp, pn, ps, dw, wpb, need_d_input=True):
        self.calls.append(("backward", need_d_input))
        dX, dW = oracle.backward(*self._np(dO, X, W, rp, ci, deg, pp, pn))
        return [torch.from_numpy(dX) if need_d_input else None, torch.from_numpy(dW)]

    def forward_gin(self, X, W, rp, ci, eps, pp, pn, ps, dw, wpb):
        self.calls.append(("forward_gin", eps))
        out, agg = oracle.forward_gin(*self._np(X, W, rp, ci), eps, *self._np(pp, pn))
        return [torch.from_numpy(out), torch.from_numpy(agg)]

    def backward_gin(self, dO, X_agg, W, rp, ci, eps, pp, pn, ps, dw, wpb, need_d_input=True):
        self.calls.append(("backward_gin", need_d_input))
        dX, dW = oracle.backward_gin(*self._np(dO, X_agg, W, rp, ci), eps, *self._np(pp, pn))
        return [torch.from_numpy(dX) if need_d_input else None, torch.from_numpy(dW)]

    # the mixed-precision / fused names resolve to the fp32 operators here: routing is what the tests look at
    def forward_mixed(self, *a):
        self.calls.append(("forward_mixed",))
        return self.forward(*a)

    def backward_mixed(self, *a, **k):
        self.calls.append(("backward_mixed",))
        return self.backward(*a, **k)

    def forward_gin_mixed(self, *a):
        self.calls.append(("forward_gin_mixed",))
        return self.forward_gin(*a)

    def backward_gin_mixed(self, *a, **k):
        self.calls.append(("backward_gin_mixed",))
        return self.backward_gin(*a, **k)

    def forward_gin_fused(self, *a):
        self.calls.append(("forward_gin_fused",))
        return self.forward_gin(*a)

    def backward_fused(self, dO, X, W, rp, ci, deg, pp, pn, ps, dw, wpb, gather_bf16=False):
        self.calls.append(("backward_fused", gather_bf16))
        return self.backward(dO, X, W, rp, ci, deg, pp, pn, ps, dw, wpb)

    def scale_rows_bf16(self, X, degrees=None):
        self.calls.append(("scale_rows_bf16",))
        return X

    def names(self):
        return [c[0] for c in self.calls]


@pytest.fixture()
def setup(monkeypatch):
    surface = OracleSurface()
    monkeypatch.setattr(layers, "GNNA", surface)
    n = 200
    rp, ci = graph.synth_graph(n, 2400, kind="rmat", seed=12)
    pp, pn = oracle.build_part(4, rp.numpy(), exact=True)

    class Info:
        pass
    info = Info()
    info.row_pointers, info.column_index = rp, ci
    info.degrees = torch.from_numpy(oracle.degrees(rp.numpy()))
    info.partPtr, info.part2Node = torch.from_numpy(pp), torch.from_numpy(pn)
    info.partSize, info.dimWorker, info.warpPerBlock = 4, 16, 2
    A = torch.zeros(n, n, dtype=torch.float64)
    A[np.repeat(np.arange(n), np.diff(rp.numpy())), ci.numpy().astype(np.int64)] = 1.0
    dn = info.degrees.double()
    return surface, info, A, dn[:, None] * A * dn[None, :], n


def _grads_close(params, refs):
    for p, q in zip(params, refs):
        assert torch.allclose(p.grad.double(), q.grad, rtol=2e-3, atol=2e-4 * float(q.grad.abs().max()))


def test_gcn_model_gradients_and_skipped_input_gradient(setup):
    """GNNA_main.py:142-156: conv -> relu -> conv -> log_softmax; gnn_conv.py:31-78."""
    surface, info, _, Ah, n = setup
    torch.manual_seed(0)
    x = torch.randn(n, 12)
    y = torch.randint(0, 4, (n,))
    c1, c2 = layers.GCNConv(12, 8), layers.GCNConv(8, 4)
    bound = 1.0 / np.sqrt(8)
    assert float(c1.weights.abs().max()) <= bound and float(c1.weights.abs().max()) > 0.5 * bound      # gnn_conv.py:86-88
    loss = torch.nn.functional.nll_loss(torch.log_softmax(c2(torch.relu(c1(x, info)), info), dim=1), y)
    loss.backward()
    w = [c1.weights.detach().double().requires_grad_(True), c2.weights.detach().double().requires_grad_(True)]
    ref = torch.nn.functional.nll_loss(torch.log_softmax(Ah @ (torch.relu(Ah @ (x.double() @ w[0])) @ w[1]), dim=1), y)
    ref.backward()
    assert abs(float(loss) - float(ref)) < 1e-5 * abs(float(ref))
    _grads_close([c1.weights, c2.weights], w)
    # the last layer is differentiated first and needs its input gradient; the first layer's features need none
    assert [c for c in surface.calls if c[0] == "backward"] == [("backward", True), ("backward", False)]
    assert ("forward", 4, 16, 2) in surface.calls              # the three knobs reach the extension as they are


def test_gin_model_gradients_and_saved_aggregate(setup):
    """gnn_conv.py:101-147: eps = 0.5, X_agg saved for the backward pass, no self term."""
    surface, info, A, _, n = setup
    torch.manual_seed(1)
    x = torch.randn(n, 10) * 0.1
    y = torch.randint(0, 3, (n,))
    convs = [layers.GINConv(10, 6), layers.GINConv(6, 6), layers.GINConv(6, 3)]
    assert all(c.eplison == 0.5 for c in convs)                # the reference's spelling and value (gnn_conv.py:132)
    h = x
    for i, c in enumerate(convs):
        h = c(h, info)
        if i < 2:
            h = torch.relu(h)
    loss = torch.nn.functional.nll_loss(torch.log_softmax(h, dim=1), y)
    loss.backward()
    w = [c.weights.detach().double().requires_grad_(True) for c in convs]
    r = x.double()
    for i in range(3):
        r = (0.5 * (A @ r)) @ w[i]
        if i < 2:
            r = torch.relu(r)
    ref = torch.nn.functional.nll_loss(torch.log_softmax(r, dim=1), y)
    ref.backward()
    assert abs(float(loss) - float(ref)) < 1e-5 * abs(float(ref))
    _grads_close([c.weights for c in convs], w)
    assert [c for c in surface.calls if c[0] == "backward_gin"] == [("backward_gin", True), ("backward_gin", True), ("backward_gin", False)]
    assert all(c[1] == 0.5 for c in surface.calls if c[0] == "forward_gin")


def test_scatter_and_gather_is_its_own_backward(setup):
    """gnn_conv.py:7-28: d_input = SAG(d_output) on the same (symmetric) graph."""
    surface, info, A, _, n = setup
    x = torch.randn(n, 5, requires_grad=True)
    out = layers.ScatterAndGather.apply(x, info)
    g = torch.randn(n, 5)
    out.backward(g)
    assert torch.allclose(out.detach().double(), A @ x.detach().double(), atol=1e-4)
    assert torch.allclose(x.grad.double(), A.t() @ g.double(), atol=1e-4)
    assert surface.names() == ["SAG", "SAG"]


def test_layer_switches_route_to_the_right_operators(setup):
    surface, info, _, _, n = setup
    x = torch.randn(n, 64, requires_grad=True)
    # bf16 gathered rows
    layers.GCNConv(64, 8, gather_dtype="bf16")(x, info).sum().backward()
    assert surface.names()[:4] == ["forward_mixed", "forward", "backward_mixed", "backward"]
    surface.calls.clear()
    # fused GIN forward: the aggregated matrix is the INPUT (width 64 has a tile, 48 has none -> silently unfused)
    layers.GINConv(64, 8, fused=True)(x, info).sum().backward()
    assert "forward_gin_fused" in surface.names() and "backward_gin" in surface.names()
    surface.calls.clear()
    layers.GINConv(48, 8, fused=True)(torch.randn(n, 48), info)
    assert "forward_gin_fused" not in surface.names() and "forward_gin" in surface.names()
    surface.calls.clear()
    layers.GINConv(64, 8, fused=True, gather_dtype="bf16")(x, info).sum().backward()
    assert surface.names()[0] == "scale_rows_bf16" and "forward_gin_fused" in surface.names() and "backward_gin_mixed" in surface.names()
    surface.calls.clear()
    # fused GCN backward: the aggregated matrix is d_output (width = output_dim); only when the input needs a gradient
    layers.GCNConv(64, 32, fused=True)(x, info).sum().backward()
    assert ("backward_fused", False) in surface.calls
    surface.calls.clear()
    layers.GCNConv(64, 32, fused=True)(x.detach(), info).sum().backward()
    assert "backward_fused" not in surface.names() and ("backward", False) in surface.calls
    surface.calls.clear()
    layers.GCNConv(64, 41, fused=True)(x, info).sum().backward()          # 41 classes: no tile of that width
    assert "backward_fused" not in surface.names()
    assert layers.fused_tile_supported(256, "bf16") and not layers.fused_tile_supported(256, "fp32")
    with pytest.raises(ValueError):
        layers.GCNConv(4, 4, gather_dtype="fp16")


REF_PY = "/root/reference/GNNAdvisor"


@pytest.mark.skipif(not __import__("os").path.isdir(REF_PY), reason="the reference tree is only mounted in the authoring container")
def test_same_seed_gives_the_reference_layers_initial_weights():
    """gnn_conv.py:80-88, 128-138 imported unchanged (compat/ provides `GNNAdvisor`): a model built after
    torch.manual_seed(s) starts from the same weights here and there, layer by layer."""
    import importlib
    import os
    import sys
    from gnnadvisor_osdi21_b200 import sharded
    compat = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gnnadvisor_osdi21_b200", "compat")
    saved = {n: sys.modules.pop(n) for n in ("GNNAdvisor", "gnn_conv") if n in sys.modules}
    sys.path[:0] = [compat, REF_PY]
    try:
        ref = importlib.import_module("gnn_conv")
        for ours, theirs in ((layers.GCNConv, ref.GCNConv), (layers.GINConv, ref.GINConv),
                             (sharded.ShardedGCNConv, ref.GCNConv), (sharded.ShardedGINConv, ref.GINConv)):
            torch.manual_seed(123)
            a = [ours(7, 5), ours(5, 3)]
            torch.manual_seed(123)
            b = [theirs(7, 5), theirs(5, 3)]
            for x, y in zip(a, b):
                assert torch.equal(x.weights, y.weights)
            assert getattr(a[0], "eplison", None) == getattr(b[0], "eplison", None)
    finally:
        del sys.path[:2]
        for n in ("GNNAdvisor", "gnn_conv"):
            sys.modules.pop(n, None)
        sys.modules.update(saved)


@pytest.mark.skipif(not __import__("os").path.isdir(REF_PY), reason="the reference tree is only mounted in the authoring container")
@pytest.mark.parametrize("model", ["gcn", "gin"])
def test_models_equal_the_reference_layers_run_live(setup, model):
    """The reference's gnn_conv.py (unchanged) and this package's layers.py, both on the same extension surface (the CPU
    oracle), same seed: GNNA_main.py's two models give the same output, loss and weight gradients -- the layers call the
    extension with the same operands in the same order.  (ScatterAndGather too.)"""
    import importlib
    import os
    import sys
    surface, info, _, _, n = setup
    compat = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gnnadvisor_osdi21_b200", "compat")
    saved = {k: sys.modules.pop(k) for k in ("GNNAdvisor", "gnn_conv") if k in sys.modules}
    sys.path[:0] = [compat, REF_PY]
    try:
        ref = importlib.import_module("gnn_conv")
        ref.GNNA = surface                                       # the reference's `import GNNAdvisor as GNNA`
        dims = [12, 8, 4] if model == "gcn" else [12, 8, 8, 8, 8, 4]            # GNNA_main.py:142-171
        x = torch.randn(n, 12, generator=torch.Generator().manual_seed(3)) * 0.1
        y = torch.randint(0, 4, (n,), generator=torch.Generator().manual_seed(4))
        results = []
        for mod in (layers, ref):
            torch.manual_seed(11)
            conv = mod.GCNConv if model == "gcn" else mod.GINConv
            convs = [conv(a, b) for a, b in zip(dims[:-1], dims[1:])]
            h = x
            for i, c in enumerate(convs):
                h = c(h, info)
                if i < len(convs) - 1:
                    h = torch.relu(h)
            loss = torch.nn.functional.nll_loss(torch.log_softmax(h, dim=1), y)
            loss.backward()
            results.append((h.detach(), float(loss), [c.weights.grad.clone() for c in convs]))
        (h_a, l_a, g_a), (h_b, l_b, g_b) = results
        assert torch.equal(h_a, h_b) and l_a == l_b
        for a, b in zip(g_a, g_b):
            assert torch.equal(a, b)
        # ScatterAndGather: forward only -- the reference's backward returns ONE gradient for two inputs (gnn_conv.py:21-28),
        # which autograd rejects ("incorrect number of gradients"); ours returns (d_input, None)
        xs = torch.randn(n, 6)
        assert torch.equal(layers.ScatterAndGather.apply(xs, info), ref.ScatterAndGather.apply(xs, info))
        with pytest.raises(RuntimeError, match="incorrect number of gradients"):
            ref.ScatterAndGather.apply(xs.clone().requires_grad_(True), info).sum().backward()
    finally:
        del sys.path[:2]
        for k in ("GNNAdvisor", "gnn_conv"):
            sys.modules.pop(k, None)
        sys.modules.update(saved)
