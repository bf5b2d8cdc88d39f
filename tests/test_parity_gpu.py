"""GPU parity tests: the CUDA path (through the ctypes binding of the C ABI) against the CPU oracle on
the same seeded inputs, against golden outputs of the reference's own CUDA kernels, and -- at the
sizes of BASELINE.json's configurations -- through size-independent properties.

Tolerance: 1e-4 relative fp32 (BASELINE.json north_star), see helpers.assert_close.
Integer work (build_part on the device) is compared bit for bit."""
import os

import numpy as np
import pytest
import torch

import oracle
from helpers import assert_close, golden_case, golden_terms, make_graph, rand_features, rand_weight
from gnnadvisor_osdi21_b200 import graph, ops, layers

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _general_kernels_unless_asked(request):
    """The graphs of this file are small enough for the single-launch path (csrc/aggregate_small.cu), which would leave the
    general kernels untested.  Every test runs with that path OFF unless it is marked `small_path`."""
    from gnnadvisor_osdi21_b200 import _lib
    prev = _lib.set_small_parts(16384 if request.node.get_closest_marker("small_path") else 0)
    yield
    _lib.set_small_parts(prev)


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(DEV)


class G:
    """A graph + its group table on the device, with the numpy copies for the oracle."""

    def __init__(self, rp, ci, ps, exact=True):
        self.rp, self.ci, self.ps = rp, ci, ps
        self.pp, self.pn = oracle.build_part(ps, rp, exact=exact)
        self.deg = oracle.degrees(rp)
        self.n = len(rp) - 1
        self.d_rp, self.d_ci, self.d_pp, self.d_pn, self.d_deg = dev(rp), dev(ci), dev(self.pp), dev(self.pn), dev(self.deg)

    def gargs(self):
        return (self.d_rp, self.d_ci)

    def terms(self, mode, X, eps=0.5):
        """Sum of the absolute values of the terms each output element adds up (see helpers.assert_close)."""
        return oracle.aggregate(mode, np.abs(X), self.ci, self.deg, eps, self.pp, self.pn)

    def pargs(self):
        return (self.d_pp, self.d_pn)


GRAPHS = {
    "uniform": lambda: make_graph("uniform", 700, 4000, 21),
    "rmat": lambda: make_graph("rmat", 1500, 40000, 22),
}


@pytest.mark.parametrize("gname", sorted(GRAPHS))
@pytest.mark.parametrize("dim", [1, 2, 3, 4, 6, 7, 16, 32, 41, 47, 64, 100, 128, 172, 300])
def test_aggregation_modes_all_dims(gname, dim):
    """SAG / GCN / GIN aggregation at the dims of every BASELINE.json config (and awkward ones)."""
    rp, ci = GRAPHS[gname]()
    g = G(rp, ci, 32)
    X = rand_features(g.n, dim, 100 + dim)
    dX = dev(X)
    out = ops.SAG(dX, *g.gargs(), g.d_deg, *g.pargs(), 32, 32, 8)
    assert_close(out.cpu().numpy(), oracle.aggregate(0, X, ci, None, 1.0, g.pp, g.pn), what="SAG", terms=g.terms(0, X))
    lib_gcn = _gcn_agg(dX, g, 32, 8)
    assert_close(lib_gcn, oracle.aggregate(1, X, ci, g.deg, 1.0, g.pp, g.pn), what="GCN", terms=g.terms(1, X))
    lib_gin = _gin_agg(dX, g, 0.5, 32, 8)
    assert_close(lib_gin, oracle.aggregate(2, X, ci, None, 0.5, g.pp, g.pn), what="GIN", terms=g.terms(2, X))


def _gcn_agg(dX, g, dw, wpb):
    """GCN aggregation alone: forward() with W = identity would add a GEMM, so call the C ABI entry."""
    import ctypes
    from gnnadvisor_osdi21_b200 import _lib
    out = torch.empty_like(dX)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    _lib.check(_lib.load().gnna_gcn_aggregate_f32(p(dX), p(out), p(g.d_rp), p(g.d_ci), p(g.d_deg), p(g.d_pp), p(g.d_pn),
                                                  g.n, dX.shape[1], g.d_pn.numel(), g.ps, dw, wpb,
                                                  ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "gcn")
    return out.cpu().numpy()


def _gin_agg(dX, g, eps, dw, wpb):
    import ctypes
    from gnnadvisor_osdi21_b200 import _lib
    out = torch.empty_like(dX)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    _lib.check(_lib.load().gnna_gin_aggregate_f32(p(dX), p(out), p(g.d_rp), p(g.d_ci), eps, p(g.d_pp), p(g.d_pn),
                                                  g.n, dX.shape[1], g.d_pn.numel(), g.ps, dw, wpb,
                                                  ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "gin")
    return out.cpu().numpy()


@pytest.mark.parametrize("ps", [1, 2, 5, 16, 32, 64, 512])
@pytest.mark.parametrize("dw", [1, 2, 4, 8, 16, 32])
def test_partsize_dimworker_sweep(ps, dw):
    """The study scripts sweep partSize 2..512 and dimWorker 1..32 (s7-4_1, s7-4_2): any combination is valid."""
    rp, ci = GRAPHS["rmat"]()
    g = G(rp, ci, ps)
    X = rand_features(g.n, 64, 7)
    assert_close(_gcn_agg(dev(X), g, dw, 4), oracle.aggregate(1, X, ci, g.deg, 1.0, g.pp, g.pn), what="ps%d dw%d" % (ps, dw),
                 terms=g.terms(1, X))


@pytest.mark.parametrize("wpb", [1, 2, 3, 8, 16, 32])
def test_warp_per_block_sweep(wpb):
    rp, ci = GRAPHS["uniform"]()
    g = G(rp, ci, 8)
    X = rand_features(g.n, 16, 8)
    out = ops.SAG(dev(X), *g.gargs(), g.d_deg, *g.pargs(), 8, 16, wpb)
    assert_close(out.cpu().numpy(), oracle.aggregate(0, X, ci, None, 1.0, g.pp, g.pn), what="wpb%d" % wpb)


def test_single_group_nodes_are_bit_identical():
    """A node whose neighbours fit one group is summed in the reference's order: exact equality for
    SAG / GIN always, and for GCN in the reference-rounding mode (gnna_set_gcn_exact)."""
    from gnnadvisor_osdi21_b200 import _lib
    rp, ci = make_graph("uniform", 2000, 12000, 31)
    assert (rp[1:] - rp[:-1]).max() <= 32
    g = G(rp, ci, 32)
    for dim in (64, 41):
        X = rand_features(g.n, dim, 9)
        assert np.array_equal(_gin_agg(dev(X), g, 0.5, 32, 8), oracle.aggregate(2, X, ci, None, 0.5, g.pp, g.pn))
        assert np.array_equal(ops.SAG(dev(X), *g.gargs(), g.d_deg, *g.pargs(), 32, 32, 8).cpu().numpy(),
                              oracle.aggregate(0, X, ci, None, 1.0, g.pp, g.pn))
        ref = oracle.aggregate(1, X, ci, g.deg, 1.0, g.pp, g.pn)
        prev = _lib.set_gcn_exact(True)
        try:
            assert np.array_equal(_gcn_agg(dev(X), g, 32, 8), ref)
        finally:
            _lib.set_gcn_exact(prev)
        assert_close(_gcn_agg(dev(X), g, 32, 8), ref, rtol=1e-5, what="prescaled GCN", terms=g.terms(1, X))   # a few ulp, not bitwise


@pytest.mark.parametrize("dim", [7, 16, 64, 100])
def test_gcn_exact_mode_all_paths(dim):
    """The per-edge-rounding kernel (WEIGHTED template) stays covered: multi-group nodes, any dim."""
    from gnnadvisor_osdi21_b200 import _lib
    rp, ci = GRAPHS["rmat"]()
    g = G(rp, ci, 16)
    X = rand_features(g.n, dim, 10)
    prev = _lib.set_gcn_exact(True)
    try:
        assert_close(_gcn_agg(dev(X), g, 32, 8), oracle.aggregate(1, X, ci, g.deg, 1.0, g.pp, g.pn), what="exact", terms=g.terms(1, X))
    finally:
        _lib.set_gcn_exact(prev)


def test_layer_operators_against_oracle():
    """forward / backward / forward_gin / backward_gin incl. the dense products (cuBLAS SGEMM)."""
    for (n, e, din, dout, ps) in [(900, 12000, 96, 16, 32), (900, 12000, 16, 7, 8), (600, 9000, 64, 64, 32)]:
        rp, ci = make_graph("rmat", n, e, 41)
        g = G(rp, ci, ps)
        X, W, dO = rand_features(n, din, 42), rand_weight(din, dout, 43), rand_features(n, dout, 44)
        dXd, dWd, dOd = dev(X), dev(W), dev(dO)
        out, = ops.forward(dXd, dWd, *g.gargs(), g.d_deg, *g.pargs(), ps, 32, 8)
        assert_close(out.cpu().numpy(), oracle.forward(X, W, rp, ci, g.deg, g.pp, g.pn)[0], what="forward",
                     terms=g.terms(1, np.abs(X).astype(np.float64) @ np.abs(W)))
        dX, dW = ops.backward(dOd, dXd, dWd, *g.gargs(), g.d_deg, *g.pargs(), ps, 32, 8)
        odX, odW = oracle.backward(dO, X, W, rp, ci, g.deg, g.pp, g.pn)
        aG = np.abs(oracle.aggregate(1, dO, ci, g.deg, 1.0, g.pp, g.pn)).astype(np.float64)
        assert_close(dX.cpu().numpy(), odX, what="backward dX", terms=aG @ np.abs(W.T))
        assert_close(dW.cpu().numpy(), odW, what="backward dW", terms=np.abs(X.T).astype(np.float64) @ aG)
        o, S = ops.forward_gin(dXd, dWd, *g.gargs(), 0.5, *g.pargs(), ps, 32, 2)
        oo, oS = oracle.forward_gin(X, W, rp, ci, 0.5, g.pp, g.pn)
        assert_close(S.cpu().numpy(), oS, what="gin agg")
        assert_close(o.cpu().numpy(), oo, what="gin out")
        dXg, dWg = ops.backward_gin(dOd, S, dWd, *g.gargs(), 0.5, *g.pargs(), ps, 32, 2)
        odXg, odWg = oracle.backward_gin(dO, oS, W, rp, ci, 0.5, g.pp, g.pn)
        aPm = np.abs(dO).astype(np.float64) @ np.abs(W.T)
        assert_close(dXg.cpu().numpy(), odXg, what="gin dX", terms=oracle.closed_form(2, aPm, rp, ci, 0.5))
        assert_close(dWg.cpu().numpy(), odWg, what="gin dW", terms=np.abs(oS.T).astype(np.float64) @ np.abs(dO))


def test_reference_cuda_golden(golden_dir):
    """Outputs of the reference's own kernels on a B200 (tests/golden/refgpu.npz)."""
    path = os.path.join(golden_dir, "refgpu.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/refgpu.npz not generated yet")
    gz = np.load(path)
    for c in range(len({k.split("/")[0] for k in gz.files})):
        k = "case%d/" % c
        din, dout, ps, dw, wpb = [int(v) for v in gz[k + "meta"]]
        rp, ci = gz[k + "row_ptr"], gz[k + "col_idx"]
        pp, pn = dev(gz[k + "partPtr"]), dev(gz[k + "part2Node"])
        d_rp, d_ci = dev(rp), dev(ci)
        deg = ops.degrees_from_row_ptr(d_rp)
        X, W, dO, what = golden_case(gz, k)
        X, W, dO = dev(X), dev(W), dev(dO)
        t = golden_terms(gz, k, oracle)
        if "SAG" in what:
            assert_close(ops.SAG(X, d_rp, d_ci, deg, pp, pn, ps, dw, wpb).cpu().numpy(), gz[k + "SAG"], what=k + "SAG")
        if "gcn" in what:
            assert_close(ops.forward(X, W, d_rp, d_ci, deg, pp, pn, ps, dw, wpb)[0].cpu().numpy(), gz[k + "forward"],
                         what=k + "forward", terms=t["fwd"])
            dX, dW = ops.backward(dO, X, W, d_rp, d_ci, deg, pp, pn, ps, dw, wpb)
            assert_close(dX.cpu().numpy(), gz[k + "backward_dX"], what=k + "dX", terms=t["dX"])
            assert_close(dW.cpu().numpy(), gz[k + "backward_dW"], what=k + "dW", terms=t["dW"])
        if "gin" in what:
            o, S = ops.forward_gin(X, W, d_rp, d_ci, 0.5, pp, pn, ps, dw, wpb)
            assert_close(S.cpu().numpy(), gz[k + "forward_gin_agg"], what=k + "gin agg")
            assert_close(o.cpu().numpy(), gz[k + "forward_gin"], what=k + "gin out", terms=t["gin_out"])
            dXg, dWg = ops.backward_gin(dO, dev(gz[k + "forward_gin_agg"]), W, d_rp, d_ci, 0.5, pp, pn, ps, dw, wpb)
            assert_close(dXg.cpu().numpy(), gz[k + "backward_gin_dX"], what=k + "gin dX", terms=t["gin_dX"])
            assert_close(dWg.cpu().numpy(), gz[k + "backward_gin_dW"], what=k + "gin dW", terms=t["gin_dW"])


# ------------------------------------------------------------------------------------------ edge cases
def test_empty_and_degenerate_graphs():
    # no edges at all: every output row is zero, no groups
    rp = np.zeros(11, dtype=np.int32)
    g = G(rp, np.zeros(0, dtype=np.int32), 32)
    X = rand_features(10, 16, 1)
    assert g.pn.size == 0
    out = ops.SAG(dev(X), *g.gargs(), g.d_deg, *g.pargs(), 32, 32, 4)
    assert torch.count_nonzero(out).item() == 0
    # one node with a self loop
    g = G(np.array([0, 1], dtype=np.int32), np.array([0], dtype=np.int32), 32)
    X = rand_features(1, 5, 2)
    assert np.array_equal(ops.SAG(dev(X), *g.gargs(), g.d_deg, *g.pargs(), 32, 32, 4).cpu().numpy(), X)
    # zero nodes
    e = torch.empty(0, 8, device=DEV)
    z = torch.zeros(1, dtype=torch.int32, device=DEV)
    zi = torch.empty(0, dtype=torch.int32, device=DEV)
    assert ops.SAG(e, z, zi, torch.empty(0, device=DEV), z, zi, 32, 32, 4).shape == (0, 8)


def test_isolated_nodes_and_f6_table():
    """Last node isolated: the reference table's terminal is 0 (F6), which makes the last group empty
    (partEnd <= partBeg contributes nothing, kernel.cu:383).  With the exact table it is aggregated."""
    rng = np.random.default_rng(3)
    deg = rng.integers(0, 50, 400); deg[::5] = 0; deg[-1] = 0; deg[-2] = 40
    rp = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
    ci = rng.integers(0, 400, rp[-1]).astype(np.int32)
    X = rand_features(400, 32, 4)
    for exact in (True, False):
        g = G(rp, ci, 32, exact=exact)
        if not exact:
            assert g.pp[-1] == 0
        got = _gcn_agg(dev(X), g, 32, 4)
        assert_close(got, oracle.aggregate(1, X, ci, g.deg, 1.0, g.pp, g.pn), what="exact=%s" % exact, terms=g.terms(1, X))
        assert np.count_nonzero(got[deg == 0]) == 0
    # the F6 table drops the last group's neighbours, the exact one does not
    ge, gc = G(rp, ci, 32, exact=True), G(rp, ci, 32, exact=False)
    a, b = _gcn_agg(dev(X), ge, 32, 4), _gcn_agg(dev(X), gc, 32, 4)
    assert not np.allclose(a[-2], b[-2]) and np.array_equal(a[:-2], b[:-2])


def test_hub_node_many_groups():
    """One node with 50 000 neighbours (1 563 groups merged by vector reductions) among small ones."""
    n = 60000
    src = np.concatenate([np.zeros(50000, dtype=np.int64), np.arange(1, 50001)])
    dst = np.concatenate([np.arange(1, 50001), np.zeros(50000, dtype=np.int64)])
    rp, ci = graph.csr_from_edges(torch.from_numpy(src), torch.from_numpy(dst), n)
    rp, ci = rp.numpy(), ci.numpy()
    g = G(rp, ci, 32)
    X = np.abs(rand_features(n, 16, 5)) + 0.5
    assert_close(ops.SAG(dev(X), *g.gargs(), g.d_deg, *g.pargs(), 32, 32, 8).cpu().numpy(),
                 oracle.aggregate(0, X, ci, None, 1.0, g.pp, g.pn, threads=-1), what="hub")


def test_noncontiguous_and_wrong_dtype_inputs_raise():
    rp, ci = GRAPHS["uniform"]()
    g = G(rp, ci, 32)
    X = dev(rand_features(g.n, 32, 1))
    with pytest.raises(RuntimeError, match="input must be contiguous"):
        ops.SAG(X[:, ::2], *g.gargs(), g.d_deg, *g.pargs(), 32, 32, 4)
    with pytest.raises(RuntimeError, match="dtype"):
        ops.SAG(X.double(), *g.gargs(), g.d_deg, *g.pargs(), 32, 32, 4)
    with pytest.raises(RuntimeError, match="column_index must be a CUDA tensor"):
        ops.SAG(X, g.d_rp, g.d_ci.cpu(), g.d_deg, *g.pargs(), 32, 32, 4)
    with pytest.raises(RuntimeError, match="size mismatch"):
        ops.forward(X, torch.zeros(5, 3, device=DEV), *g.gargs(), g.d_deg, *g.pargs(), 32, 32, 4)


def test_inputs_are_not_modified_and_nonzero_stream():
    rp, ci = GRAPHS["rmat"]()
    g = G(rp, ci, 32)
    X = rand_features(g.n, 64, 6)
    dX = dev(X)
    keep = dX.clone()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        out = ops.SAG(dX, *g.gargs(), g.d_deg, *g.pargs(), 32, 32, 8)
    s.synchronize()
    assert torch.equal(dX, keep)
    assert_close(out.cpu().numpy(), oracle.aggregate(0, X, ci, None, 1.0, g.pp, g.pn), what="stream")


# ------------------------------------------------------------------------------------------ device build_part / degrees
def test_build_part_device_bit_exact():
    rng = np.random.default_rng(11)
    for n, hi in ((1, 5), (1000, 90), (300000, 40)):
        deg = rng.integers(0, hi, n)
        if n > 1000:
            deg[rng.integers(0, n, 20)] = 30000
        rp = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
        for ps in (1, 3, 32):
            pp, pn = ops.build_part(ps, dev(rp))
            opp, opn = oracle.build_part(ps, rp, exact=True)
            assert pp.is_cuda and pp.dtype == torch.int32
            assert np.array_equal(pp.cpu().numpy(), opp) and np.array_equal(pn.cpu().numpy(), opn)
    rp = np.array([0, 0, 3, 3, 10], dtype=np.int32)
    assert np.array_equal(ops.degrees_from_row_ptr(dev(rp)).cpu().numpy(), oracle.degrees(rp))


# ------------------------------------------------------------------------------------------ bf16 storage path
@pytest.mark.parametrize("dim", [8, 16, 41, 64, 128, 200])
def test_bf16_gather_path(dim):
    """bf16 neighbour rows, fp32 accumulate: compared with the fp32 oracle on the SAME bf16-rounded
    inputs (no reference of this precision exists) -- so the fp32 tolerance applies."""
    rp, ci = GRAPHS["rmat"]()
    g = G(rp, ci, 32)
    Xb = torch.from_numpy(rand_features(g.n, dim, 12)).to(torch.bfloat16)
    Xr = Xb.float().numpy()
    for mode in (0, 1, 2):
        got = ops.aggregate_bf16(mode, Xb.to(DEV), *g.gargs(), g.d_deg, 0.5, *g.pargs(), 32, 32, 8).cpu().numpy()
        assert_close(got, oracle.aggregate(mode, Xr, ci, g.deg, 0.5, g.pp, g.pn), what="bf16 mode %d" % mode, terms=g.terms(mode, Xr))
    # mode 3: the caller pre-scales by n_j (and rounds THAT to bf16); out_i = n_i * sum_j Xs_j
    Xs = (torch.from_numpy(Xr) * torch.from_numpy(g.deg)[:, None]).to(torch.bfloat16)
    got = ops.aggregate_bf16(3, Xs.to(DEV), *g.gargs(), g.d_deg, 0.5, *g.pargs(), 32, 32, 8).cpu().numpy()
    ref = g.deg[:, None].astype(np.float64) * oracle.closed_form(0, Xs.float().numpy(), rp, ci)
    assert_close(got, ref, what="bf16 prescaled", terms=g.deg[:, None] * oracle.closed_form(0, np.abs(Xs.float().numpy()), rp, ci))


# ------------------------------------------------------------------------------------------ mixed-precision GCN layer
@pytest.mark.parametrize("dim", [1, 7, 8, 41, 47, 64, 100])
def test_scale_rows_bf16_is_bit_exact_and_padded(dim):
    """bf16(fl(n_i * x)) with round-to-nearest-even, rows zero-padded to whole 16-byte chunks: integer/bit work, compared
    bit for bit with the same two roundings done by torch on the CPU."""
    n = 333
    X = torch.from_numpy(rand_features(n, dim, 61))
    deg = torch.from_numpy(oracle.degrees(make_graph("rmat", n, 4000, 62)[0]))
    for d in (deg, None):
        got = ops.scale_rows_bf16(X.to(DEV), None if d is None else d.to(DEV)).cpu()
        ld = (dim + 7) // 8 * 8
        assert got.shape == (n, ld) and got.dtype == torch.bfloat16
        want = (X if d is None else X * d[:, None]).to(torch.bfloat16)
        assert torch.equal(got[:, :dim].view(torch.int16), want.view(torch.int16))
        assert (got[:, dim:].view(torch.int16) == 0).all()


@pytest.mark.parametrize("dout", [16, 41, 64, 47])
def test_mixed_precision_gcn_operators(dout):
    """forward_mixed / backward_mixed (bf16 gathered rows, everything else fp32).
    (1) the padded-row bf16 gather against the fp32 oracle on the SAME rounded rows: fp32 tolerance (1e-4);
    (2) the two operators against the fp32 oracle on the unrounded inputs: 1e-2 relative, the bound stated for bf16
        storage in SURVEY.md 8c (one bf16 rounding per gathered term, 2^-9 relative each)."""
    n, din = 1500, 40
    rp, ci = GRAPHS["rmat"]()
    g = G(rp, ci, 32)
    X, W, dO = rand_features(n, din, 63), rand_weight(din, dout, 64), rand_features(n, dout, 65)
    a = (*g.gargs(), g.d_deg, *g.pargs(), 32, 32, 4)
    # (1)
    Tb = ops.scale_rows_bf16(dev(dO), g.d_deg)
    got = ops.aggregate_bf16(3, Tb, *g.gargs(), g.d_deg, 0.5, *g.pargs(), 32, 32, 4, dim=dout).cpu().numpy()
    Tr = Tb[:, :dout].float().cpu().numpy()
    ref = g.deg[:, None].astype(np.float64) * oracle.closed_form(0, Tr, rp, ci)
    assert_close(got, ref, what="padded bf16 gather", terms=g.deg[:, None] * oracle.closed_form(0, np.abs(Tr), rp, ci))
    # (2) bound: |err| <= u * sum|terms| with u = 2^-8, the unit roundoff of bf16 (8 significant bits) -- written as
    # rtol 1e-2 on max(|ref|, 0.4 * sum|terms|)
    aX, aW, adO = np.abs(X).astype(np.float64), np.abs(W).astype(np.float64), np.abs(dO).astype(np.float64)
    y, = ops.forward_mixed(dev(X), dev(W), *a)
    assert_close(y.cpu().numpy(), oracle.forward(X, W, rp, ci, g.deg, g.pp, g.pn)[0], rtol=1e-2, what="mixed forward",
                 terms=4 * g.terms(1, aX @ aW))
    dX, dW = ops.backward_mixed(dev(dO), dev(X), dev(W), *a)
    odX, odW = oracle.backward(dO, X, W, rp, ci, g.deg, g.pp, g.pn)
    aG = g.terms(1, adO).astype(np.float64)
    assert_close(dX.cpu().numpy(), odX, rtol=1e-2, what="mixed dX", terms=4 * (aG @ aW.T))
    assert_close(dW.cpu().numpy(), odW, rtol=1e-2, what="mixed dW", terms=4 * (aX.T @ aG))
    dX2, dW2 = ops.backward_mixed(dev(dO), dev(X), dev(W), *a, need_d_input=False)
    assert dX2 is None
    assert_close(dW2.cpu().numpy(), dW.cpu().numpy(), what="mixed dW without dX", terms=aX.T @ aG)


@pytest.mark.parametrize("din,dout", [(100, 64), (64, 47), (24, 16)])
def test_mixed_precision_gin_operators(din, dout):
    """forward_gin_mixed / backward_gin_mixed against the fp32 oracle, bound u_bf16 * sum|terms| (see above); the saved
    X_agg against the oracle on the bf16-rounded X at the fp32 tolerance."""
    n = 1500
    rp, ci = GRAPHS["rmat"]()
    g = G(rp, ci, 32)
    X, W, dO = rand_features(n, din, 67), rand_weight(din, dout, 68), rand_features(n, dout, 69)
    a = (*g.gargs(), 0.5, *g.pargs(), 32, 32, 4)
    aX, aW, adO = np.abs(X).astype(np.float64), np.abs(W).astype(np.float64), np.abs(dO).astype(np.float64)
    y, S = ops.forward_gin_mixed(dev(X), dev(W), *a)
    Xr = torch.from_numpy(X).to(torch.bfloat16).float().numpy()
    assert_close(S.cpu().numpy(), oracle.aggregate(2, Xr, ci, g.deg, 0.5, g.pp, g.pn), what="mixed gin X_agg", terms=g.terms(2, Xr))
    oy, oS = oracle.forward_gin(X, W, rp, ci, 0.5, g.pp, g.pn)
    assert_close(y.cpu().numpy(), oy, rtol=1e-2, what="mixed gin out", terms=4 * (g.terms(2, aX).astype(np.float64) @ aW))
    dX, dW = ops.backward_gin_mixed(dev(dO), S, dev(W), *a)
    Sn = S.cpu().numpy()
    odX, odW = oracle.backward_gin(dO, Sn, W, rp, ci, 0.5, g.pp, g.pn)
    assert_close(dX.cpu().numpy(), odX, rtol=1e-2, what="mixed gin dX", terms=4 * g.terms(2, adO @ aW.T))
    assert_close(dW.cpu().numpy(), odW, what="mixed gin dW", terms=np.abs(Sn.T).astype(np.float64) @ adO)
    dX2, dW2 = ops.backward_gin_mixed(dev(dO), S, dev(W), *a, need_d_input=False)
    assert dX2 is None
    assert_close(dW2.cpu().numpy(), odW, what="mixed gin dW only", terms=np.abs(Sn.T).astype(np.float64) @ adO)


@pytest.mark.parametrize("model", ["gcn", "gin"])
def test_mixed_precision_layers_train(model):
    """GCNConv / GINConv(gather_dtype="bf16") in a 2-layer model with autograd: output and weight gradients within 1e-2
    (relative to the largest element) of the fp32 layers.  The loss is a fixed random projection of the output, so the
    gradients are well conditioned (nll_loss on near-uniform logits makes them sums of cancelling terms), and there is no
    ReLU between the layers: a pre-activation that changes sign under a 2^-9 perturbation flips its whole gradient term,
    which is a property of ReLU, not an error of the operators."""
    n, din, hid, cls = 1500, 32, 64, 41
    rp, ci = GRAPHS["rmat"]()
    g = G(rp, ci, 32)

    class Info:
        pass
    info = Info()
    info.row_pointers, info.column_index, info.degrees = g.d_rp, g.d_ci, 1.0 / g.d_deg   # keep activations O(1)
    info.partPtr, info.part2Node = g.d_pp, g.d_pn
    info.partSize, info.dimWorker, info.warpPerBlock = 32, 32, 4
    x = dev(rand_features(n, din, 66))
    R = dev(rand_features(n, cls, 70))
    res = {}
    for kind in ("fp32", "bf16"):
        torch.manual_seed(7)
        conv = layers.GCNConv if model == "gcn" else layers.GINConv
        c1, c2 = conv(din, hid, gather_dtype=kind).to(DEV), conv(hid, cls, gather_dtype=kind).to(DEV)
        if model == "gin":
            c1.eplison = c2.eplison = 0.02           # keep activations O(1) (no degree normalisation in GIN)
        h = c2(c1(x, info), info)
        (h * R).sum().backward()
        res[kind] = (h.detach().cpu().numpy(), c1.weights.grad.cpu().numpy(), c2.weights.grad.cpu().numpy())
    for k, what in enumerate(("output", "grad of layer 1", "grad of layer 2")):
        ref = res["fp32"][k]
        assert np.abs(res["bf16"][k] - ref).max() <= 1e-2 * np.abs(ref).max(), what


# ------------------------------------------------------------------------------------------ autograd layers
def test_gcn_and_gin_layers_autograd():
    """GCNConv / GINConv modules: forward value and the gradients the reference's backward defines."""
    n, din, hid = 500, 24, 16
    rp, ci = make_graph("rmat", n, 6000, 51)
    g = G(rp, ci, 16)

    class Info:
        pass
    info = Info()
    info.row_pointers, info.column_index, info.degrees = g.d_rp, g.d_ci, g.d_deg
    info.partPtr, info.part2Node = g.d_pp, g.d_pn
    info.partSize, info.dimWorker, info.warpPerBlock = 16, 16, 4
    X = rand_features(n, din, 52)
    dO = rand_features(n, hid, 53)
    for conv, fwd, bwd in ((layers.GCNConv(din, hid), oracle.forward, oracle.backward),
                           (layers.GINConv(din, hid), None, None)):
        conv = conv.to(DEV)
        W = conv.weights.detach().cpu().numpy()
        assert np.abs(W).max() <= 1.0 / np.sqrt(hid) + 1e-7
        x = dev(X).requires_grad_(True)
        y = conv(x, info)
        y.backward(dev(dO))
        aX, aW, adO = np.abs(X).astype(np.float64), np.abs(W).astype(np.float64), np.abs(dO).astype(np.float64)
        if fwd is not None:
            assert_close(y.detach().cpu().numpy(), oracle.forward(X, W, rp, ci, g.deg, g.pp, g.pn)[0], what="gcn y",
                         terms=g.terms(1, aX @ aW))
            odX, odW = oracle.backward(dO, X, W, rp, ci, g.deg, g.pp, g.pn)
            aG = g.terms(1, adO).astype(np.float64)
            tX, tW = aG @ aW.T, aX.T @ aG
        else:
            oy, oS = oracle.forward_gin(X, W, rp, ci, 0.5, g.pp, g.pn)
            assert_close(y.detach().cpu().numpy(), oy, what="gin y", terms=g.terms(2, aX).astype(np.float64) @ aW)
            odX, odW = oracle.backward_gin(dO, oS, W, rp, ci, 0.5, g.pp, g.pn)
            tX, tW = g.terms(2, adO @ aW.T), np.abs(oS.T).astype(np.float64) @ adO
        assert_close(x.grad.cpu().numpy(), odX, what="dX", terms=tX)
        assert_close(conv.weights.grad.cpu().numpy(), odW, what="dW", terms=tW)


# ------------------------------------------------------------------------------------------ full-size properties
@pytest.mark.parametrize("name,dim,scale", [("reddit", 64, 0.25), ("ogbn-products", 64, 0.1)])
def test_full_size_properties(name, dim, scale):
    """At (a fraction of) BASELINE.json's sizes the oracle is too slow to run in full; check
    size-independent properties instead: SAG(ones) == degree exactly (the reference's own
    verification, unitest.py:54-63), linearity, GCN closed form on sampled rows, and agreement with
    the multi-threaded oracle on the whole output."""
    gr = graph.lookalike(name, device=DEV, scale=scale)
    rp, ci, n = gr["row_ptr"], gr["col_idx"], gr["num_nodes"]
    pp, pn = ops.build_part(32, rp)
    deg = ops.degrees_from_row_ptr(rp)
    ones = torch.ones(n, dim, device=DEV)
    out = ops.SAG(ones, rp, ci, deg, pp, pn, 32, 32, 8)
    want = (rp[1:] - rp[:-1]).float()[:, None].expand(n, dim)
    assert torch.equal(out, want)
    gen = torch.Generator(device=DEV).manual_seed(3)
    A = torch.randn(n, dim, device=DEV, generator=gen)
    B = torch.randn(n, dim, device=DEV, generator=gen)
    sa, sb = ops.SAG(A, rp, ci, deg, pp, pn, 32, 32, 8), ops.SAG(B, rp, ci, deg, pp, pn, 32, 32, 8)
    sab = ops.SAG(A + 2 * B, rp, ci, deg, pp, pn, 32, 32, 8)
    lin = sa + 2 * sb
    assert ((sab - lin).abs().max() / lin.abs().max()).item() < 1e-5
    # whole output against the multi-threaded oracle
    got = ops.forward(A, torch.eye(dim, device=DEV), rp, ci, deg, pp, pn, 32, 32, 8)[0].cpu().numpy()
    ref = oracle.aggregate(1, A.cpu().numpy(), ci.cpu().numpy(), deg.cpu().numpy(), 1.0,
                           pp.cpu().numpy(), pn.cpu().numpy(), threads=-1)
    terms = oracle.aggregate(1, A.abs().cpu().numpy(), ci.cpu().numpy(), deg.cpu().numpy(), 1.0,
                             pp.cpu().numpy(), pn.cpu().numpy(), threads=-1)
    assert_close(got, ref, what=name, terms=terms)


# ------------------------------------------------------------------------------------------ fused tcgen05 tile
@pytest.mark.parametrize("din,dout", [(64, 64), (64, 41), (64, 16), (128, 64), (32, 7), (128, 172)])
@pytest.mark.parametrize("xdtype", ["f32", "bf16"])
def test_fused_aggregate_gemm_tcgen05(din, dout, xdtype):
    """aggregate -> X*W in one kernel, the product on the tensor cores with bf16 operands and fp32
    accumulation in TMEM (csrc/fused_gemm.cu).  The aggregated features must match the fp32 oracle at the
    fp32 tolerance; the product is compared with the oracle's product of the same aggregated features
    within the bf16 operand rounding bound (2 x 2^-9 of the absolute terms; no reference of this
    precision exists)."""
    if xdtype == "bf16" and din == 32:
        pytest.skip("bf16 rows of 32 elements have no fused tile")
    rp, ci = GRAPHS["rmat"]()
    g = G(rp, ci, 32)
    X = rand_features(g.n, din, 300 + din)
    W = rand_weight(din, dout, 301 + dout)
    if xdtype == "bf16":
        Xt = torch.from_numpy(X).to(torch.bfloat16)
        X = Xt.float().numpy()
        dX = Xt.to(DEV)
    else:
        dX = dev(X)
    for mode in (2, 0, 3):
        if mode == 3:
            Xin = X * g.deg[:, None]
            if xdtype == "bf16":
                Xb = torch.from_numpy(Xin).to(torch.bfloat16)
                Xin, dXin = Xb.float().numpy(), Xb.to(DEV)
            else:
                dXin = dev(Xin)
            S_ref = (g.deg[:, None].astype(np.float64) * oracle.closed_form(0, Xin, rp, ci)).astype(np.float32)
            S_terms = g.deg[:, None] * oracle.closed_form(0, np.abs(Xin), rp, ci)
        else:
            Xin, dXin = X, dX
            S_ref = oracle.aggregate(mode, X, ci, g.deg, 0.5, g.pp, g.pn)
            S_terms = g.terms(mode, X)
        out, S = ops.aggregate_gemm_fused(mode, dXin, dev(W), *g.gargs(), g.d_deg, 0.5, *g.pargs(), 32, 32, 8)
        assert_close(S.cpu().numpy(), S_ref, what="fused x_agg mode %d" % mode, terms=S_terms)
        ref = S_ref.astype(np.float64) @ W.astype(np.float64)
        terms = np.abs(S_ref).astype(np.float64) @ np.abs(W).astype(np.float64)
        assert_close(out.cpu().numpy(), ref, rtol=5e-2, what="fused out mode %d" % mode, terms=terms)
    # same answer as the unfused fp32 operator, within the bf16 bound
    if xdtype == "f32":
        o2, S2 = ops.forward_gin(dX, dev(W), *g.gargs(), 0.5, *g.pargs(), 32, 32, 8)
        o1, S1 = ops.forward_gin_fused(dX, dev(W), *g.gargs(), 0.5, *g.pargs(), 32, 32, 8)
        assert_close(S1.cpu().numpy(), S2.cpu().numpy(), what="fused vs unfused x_agg", terms=g.terms(2, X))
        terms = np.abs(S2.cpu().numpy()).astype(np.float64) @ np.abs(W).astype(np.float64)
        assert_close(o1.cpu().numpy(), o2.cpu().numpy(), rtol=5e-2, what="fused vs unfused out", terms=terms)


def test_fused_tile_edge_cases():
    """Tail tile (N % 128 != 0), rows without neighbours, unsupported widths."""
    rng = np.random.default_rng(9)
    deg = rng.integers(0, 40, 300); deg[::7] = 0; deg[-1] = 3
    rp = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
    ci = rng.integers(0, 300, rp[-1]).astype(np.int32)
    g = G(rp, ci, 32)
    X, W = rand_features(300, 64, 1), rand_weight(64, 24, 2)
    out, S = ops.aggregate_gemm_fused(2, dev(X), dev(W), *g.gargs(), g.d_deg, 0.5, *g.pargs(), 32, 32, 8)
    S_ref = oracle.aggregate(2, X, ci, None, 0.5, g.pp, g.pn)
    assert_close(S.cpu().numpy(), S_ref, what="x_agg", terms=g.terms(2, X))
    assert np.count_nonzero(out.cpu().numpy()[deg == 0]) == 0
    terms = np.abs(S_ref).astype(np.float64) @ np.abs(W).astype(np.float64)
    assert_close(out.cpu().numpy(), S_ref.astype(np.float64) @ W, rtol=5e-2, what="out", terms=terms)
    with pytest.raises(RuntimeError, match="no fused tile"):
        ops.aggregate_gemm_fused(2, dev(rand_features(300, 48, 3)), dev(rand_weight(48, 8, 4)), *g.gargs(), g.d_deg, 0.5,
                                 *g.pargs(), 32, 32, 8)


@pytest.mark.parametrize("gather_dtype", ["fp32", "bf16"])
def test_fused_layers_in_the_product_path(gather_dtype):
    """GINConv / GCNConv(fused=True): the fused tcgen05 tile behind the layer API (forward of GIN, backward of GCN),
    compared with the unfused layers on the same weights: outputs and gradients agree within the bf16 operand bound
    (5e-2 of the absolute terms, as test_fused_aggregate_gemm_tcgen05), the aggregated matrices (saved X_agg / G, which feed
    dW) at the fp32 tolerance for fp32 rows.  A width without a fused tile silently keeps the unfused operators."""
    rp, ci = GRAPHS["rmat"]()
    g = G(rp, ci, 32)
    info = type("Info", (), {})()
    info.row_pointers, info.column_index, info.degrees, info.partPtr, info.part2Node = g.d_rp, g.d_ci, g.d_deg, g.d_pp, g.d_pn
    info.partSize, info.dimWorker, info.warpPerBlock = 32, 32, 8
    X = dev(rand_features(g.n, 64, 11) * 0.1)
    dO = dev(rand_features(g.n, 64, 12))

    def run(layer_cls, fused, din, dout, x):
        torch.manual_seed(5)
        layer = layer_cls(din, dout, gather_dtype=gather_dtype, fused=fused).to(DEV)
        xin = x.clone().requires_grad_(True)
        out = layer(xin, info)
        out.backward(dO[:, :dout].contiguous())
        return out.detach(), xin.grad.detach(), layer.weights.grad.detach()

    for cls, din, dout in ((layers.GINConv, 64, 64), (layers.GCNConv, 64, 64), (layers.GINConv, 64, 41)):
        o_f, dx_f, dw_f = run(cls, True, din, dout, X)
        o_u, dx_u, dw_u = run(cls, False, din, dout, X)
        scale = lambda t: float(t.abs().max())   # noqa: E731
        assert float((o_f - o_u).abs().max()) <= 5e-2 * scale(o_u), cls.__name__
        assert float((dx_f - dx_u).abs().max()) <= 5e-2 * scale(dx_u), cls.__name__
        assert float((dw_f - dw_u).abs().max()) <= 5e-2 * scale(dw_u), cls.__name__
        assert float((o_f - o_u).abs().max()) > 0 or float((dx_f - dx_u).abs().max()) > 0      # the fused kernel did run
    # 48-wide aggregated matrix: no fused tile -> identical to the unfused layer
    X48 = dev(rand_features(g.n, 48, 13) * 0.1)
    o_f, dx_f, dw_f = run(layers.GINConv, True, 48, 64, X48)
    o_u, dx_u, dw_u = run(layers.GINConv, False, 48, 64, X48)
    assert_close(o_f.cpu().numpy(), o_u.cpu().numpy(), what="fallback out")


def test_backward_fused_against_oracle():
    """ops.backward_fused: G = Ahat dOut and dX = G W^T in one kernel, dW = X^T G.  G-dependent outputs against the oracle."""
    rp, ci = GRAPHS["rmat"]()
    g = G(rp, ci, 32)
    X, W, dO = rand_features(g.n, 128, 1), rand_weight(128, 64, 2), rand_features(g.n, 64, 3)
    dX, dW = ops.backward_fused(dev(dO), dev(X), dev(W), *g.gargs(), g.d_deg, *g.pargs(), 32, 32, 8)
    odX, odW = oracle.backward(dO, X, W, rp, ci, g.deg, g.pp, g.pn)
    aG = oracle.closed_form(1, np.abs(dO).astype(np.float64), rp, ci)
    assert_close(dW.cpu().numpy(), odW, what="dW (fp32 G)", terms=np.abs(X.T).astype(np.float64) @ aG)
    assert_close(dX.cpu().numpy(), odX, rtol=5e-2, what="dX (bf16 tile)", terms=aG @ np.abs(W.T))


# ------------------------------------------------------------------------------------------ split CSRs, host pipeline
def test_aggregation_over_split_csrs_accumulates():
    """gnna_aggregate_part_f32_ex: the edges of a graph split over two CSRs (by column range, as the sharded
    path splits them by owner) and aggregated one after the other into the same output equal the whole graph."""
    import ctypes
    from gnnadvisor_osdi21_b200 import _lib
    rp, ci = GRAPHS["rmat"]()
    n = len(rp) - 1
    g = G(rp, ci, 32)
    X = rand_features(n, 64, 77)
    rows = np.repeat(np.arange(n), rp[1:] - rp[:-1])
    out = torch.empty(n, 64, device=DEV)
    p = lambda t: ctypes.c_void_p(t.data_ptr() if t.numel() else 0)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for mode in (0, 2, 3):
        Xin = X * g.deg[:, None] if mode == 3 else X
        dX = dev(Xin.astype(np.float32))
        for k, mask in enumerate((ci < n // 3, ci >= n // 3)):
            rp_k = np.concatenate([[0], np.cumsum(np.bincount(rows[mask], minlength=n))]).astype(np.int32)
            ci_k = ci[mask]
            pp_k, pn_k = oracle.build_part(32, rp_k, exact=True)
            d_rp, d_ci, d_pp, d_pn = dev(rp_k), dev(ci_k), dev(pp_k), dev(pn_k)
            _lib.check(_lib.load().gnna_aggregate_part_f32_ex(mode, k, p(dX), n, p(out), n, p(d_rp), p(d_ci),
                                                              p(g.d_deg) if mode == 3 else ctypes.c_void_p(0), 0.5, p(d_pp), p(d_pn),
                                                              64, d_pn.numel(), 32, 0, 0, st), "part")
        ref_mode = 1 if mode == 3 else mode
        assert_close(out.cpu().numpy(), oracle.aggregate(ref_mode, X, ci, g.deg, 0.5, g.pp, g.pn), what="split mode %d" % mode,
                     terms=g.terms(ref_mode, X))


def test_host_aggregator_pipeline():
    """HostAggregator: pinned host in / out, three overlapped stages, several steps with different inputs."""
    from gnnadvisor_osdi21_b200.host_pipeline import HostAggregator
    rp, ci = GRAPHS["rmat"]()
    g = G(rp, ci, 32)
    pipe = HostAggregator(g.d_rp, g.d_ci, g.d_deg, g.d_pp, g.d_pn, g.n, 32, mode=1)
    xs = [torch.from_numpy(rand_features(g.n, 32, 900 + i)).pin_memory() for i in range(5)]
    outs = [torch.empty(g.n, 32).pin_memory() for _ in range(5)]
    for x, o in zip(xs, outs):
        pipe.submit(x, o)
    pipe.drain()
    for x, o in zip(xs, outs):
        assert_close(o.numpy(), oracle.aggregate(1, x.numpy(), ci, g.deg, 1.0, g.pp, g.pn), what="host pipeline",
                     terms=g.terms(1, x.numpy()))


# ------------------------------------------------------------------------------------------ TMA-staged persistent kernel
@pytest.mark.parametrize("dim", [8, 16, 32, 64, 100, 128])
@pytest.mark.parametrize("ps", [3, 32, 64])
def test_staged_tma_kernel(dim, ps):
    """csrc/aggregate_staged.cu: group table + column indices streamed through cp.async.bulk into a shared-memory
    ring, persistent CTAs.  Same results as the oracle (bit-identical where a row is one group)."""
    from gnnadvisor_osdi21_b200 import _lib
    prev = _lib.set_staged(True)
    try:
        for gname in ("rmat", "uniform"):
            rp, ci = GRAPHS[gname]()
            g = G(rp, ci, ps)
            X = rand_features(g.n, dim, 40 + dim)
            dX = dev(X)
            assert_close(ops.SAG(dX, *g.gargs(), g.d_deg, *g.pargs(), ps, 32, 4).cpu().numpy(),
                         oracle.aggregate(0, X, ci, None, 1.0, g.pp, g.pn), what="staged SAG", terms=g.terms(0, X))
            assert_close(_gcn_agg(dX, g, 32, 4), oracle.aggregate(1, X, ci, g.deg, 1.0, g.pp, g.pn), what="staged GCN",
                         terms=g.terms(1, X))
            assert_close(_gin_agg(dX, g, 0.5, 32, 4), oracle.aggregate(2, X, ci, None, 0.5, g.pp, g.pn), what="staged GIN",
                         terms=g.terms(2, X))
        # single-group rows stay bit-identical; the F6 table (terminal 0) and isolated nodes go through the direct-load tiles
        rp, ci = make_graph("uniform", 3000, 20000, 33)
        g = G(rp, ci, 64)
        X = rand_features(g.n, dim, 41)
        assert np.array_equal(_gin_agg(dev(X), g, 0.5, 32, 4), oracle.aggregate(2, X, ci, None, 0.5, g.pp, g.pn))
        rng = np.random.default_rng(3)
        deg = rng.integers(0, 50, 900); deg[::5] = 0; deg[-1] = 0; deg[-2] = 40
        rp = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
        ci = rng.integers(0, 900, rp[-1]).astype(np.int32)
        X = rand_features(900, dim, 42)
        for exact in (True, False):
            g = G(rp, ci, 32, exact=exact)
            assert_close(ops.SAG(dev(X), *g.gargs(), g.d_deg, *g.pargs(), 32, 32, 4).cpu().numpy(),
                         oracle.aggregate(0, X, ci, None, 1.0, g.pp, g.pn), what="staged F6 exact=%s" % exact, terms=g.terms(0, X))
    finally:
        _lib.set_staged(prev)


# ------------------------------------------------------------------------------------------ run-based pipelined kernel
@pytest.mark.parametrize("dim", [8, 16, 32, 64, 100, 128])
@pytest.mark.parametrize("run,ps", [(1, 32), (4, 32), (8, 3), (8, 64), (64, 32)])
def test_run_based_kernel(dim, run, ps):
    """csrc/aggregate_runs.cu: a sub-warp owns `run` consecutive groups, prefetches the next group's table entries and
    ids, merges groups of one node in registers.  Same results as the oracle; rows that are one group stay bit-identical."""
    from gnnadvisor_osdi21_b200 import _lib
    prev = _lib.set_runs(run)
    try:
        for gname in ("rmat", "uniform"):
            rp, ci = GRAPHS[gname]()
            g = G(rp, ci, ps)
            X = rand_features(g.n, dim, 70 + dim)
            dX = dev(X)
            assert_close(ops.SAG(dX, *g.gargs(), g.d_deg, *g.pargs(), ps, 32, 4).cpu().numpy(),
                         oracle.aggregate(0, X, ci, None, 1.0, g.pp, g.pn), what="runs SAG", terms=g.terms(0, X))
            assert_close(_gcn_agg(dX, g, 32, 4), oracle.aggregate(1, X, ci, g.deg, 1.0, g.pp, g.pn), what="runs GCN",
                         terms=g.terms(1, X))
            assert_close(_gin_agg(dX, g, 0.5, 32, 4), oracle.aggregate(2, X, ci, None, 0.5, g.pp, g.pn), what="runs GIN",
                         terms=g.terms(2, X))
            if dim % 8 == 0:
                Xb = torch.from_numpy(X).to(torch.bfloat16)
                got = ops.aggregate_bf16(2, Xb.to(DEV), *g.gargs(), g.d_deg, 0.5, *g.pargs(), ps, 32, 4).cpu().numpy()
                Xr = Xb.float().numpy()
                assert_close(got, oracle.aggregate(2, Xr, ci, g.deg, 0.5, g.pp, g.pn), what="runs bf16 GIN", terms=g.terms(2, Xr))
        # single-group rows stay bit-identical; F6 table (terminal 0), isolated nodes, last node isolated
        rp, ci = make_graph("uniform", 3000, 20000, 33)
        g = G(rp, ci, 64)
        X = rand_features(g.n, dim, 71)
        assert np.array_equal(_gin_agg(dev(X), g, 0.5, 32, 4), oracle.aggregate(2, X, ci, None, 0.5, g.pp, g.pn))
        rng = np.random.default_rng(3)
        deg = rng.integers(0, 50, 900); deg[::5] = 0; deg[-1] = 0; deg[-2] = 40
        rp = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
        ci = rng.integers(0, 900, rp[-1]).astype(np.int32)
        X = rand_features(900, dim, 72)
        for exact in (True, False):
            g = G(rp, ci, 32, exact=exact)
            assert_close(ops.SAG(dev(X), *g.gargs(), g.d_deg, *g.pargs(), 32, 32, 4).cpu().numpy(),
                         oracle.aggregate(0, X, ci, None, 1.0, g.pp, g.pn), what="runs F6 exact=%s" % exact, terms=g.terms(0, X))
    finally:
        _lib.set_runs(prev)


def test_run_based_kernel_auto_rule_on_a_dense_graph():
    """Default setting (-1): the library routes bf16 rows and wide fp32 rows of graphs with >= 4 groups per node through
    the run-based kernel (csrc/aggregate_runs.cu auto_runs).  Same tolerance as everywhere; also the mixed-precision
    GCN operator on top of it with an odd width (padded bf16 rows)."""
    from gnnadvisor_osdi21_b200 import _lib
    assert _lib.set_runs(-1) == -1, "the default must be the library's own choice"
    rp, ci = make_graph("rmat", 400, 120000, 81)          # ~300 neighbours per node: ~10 groups per node
    g = G(rp, ci, 32)
    assert len(g.pn) >= 4 * g.n
    for dim in (16, 64, 128):
        X = rand_features(g.n, dim, 82 + dim)
        assert_close(ops.SAG(dev(X), *g.gargs(), g.d_deg, *g.pargs(), 32, 32, 4).cpu().numpy(),
                     oracle.aggregate(0, X, ci, None, 1.0, g.pp, g.pn), what="auto SAG %d" % dim, terms=g.terms(0, X))
        Xb = torch.from_numpy(X).to(torch.bfloat16)
        Xr = Xb.float().numpy()
        for mode in (0, 2):
            got = ops.aggregate_bf16(mode, Xb.to(DEV), *g.gargs(), g.d_deg, 0.5, *g.pargs(), 32, 32, 4).cpu().numpy()
            assert_close(got, oracle.aggregate(mode, Xr, ci, g.deg, 0.5, g.pp, g.pn), what="auto bf16 mode %d dim %d" % (mode, dim),
                         terms=g.terms(mode, Xr))
    X, W = rand_features(g.n, 24, 90), rand_weight(24, 41, 91)
    y, = ops.forward_mixed(dev(X), dev(W), *g.gargs(), 1.0 / g.d_deg, *g.pargs(), 32, 32, 4)
    inv = (1.0 / g.deg).astype(np.float32)
    aX, aW = np.abs(X).astype(np.float64), np.abs(W).astype(np.float64)
    ref = oracle.forward(X, W, rp, ci, inv, g.pp, g.pn)[0]
    terms = oracle.aggregate(1, aX @ aW, ci, inv, 1.0, g.pp, g.pn)
    assert_close(y.cpu().numpy(), ref, rtol=1e-2, what="auto mixed forward", terms=4 * terms)


@pytest.mark.parametrize("kind,n", [("rmat", 232965), ("rmat", 111059956), ("uniform", 3327), ("rmat", 2)])
def test_pair_stream_kernel_equals_the_torch_definition(kind, n):
    """csrc/graphgen.cu against graph.stream_pairs' torch integer ops on the CPU: the same pairs bit for bit, so a graph is
    the same graph wherever it is generated (the reference arm builds it on the host, the ranks shard by shard on GPUs)."""
    for start, count in ((0, 100000), (123456789012, 50000)):
        cs, cd = graph.stream_pairs(n, start, count, kind=kind, seed=20211, device="cpu")
        gs, gd = graph.stream_pairs(n, start, count, kind=kind, seed=20211, device=DEV)
        assert torch.equal(gs.cpu(), cs) and torch.equal(gd.cpu(), cd) and (cs.numel() > 0 or n == 2)
    rp_c, ci_c = graph.synth_graph(5000, 160000, kind=kind, seed=9)
    rp_g, ci_g = graph.synth_graph(5000, 160000, kind=kind, seed=9, device=DEV)
    assert torch.equal(rp_g.cpu(), rp_c) and torch.equal(ci_g.cpu(), ci_c)


def _sgemm(A, B, ta=False):
    import ctypes
    from gnnadvisor_osdi21_b200 import _lib
    m = A.shape[1] if ta else A.shape[0]
    k = A.shape[0] if ta else A.shape[1]
    n = B.shape[1]
    C = torch.full((m, n), float("nan"), device=DEV)
    p = lambda t: ctypes.c_void_p(t.data_ptr())   # noqa: E731
    _lib.check(_lib.load().gnna_sgemm_f32(int(ta), 0, m, n, k, p(A), p(B), p(C), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "sgemm")
    return C


@pytest.fixture(params=[1, 2], ids=["lockstep", "warp-specialised"])
def tc_gemm(request):
    """Force both tall-skinny products onto the tensor cores (GNNA_TC_GEMM / gnna_set_tc_gemm): 1 = lockstep kernel,
    2 = warp-specialised (the default mode 3 uses it for X^T*G only)."""
    from gnnadvisor_osdi21_b200 import _lib
    prev = _lib.set_tc_gemm(request.param)
    yield request.param
    _lib.set_tc_gemm(prev)


@pytest.mark.parametrize("m,k,n", [(10000, 602, 64), (9000, 300, 47), (8200, 257, 16), (8200, 1001, 8), (8229, 256, 128),
                                   (58241, 602, 41), (8192, 1433, 16)])
def test_tensor_core_gemm_nn_is_fp32_grade(m, k, n, tc_gemm):
    """X*W on the tensor cores (csrc/gemm_tf32x3.cu: tcgen05 kind::tf32, 3xTF32 split) against a float64 product:
    the error is bounded by 4e-6 of the absolute terms -- SGEMM-grade, 25x inside the 1e-4 parity bar -- for every copy
    width (K % 4 == 0, even, odd), ragged row tiles, N not a multiple of 16, and it agrees with the cuBLAS path."""
    from gnnadvisor_osdi21_b200 import _lib
    g = torch.Generator(device=DEV).manual_seed(m + k + n)
    A = torch.randn(m, k, device=DEV, generator=g)
    B = torch.randn(k, n, device=DEV, generator=g)
    _lib.launch_count(reset=True)
    C = _sgemm(A, B)
    assert _lib.launch_count() == 1                      # our kernel, not a cuBLAS call
    ref = A.double() @ B.double()
    terms = A.abs().double() @ B.abs().double()
    assert bool(torch.isfinite(C).all())
    assert float(((C.double() - ref).abs() / terms).max()) <= 4e-6
    prev = _lib.set_tc_gemm(0)
    try:
        Cb = _sgemm(A, B)
    finally:
        _lib.set_tc_gemm(prev)
    assert float(((Cb.double() - ref).abs() / terms).max()) <= 4e-6
    assert float(((C - Cb).abs().double() / terms).max()) <= 4e-6


@pytest.mark.parametrize("rows,m,n", [(12000, 602, 64), (9000, 300, 47), (8200, 256, 41), (60000, 257, 7), (8192, 384, 128)])
def test_tensor_core_gemm_tn_is_fp32_grade(rows, m, n, tc_gemm):
    """X^T*G (reduced over the node dimension, split over the SMs, merged with reductions) against a float64 product."""
    from gnnadvisor_osdi21_b200 import _lib
    g = torch.Generator(device=DEV).manual_seed(rows + m + n)
    A = torch.randn(rows, m, device=DEV, generator=g)
    B = torch.randn(rows, n, device=DEV, generator=g)
    _lib.launch_count(reset=True)
    C = _sgemm(A, B, ta=True)
    assert _lib.launch_count() == 1
    ref = A.double().t() @ B.double()
    terms = A.abs().double().t() @ B.abs().double()
    assert bool(torch.isfinite(C).all())
    assert float(((C.double() - ref).abs() / terms).max()) <= 4e-6


def test_default_product_routing():
    """Default mode 3: X^T*G runs on the warp-specialised tcgen05 kernel (one launch of ours), X*W on cuBLAS (none)."""
    from gnnadvisor_osdi21_b200 import _lib
    assert _lib.set_tc_gemm(3) == 3
    X, W, G = torch.randn(12000, 602, device=DEV), torch.randn(602, 64, device=DEV), torch.randn(12000, 64, device=DEV)
    _lib.launch_count(reset=True)
    C = _sgemm(X, G, ta=True)
    assert _lib.launch_count() == 1
    terms = X.abs().double().t() @ G.abs().double()
    assert float(((C.double() - X.double().t() @ G.double()).abs() / terms).max()) <= 4e-6
    _lib.launch_count(reset=True)
    _sgemm(X, W)
    assert _lib.launch_count() == 0


def test_small_or_odd_products_stay_on_cublas(tc_gemm):
    from gnnadvisor_osdi21_b200 import _lib
    A, B = torch.randn(500, 48, device=DEV), torch.randn(48, 16, device=DEV)
    _lib.launch_count(reset=True)
    C = _sgemm(A, B)
    assert _lib.launch_count() == 0                      # cuBLAS launches are not counted as ours
    assert torch.allclose(C, A @ B, rtol=1e-4, atol=1e-4)


# ------------------------------------------------------------------------------------------ single-launch path (small graphs)
def _launches(fn):
    from gnnadvisor_osdi21_b200 import _lib
    _lib.launch_count(reset=True)
    r = fn()
    return r, _lib.launch_count()


@pytest.mark.small_path
@pytest.mark.parametrize("gname", sorted(GRAPHS))
@pytest.mark.parametrize("dim", [1, 2, 3, 4, 6, 7, 16, 32, 41, 47, 64, 100, 128, 172, 300, 512])
def test_small_path_all_modes_bit_identical_to_the_oracle(gname, dim):
    """One launch per aggregation; groups of a row are merged in ascending order in registers, which is the oracle's order:
    SAG, GIN and exact-rounding GCN are BIT-identical to it on every row, default GCN within the tolerance."""
    from gnnadvisor_osdi21_b200 import _lib
    rp, ci = GRAPHS[gname]()
    g = G(rp, ci, 32)
    X = rand_features(g.n, dim, 300 + dim)
    dX = dev(X)
    out, n = _launches(lambda: ops.SAG(dX, *g.gargs(), g.d_deg, *g.pargs(), 32, 32, 8))
    assert n == 1
    assert np.array_equal(out.cpu().numpy(), oracle.aggregate(0, X, ci, None, 1.0, g.pp, g.pn))
    gin, n = _launches(lambda: _gin_agg(dX, g, 0.37, 32, 8))
    assert n == 1 and np.array_equal(gin, oracle.aggregate(2, X, ci, None, 0.37, g.pp, g.pn))
    gcn, n = _launches(lambda: _gcn_agg(dX, g, 32, 8))
    assert n == 1
    assert_close(gcn, oracle.aggregate(1, X, ci, g.deg, 1.0, g.pp, g.pn), what="GCN", terms=g.terms(1, X))
    prev = _lib.set_gcn_exact(True)
    try:
        assert np.array_equal(_gcn_agg(dX, g, 32, 8), oracle.aggregate(1, X, ci, g.deg, 1.0, g.pp, g.pn))
    finally:
        _lib.set_gcn_exact(prev)


@pytest.mark.small_path
@pytest.mark.parametrize("ps", [1, 2, 5, 32, 512])
def test_small_path_every_row_is_written_once(ps):
    """No zero-fill launch: rows without neighbours at the head, in the middle and at the tail must come out zero even
    when the output buffer held garbage; F6 tables (terminal 0) drop the last group as the reference does."""
    from gnnadvisor_osdi21_b200 import _lib as _l
    _l.set_small_parts(1 << 20)                         # partSize 1 makes a 17 K-group table: keep it on the single-launch path
    rng = np.random.default_rng(31)
    deg = rng.integers(0, 70, 500)
    deg[:7] = 0; deg[100:160] = 0; deg[-9:] = 0; deg[-10] = 45; deg[250] = 1900          # a hub of many groups too
    rp = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
    ci = rng.integers(0, 500, rp[-1]).astype(np.int32)
    X = rand_features(500, 24, 32)
    for exact in (True, False):
        g = G(rp, ci, ps, exact=exact)
        if not exact:
            assert g.pp[-1] == 0
        import ctypes
        from gnnadvisor_osdi21_b200 import _lib
        out = torch.full((500, 24), float("nan"), device=DEV)
        p = lambda t: ctypes.c_void_p(t.data_ptr())   # noqa: E731
        _lib.check(_lib.load().gnna_sag_f32(p(dev(X)), p(out), p(g.d_rp), p(g.d_ci), p(g.d_pp), p(g.d_pn), 500, 24, g.d_pn.numel(),
                                            ps, 32, 4, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "sag")
        got = out.cpu().numpy()
        assert np.array_equal(got, oracle.aggregate(0, X, ci, None, 1.0, g.pp, g.pn)), "exact=%s" % exact
        assert np.count_nonzero(got[deg == 0]) == 0


@pytest.mark.small_path
def test_small_path_layer_operators_and_limit():
    """The four layer operators on a Cora-sized graph (the path BASELINE.json's configs #1/#2 take), and the switch:
    above the limit the general path (memset + pre-scale + gather: more launches) runs and agrees."""
    from gnnadvisor_osdi21_b200 import _lib
    rp, ci = make_graph("uniform", 2708, 10556, 20211)
    g = G(rp, ci, 32)
    X, W, dO = rand_features(g.n, 48, 1), rand_weight(48, 16, 2), rand_features(g.n, 16, 3)
    a = (*g.gargs(), g.d_deg, *g.pargs(), 32, 16, 8)
    out, = ops.forward(dev(X), dev(W), *a)
    assert_close(out.cpu().numpy(), oracle.forward(X, W, rp, ci, g.deg, g.pp, g.pn)[0], what="fwd",
                 terms=oracle.closed_form(1, np.abs(X).astype(np.float64) @ np.abs(W), rp, ci))
    dX, dW = ops.backward(dev(dO), dev(X), dev(W), *a)
    odX, odW = oracle.backward(dO, X, W, rp, ci, g.deg, g.pp, g.pn)
    aG = oracle.closed_form(1, np.abs(dO).astype(np.float64), rp, ci)
    assert_close(dX.cpu().numpy(), odX, what="dX", terms=aG @ np.abs(W.T))
    assert_close(dW.cpu().numpy(), odW, what="dW", terms=np.abs(X.T).astype(np.float64) @ aG)
    o, S = ops.forward_gin(dev(X), dev(W), *g.gargs(), 0.5, *g.pargs(), 32, 16, 2)
    oo, oS = oracle.forward_gin(X, W, rp, ci, 0.5, g.pp, g.pn)
    assert np.array_equal(S.cpu().numpy(), oS)
    assert_close(o.cpu().numpy(), oo, what="gin out", terms=np.abs(oS).astype(np.float64) @ np.abs(W))
    small, n_small = _launches(lambda: _gcn_agg(dev(X), g, 16, 8))
    prev = _lib.set_small_parts(100)                      # the table has ~2.9 K groups: now above the limit
    try:
        general, n_general = _launches(lambda: _gcn_agg(dev(X), g, 16, 8))
    finally:
        _lib.set_small_parts(prev)
    assert n_small == 1 and n_general >= 2
    assert_close(small, general, what="small vs general path", terms=g.terms(1, X))
