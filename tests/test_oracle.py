"""CPU tests of the checker itself: the oracle against the reference's golden vectors and against
independent closed forms.  No GPU, no product code under test."""
import os

import numpy as np
import pytest

import oracle
from helpers import assert_close, golden_case, golden_terms, make_graph, rand_features, rand_weight

PART_SIZES = (1, 2, 3, 8, 32, 64)


@pytest.fixture(scope="module")
def bp_golden(golden_dir):
    return np.load(os.path.join(golden_dir, "build_part.npz"))


def _names(g):
    return sorted({k.split("/")[1] for k in g.files if k.startswith("partPtr/")})


def test_build_part_oracle_matches_reference_bit_exact(bp_golden):
    """oracle_build_part_f32 == the reference's build_part (float32 tensors, F5/F6 included)."""
    names = _names(bp_golden)
    assert "last_isolated" in names and "chesapeake" in names
    for name in names:
        indptr = bp_golden["indptr/" + name]
        for ps in PART_SIZES:
            pp, pn = oracle.build_part_f32(ps, indptr)
            ref_pp, ref_pn = bp_golden["partPtr/%s/%d" % (name, ps)], bp_golden["part2Node/%s/%d" % (name, ps)]
            assert pp.dtype == np.float32 and np.array_equal(pp, ref_pp), (name, ps)
            assert np.array_equal(pn, ref_pn), (name, ps)
            ipp, ipn = oracle.build_part(ps, indptr)                      # what `.int()` makes of it
            assert np.array_equal(ipp, ref_pp.astype(np.int32)) and np.array_equal(ipn, ref_pn.astype(np.int32))


def test_build_part_f6_terminal_rule(bp_golden):
    """Last node isolated: the reference leaves the terminal at 0; exact mode fixes it."""
    indptr = bp_golden["indptr/last_isolated"]
    pp, _ = oracle.build_part(32, indptr)
    assert pp[-1] == 0 and indptr[-1] > 0
    epp, _ = oracle.build_part(32, indptr, exact=True)
    assert epp[-1] == indptr[-1]
    assert np.array_equal(pp[:-1], epp[:-1])


def test_build_part_f5_float_rounding():
    """Offsets above 2^24 are rounded by the reference's float32 tables."""
    deg = np.array([2 ** 24 + 1, 3], dtype=np.int64)
    indptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
    ps = 2 ** 23
    pp, _ = oracle.build_part(ps, indptr)
    epp, _ = oracle.build_part(ps, indptr, exact=True)
    assert epp[-1] == 2 ** 24 + 4 and epp[3] == 2 ** 24 + 1
    assert pp[3] == 2 ** 24          # 16777217 -> 16777216
    assert not np.array_equal(pp, epp)


def test_degrees():
    indptr = np.array([0, 0, 1, 5, 5, 14], dtype=np.int32)
    d = oracle.degrees(indptr)
    assert d.dtype == np.float32
    assert np.array_equal(d, np.sqrt(np.array([1, 1, 4, 1, 9], dtype=np.float32)))


@pytest.mark.parametrize("kind,n,e", [("uniform", 300, 1500), ("rmat", 500, 9000)])
@pytest.mark.parametrize("ps", [1, 4, 32])
def test_aggregate_matches_closed_form(kind, n, e, ps):
    """SAG = A@X, GCN = diag(n) A diag(n) @ X, GIN = eps*A@X in float64 (SURVEY.md 8c closed forms)."""
    rp, ci = make_graph(kind, n, e, seed=3)
    X = rand_features(n, 24, seed=4)
    pp, pn = oracle.build_part(ps, rp, exact=True)
    deg = oracle.degrees(rp)
    for mode in (oracle.MODE_SAG, oracle.MODE_GCN, oracle.MODE_GIN):
        got = oracle.aggregate(mode, X, ci, deg, 0.5, pp, pn)
        assert_close(got, oracle.closed_form(mode, X, rp, ci, 0.5), rtol=1e-4, what="mode %d" % mode)


def test_aggregate_mt_is_bitwise_serial():
    rp, ci = make_graph("rmat", 2000, 60000, seed=5)
    X = rand_features(2000, 17, seed=6)
    deg = oracle.degrees(rp)
    for ps in (2, 32):
        pp, pn = oracle.build_part(ps, rp, exact=True)
        for mode in (0, 1, 2):
            a = oracle.aggregate(mode, X, ci, deg, 0.5, pp, pn, threads=0)
            b = oracle.aggregate(mode, X, ci, deg, 0.5, pp, pn, threads=-1)
            c = oracle.aggregate(mode, X, ci, deg, 0.5, pp, pn, threads=3)
            assert np.array_equal(a, b) and np.array_equal(a, c)


def test_gcn_product_is_rounded_before_the_add():
    """fl(fl(n_i*n_j)*x) + acc, never an FMA (kernel.cu:389,403): one hand-computed row."""
    rp = np.array([0, 2, 3, 4], dtype=np.int32)
    ci = np.array([1, 2, 0, 0], dtype=np.int32)
    X = np.array([[1.0000001], [3.3333333], [7.7777777]], dtype=np.float32)
    deg = oracle.degrees(rp)
    pp, pn = oracle.build_part(32, rp, exact=True)
    got = oracle.aggregate(oracle.MODE_GCN, X, ci, deg, 1.0, pp, pn)
    f = np.float32
    w1, w2 = f(deg[0] * deg[1]), f(deg[0] * deg[2])
    want = f(f(f(0) + f(w1 * X[1, 0])) + f(w2 * X[2, 0]))
    assert got[0, 0] == want


def test_sag_on_ones_is_degree():
    """The reference's only own check (unitest.py:27,54-63): SAG(ones) == deduplicated row degree."""
    rp, ci = make_graph("rmat", 1000, 20000, seed=9)
    pp, pn = oracle.build_part(32, rp)
    out = oracle.SAG(np.ones((1000, 16), dtype=np.float32), rp, ci, None, pp, pn)
    deg = (rp[1:] - rp[:-1]).astype(np.float32)
    assert np.array_equal(out, np.repeat(deg[:, None], 16, axis=1))


def test_layer_ops_match_closed_forms():
    n, din, dout = 400, 33, 12
    rp, ci = make_graph("rmat", n, 5000, seed=11)
    X, W, dO = rand_features(n, din, 12), rand_weight(din, dout, 13), rand_features(n, dout, 14)
    pp, pn = oracle.build_part(8, rp)
    deg = oracle.degrees(rp)
    X64, W64, dO64 = X.astype(np.float64), W.astype(np.float64), dO.astype(np.float64)
    out, = oracle.forward(X, W, rp, ci, deg, pp, pn)
    assert_close(out, oracle.closed_form(1, X64 @ W64, rp, ci), rtol=1e-4, what="forward")
    dX, dW = oracle.backward(dO, X, W, rp, ci, deg, pp, pn)
    G = oracle.closed_form(1, dO64, rp, ci)
    assert_close(dX, G @ W64.T, rtol=1e-4, what="backward dX")
    assert_close(dW, X64.T @ G, rtol=1e-4, what="backward dW")
    o, S = oracle.forward_gin(X, W, rp, ci, 0.5, pp, pn)
    S64 = oracle.closed_form(2, X64, rp, ci, 0.5)
    assert_close(S, S64, rtol=1e-4, what="gin agg")
    assert_close(o, S64 @ W64, rtol=1e-4, what="gin out")
    dXg, dWg = oracle.backward_gin(dO, S, W, rp, ci, 0.5, pp, pn)
    assert_close(dWg, S64.T @ dO64, rtol=1e-4, what="gin dW")
    assert_close(dXg, oracle.closed_form(2, dO64 @ W64.T, rp, ci, 0.5), rtol=1e-4, what="gin dX")


def test_refgpu_golden_pins_the_oracle(golden_dir):
    """Outputs of the reference CUDA kernels run on a B200 (oracle/make_golden_refgpu.py)."""
    path = os.path.join(golden_dir, "refgpu.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/refgpu.npz not generated yet")
    g = np.load(path)
    ncases = len({k.split("/")[0] for k in g.files})
    assert ncases >= 4
    for c in range(ncases):
        k = "case%d/" % c
        din, dout, ps, dw, wpb = [int(v) for v in g[k + "meta"]]
        rp, ci, pp, pn = g[k + "row_ptr"], g[k + "col_idx"], g[k + "partPtr"], g[k + "part2Node"]
        X, W, dO, what = golden_case(g, k)
        deg = oracle.degrees(rp)
        opp, opn = oracle.build_part(ps, rp)
        assert np.array_equal(opp, pp) and np.array_equal(opn, pn)
        t = golden_terms(g, k, oracle)
        if "SAG" in what:
            assert_close(oracle.SAG(X, rp, ci, deg, pp, pn), g[k + "SAG"], what=k + "SAG")
        if "gcn" in what:
            assert_close(oracle.forward(X, W, rp, ci, deg, pp, pn)[0], g[k + "forward"], what=k + "forward", terms=t["fwd"])
            dX, dW = oracle.backward(dO, X, W, rp, ci, deg, pp, pn)
            assert_close(dX, g[k + "backward_dX"], what=k + "backward_dX", terms=t["dX"])
            assert_close(dW, g[k + "backward_dW"], what=k + "backward_dW", terms=t["dW"])
        if "gin" in what:
            o, S = oracle.forward_gin(X, W, rp, ci, 0.5, pp, pn)
            assert_close(S, g[k + "forward_gin_agg"], what=k + "gin agg")
            assert_close(o, g[k + "forward_gin"], what=k + "gin out", terms=t["gin_out"])
            dXg, dWg = oracle.backward_gin(dO, g[k + "forward_gin_agg"], W, rp, ci, 0.5, pp, pn)
            assert_close(dXg, g[k + "backward_gin_dX"], what=k + "gin dX", terms=t["gin_dX"])
            assert_close(dWg, g[k + "backward_gin_dW"], what=k + "gin dW", terms=t["gin_dW"])
