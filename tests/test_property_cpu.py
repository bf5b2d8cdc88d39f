"""Property-based (hypothesis) checks of the integer host code of the path against its references: the group table
against the oracle's restatement of GNNAdvisor.cpp:210-251, the CSR builder against scipy's coo -> csr
(dataset.py:108-111), the text loader against the reference's per-line loop (dataset.py:62-72).  Bit-exact, CPU only."""
import numpy as np
import scipy.sparse as sp
import torch
from hypothesis import given, settings, strategies as st

import oracle
from gnnadvisor_osdi21_b200 import graph, ops

degree_lists = st.lists(st.one_of(st.just(0), st.integers(0, 9), st.integers(10, 300)), min_size=1, max_size=60)


@settings(max_examples=60, deadline=None, derandomize=True)
@given(degs=degree_lists, ps=st.sampled_from([1, 2, 3, 7, 32, 64, 512]))
def test_build_part_host_equals_the_oracle_for_any_degree_sequence(degs, ps):
    rp = np.concatenate([[0], np.cumsum(degs)]).astype(np.int32)
    t = torch.from_numpy(rp)
    # exact table (what this runtime consumes): terminal always indptr[-1]
    pp, pn = ops.build_part_exact(ps, t)
    opp, opn = oracle.build_part(ps, rp, exact=True)
    assert np.array_equal(pp.numpy(), opp) and np.array_equal(pn.numpy(), opn)
    assert len(pn) == sum(-(-d // ps) for d in degs) and int(pp[-1]) == int(rp[-1])
    assert all(0 < int(pp[i + 1]) - int(pp[i]) <= ps for i in range(len(pn)))        # every group has 1..ps neighbours
    # compat table: the reference's float32 tensors bit for bit, F6 (terminal 0 after an isolated last node) included
    cpp, cpn = ops.build_part(ps, t, compat=True)
    fpp, fpn = oracle.build_part_f32(ps, rp)
    assert cpp.dtype == torch.float32 and np.array_equal(cpp.numpy(), fpp) and np.array_equal(cpn.numpy(), fpn)
    if degs[-1] == 0 and len(pn):
        assert float(cpp[-1]) == 0.0


@settings(max_examples=60, deadline=None, derandomize=True)
@given(n=st.integers(1, 40), pairs=st.lists(st.tuples(st.integers(0, 39), st.integers(0, 39)), max_size=200),
       dup=st.integers(1, 3))
def test_native_csr_equals_scipy_for_any_edge_list(n, pairs, dup):
    pairs = [(a % n, b % n) for a, b in pairs] * dup
    src = np.array([p[0] for p in pairs], dtype=np.int64)
    dst = np.array([p[1] for p in pairs], dtype=np.int64)
    rp, ci = graph.csr_from_edges(src, dst, n, native=True)
    csr = sp.coo_matrix((np.ones(len(src)), (src, dst)), shape=(n, n)).tocsr()
    csr.sort_indices()
    assert np.array_equal(rp.numpy(), csr.indptr) and np.array_equal(ci.numpy(), csr.indices)
    rp2, ci2 = graph.csr_from_edges(src, dst, n, native=False)
    assert torch.equal(rp, rp2) and torch.equal(ci, ci2)


blank = st.sampled_from([" ", "\t", "  ", " \t "])


@settings(max_examples=40, deadline=None, derandomize=True)
@given(edges=st.lists(st.tuples(st.integers(0, 10 ** 9), st.integers(0, 10 ** 9), blank, st.sampled_from(["", " ", "\t"]),
                                st.sampled_from(["", " "])), max_size=50),
       crlf=st.booleans(), final_newline=st.booleans())
def test_text_loader_equals_the_reference_loop_for_any_spacing(tmp_path_factory, edges, crlf, final_newline):
    text = ("\r\n" if crlf else "\n").join("%s%d%s%d%s" % (lead, a, sep, b, trail) for a, b, sep, trail, lead in edges)
    if final_newline and edges:
        text += "\r\n" if crlf else "\n"
    path = str(tmp_path_factory.mktemp("txt") / "g.txt")
    with open(path, "w", newline="") as f:
        f.write(text)
    s, d, n = graph.load_edge_text(path)
    # dataset.py:62-72 (split() takes the same blanks, '\r' included)
    rs, rd = [], []
    with open(path, newline="") as fp:
        for line in fp:
            a, b = line.strip("\n").split()
            rs.append(int(a))
            rd.append(int(b))
    assert s.tolist() == rs and d.tolist() == rd
    assert n == (max(rs + rd) + 1 if rs else 0)


@settings(max_examples=60, deadline=None, derandomize=True)
@given(n=st.integers(1, 50), pairs=st.lists(st.tuples(st.integers(0, 49), st.integers(0, 49)), max_size=300),
       window=st.sampled_from([0, 1, 2, 5, 1000]))
def test_reorder_is_a_permutation_for_any_edge_list_and_window(n, pairs, window):
    """rabbit.reorder's contract (reorder.cpp:235-290): every vertex gets exactly one new id, isolated ones included, whatever
    the edge list holds (self loops, duplicates, one direction only); and the same call gives the same answer."""
    from gnnadvisor_osdi21_b200 import reorder
    e = torch.tensor([[a % n for a, _ in pairs], [b % n for _, b in pairs]], dtype=torch.int32).reshape(2, -1)
    perm = reorder.permutation(e, n, window=window)
    assert sorted(perm.tolist()) == list(range(n))
    assert torch.equal(perm, reorder.permutation(e, n, window=window))


_REF = None


def _reference():
    """The reference's own extension, compiled from /root/reference by oracle/build_ref.py (oracle/_ref/, git-ignored).
    build_part is host code, so it runs without a GPU."""
    global _REF
    if _REF is None:
        import build_ref
        _REF = build_ref.load_ref() or False
    return _REF


@settings(max_examples=80, deadline=None, derandomize=True)
@given(degs=degree_lists, ps=st.sampled_from([1, 2, 3, 7, 32, 64, 512]))
def test_build_part_equals_the_reference_itself_run_live(degs, ps):
    """Not a golden file: the reference's build_part (GNNAdvisor.cpp:210-251) executed here on the same indptr.  The
    compat table of the product and the oracle's float table must be its two float32 tensors bit for bit (F5 rounding,
    F6 terminal); the exact table must equal it after the caller's `.int()` (GNNA_main.py:109-110) whenever the last node
    has a neighbour."""
    ref = _reference()
    if not ref:
        import pytest
        pytest.skip("oracle/_ref/GNNAdvisor_ref.so not built (needs /root/reference)")
    rp = np.concatenate([[0], np.cumsum(degs)]).astype(np.int32)
    t = torch.from_numpy(rp)
    rpp, rpn = ref.build_part(ps, t)
    assert rpp.dtype == torch.float32 and rpn.dtype == torch.float32
    cpp, cpn = ops.build_part(ps, t, compat=True)
    assert torch.equal(cpp, rpp) and torch.equal(cpn, rpn)
    fpp, fpn = oracle.build_part_f32(ps, rp)
    assert np.array_equal(fpp, rpp.numpy()) and np.array_equal(fpn, rpn.numpy())
    epp, epn = ops.build_part_exact(ps, t)
    assert torch.equal(epn, rpn.int())
    if degs[-1] > 0:
        assert torch.equal(epp, rpp.int())
    else:                                                       # F6: the reference leaves the terminal 0
        assert torch.equal(epp[:-1], rpp.int()[:-1]) and int(epp[-1]) == int(rp[-1])


class _Dataset:
    def __init__(self, n, e, feat, span):
        self.num_nodes, self.num_features = n, feat
        self.avg_degree, self.avg_edgeSpan = e / n, span
        self.reorder_flag, self.row_pointers, self.column_index, self.reordered = None, None, None, 0

    def rabbit_reorder(self):
        self.reordered += 1


@settings(max_examples=200, deadline=None, derandomize=True)
@given(n=st.integers(2, 3_000_000), avg=st.floats(1.0, 600.0), feat=st.integers(1, 4000), hid=st.integers(1, 2048),
       smem=st.sampled_from([16, 48, 64, 100, 164, 227]), span=st.floats(0.0, 1e6), rabbit=st.booleans(), manual=st.booleans())
def test_decider_equals_the_reference_decider_run_live(n, avg, feat, hid, smem, span, rabbit, manual):
    """param.py:52-120 of the reference, imported here, against param.InputProperty on the same dataset statistics: partSize,
    the per-layer dimWorker / warpPerBlock, the reorder decision and how often the dataset is asked to reorder."""
    import importlib.util
    import os
    import pytest
    from gnnadvisor_osdi21_b200 import param
    path = "/root/reference/GNNAdvisor/param.py"
    if not os.path.exists(path):
        pytest.skip("the reference tree is only mounted in the authoring container")
    spec = importlib.util.spec_from_file_location("ref_param_live", path)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    import contextlib
    import io
    out = []
    for cls in (param.InputProperty, ref.inputProperty):
        ds = _Dataset(n, int(avg * n), feat, span)
        text = io.StringIO()
        with contextlib.redirect_stdout(text):               # verbose mode: the lines the reference's log scrapers read
            p = cls(None, None, None, 32, 32, 4, smem, hiddenDim=hid, dataset_obj=ds, enable_rabbit=rabbit, manual_mode=manual,
                    verbose=True)
            p.decider()
            a = (p.set_input().dimWorker, p.warpPerBlock)
            p.print_param()
            b = (p.set_hidden().dimWorker, p.warpPerBlock)
            p.print_param()
        out.append((p.partSize, a, b, bool(p.reorder_status), ds.reorder_flag, ds.reordered, text.getvalue()))
    assert out[0] == out[1]
