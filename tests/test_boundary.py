"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol the header
declares, the host-side build_part matches the reference bit for bit, argument checking follows the
reference's CHECK_INPUT, and the product never touches oracle/.  No GPU compute calls here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import oracle
from gnnadvisor_osdi21_b200 import _lib, ops, param

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PART_SIZES = (1, 2, 3, 8, 32, 64)


def _header_functions():
    text = open(os.path.join(ROOT, "include", "gnna_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"GNNA_API[^;(]*?\b(gnna_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = _header_functions()
    assert len(names) >= 15 and "gnna_forward_f32" in names and "gnna_build_part_host" in names
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "libgnna_b200.so does not export " + n
    # and the ctypes table mirrors the header, one entry per declared function
    assert sorted(_lib.SIGNATURES) == names
    assert _lib.load().gnna_abi_version() == 1


def test_library_has_no_torch_dependency():
    out = os.popen("ldd %s" % _lib.LIB_PATH).read()
    assert "libtorch" not in out and "libc10" not in out
    assert "libcublas" in out


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gnnadvisor_osdi21_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), f
                assert "gnna_oracle" not in text and "oracle/_" not in text, f


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libgnna_b200.so")
    with pytest.raises(ImportError, match="no CPU or torch fallback"):
        _lib.load()


@pytest.fixture(scope="module")
def bp_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "build_part.npz"))


def test_build_part_matches_reference_golden(bp_golden):
    """ops.build_part(compat=True) on a CPU IntTensor returns the reference's float32 tensors verbatim (F6 included);
    the default differs from it only in the terminal entry of a graph whose last node is isolated."""
    names = sorted({k.split("/")[1] for k in bp_golden.files if k.startswith("partPtr/")})
    for name in names:
        indptr = torch.from_numpy(bp_golden["indptr/" + name])
        for ps in PART_SIZES:
            pp, pn = ops.build_part(ps, indptr, compat=True)
            assert pp.dtype == torch.float32 and pn.dtype == torch.float32 and not pp.is_cuda
            assert np.array_equal(pp.numpy(), bp_golden["partPtr/%s/%d" % (name, ps)]), (name, ps)
            assert np.array_equal(pn.numpy(), bp_golden["part2Node/%s/%d" % (name, ps)]), (name, ps)
            dpp, dpn = ops.build_part(ps, indptr)                      # default: same table, terminal always indptr[-1]
            assert dpp.dtype == torch.float32 and np.array_equal(dpn.numpy(), pn.numpy())
            assert np.array_equal(dpp.numpy()[:-1], pp.numpy()[:-1]) and (dpp.numel() == 1 or int(dpp[-1]) == int(indptr[-1]))
            epp, epn = ops.build_part_exact(ps, indptr)
            opp, opn = oracle.build_part(ps, indptr.numpy(), exact=True)
            assert epp.dtype == torch.int32 and np.array_equal(epp.numpy(), opp) and np.array_equal(epn.numpy(), opn)


def test_build_part_large_random_vs_oracle():
    rng = np.random.default_rng(5)
    deg = rng.integers(0, 300, 200_000)
    deg[rng.integers(0, len(deg), 50)] = 20_000
    indptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
    for ps in (7, 32):
        pp, pn = ops.build_part_exact(ps, torch.from_numpy(indptr))
        opp, opn = oracle.build_part(ps, indptr, exact=True)
        assert np.array_equal(pp.numpy(), opp) and np.array_equal(pn.numpy(), opn)
        # structural properties of the table at a size the serial oracle would be slow for
        assert pp[-1] == indptr[-1] and bool((pn[1:] >= pn[:-1]).all())
        assert int((pp[1:] - pp[:-1]).max()) <= ps and int((pp[1:] - pp[:-1]).min()) >= 1


def test_build_part_beyond_2_24_is_exact_and_warns():
    deg = np.array([2 ** 24 + 1, 3], dtype=np.int64)
    indptr = torch.from_numpy(np.concatenate([[0], np.cumsum(deg)]).astype(np.int32))
    with pytest.warns(UserWarning, match="2\\^24"):
        pp, pn = ops.build_part(2 ** 23, indptr)
    assert pp.dtype == torch.int32 and int(pp[3]) == 2 ** 24 + 1 and int(pp[-1]) == 2 ** 24 + 4


def test_build_part_argument_errors():
    with pytest.raises(RuntimeError):
        ops.build_part(32, torch.zeros(4, dtype=torch.int64))       # reference: accessor<int,1> throws
    with pytest.raises(RuntimeError):
        ops.build_part(0, torch.zeros(4, dtype=torch.int32))


def test_check_input_messages_follow_the_reference():
    """GNNAdvisor.cpp:71-73: '<name> must be a CUDA tensor' / '<name> must be contiguous'."""
    x = torch.zeros(4, 8)
    idx = torch.zeros(5, dtype=torch.int32)
    with pytest.raises(RuntimeError, match="input must be a CUDA tensor"):
        ops.SAG(x, idx, idx, torch.zeros(4), idx, idx, 32, 32, 4)
    with pytest.raises(RuntimeError, match="input must be a CUDA tensor"):
        ops.forward(x, torch.zeros(8, 2), idx, idx, torch.zeros(4), idx, idx, 32, 32, 4)
    with pytest.raises(RuntimeError, match="d_output must be a CUDA tensor"):
        ops.backward_gin(x, x, torch.zeros(8, 8), idx, idx, 0.5, idx, idx, 32, 32, 4)


def test_c_abi_rejects_bad_arguments_before_touching_the_device():
    """The reference printf()s "CUDA error" and exit(-1)s after a bad launch (kernel.cu:177-181); the C ABI checks its
    arguments first and returns GNNA_ERR_INVALID with a message.  Validation happens before any CUDA call, so this runs
    without a GPU (no compute is launched here: every call below fails or is empty)."""
    lib = _lib.load()
    INVALID, OK, UNSUPPORTED = -1, 0, -4
    x = np.zeros((4, 8), np.float32)
    ix = np.zeros(8, np.int32)
    p = lambda a: ctypes.c_void_p(a.ctypes.data)   # noqa: E731  (host pointers: never dereferenced by the calls below)
    null = ctypes.c_void_p(0)

    def err():
        return lib.gnna_last_error().decode()

    # empty problems are a no-op, not an error
    assert lib.gnna_sag_f32(null, null, null, null, null, null, 0, 8, 0, 32, 32, 4, null) == OK
    assert lib.gnna_gcn_aggregate_f32(p(x), p(x), p(ix), p(ix), p(x), p(ix), p(ix), 4, 0, 3, 32, 32, 4, null) == OK
    # negative sizes, null pointers, missing degrees, unknown modes
    assert lib.gnna_sag_f32(p(x), p(x), p(ix), p(ix), p(ix), p(ix), -1, 8, 3, 32, 32, 4, null) == INVALID and "negative" in err()
    assert lib.gnna_sag_f32(null, p(x), p(ix), p(ix), p(ix), p(ix), 4, 8, 3, 32, 32, 4, null) == INVALID and "null feature" in err()
    assert lib.gnna_sag_f32(p(x), p(x), p(ix), null, p(ix), p(ix), 4, 8, 3, 32, 32, 4, null) == INVALID and "null index" in err()
    assert lib.gnna_gcn_aggregate_f32(p(x), p(x), p(ix), p(ix), null, p(ix), p(ix), 4, 8, 3, 32, 32, 4, null) == INVALID \
        and "needs degrees" in err()
    assert lib.gnna_aggregate_bf16(7, p(x), p(x), p(ix), p(ix), null, 0.0, p(ix), p(ix), 4, 8, 3, 32, 32, 4, null) == INVALID \
        and "bad mode 7" in err()
    assert lib.gnna_aggregate_f32_ex(1, p(x), 2, p(x), 4, p(ix), p(ix), p(x), 0.0, p(ix), p(ix), 8, 3, 32, 32, 4, null) == INVALID \
        and "fewer source rows" in err()
    assert lib.gnna_aggregate_part_f32_ex(1, 0, p(x), 4, p(x), 4, p(ix), p(ix), p(x), 0.0, p(ix), p(ix), 8, 3, 32, 32, 4, null) == INVALID \
        and "mode 1 not supported" in err()
    assert lib.gnna_aggregate_part_f32_ex(0, 0, p(x), 4, p(x), 4, p(ix), p(ix), p(x), 0.0, p(ix), p(ix), 6, 3, 32, 32, 4, null) == INVALID \
        and "dim % 4" in err()
    # the fused tile: modes, widths and output range it does not have
    fused = lib.gnna_aggregate_gemm_fused_bf16
    assert fused(1, p(x), 0, p(x), 0.0, p(x), null, p(ix), p(ix), p(x), p(ix), p(ix), 4, 64, 16, 3, 32, 32, 4, null) == INVALID \
        and "mode 1 not supported" in err()
    assert fused(0, p(x), 0, p(x), 0.0, p(x), null, p(ix), p(ix), null, p(ix), p(ix), 4, 64, 300, 3, 32, 32, 4, null) == INVALID \
        and "dout 300" in err()
    assert fused(0, p(x), 0, p(x), 0.0, p(x), null, p(ix), p(ix), null, p(ix), p(ix), 4, 48, 16, 3, 32, 32, 4, null) == UNSUPPORTED \
        and "no fused tile" in err()
    # host entry points
    assert lib.gnna_count_parts_host(0, p(ix), 4) < 0
    assert lib.gnna_rabbit_reorder_host(p(ix), p(ix), 2, -1, p(ix)) == INVALID
    bad_edges = np.array([0, 9], np.int32)
    assert lib.gnna_rabbit_reorder_host(p(bad_edges), p(bad_edges), 2, 4, p(ix)) == INVALID and "out of range at edge 1" in err()
    n64 = ctypes.c_int64(0)
    assert lib.gnna_csr_from_edges_host(null, null, 3, 4, p(ix), p(ix), ctypes.byref(n64)) == INVALID and "null pointer" in err()
    assert lib.gnna_edge_text_scan(b"/nonexistent/edges.txt", ctypes.byref(n64)) == INVALID and "cannot open" in err()
    info = _lib.LaunchInfo()
    assert lib.gnna_query_launch(3, 64, 100, 32, 4, ctypes.byref(info)) != OK      # element size 3


def test_launch_geometry_follows_the_three_knobs():
    """dimWorker = lanes per neighbour row (pow2, capped by the row), warpPerBlock = warps per CTA."""
    q = ops.launch_info(64, 1000, 32, 8)
    assert (q["vec_width"], q["lanes_per_row"], q["chunks_per_lane"], q["groups_per_warp"]) == (4, 16, 1, 2)
    assert q["warps_per_block"] == 8 and q["grid_x"] == (1000 + 15) // 16 and q["grid_y"] == 1
    q = ops.launch_info(64, 1000, 4, 2)
    assert (q["lanes_per_row"], q["chunks_per_lane"], q["groups_per_warp"], q["warps_per_block"]) == (4, 4, 8, 2)
    q = ops.launch_info(16, 1000, 32, 4)
    assert (q["vec_width"], q["lanes_per_row"], q["groups_per_warp"]) == (4, 4, 8)
    q = ops.launch_info(41, 10, 32, 4)          # fp32 rows are re-packed to 48 floats (whole sectors): 12 x 128-bit chunks
    assert (q["vec_width"], q["lanes_per_row"], q["chunks_per_lane"], q["grid_y"]) == (4, 16, 1, 1)
    q = ops.launch_info(3703, 10, 32, 2)        # citeseer GIN layer 1: 926 chunks, 128 per d-tile
    assert q["vec_width"] == 4 and q["grid_y"] == (926 + 127) // 128
    q = ops.launch_info(64, 1000, 32, 8, elem_bytes=2)
    assert (q["vec_width"], q["lanes_per_row"]) == (8, 8)


class _FakeDataset:
    def __init__(self, n, e, feat, span):
        self.num_nodes, self.num_features = n, feat
        self.avg_degree, self.avg_edgeSpan = e / n, span
        self.reorder_flag = None
        self.row_pointers = self.column_index = None
        self.reordered = 0

    def rabbit_reorder(self):
        self.reordered += 1


def _load_reference_param():
    import importlib.util
    path = "/root/reference/GNNAdvisor/param.py"
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location("ref_param", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# decisions of the reference Decider (param.py:71-120), generated by importing the reference here:
#   (nodes, edges, in_dim, hidden, sharedMem, avgEdgeSpan) -> (partSize, dw_in, dw_hid, wpb_in, wpb_hid, reorder)
DECIDER_GOLDEN = [
    ((2708, 10556, 1433, 16, 100, 500.0), (3, 32, 16, 6, 8, True)),
    ((3327, 9104, 3703, 16, 100, 3.0), (2, 32, 16, 2, 8, True)),
    ((232965, 114615892, 602, 64, 100, 70000.0), (491, 32, 32, 8, 8, True)),
    ((2449029, 123718280, 100, 64, 64, 2.0), (50, 32, 32, 8, 8, False)),
    ((410236, 4878874, 96, 16, 100, 30.0), (11, 32, 16, 8, 8, False)),
]


@pytest.mark.parametrize("cfg,want", DECIDER_GOLDEN)
def test_decider_matches_reference(cfg, want):
    n, e, feat, hid, smem, span = cfg
    ds = _FakeDataset(n, e, feat, span)
    p = param.InputProperty(None, None, None, 32, 32, 4, smem, hiddenDim=hid, dataset_obj=ds,
                            enable_rabbit=True, manual_mode=False)
    p.decider()
    got = (p.partSize, p.dimWorker_input, p.dimWorker_hidden, p.warpPerBlock_input, p.warpPerBlock_hidden,
           bool(p.reorder_status))
    assert got == want
    assert ds.reordered == 1
    ref = _load_reference_param()
    if ref is not None:      # authoring container: check the golden row against the reference itself
        ds2 = _FakeDataset(n, e, feat, span)
        r = ref.inputProperty(None, None, None, 32, 32, 4, smem, hiddenDim=hid, dataset_obj=ds2,
                              enable_rabbit=True, manual_mode=False)
        r.decider()
        assert want == (r.partSize, r.dimWorker_input, r.dimWorker_hidden, r.warpPerBlock_input,
                        r.warpPerBlock_hidden, bool(r.reorder_status))


def test_param_layer_switch_and_manual_mode():
    ds = _FakeDataset(100, 1000, 50, 5.0)
    p = param.inputProperty(None, None, None, 16, 8, 2, 100, hiddenDim=16, dataset_obj=ds, manual_mode=True)
    p.decider()
    assert (p.partSize, p.dimWorker, p.warpPerBlock) == (16, 8, 2) and ds.reordered == 0
    p.dimWorker_input, p.warpPerBlock_hidden = 4, 7
    assert p.set_input().dimWorker == 4 and p.state_set_input
    assert p.set_hidden().warpPerBlock == 7 and not p.state_set_input
    with pytest.raises(ValueError):
        param.inputProperty(dataset_obj=None)


def test_b200_decider_choices():
    """re-tuned (partSize, dimWorker, warpPerBlock) from the B200 parameter studies."""
    c = param.InputProperty.b200_choice
    assert c(492.0, 64) == (32, 8, 4)       # Reddit hidden layer
    assert c(50.5, 128) == (32, 16, 4)
    assert c(11.9, 16) == (16, 4, 4)        # amazon0505 hidden 16
    assert c(3.9, 1433) == (16, 16, 4)      # Cora input layer
    ds = _FakeDataset(232965, 114615892, 602, 70000.0)
    p = param.InputProperty(None, None, None, 32, 32, 4, 100, hiddenDim=64, dataset_obj=ds, enable_rabbit=True, manual_mode=False)
    p.decider_b200()
    assert (p.partSize, p.dimWorker_input, p.dimWorker_hidden, p.warpPerBlock_input) == (32, 16, 8, 4)
    assert ds.reordered == 1 and p.reorder_status


def test_run_based_kernel_rule_and_switch():
    """gnna_set_runs / gnna_query_runs (host-only): default = the library's rule from the B200 measurements
    (csrc/aggregate_runs.cu auto_runs): dense graphs -> every bf16 width and fp32 rows that are not whole 128-byte
    lines; sparse graphs -> bf16 rows of >= 6 chunks; forced / forbidden settings override it."""
    from gnnadvisor_osdi21_b200 import _lib
    prev = _lib.set_runs(-1)
    try:
        reddit = (232965, 3696299)            # ~16 groups per node
        products = (2449029, 4900000)         # ~2 groups per node
        q = _lib.query_runs
        assert q(4, 64, *reddit) == 0 and q(4, 32, *reddit) == 0          # 256- and 128-byte fp32 rows: default kernel
        assert q(4, 128, *reddit) == 4 and q(4, 48, *reddit) == 4 and q(4, 16, *reddit) == 4
        assert q(2, 64, *reddit) == 4 and q(2, 128, *reddit) == 4 and q(2, 16, *reddit) == 8
        assert q(4, 64, *products) == 0 and q(4, 128, *products) == 0
        assert q(2, 64, *products) == 4 and q(2, 32, *products) == 0
        assert q(4, 41, *reddit) == 0 and q(2, 100, *reddit) == 0         # not whole 16-byte chunks: not eligible
        assert q(4, 256, *reddit) == 0                                    # more than 32 chunks per row
        assert _lib.set_runs(0) == -1 and q(2, 64, *reddit) == 0
        assert _lib.set_runs(8) == 0 and q(4, 64, *products) == 8
    finally:
        _lib.set_runs(prev)
