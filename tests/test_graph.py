"""CSR construction and the dataset object against what the reference does (GNNAdvisor/dataset.py:55-122):
scipy coo_matrix(...).tocsr() on the raw edge list (duplicates merged, columns sorted), degrees = sqrt(max(deg, 1)),
the reference's constructor signature and masks; plus the device-independent, shardable graph generator."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from gnnadvisor_osdi21_b200 import graph


def _scipy_csr(src, dst, n):
    csr = sp.coo_matrix((np.ones(len(src)), (src, dst)), shape=(n, n)).tocsr()     # dataset.py:110-111
    csr.sort_indices()
    return csr.indptr.astype(np.int32), csr.indices.astype(np.int32)


@pytest.mark.parametrize("n,e,seed", [(50, 400, 0), (1000, 30000, 1), (7, 0, 2), (300, 5000, 3)])
def test_csr_from_edges_equals_scipy_coo_to_csr(n, e, seed):
    rng = np.random.default_rng(seed)
    src = rng.integers(0, n, e)
    dst = rng.integers(0, max(n // 3, 1), e)                  # a third of the id range: many duplicate edges
    if e:
        src[: e // 4] = src[e // 4: 2 * (e // 4)]             # exact duplicates, self loops stay in
        dst[: e // 4] = dst[e // 4: 2 * (e // 4)]
    erp, eci = _scipy_csr(src, dst, n)
    for native in (True, False):                               # csrc/dataset.cu on the host threads / torch ops
        rp, ci = graph.csr_from_edges(src, dst, n, native=native)
        assert rp.dtype == torch.int32 and ci.dtype == torch.int32
        assert np.array_equal(rp.numpy(), erp) and np.array_equal(ci.numpy(), eci)
    # degrees: sqrt(max(deg, 1)) in float32 (dataset.py:11-18,121-122)
    deg = graph.degrees_from_row_ptr_host(rp).numpy()
    d = np.diff(erp).astype(np.float32)
    assert np.array_equal(deg, np.sqrt(np.maximum(d, 1).astype(np.float32)))


def test_native_csr_builder_on_skewed_and_bucketed_inputs():
    """Enough edges for several buckets and all host threads; a hub row that is most of one bucket; isolated rows at both
    ends; every edge duplicated many times."""
    rng = np.random.default_rng(11)
    n, e = 70001, 400000
    src = rng.integers(1, n - 1, e)
    dst = rng.integers(1, n - 1, e)
    src[:150000] = 4097                                        # hub
    dst[150000:300000] = dst[:150000]
    src[300000:] = src[299999]                                 # one row, random columns
    src2, dst2 = np.concatenate([src, src[::-1]]), np.concatenate([dst, dst[::-1]])
    rp, ci = graph.csr_from_edges(src2, dst2, n, native=True)
    erp, eci = _scipy_csr(src2, dst2, n)
    assert np.array_equal(rp.numpy(), erp) and np.array_equal(ci.numpy(), eci)
    assert int(rp[1]) == 0 and int(rp[-1]) == int(rp[-2])      # rows 0 and n-1 are empty
    one = graph.csr_from_edges([3], [3], 5, native=True)       # a single self loop stays in, like scipy keeps it
    assert one[0].tolist() == [0, 0, 0, 0, 1, 1] and one[1].tolist() == [3]


@pytest.mark.parametrize("native", [True, False])
def test_csr_from_edges_rejects_out_of_range_ids(native):
    with pytest.raises(ValueError, match="outside"):
        graph.csr_from_edges([0, 5], [1, 2], 5, native=native)               # scipy raises as well
    with pytest.raises(ValueError, match="outside"):
        graph.csr_from_edges([0, 1], [-1, 2], 5, native=native)
    with pytest.raises(ValueError):
        sp.coo_matrix((np.ones(2), ([0, 5], [1, 2])), shape=(5, 5))


def test_custom_dataset_has_the_reference_constructor(tmp_path):
    rng = np.random.default_rng(4)
    src, dst = rng.integers(0, 40, 300), rng.integers(0, 40, 300)
    npz = str(tmp_path / "g.npz")
    graph.save_npz(npz, src, dst, 40)
    ds = graph.custom_dataset(npz, 16, 10, load_from_txt=False, verbose=False, device="cpu")    # dataset.py:24
    erp, eci = _scipy_csr(src, dst, 40)
    assert np.array_equal(ds.row_pointers.numpy(), erp) and np.array_equal(ds.column_index.numpy(), eci)
    assert ds.num_nodes == 40 and ds.num_edges == 300 and ds.num_features == 16 and ds.num_classes == 10
    assert ds.x.shape == (40, 16) and ds.y.dtype == torch.long and bool((ds.y == 1).all())
    assert abs(ds.avg_degree - 300 / 40) < 1e-12 and abs(ds.avg_edgeSpan - np.mean(np.abs(src - dst))) < 1e-9
    assert int(ds.train_mask.sum()) == 40 and int(ds.val_mask.sum()) == 12 and int(ds.test_mask.sum()) == 4   # :44-53
    txt = str(tmp_path / "g.txt")
    with open(txt, "w") as f:
        for a, b in zip(src, dst):
            f.write("%d %d\n" % (a, b))
    dt = graph.custom_dataset(txt, 16, 10, load_from_txt=True, device="cpu")
    assert dt.num_nodes == int(max(src.max(), dst.max())) + 1
    assert np.array_equal(dt.column_index.numpy(), _scipy_csr(src, dst, dt.num_nodes)[1])
    with pytest.raises(ValueError, match=".npz"):
        graph.custom_dataset(txt, 16, 10, load_from_txt=False, device="cpu")


def test_verbose_dataset_prints_the_reference_phase_lines(tmp_path, capsys):
    """dataset.py's verbose phase timers (:77-79, :91-93, :99-102, :113, :152-175), which the reference's log scrapers read."""
    rng = np.random.default_rng(6)
    src, dst = rng.integers(0, 60, 500), rng.integers(0, 60, 500)
    txt = str(tmp_path / "g.txt")
    with open(txt, "w") as f:
        f.write("".join("%d %d\n" % (a, b) for a, b in zip(src, dst)))
    ds = graph.custom_dataset(txt, 8, 3, load_from_txt=True, verbose=True, device="cpu")
    ds.rabbit_reorder()                                       # flag not set
    ds.reorder_flag = True
    before = ds.column_index.clone()
    ds.rabbit_reorder()
    out = capsys.readouterr().out
    for needle in ("# Loading (txt) ", "# nodes: %d" % ds.num_nodes, "# avg_degree: ", "# avg_edgeSpan: ", "# Build CSR after reordering (s): ",
                   "Reorder flag is not set. Skipped...", "Reorder flag is set. Continue...", "# Reorder time (s): ", "# Re-Build CSR (s): "):
        assert needle in out, needle
    assert ds.column_index.numel() == before.numel() and int(ds.row_pointers[-1]) == before.numel()
    npz = str(tmp_path / "g.npz")
    graph.save_npz(npz, src, dst, 60)
    graph.custom_dataset(npz, 8, 3, load_from_txt=False, verbose=True, device="cpu")
    assert "# Loading (npz)(s): " in capsys.readouterr().out


def _reference_text_loop(path):
    """dataset.py:62-72 as written: the per-line loop of the reference."""
    src_li, dst_li, nodes = [], [], set()
    with open(path) as fp:
        for line in fp:
            src, dst = line.strip('\n').split()
            src, dst = int(src), int(dst)
            src_li.append(src)
            dst_li.append(dst)
            nodes.add(src)
            nodes.add(dst)
    return np.asarray(src_li, dtype=np.int64), np.asarray(dst_li, dtype=np.int64), max(nodes) + 1


def test_text_loader_equals_the_reference_loop(tmp_path):
    rng = np.random.default_rng(5)
    e = 300000                                                 # several MB: many pieces, all host threads
    src, dst = rng.integers(0, 1 << 20, e), rng.integers(0, 977, e)
    plain = str(tmp_path / "plain.txt")
    with open(plain, "w") as f:
        f.write("".join("%d %d\n" % (a, b) for a, b in zip(src, dst)))
    s, d, n = graph.load_edge_text(plain)
    rs, rd, rn = _reference_text_loop(plain)
    assert s.dtype == np.int64 and np.array_equal(s, rs) and np.array_equal(d, rd) and n == rn
    # what real files contain and the reference's loop also takes: tabs, several blanks, CRLF, no newline at the end
    messy = str(tmp_path / "messy.txt")
    with open(messy, "w", newline="") as f:
        f.write("0\t5\n 7   2 \n3 4\r\n+6 1\n9 9")
    s, d, n = graph.load_edge_text(messy)
    rs, rd, rn = _reference_text_loop(messy)
    assert s.tolist() == rs.tolist() == [0, 7, 3, 6, 9] and d.tolist() == rd.tolist() == [5, 2, 4, 1, 9] and n == rn == 10
    # beyond the reference: comment lines, blank lines and a weight column are skipped / ignored
    snap = str(tmp_path / "snap.txt")
    with open(snap, "w") as f:
        f.write("# Directed graph\n% mtx style\n\n1 2 0.5\n  \n2 0\n")
    s, d, n = graph.load_edge_text(snap)
    assert s.tolist() == [1, 2] and d.tolist() == [2, 0] and n == 3
    empty = str(tmp_path / "empty.txt")
    open(empty, "w").close()
    s, d, n = graph.load_edge_text(empty)
    assert len(s) == 0 and len(d) == 0 and n == 0


@pytest.mark.parametrize("text,line", [("1 2\n3\n", 2), ("1 2\n3 4\nfoo bar\n", 3), ("1.5 2\n", 1), ("1 2x\n", 1),
                                       ("1 99999999999999999999\n", 1)])
def test_text_loader_names_the_malformed_line(tmp_path, text, line):
    path = str(tmp_path / "bad.txt")
    with open(path, "w") as f:
        f.write(text)
    with pytest.raises(ValueError, match="line %d " % line):    # the reference's unpack / int() raise ValueError here too
        graph.load_edge_text(path)
    with pytest.raises((ValueError, OverflowError)):            # OverflowError: the 20-digit id, when numpy sees it
        _reference_text_loop(path)


def test_text_loader_missing_file(tmp_path):
    with pytest.raises(OSError, match="cannot open"):
        graph.load_edge_text(str(tmp_path / "nope.txt"))


def test_splitmix64_matches_the_published_algorithm():
    M = (1 << 64) - 1

    def ref(x):
        x &= M
        x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M
        x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M
        return x ^ (x >> 31)
    vals = [0, 1, 12345, (1 << 63) - 1, -1, -(1 << 63), 0x9E3779B97F4A7C15 - (1 << 64)]
    got = graph._mix64(torch.tensor(vals, dtype=torch.int64))
    assert [int(g) & M for g in got] == [ref(v) for v in vals]


@pytest.mark.parametrize("kind", ["rmat", "uniform"])
def test_shards_of_the_pair_stream_tile_the_whole_graph(kind):
    n, e, seed = 5000, 160000, 9
    rp, ci = graph.synth_graph(n, e, kind=kind, seed=seed, exact=False)
    # symmetric, no self loops, sorted unique columns
    A = sp.csr_matrix((np.ones(ci.numel()), ci.numpy(), rp.numpy()), shape=(n, n))
    assert (A != A.T).nnz == 0 and A.diagonal().sum() == 0
    est = graph.stream_degree_estimate(n, e // 2, kind=kind, seed=seed, chunk=30011)
    assert int(est.sum()) >= int(rp[-1]) and bool((est >= (rp[1:] - rp[:-1]).long()).all())   # pre-dedup counts bound the rows
    cuts = [0, 1, 1200, 3100, 4999, 5000]
    for v0, v1 in zip(cuts[:-1], cuts[1:]):
        r, c = graph.synth_graph_shard(n, e, v0, v1, kind=kind, seed=seed, chunk=30011)      # chunking must not matter
        assert torch.equal(r, rp[v0:v1 + 1].long() - int(rp[v0]))
        assert torch.equal(c, ci[int(rp[v0]):int(rp[v1])])


def test_exact_lookalike_sizes_and_last_node():
    rp, ci = graph.synth_graph(2708, 10556, kind="uniform", seed=20211)
    assert int(rp[-1]) == 10556 and int(rp[-1] - rp[-2]) > 0
    rp2, ci2 = graph.synth_graph(2708, 10556, kind="uniform", seed=20211)
    assert torch.equal(rp, rp2) and torch.equal(ci, ci2)


REF_PY = "/root/reference/GNNAdvisor"


@pytest.mark.skipif(not __import__("os").path.isdir(REF_PY), reason="the reference tree is only mounted in the authoring container")
def test_dataset_equals_the_reference_loader_run_live(tmp_path, monkeypatch):
    """The reference's dataset.py, imported UNCHANGED (compat/ provides dgl and rabbit; `.cuda()` made a no-op because this
    container has no GPU), builds its custom_dataset from the same .txt and .npz files: every field the training script
    reads must be equal -- sizes, statistics, CSR, degrees, masks -- before and after rabbit_reorder()."""
    import importlib
    import os
    import sys
    compat = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gnnadvisor_osdi21_b200", "compat")
    rng = np.random.default_rng(21)
    n, e = 400, 6000
    c = rng.integers(0, n // 20, e)                              # 20 communities of 20, most edges inside: reordering has work to do
    src = c * 20 + rng.integers(0, 20, e)
    dst = np.where(rng.random(e) < 0.8, c * 20 + rng.integers(0, 20, e), rng.integers(0, n, e))
    shuffle = rng.permutation(n)
    src, dst = shuffle[src], shuffle[dst]
    src[0], dst[0] = n - 1, 0                                    # the largest id occurs: num_nodes = n in both formats
    txt, npz = str(tmp_path / "g.txt"), str(tmp_path / "g.npz")
    with open(txt, "w") as f:
        f.write("".join("%d %d\n" % (a, b) for a, b in zip(src, dst)))
    graph.save_npz(npz, src, dst, n)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    names = ("dataset", "dgl", "rabbit", "GNNAdvisor")
    saved = {k: sys.modules.pop(k) for k in names if k in sys.modules}
    sys.path[:0] = [compat, REF_PY]
    try:
        ref_mod = importlib.import_module("dataset")
        assert ref_mod.__file__.startswith(REF_PY)
        for path, from_txt in ((txt, True), (npz, False)):
            ref = ref_mod.custom_dataset(path, 16, 5, load_from_txt=from_txt, verbose=False)
            ours = graph.custom_dataset(path, 16, 5, load_from_txt=from_txt, verbose=False, device="cpu")

            def same():
                assert int(ref.num_nodes) == ours.num_nodes and ref.num_edges == ours.num_edges
                assert ref.num_features == ours.num_features and ref.num_classes == ours.num_classes
                assert abs(ref.avg_degree - ours.avg_degree) < 1e-12 and abs(float(ref.avg_edgeSpan) - ours.avg_edgeSpan) < 1e-9
                assert np.array_equal(np.asarray(ref.edge_index), np.asarray(ours.edge_index))
                assert torch.equal(ref.row_pointers, ours.row_pointers) and torch.equal(ref.column_index, ours.column_index)
                assert ref.row_pointers.dtype == ours.row_pointers.dtype == torch.int32
                assert torch.equal(ref.degrees, ours.degrees)
            same()
            assert len(ref.val) == len(ours.val) == e and set(ref.val) == {1} and bool((ours.val == 1).all())
            assert ref.x.shape == ours.x.shape and torch.equal(ref.y, ours.y)
            for m in ("train_mask", "val_mask", "test_mask"):
                assert torch.equal(getattr(ref, m), getattr(ours, m)), m
            ref.rabbit_reorder(); ours.rabbit_reorder()          # flag not set: nothing happens
            same()
            ref.reorder_flag = ours.reorder_flag = True
            ref.rabbit_reorder(); ours.rabbit_reorder()          # same permutation on both sides (deterministic reordering)
            same()
            assert float(ours.avg_edgeSpan) > 0 and not np.array_equal(np.asarray(ours.edge_index), np.stack([src, dst]))
    finally:
        del sys.path[:2]
        for k in names:
            sys.modules.pop(k, None)
        sys.modules.update(saved)
