"""Graph inputs of the hot path: CSR construction, degree vector, synthetic look-alike graphs and the
reference's on-disk formats.  Host-side mirror of GNNAdvisor/dataset.py (custom_dataset).

Everything is written with torch ops so the same code builds a 10-edge test graph on the CPU and a
115 M-edge Reddit look-alike on the GPU (there are no dataset files and no network in this
environment).  CSR semantics follow dataset.py:108-122: scipy's coo->csr sums
duplicate edges, i.e. the column indices of a row are UNIQUE and SORTED; degrees are
sqrt(max(deg, 1)) in float32.
"""
import math
import os
import time

import numpy as np
import torch


# ------------------------------------------------------------------------------------------ CSR
def csr_from_edges(src, dst, num_nodes, native=None):
    """(row_ptr int32 [N+1], col_idx int32 [E]) with duplicates merged and columns sorted, on the
    device `src` lives on.  Equivalent to scipy coo_matrix((1,(src,dst))).tocsr() (dataset.py:110-111).
    native: host edge lists go through the multi-threaded builder of libgnna_b200.so (None: yes unless
    GNNA_CSR_NATIVE=0); False keeps the torch ops -- what the synthetic generator uses, so bench.py's CPU reference arm
    never loads the product's library."""
    src = torch.as_tensor(src).to(torch.int64).reshape(-1)
    dst = torch.as_tensor(dst).to(torch.int64).reshape(-1)
    num_nodes = int(num_nodes)
    if src.numel() != dst.numel():
        raise ValueError("src and dst differ in length (%d vs %d)" % (src.numel(), dst.numel()))
    if native is None:
        native = os.environ.get("GNNA_CSR_NATIVE", "1") == "1"
    if native and src.device.type == "cpu":
        return _csr_from_edges_native(src, dst, num_nodes)
    if src.numel() and (int(torch.minimum(src.min(), dst.min())) < 0 or int(torch.maximum(src.max(), dst.max())) >= num_nodes):
        # scipy's coo_matrix raises here too ("row index exceeds matrix dimensions"); without the check an id
        # >= num_nodes would alias into another row of the src*N+dst key
        raise ValueError("edge endpoint outside [0, %d)" % num_nodes)
    key = torch.unique(src * num_nodes + dst)           # sorted by (src, dst), duplicates merged
    rows = torch.div(key, num_nodes, rounding_mode="floor")
    cols = (key - rows * num_nodes).to(torch.int32)
    counts = torch.bincount(rows, minlength=num_nodes)
    row_ptr = torch.zeros(num_nodes + 1, dtype=torch.int64, device=key.device)
    row_ptr[1:] = torch.cumsum(counts, 0)
    if int(row_ptr[-1]) >= 2 ** 31:
        raise ValueError("graph has %d edges; the int32 CSR contract of the reference stops at 2^31-1" % int(row_ptr[-1]))
    return row_ptr.to(torch.int32), cols


def _csr_from_edges_native(src, dst, num_nodes):
    """Host edge lists: the bucketed multi-threaded builder of libgnna_b200.so (csrc/dataset.cu) -- the same CSR bit for bit
    (tests/test_graph.py compares both paths with scipy); GNNA_CSR_NATIVE=0 keeps the torch ops."""
    import ctypes
    from . import _lib
    src, dst = src.contiguous(), dst.contiguous()
    row_ptr = torch.empty(num_nodes + 1, dtype=torch.int32)
    col = torch.empty(src.numel(), dtype=torch.int32)
    nnz = ctypes.c_int64(0)
    rc = _lib.load().gnna_csr_from_edges_host(ctypes.c_void_p(src.data_ptr()), ctypes.c_void_p(dst.data_ptr()), src.numel(),
                                              num_nodes, ctypes.c_void_p(row_ptr.data_ptr()), ctypes.c_void_p(col.data_ptr()),
                                              ctypes.byref(nnz))
    if rc != 0:
        raise ValueError(_lib.load().gnna_last_error().decode())      # out-of-range endpoint / >= 2^31 edges, like the torch path
    return row_ptr, col[:nnz.value].clone() if nnz.value < col.numel() else col


def degrees_from_row_ptr_host(row_ptr):
    """sqrt(max(deg,1)) float32 with torch ops (CPU or CUDA).  dataset.py:11-18,121-122."""
    deg = (row_ptr[1:] - row_ptr[:-1]).to(torch.float32)
    return torch.sqrt(torch.clamp(deg, min=1.0))


# ------------------------------------------------------------------------------------------ synthetic graphs
RMAT_DEFAULT = (0.48, 0.14, 0.14, 0.24)    # (a+b)^18 * 115e6 ~ 2e4: hub degree like Reddit's 21 657
# (SURVEY.md 8d suggests (0.57, 0.19, 0.19, 0.05); that puts 10x Reddit's hub degree on node 0, so the look-alikes use
# the flatter set above.  bench.py prints the parameters in config.)

# The pair stream is COUNTER-BASED: pair k of a graph is a pure function of (seed, k) -- splitmix64 over int64 tensors,
# wrapping arithmetic -- so the same graph comes out on the CPU and on any GPU, any slice [k0, k1) of the stream can be
# generated on its own, and N ranks can each generate exactly the rows they own without anyone holding the whole graph.
_M64 = (1 << 64) - 1


def _s64(c):
    c &= _M64
    return c - (1 << 64) if c >= (1 << 63) else c


_C1, _C2, _GOLD = _s64(0xBF58476D1CE4E5B9), _s64(0x94D049BB133111EB), 0x9E3779B97F4A7C15


def _mix64(x):
    """splitmix64 finaliser on an int64 tensor (logical shifts emulated by masking the sign extension)."""
    x = x ^ ((x >> 30) & 0x3FFFFFFFF)
    x = x * _C1
    x = x ^ ((x >> 27) & 0x1FFFFFFFFF)
    x = x * _C2
    return x ^ ((x >> 31) & 0x1FFFFFFFF)


def stream_pairs(num_nodes, start, count, kind="rmat", seed=20211, device="cpu", rmat=RMAT_DEFAULT):
    """Pairs [start, start+count) of the directed pair stream of graph (num_nodes, kind, seed): (src, dst) int64
    tensors with self loops and out-of-range ids already dropped (so fewer than `count` come back)."""
    N = int(num_nodes)
    if torch.device(device).type == "cuda" and os.environ.get("GNNA_GRAPHGEN_KERNEL", "1") == "1":
        return _stream_pairs_cuda(N, int(start), int(count), kind, int(seed), torch.device(device), rmat)
    idx = torch.arange(int(start), int(start) + int(count), dtype=torch.int64, device=device)
    base = _mix64(idx ^ _s64(int(seed) * _GOLD + 0x1234567))
    if kind == "uniform":
        h = _mix64(base + _s64(_GOLD))
        src = (((h >> 32) & 0xFFFFFFFF) * N) >> 32
        dst = ((h & 0xFFFFFFFF) * N) >> 32
    else:
        a, b, c, _ = rmat
        t_a, t_ab, t_abc = int(a * 2 ** 32), int((a + b) * 2 ** 32), int((a + b + c) * 2 ** 32)
        bits = max(1, math.ceil(math.log2(max(N, 2))))
        src = torch.zeros_like(idx)
        dst = torch.zeros_like(idx)
        for lvl in range(bits):
            if lvl % 2 == 0:
                h = _mix64(base + _s64((lvl // 2 + 1) * _GOLD))
                u = (h >> 32) & 0xFFFFFFFF
            else:
                u = h & 0xFFFFFFFF
            src = src * 2 + (u >= t_ab).to(torch.int64)
            dst = dst * 2 + (((u >= t_a) & (u < t_ab)) | (u >= t_abc)).to(torch.int64)
    keep = (src < N) & (dst < N) & (src != dst)
    return src[keep], dst[keep]


def stream_features(ids, dim, seed=20212):
    """Feature rows of the given GLOBAL node ids: x[i, d] = a pure function of (seed, i, d), uniform in [-sqrt(3), sqrt(3))
    (unit variance like dataset.py:129's randn) and exactly the same on every device, so a rank can produce the rows it
    owns -- and a checker any row it likes -- without a feature file or an exchange.  Returns float32 [len(ids), dim]."""
    ids = torch.as_tensor(ids).to(torch.int64).reshape(-1, 1)
    cols = torch.arange(int(dim), dtype=torch.int64, device=ids.device).reshape(1, -1)
    h = _mix64(_mix64(ids ^ _s64(int(seed) * _GOLD + 0x7654321)) + cols * _s64(_GOLD))
    u = ((h >> 40) & 0xFFFFFF).to(torch.float32)                 # 24 bits: exact in float32
    return (u * (2.0 ** -23) - 1.0) * 1.7320508


def _stream_pairs_cuda(N, start, count, kind, seed, device, rmat):
    """stream_pairs as ONE kernel of libgnna_b200.so (csrc/graphgen.cu): the same pairs, bit for bit."""
    import ctypes
    from . import _lib
    a, b, c, _ = rmat
    bits = max(1, math.ceil(math.log2(max(N, 2))))
    src = torch.empty(count, dtype=torch.int64, device=device)
    dst = torch.empty(count, dtype=torch.int64, device=device)
    with torch.cuda.device(device):
        _lib.check(_lib.load().gnna_stream_pairs(
            start, count, (int(seed) * _GOLD + 0x1234567) & _M64, N, 1 if kind == "uniform" else 0, bits,
            int(a * 2 ** 32), int((a + b) * 2 ** 32), int((a + b + c) * 2 ** 32),
            ctypes.c_void_p(src.data_ptr()), ctypes.c_void_p(dst.data_ptr()),
            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "stream_pairs")
    keep = src >= 0
    return src[keep], dst[keep]


def synth_graph(num_nodes, num_edges, kind="rmat", seed=20211, device="cpu", rmat=RMAT_DEFAULT, exact=True):
    """Symmetric graph with exactly 2*floor(num_edges/2) directed edges (when that many distinct
    pairs exist), no self loops, last node non-isolated.  Returns (row_ptr, col_idx)
    int32 on `device`.  kind: "rmat" (skewed, Reddit/products/papers look-alikes) or "uniform"
    (Cora/citeseer look-alikes).  Deterministic for a given seed on every device.
    exact=False: the first floor(num_edges/2) pairs of the stream, duplicates merged ("as generated", what
    synth_graph_shard produces piecewise)."""
    device = torch.device(device)
    want_pairs = max(1, int(num_edges) // 2)
    N = int(num_nodes)
    und = torch.empty(0, dtype=torch.int64, device=device)      # undirected keys lo*N+hi, lo<hi
    pos, draw = 0, (int(want_pairs * 1.1) + 16 if exact else want_pairs)
    for _ in range(64):
        s, d = stream_pairs(N, pos, draw, kind, seed, device, rmat)
        pos += draw
        lo, hi = torch.minimum(s, d), torch.maximum(s, d)
        und = torch.unique(torch.cat([und, lo * N + hi]))
        if not exact or und.numel() >= want_pairs:
            break
        draw = int((want_pairs - und.numel()) * 1.5) + 16
    if und.numel() > want_pairs:       # keep the pairs with the smallest hashes: a fixed, device-independent subset
        order = torch.sort(_mix64(und + _s64(int(seed) * _GOLD)))[1][:want_pairs]
        und = und[order]
    lo = torch.div(und, N, rounding_mode="floor")
    hi = und - lo * N
    # make sure the last node has a neighbour (its partPtr terminal is then written by the reference too)
    if exact and not bool(((lo == N - 1) | (hi == N - 1)).any()) and N > 1:
        lo = torch.cat([lo, torch.tensor([0], device=device)])
        hi = torch.cat([hi, torch.tensor([N - 1], device=device)])
    src = torch.cat([lo, hi])
    dst = torch.cat([hi, lo])
    return csr_from_edges(src, dst, N, native=False)


def stream_degree_estimate(num_nodes, num_pairs, kind="rmat", seed=20211, device="cpu", rmat=RMAT_DEFAULT,
                           chunk=1 << 25, every=1):
    """Row degrees of the symmetrised pair stream BEFORE duplicate merging (int64 [N]); what the sharded generator cuts
    its vertex ranges from.  every=k looks at one chunk in k (a k-fold cheaper estimate, scaled back up)."""
    N = int(num_nodes)
    deg = torch.zeros(N, dtype=torch.int64, device=device)
    one = None
    for ci, pos in enumerate(range(0, int(num_pairs), chunk)):
        if ci % every:
            continue
        s, d = stream_pairs(N, pos, min(chunk, int(num_pairs) - pos), kind, seed, device, rmat)
        if one is None or one.numel() < s.numel():
            one = torch.ones(s.numel(), dtype=torch.int64, device=device)
        deg.index_add_(0, s, one[:s.numel()])
        deg.index_add_(0, d, one[:s.numel()])
    return deg * every


def synth_graph_shard(num_nodes, num_edges, v0, v1, kind="rmat", seed=20211, device="cpu", rmat=RMAT_DEFAULT,
                      chunk=1 << 25):
    """Rows [v0, v1) of synth_graph(num_nodes, num_edges, exact=False) without ever holding the rest:
    every rank walks the same counter-based pair stream and keeps the directed edges whose SOURCE it owns.
    Returns (row_ptr int64 [v1-v0+1] local offsets, col int32 [E_local] GLOBAL ids), duplicates merged, sorted."""
    N, v0, v1 = int(num_nodes), int(v0), int(v1)
    want_pairs = max(1, int(num_edges) // 2)
    keys = []
    for pos in range(0, want_pairs, chunk):
        s, d = stream_pairs(N, pos, min(chunk, want_pairs - pos), kind, seed, device, rmat)
        for a, b in ((s, d), (d, s)):
            m = (a >= v0) & (a < v1)
            keys.append((a[m] - v0) * N + b[m])
    key = torch.unique(torch.cat(keys)) if keys else torch.empty(0, dtype=torch.int64, device=device)
    del keys
    rows = torch.div(key, N, rounding_mode="floor")
    cols = (key - rows * N).to(torch.int32)
    row_ptr = torch.zeros(v1 - v0 + 1, dtype=torch.int64, device=key.device)
    row_ptr[1:] = torch.cumsum(torch.bincount(rows, minlength=v1 - v0), 0)
    return row_ptr, cols


# sizes of the BASELINE.json configurations (public dataset statistics)
LOOKALIKES = {
    #  name            nodes        edges        in    hidden classes kind
    "cora":           (2708,        10556,       1433, 16,  7,  "uniform"),
    "citeseer":       (3327,        9104,        3703, 16,  6,  "uniform"),
    "reddit":         (232965,      114615892,   602,  64,  41, "rmat"),
    "ogbn-products":  (2449029,     123718280,   100,  64,  47, "rmat"),
    "amazon0505":     (410236,      4878874,     96,   16,  22, "rmat"),
    # config #5; use scale < 1 on a single GPU (the full graph is 1.6 G edges: sharded)
    "ogbn-papers100M": (111059956,  1615685872,  128,  128, 172, "rmat"),
}


def lookalike(name, device="cpu", seed=20211, scale=1.0):
    """Synthetic graph with the node/edge counts of a named dataset (optionally scaled down)."""
    n, e, din, hid, cls, kind = LOOKALIKES[name]
    n, e = max(2, int(n * scale)), max(2, int(e * scale))
    row_ptr, col_idx = synth_graph(n, e, kind=kind, seed=seed, device=device)
    return {"name": name, "num_nodes": n, "row_ptr": row_ptr, "col_idx": col_idx,
            "in_dim": din, "hidden": hid, "classes": cls}


# ------------------------------------------------------------------------------------------ dataset object
class GraphDataset(torch.nn.Module):
    """What the reference's custom_dataset exposes (dataset.py:24-53): num_nodes, num_edges,
    num_features, num_classes, avg_degree, avg_edgeSpan, edge_index, row_pointers, column_index
    (CPU int32, as GNNA_main.py:68-70 expects), degrees / x / y on the compute device.

    Sources: an edge list (src, dst, num_nodes), a reference-format .npz (src_li, dst_li, num_nodes;
    dataset.py:87-91), a whitespace `src dst` text file (:64-70), or a ready CSR."""

    def __init__(self, dim, num_class, edges=None, path=None, csr=None, device="cuda", verbose=False, seed=20212, text=None):
        super().__init__()
        self.num_features, self.num_classes = int(dim), int(num_class)
        self.verbose_flag = verbose
        self.reorder_flag = False
        self.device = torch.device(device)
        if path is not None:
            start = time.perf_counter()
            edges = load_edge_file(path, text=text)
            if verbose:                                        # the reference's phase timers (dataset.py:74-79, 91-93)
                is_npz = text is False or (text is None and str(path).endswith(".npz"))
                print("# Loading (npz)(s): {:.3f}".format(time.perf_counter() - start) if is_npz
                      else "# Loading (txt) {:.3f}s ".format(time.perf_counter() - start))
        if edges is not None:
            src, dst, n = edges
            src, dst = np.asarray(src), np.asarray(dst)
            self.num_nodes = int(n)
            self.num_edges = len(src)
            self.edge_index = np.stack([src, dst])
            self.avg_degree = self.num_edges / self.num_nodes
            self.avg_edgeSpan = float(np.mean(np.abs(src.astype(np.int64) - dst.astype(np.int64)))) if len(src) else 0.0
            if verbose:                                        # dataset.py:99-102
                print('# nodes: {}'.format(self.num_nodes))
                print("# avg_degree: {:.2f}".format(self.avg_degree))
                print("# avg_edgeSpan: {}".format(int(self.avg_edgeSpan)))
            start = time.perf_counter()
            rp, ci = csr_from_edges(torch.from_numpy(src), torch.from_numpy(dst), self.num_nodes)
            if verbose:
                print("# Build CSR after reordering (s): {:.3f}".format(time.perf_counter() - start))   # :113 (its wording)
        elif csr is not None:
            rp, ci = csr[0].cpu(), csr[1].cpu()
            self.num_nodes = rp.numel() - 1
            self.num_edges = ci.numel()
            rows = torch.repeat_interleave(torch.arange(self.num_nodes), (rp[1:] - rp[:-1]).long())
            self.edge_index = np.stack([rows.numpy(), ci.numpy().astype(np.int64)])
            self.avg_degree = self.num_edges / max(self.num_nodes, 1)
            self.avg_edgeSpan = float((rows - ci.long()).abs().float().mean()) if self.num_edges else 0.0
        else:
            raise ValueError("GraphDataset needs edges=, path= or csr=")
        self.row_pointers, self.column_index = rp.cpu(), ci.cpu()
        self._refresh_degrees()
        gen = torch.Generator().manual_seed(seed)
        self.x = torch.randn(self.num_nodes, self.num_features, generator=gen).to(self.device)      # dataset.py:129
        self.y = torch.ones(self.num_nodes, dtype=torch.long, device=self.device)                   # dataset.py:136

    @property
    def val(self):
        """The edge values the reference keeps as a Python list `[1] * num_edges` (dataset.py:106) and hands to
        unitest.Verification.reference (GNNA_main.py:122); a float32 array of ones serves the same callers."""
        return np.ones(self.num_edges, dtype=np.float32)

    def _refresh_degrees(self):
        self.degrees = degrees_from_row_ptr_host(self.row_pointers).to(self.device)

    def rabbit_reorder(self):
        """Renumber the vertices for locality and rebuild the CSR and the degree vector (dataset.py:138-175)."""
        if not self.reorder_flag:
            if self.verbose_flag:
                print("Reorder flag is not set. Skipped...")
            return
        from . import reorder as _reorder
        if self.verbose_flag:
            print("Reorder flag is set. Continue...")
        start = time.perf_counter()
        new_edges = _reorder.reorder(torch.as_tensor(self.edge_index).to(torch.int32))
        if self.verbose_flag:
            print("# Reorder time (s): {}".format(time.perf_counter() - start))
        self.edge_index = new_edges.numpy()
        start = time.perf_counter()
        rp, ci = csr_from_edges(new_edges[0], new_edges[1], self.num_nodes)
        self.row_pointers, self.column_index = rp, ci
        self._refresh_degrees()
        if self.verbose_flag:
            print("# Re-Build CSR (s): {:.3f}".format(time.perf_counter() - start))


class custom_dataset(GraphDataset):
    """The reference's constructor (dataset.py:24: custom_dataset(path, dim, num_class, load_from_txt=True, verbose=False))
    and its train/val/test masks (:44-53: the first 100 % / 30 % / 10 % of the nodes)."""

    def __init__(self, path, dim, num_class, load_from_txt=True, verbose=False, device="cuda"):
        if not load_from_txt and not str(path).endswith(".npz"):
            raise ValueError("graph file must be a .npz file")          # dataset.py:83-84
        super().__init__(dim, num_class, path=path, text=bool(load_from_txt), device=device, verbose=verbose)
        self.load_from_txt = load_from_txt
        n = self.num_nodes
        idx = torch.arange(n, device=self.device)
        self.train_mask = idx < int(n * 1)
        self.val_mask = idx < int(n * 0.3)
        self.test_mask = idx < int(n * 0.1)



def load_edge_file(path, text=None):
    """(src, dst, num_nodes) from a reference-format .npz or a text edge list (dataset.py:58-94).
    text: None = by extension, True/False = the reference's load_from_txt switch."""
    if (text is None and path.endswith(".npz")) or text is False:
        obj = np.load(path)
        return obj["src_li"], obj["dst_li"], int(obj["num_nodes"])
    return load_edge_text(path)


def load_edge_text(path):
    """Whitespace `src dst` text file -> (src int64, dst int64, num_nodes = largest id + 1), edges in file order
    (dataset.py:62-72), parsed by all host threads in libgnna_b200.so (csrc/dataset.cu) instead of a Python loop over the
    lines.  Blank lines and lines starting with '#' or '%' are skipped, tokens after the second ignored; any other line
    that is not two integers raises ValueError with its line number (the reference's int()/unpack raises there too)."""
    import ctypes
    from . import _lib
    lib = _lib.load()
    cpath = os.fsencode(path)
    cap = ctypes.c_int64(0)
    if lib.gnna_edge_text_scan(cpath, ctypes.byref(cap)) != 0:
        raise OSError(lib.gnna_last_error().decode())
    src = np.empty(cap.value, dtype=np.int64)
    dst = np.empty(cap.value, dtype=np.int64)
    e, n = ctypes.c_int64(0), ctypes.c_int64(0)
    if lib.gnna_edge_text_parse(cpath, ctypes.c_void_p(src.ctypes.data), ctypes.c_void_p(dst.ctypes.data), cap.value,
                                ctypes.byref(e), ctypes.byref(n)) != 0:
        raise ValueError(lib.gnna_last_error().decode())
    if e.value < cap.value:                                   # comments / blank lines: give the slack back
        src, dst = src[:e.value].copy(), dst[:e.value].copy()
    return src, dst, int(n.value)


def save_npz(path, src, dst, num_nodes):
    """Write the reference's dataset format (np.load(path)['src_li','dst_li','num_nodes'], dataset.py:87-91)."""
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    np.savez(path, src_li=np.asarray(src), dst_li=np.asarray(dst), num_nodes=np.int64(num_nodes))
