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

import numpy as np
import torch


# ------------------------------------------------------------------------------------------ CSR
def csr_from_edges(src, dst, num_nodes):
    """(row_ptr int32 [N+1], col_idx int32 [E]) with duplicates merged and columns sorted, on the
    device `src` lives on.  Equivalent to scipy coo_matrix((1,(src,dst))).tocsr() (dataset.py:110-111)."""
    src = torch.as_tensor(src).to(torch.int64).reshape(-1)
    dst = torch.as_tensor(dst).to(torch.int64).reshape(-1)
    num_nodes = int(num_nodes)
    key = torch.unique(src * num_nodes + dst)           # sorted by (src, dst), duplicates merged
    rows = torch.div(key, num_nodes, rounding_mode="floor")
    cols = (key - rows * num_nodes).to(torch.int32)
    counts = torch.bincount(rows, minlength=num_nodes)
    row_ptr = torch.zeros(num_nodes + 1, dtype=torch.int64, device=key.device)
    row_ptr[1:] = torch.cumsum(counts, 0)
    if int(row_ptr[-1]) >= 2 ** 31:
        raise ValueError("graph has %d edges; the int32 CSR contract of the reference stops at 2^31-1" % int(row_ptr[-1]))
    return row_ptr.to(torch.int32), cols


def degrees_from_row_ptr_host(row_ptr):
    """sqrt(max(deg,1)) float32 with torch ops (CPU or CUDA).  dataset.py:11-18,121-122."""
    deg = (row_ptr[1:] - row_ptr[:-1]).to(torch.float32)
    return torch.sqrt(torch.clamp(deg, min=1.0))


# ------------------------------------------------------------------------------------------ synthetic graphs
RMAT_DEFAULT = (0.48, 0.14, 0.14, 0.24)    # (a+b)^18 * 115e6 ~ 2e4: hub degree like Reddit's 21 657


def _rmat_pairs(num_nodes, count, params, gen, device):
    a, b, c, _ = params
    bits = max(1, math.ceil(math.log2(max(num_nodes, 2))))
    src = torch.zeros(count, dtype=torch.int64, device=device)
    dst = torch.zeros(count, dtype=torch.int64, device=device)
    for _ in range(bits):
        r = torch.rand(count, generator=gen, device=device)
        sbit = (r >= a + b).to(torch.int64)
        dbit = (((r >= a) & (r < a + b)) | (r >= a + b + c)).to(torch.int64)
        src = src * 2 + sbit
        dst = dst * 2 + dbit
    keep = (src < num_nodes) & (dst < num_nodes) & (src != dst)
    return src[keep], dst[keep]


def _uniform_pairs(num_nodes, count, gen, device):
    src = torch.randint(0, num_nodes, (count,), generator=gen, device=device)
    dst = torch.randint(0, num_nodes, (count,), generator=gen, device=device)
    keep = src != dst
    return src[keep], dst[keep]


def synth_graph(num_nodes, num_edges, kind="rmat", seed=20211, device="cpu", rmat=RMAT_DEFAULT):
    """Symmetric graph with exactly 2*floor(num_edges/2) directed edges (when that many distinct
    pairs exist), no self loops, last node non-isolated.  Returns (row_ptr, col_idx)
    int32 on `device`.  kind: "rmat" (skewed, Reddit/products/papers look-alikes) or "uniform"
    (Cora/citeseer look-alikes).  Deterministic for a given (seed, device type)."""
    device = torch.device(device)
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    want_pairs = max(1, int(num_edges) // 2)
    N = int(num_nodes)
    und = torch.empty(0, dtype=torch.int64, device=device)      # undirected keys lo*N+hi, lo<hi
    draw = int(want_pairs * 1.1) + 16
    for _ in range(64):
        s, d = _rmat_pairs(N, draw, rmat, gen, device) if kind == "rmat" else _uniform_pairs(N, draw, gen, device)
        lo, hi = torch.minimum(s, d), torch.maximum(s, d)
        und = torch.unique(torch.cat([und, lo * N + hi]))
        if und.numel() >= want_pairs:
            break
        draw = int((want_pairs - und.numel()) * 1.5) + 16
    if und.numel() > want_pairs:
        perm = torch.randperm(und.numel(), generator=gen, device=device)[:want_pairs]
        und = und[perm]
    lo = torch.div(und, N, rounding_mode="floor")
    hi = und - lo * N
    # make sure the last node has a neighbour (its partPtr terminal is then written by the reference too)
    if not bool(((lo == N - 1) | (hi == N - 1)).any()) and N > 1:
        lo = torch.cat([lo, torch.tensor([0], device=device)])
        hi = torch.cat([hi, torch.tensor([N - 1], device=device)])
    src = torch.cat([lo, hi])
    dst = torch.cat([hi, lo])
    return csr_from_edges(src, dst, N)


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

    def __init__(self, dim, num_class, edges=None, path=None, csr=None, device="cuda", verbose=False, seed=20212):
        super().__init__()
        self.num_features, self.num_classes = int(dim), int(num_class)
        self.verbose_flag = verbose
        self.reorder_flag = False
        self.device = torch.device(device)
        if path is not None:
            edges = load_edge_file(path)
        if edges is not None:
            src, dst, n = edges
            src, dst = np.asarray(src), np.asarray(dst)
            self.num_nodes = int(n)
            self.num_edges = len(src)
            self.edge_index = np.stack([src, dst])
            self.avg_degree = self.num_edges / self.num_nodes
            self.avg_edgeSpan = float(np.mean(np.abs(src.astype(np.int64) - dst.astype(np.int64)))) if len(src) else 0.0
            rp, ci = csr_from_edges(torch.from_numpy(src), torch.from_numpy(dst), self.num_nodes)
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

    def _refresh_degrees(self):
        self.degrees = degrees_from_row_ptr_host(self.row_pointers).to(self.device)

    def rabbit_reorder(self):
        """Renumber the vertices for locality and rebuild the CSR (dataset.py:138-175).  Unlike the
        reference the degree vector is refreshed too."""
        if not self.reorder_flag:
            return
        from . import reorder as _reorder
        new_edges = _reorder.reorder(torch.as_tensor(self.edge_index).to(torch.int32))
        self.edge_index = new_edges.numpy()
        rp, ci = csr_from_edges(new_edges[0], new_edges[1], self.num_nodes)
        self.row_pointers, self.column_index = rp, ci
        self._refresh_degrees()


custom_dataset = GraphDataset


def load_edge_file(path):
    """(src, dst, num_nodes) from a reference-format .npz or a text edge list (dataset.py:58-94)."""
    if path.endswith(".npz"):
        obj = np.load(path)
        return obj["src_li"], obj["dst_li"], int(obj["num_nodes"])
    src, dst = [], []
    with open(path) as f:
        for line in f:
            parts = line.split()
            if len(parts) >= 2 and not line.startswith(("#", "%")):
                src.append(int(parts[0]))
                dst.append(int(parts[1]))
    src, dst = np.asarray(src, dtype=np.int64), np.asarray(dst, dtype=np.int64)
    n = int(max(src.max(), dst.max())) + 1 if len(src) else 0
    return src, dst, n


def save_npz(path, src, dst, num_nodes):
    """Write the reference's dataset format (np.load(path)['src_li','dst_li','num_nodes'], dataset.py:87-91)."""
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    np.savez(path, src_li=np.asarray(src), dst_li=np.asarray(dst), num_nodes=np.int64(num_nodes))
