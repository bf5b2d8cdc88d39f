"""Multi-GPU sharding of the aggregation path: 1-D vertex-range partition + halo exchange.

The reference is single-GPU (no NCCL/MPI call site anywhere); this is new functionality
asked for by BASELINE.json's north_star ("graphs ... 1-D vertex-range partitioned across the 8xB200
box with a single NCCL all-to-all of halo neighbor features per layer over NVLink").

Layout (one process per GPU, torch.distributed; NCCL on GPUs, gloo in the CPU tests):
  * rank r owns the contiguous vertex range [v_r, v_{r+1}), cut so every rank holds ~E/world edges
    (balanced by edge count, not node count: the graphs are skewed);
  * it keeps the CSR rows of its range with column ids remapped to a LOCAL index space
    [0, n_local) = own rows, [n_local, n_local + n_halo) = halo = sorted unique remote neighbours
    (sorted by global id => automatically grouped by owner, ranges being contiguous);
  * features live in ONE buffer X_ext [n_local + n_halo, D]; the first n_local rows are the rank's own
    features (a view, no copy), the halo rows are filled by `exchange()`: each owner gathers the rows
    its peers asked for (index lists swapped once at setup) and ONE all_to_all_single moves them;
    every remote row crosses NVLink once per aggregation no matter how many local edges use it;
  * the aggregation kernel then runs unchanged on the local CSR over X_ext.  F4 (the backward reuses the
    same CSR) means forward and backward need exactly the same exchange, no transpose.
  * `degrees` of the halo nodes are fetched once at setup with the same exchange.
"""
import ctypes
import os

import torch
import torch.distributed as dist


def default_row_weight(world, avg_degree=None):
    """Cost of owning one row, in edge-equivalents, when cutting the vertex ranges.  A rank's step is aggregation
    (proportional to its edges) plus halo traffic.  When the halo rows are gathered and pushed by a kernel (sparse halos)
    every row a rank owns costs its SMs time once per peer that needs it -- measured on 8xB200 (D=64): 0.0154 ns per edge
    against ~4 ns per row at 7 peers, i.e. ~40 edge-equivalents per row and peer.  On DENSE graphs (avg_degree / world >= 16:
    every peer needs nearly every row) that weight shifts so many edges onto the hub ranks that their kernel becomes the
    step (0.24 vs 0.20 ms at 8 GPUs); a quarter of it is the measured optimum there."""
    if world <= 1:
        return 0
    if avg_degree is not None and avg_degree / world >= 16:
        return 10 * (world - 1)          # measured at 4 and 8 GPUs (profiles/r02_*sweep*): 10 beats both 0 and 40 per peer
    return 40 * (world - 1)


def partition_ranges(row_ptr, world, row_weight=0):
    """Vertex boundaries v[0..world] with ~equal cost: cost(rows [a,b)) = edges + row_weight * (b-a).
    row_weight = 0 balances edges only: v_g = first row whose offset >= g*E/world."""
    rp = row_ptr.to(torch.int64)
    n = rp.numel() - 1
    if row_weight:
        rp = rp + int(row_weight) * torch.arange(n + 1, dtype=torch.int64, device=rp.device)
    E = int(rp[-1])
    targets = torch.tensor([(E * g) // world for g in range(world + 1)], dtype=torch.int64, device=rp.device)
    v = torch.searchsorted(rp, targets, right=False).clamp_(0, n)
    v[0], v[-1] = 0, n
    v = torch.cummax(v, 0).values
    return [int(x) for x in v.cpu()]


def _a2a(out, inp, out_splits, in_splits, group):
    """all_to_all_single, or point-to-point exchanges on backends without it (gloo)."""
    backend = dist.get_backend(group)
    if backend == "nccl":
        dist.all_to_all_single(out, inp, out_splits, in_splits, group=group)
        return
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if out.is_cuda:      # gloo has no CUDA point-to-point: stage through the host (tests that run the CUDA kernels
        o_h = torch.empty(out.shape, dtype=out.dtype)          # of several ranks on ONE GPU use this)
        _a2a(o_h, inp.cpu(), out_splits, in_splits, group)
        out.copy_(o_h)
        return
    ooff = [0]
    ioff = [0]
    for s in out_splits:
        ooff.append(ooff[-1] + s)
    for s in in_splits:
        ioff.append(ioff[-1] + s)
    reqs = []
    for p in range(world):
        if p == rank:
            out[ooff[p]:ooff[p + 1]].copy_(inp[ioff[p]:ioff[p + 1]])
            continue
        if in_splits[p]:
            reqs.append(dist.isend(inp[ioff[p]:ioff[p + 1]].contiguous(), dist.get_global_rank(group, p) if group else p, group=group))
        if out_splits[p]:
            reqs.append(dist.irecv(out[ooff[p]:ooff[p + 1]], dist.get_global_rank(group, p) if group else p, group=group))
    for r in reqs:
        r.wait()


class ShardedGraph:
    """This rank's shard of a graph every rank can see (replicated CSR in, sharded tables out)."""

    def __init__(self, row_ptr, col_idx, part_size, device=None, group=None, ranges=None, row_weight=0, dense_halo=None):
        world = dist.get_world_size(group)
        rank = dist.get_rank(group)
        device = torch.device(device) if device is not None else row_ptr.device
        ranges = ranges if ranges is not None else partition_ranges(row_ptr, world, row_weight)
        v0, v1 = ranges[rank], ranges[rank + 1]
        e0, e1 = int(row_ptr[v0]), int(row_ptr[v1])
        rp_local = (row_ptr[v0:v1 + 1].to(torch.int64) - e0).to(device)
        self._setup(ranges, rp_local, col_idx[e0:e1].to(device), part_size, device, group,
                    num_nodes_global=row_ptr.numel() - 1, num_edges_global=int(row_ptr[-1]), dense_halo=dense_halo)

    @classmethod
    def from_rows(cls, ranges, row_ptr_local, cols_global, part_size, device=None, group=None, dense_halo=None):
        """A shard built from the rank's OWN rows only (graph.synth_graph_shard, or a loader that reads a vertex range):
        row_ptr_local [n_local+1] offsets from 0, cols_global [E_local] GLOBAL neighbour ids.  No rank ever holds the
        whole graph; the global edge count is an all-reduce of the local ones (int64: it exceeds 2^31 for
        ogbn-papers100M-size graphs while every local CSR stays int32)."""
        self = cls.__new__(cls)
        device = torch.device(device) if device is not None else row_ptr_local.device
        e_local = torch.tensor([int(row_ptr_local[-1])], dtype=torch.int64, device=device)
        if dist.get_world_size(group) > 1:
            dist.all_reduce(e_local, group=group)
        self._setup(list(ranges), row_ptr_local.to(device), cols_global.to(device), part_size, device, group,
                    num_nodes_global=int(ranges[-1]), num_edges_global=int(e_local.item()), dense_halo=dense_halo)
        return self

    def _setup(self, ranges, rp_local, cols_global, part_size, device, group, num_nodes_global, num_edges_global,
               dense_halo=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = device
        self.part_size = int(part_size)
        self.ranges = ranges
        v0, v1 = self.ranges[self.rank], self.ranges[self.rank + 1]
        self.v0, self.n_local = v0, v1 - v0
        self.num_nodes_global = int(num_nodes_global)
        self.num_edges_local = int(rp_local[-1])
        self.num_edges_global = int(num_edges_global)
        if self.num_edges_local >= 2 ** 31:
            raise ValueError("shard holds %d edges; a local CSR is int32 -- use more ranks" % self.num_edges_local)
        cols = cols_global.to(torch.int64)
        bounds = torch.tensor(self.ranges, dtype=torch.int64, device=device)
        remote = (cols < v0) | (cols >= v1)
        halo = torch.unique(cols[remote])                                   # sorted => grouped by owner
        owner = torch.searchsorted(bounds, halo, right=True) - 1
        needed = [int(x) for x in torch.bincount(owner, minlength=self.world).cpu()]
        # DENSE halos: when this rank needs at least `dense_halo` (default 0.75, GNNA_DENSE_HALO; 0 = never) of an owner's
        # rows -- every pair does on a dense graph: 92 % on the Reddit look-alike at 8 GPUs -- it asks for the owner's WHOLE
        # range.  The owner's block is then its local rows as they lie in memory and travels as one copy-engine memcpy
        # over NVLink (csrc/halo.cu: gnna_halo_push_ce) instead of a gather kernel over an index list.
        # Measured on 4xB200 (profiles/r02_v2_sweep_4gpu.txt): the copy-engine exchange frees the SMs (step 0.439 vs 0.448 ms)
        # but competes with the H2D/D2H copies of an end-to-end step (e2e 0.89 vs 0.69 ms), so it is opt-in:
        # GNNA_HALO_CE=1 (and then GNNA_DENSE_HALO, default 0.75, is the density from which a whole range is requested).
        if dense_halo is None:
            dense_halo = float(os.environ.get("GNNA_DENSE_HALO", "0.75")) if os.environ.get("GNNA_HALO_CE", "0") == "1" else 0.0
        self.dense_from = [q != self.rank and dense_halo > 0 and self.ranges[q + 1] > self.ranges[q]
                           and needed[q] >= dense_halo * (self.ranges[q + 1] - self.ranges[q]) for q in range(self.world)]
        if any(self.dense_from):
            halo = torch.unique(torch.cat([halo] + [torch.arange(self.ranges[q], self.ranges[q + 1], dtype=torch.int64, device=device)
                                                    for q in range(self.world) if self.dense_from[q]]))
            owner = torch.searchsorted(bounds, halo, right=True) - 1
        self.halo_rows_needed = needed
        self.n_halo = int(halo.numel())
        self.halo_ids = halo
        self.recv_counts = [int(x) for x in torch.bincount(owner, minlength=self.world).cpu()]
        local_cols = torch.where(remote, self.n_local + torch.searchsorted(halo, cols), cols - v0).to(torch.int32)
        del cols, remote
        self.row_ptr = rp_local.to(torch.int32).contiguous()
        self.col_idx = local_cols.contiguous()
        # tell every owner which of ITS rows we need (as row indices local to the owner)
        want = (halo - bounds[owner]).to(torch.int64)
        rc = torch.tensor(self.recv_counts, dtype=torch.int64, device=device)
        sc = torch.empty_like(rc)
        _a2a(sc, rc, [1] * self.world, [1] * self.world, group)
        self.send_counts = [int(x) for x in sc.cpu()]
        self.send_idx = torch.empty(sum(self.send_counts), dtype=torch.int64, device=device)
        _a2a(self.send_idx, want, self.send_counts, self.recv_counts, group)
        # a peer that wants n_local distinct rows of mine wants all of them, in order: its block is my rows verbatim
        self.dense_to = [p != self.rank and self.n_local > 0 and self.send_counts[p] == self.n_local for p in range(self.world)]
        self.n_ext = self.n_local + self.n_halo
        self._tables_built = False
        self.part_ptr = self.part2node = self.degrees_ext = None

    # -------------------------------------------------------------------------------- setup on the device
    def build_tables(self, degrees_global=None):
        """Neighbour-group table of the local CSR (device build_part) and degrees of local + halo nodes.
        The GCN degree is the GLOBAL row degree, which a rank knows for its own rows; halo degrees come
        through the exchange."""
        from . import ops
        if self.device.type == "cuda":
            self.part_ptr, self.part2node = ops.build_part_exact(self.part_size, self.row_ptr)
            deg_local = ops.degrees_from_row_ptr(self.row_ptr)
        else:   # CPU (gloo tests): host build_part of the C ABI + torch ops for the degrees
            self.part_ptr, self.part2node = ops.build_part_exact(self.part_size, self.row_ptr)
            d = (self.row_ptr[1:] - self.row_ptr[:-1]).to(torch.float32)
            deg_local = torch.sqrt(torch.clamp(d, min=1.0))
        ext = torch.empty(self.n_ext, 1, dtype=torch.float32, device=self.device)
        ext[:self.n_local, 0] = deg_local
        self.exchange(ext)
        self.degrees_ext = ext[:, 0].contiguous()
        self._tables_built = True
        return self

    def new_features(self, dim, dtype=torch.float32):
        """X_ext [n_local + n_halo, dim]; fill [:n_local] (= .local(x)) with this rank's rows."""
        return torch.empty(self.n_ext, dim, dtype=dtype, device=self.device)

    def local(self, x_ext):
        return x_ext[:self.n_local]

    # -------------------------------------------------------------------------------- per aggregation
    def exchange(self, x_ext):
        """Fill the halo rows of x_ext from their owners: gather + ONE all-to-all."""
        if self.world == 1:
            return x_ext
        send = x_ext[:self.n_local].index_select(0, self.send_idx)
        _a2a(x_ext[self.n_local:], send, self.recv_counts, self.send_counts, self.group)
        return x_ext

    # -------------------------------------------------------------------------------- overlap: per-owner sub-shards
    def build_owner_shards(self):
        """Split this rank's CSR by the OWNER of the neighbour: one sub-CSR (+ group table) per rank, listed in the
        order their rows arrive (own rows first, then rank-1, rank-2, ...: every sender serves its peers in ring
        order).  The overlapped step aggregates a sub-shard as soon as its owner's rows have landed while the
        remaining rows are still crossing NVLink."""
        from . import ops
        rank, world, n_local, dev = self.rank, self.world, self.n_local, self.device
        hoff = [0]
        for c in self.recv_counts:
            hoff.append(hoff[-1] + c)
        hoff_t = torch.tensor(hoff, dtype=torch.int64, device=dev)
        cols = self.col_idx.to(torch.int64)
        deg_rows = (self.row_ptr[1:] - self.row_ptr[:-1]).to(torch.int64)
        rows = torch.repeat_interleave(torch.arange(n_local, device=dev), deg_rows)
        owner = torch.where(cols < n_local, torch.full_like(cols, rank),
                            torch.searchsorted(hoff_t, (cols - n_local).clamp_min(0), right=True) - 1)
        self.owner_order = [(rank - k) % world for k in range(world)]
        self.owner_shards = []
        for q in self.owner_order:
            m = owner == q
            rp = torch.zeros(n_local + 1, dtype=torch.int64, device=dev)
            rp[1:] = torch.cumsum(torch.bincount(rows[m], minlength=n_local), 0)
            rp = rp.to(torch.int32).contiguous()
            ci = cols[m].to(torch.int32).contiguous()
            # a row's neighbours are split over `world` sub-CSRs, so its groups are shorter than in the whole CSR; a larger
            # group size for the sub-shards (GNNA_OWNER_PS) trades that fragmentation against balance
            pp, pn = ops.build_part_exact(int(os.environ.get("GNNA_OWNER_PS", self.part_size)), rp)
            self.owner_shards.append((rp, ci, pp, pn))
        # the same sub-shards as ONE group table, segment after segment in arrival order: what the fused
        # exchange+aggregation kernel walks (gnna_aggregate_gated_f32)
        e_off, g_off, pps, bounds = 0, 0, [], [0]
        for rp, ci, pp, pn in self.owner_shards:
            pps.append(pp[:-1].to(torch.int64) + e_off)
            e_off += int(ci.numel())
            g_off += int(pn.numel())
            bounds.append(g_off)
        self.gated_col = torch.cat([sh[1] for sh in self.owner_shards]).contiguous()
        self.gated_pp = torch.cat(pps + [torch.tensor([e_off], dtype=torch.int64, device=dev)]).to(torch.int32).contiguous()
        self.gated_pn = torch.cat([sh[3] for sh in self.owner_shards]).contiguous()
        self.gated_bounds = (ctypes.c_int64 * (world + 1))(*bounds)
        self.gated_peer = (ctypes.c_int32 * world)(*[-1 if q == rank else q for q in self.owner_order])
        return self

    def write_local(self, peer, x_local, prescale):
        """Fill the local rows of the next step's feature buffer from x_local [n_local, D] (scaled by the GCN
        degrees when `prescale`): what a layer's producer (the X*W epilogue) does in a real pipeline."""
        from . import _lib
        dst = self.local(peer.features(dim=x_local.shape[1]))
        if prescale:
            p = lambda t: ctypes.c_void_p(t.data_ptr() if t.numel() else 0)   # noqa: E731
            with torch.cuda.device(self.device):
                _lib.check(_lib.load().gnna_prescale_rows_f32(p(x_local), p(dst), p(self.degrees_ext), self.n_local, x_local.shape[1],
                                                              ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "prescale")
        else:
            dst.copy_(x_local)
        return dst

    def aggregate_overlapped(self, mode, peer, out, eps=0.5, dim_worker=32, warp_per_block=4):
        """One sharded step with the exchange hidden behind the aggregation.  The next step's buffer
        (peer.features()) must hold this rank's rows -- pre-scaled by degrees for mode 1, see write_local.
        mode: 0 SAG, 1 GCN, 2 GIN."""
        from . import _lib
        assert self._tables_built and getattr(self, "owner_shards", None), "call build_tables() and build_owner_shards() first"
        lib = _lib.load()
        torch.cuda.set_device(self.device)            # one process drives one GPU; every launch below goes there
        cur = torch.cuda.current_stream(self.device)
        if not hasattr(self, "_comm_stream"):
            self._comm_stream = torch.cuda.Stream(self.device, priority=-1)
            self._ev_ready, self._ev_pushed = torch.cuda.Event(), torch.cuda.Event()
        peer.begin_step()
        self._ev_ready.record(cur)
        self._comm_stream.wait_event(self._ev_ready)              # my rows are in the buffer, the step is advanced
        x_ext = peer.push(self._comm_stream)
        self._ev_pushed.record(self._comm_stream)
        d = x_ext.shape[1]
        kmode = 3 if mode == 1 else mode
        p = lambda t: ctypes.c_void_p(t.data_ptr() if t.numel() else 0)   # noqa: E731
        st = ctypes.c_void_p(cur.cuda_stream)
        # ONE kernel whose CTAs wait for a peer's flag inside the kernel (csrc/aggregate.cu) against one kernel per owner
        # sub-shard with wait kernels in between: measured on the Reddit look-alike (profiles/r02_*sweep*, ab_gated) the fused
        # kernel wins where the sub-kernels are short -- 8 GPUs: 0.304 vs 0.324 ms -- and loses 2-3 % at 2 and 4 GPUs (every CTA
        # of a peer segment pays the flag round trip).  GNNA_GATED=0/1 forces either.
        if os.environ.get("GNNA_GATED", "1" if self.world >= 8 else "0") == "1":
            _lib.check(lib.gnna_aggregate_gated_f32(kmode, p(x_ext), self.n_ext, p(out), self.n_local, p(self.row_ptr),
                                                    p(self.gated_col), p(self.degrees_ext) if kmode == 3 else ctypes.c_void_p(0),
                                                    float(eps), p(self.gated_pp), p(self.gated_pn), d, self.gated_pn.numel(),
                                                    self.gated_bounds, self.gated_peer, self.world, ctypes.c_void_p(peer.ctrl_ptr),
                                                    self.part_size, int(dim_worker),
                                                    int(os.environ.get("GNNA_GATED_WPB", warp_per_block)), st), "gated aggregate")
            peer.ack()
            cur.wait_event(self._ev_pushed)
            return out
        for k, (q, (rp, ci, pp, pn)) in enumerate(zip(self.owner_order, self.owner_shards)):
            if q != self.rank:
                peer.wait([q], cur)
            _lib.check(lib.gnna_aggregate_part_f32_ex(kmode, 0 if k == 0 else 1, p(x_ext), self.n_ext, p(out), self.n_local,
                                                      p(rp), p(ci), p(self.degrees_ext) if kmode == 3 else ctypes.c_void_p(0),
                                                      float(eps), p(pp), p(pn), d, pn.numel(),
                                                      self.part_size, int(dim_worker), int(warp_per_block), st), "overlapped aggregate")
        peer.ack()
        cur.wait_event(self._ev_pushed)                           # the push has read my rows: the buffer may be refilled
        return out

    def halo_bytes(self, dim, elem=4):
        return {"recv": self.n_halo * dim * elem, "send": int(self.send_idx.numel()) * dim * elem}

    def aggregate(self, mode, x_ext, out=None, eps=0.5, dim_worker=32, warp_per_block=8, do_exchange=True, peer=None):
        """out[n_local, D] = aggregation of this rank's rows (mode 0 SAG, 1 GCN, 2 GIN) over X_ext.
        peer: a PeerHalo -- x_ext is ignored, the step's buffer is peer.features(); the exchange is the
        NVLink push kernel instead of gather + all_to_all."""
        from . import _lib
        assert self._tables_built, "call build_tables() first"
        if peer is not None:
            x_ext = peer.exchange() if do_exchange else peer.features(peer.step)
        elif do_exchange:
            self.exchange(x_ext)
        d = x_ext.shape[1]
        if out is None:
            out = torch.empty(self.n_local, d, dtype=torch.float32, device=self.device)
        lib = _lib.load()
        p = lambda t: ctypes.c_void_p(t.data_ptr() if t.numel() else 0)   # noqa: E731
        P = self.part2node.numel()
        with torch.cuda.device(self.device):
            st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            rc = lib.gnna_aggregate_f32_ex(int(mode), p(x_ext), self.n_ext, p(out), self.n_local,
                                           p(self.row_ptr), p(self.col_idx),
                                           p(self.degrees_ext) if mode == 1 else ctypes.c_void_p(0), float(eps),
                                           p(self.part_ptr), p(self.part2node), d, P,
                                           self.part_size, int(dim_worker), int(warp_per_block), st)
        _lib.check(rc, "sharded aggregate")
        if peer is not None and do_exchange:
            peer.ack()
        return out


class _RawCuda:
    """A raw device pointer dressed up for torch.as_tensor (no ownership)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class PeerHalo:
    """NVLink-native halo exchange (csrc/halo.cu): every rank maps its peers' feature buffers through
    CUDA IPC, ONE kernel stores the rows a peer needs straight into that peer's halo rows and raises a
    flag there; no gather buffer, no collective launch.  Feature buffers are double-buffered by step
    parity; use `features(step)` to get the [n_ext, dim] tensor a step works on (fill its first n_local
    rows), then `sg.aggregate(..., peer=True)`.

    `dim` is the WIDEST matrix that will be exchanged; a step may exchange any narrower one (`features(dim=d)` /
    `stage(d)`): the layers of a model trade matrices of different widths through the same mapped buffers.

    Collective constructor (all ranks of the group must call it)."""

    def __init__(self, sg, dim):
        from . import _lib
        self.sg, self.dim, self.lib = sg, int(dim), _lib.load()
        self.step = 0
        self.cur_dim = self.dim
        self._views = {}
        dev, world, rank, group = sg.device, sg.world, sg.rank, sg.group
        self._own = []
        handles = torch.zeros(3, 64, dtype=torch.uint8)
        ptrs = []
        sizes = [max(sg.n_ext, 1) * self.dim * 4] * 2 + [64 * 4]
        with torch.cuda.device(dev):
            for i, nbytes in enumerate(sizes):
                ptr = ctypes.c_void_p(0)
                h = (ctypes.c_ubyte * 64)()
                _lib.check(self.lib.gnna_ipc_alloc(nbytes, ctypes.byref(ptr), h), "ipc_alloc")
                ptrs.append(ptr.value)
                self._own.append(ptr.value)
                handles[i] = torch.frombuffer(bytearray(h), dtype=torch.uint8)
        self.buf_ptr, self.ctrl_ptr = ptrs[:2], ptrs[2]
        self.bufs = [self._view(b, self.dim) for b in range(2)]
        # everyone learns everyone's handles, local row counts and per-source receive counts
        gathered = [torch.zeros_like(handles) for _ in range(world)]
        meta = torch.tensor([sg.n_local] + sg.recv_counts, dtype=torch.int64)
        metas = [torch.zeros_like(meta) for _ in range(world)]
        if dist.get_backend(group) == "nccl":
            g_h = [t.to(dev) for t in gathered]
            dist.all_gather(g_h, handles.to(dev), group=group)
            gathered = [t.cpu() for t in g_h]
            g_m = [t.to(dev) for t in metas]
            dist.all_gather(g_m, meta.to(dev), group=group)
            metas = [t.cpu() for t in g_m]
        else:
            dist.all_gather(gathered, handles, group=group)
            dist.all_gather(metas, meta, group=group)
        self.peer_buf = [[0] * world, [0] * world]
        self.peer_ctrl = [0] * world
        self._opened = []
        with torch.cuda.device(dev):
            for p in range(world):
                if p == rank:
                    self.peer_buf[0][p], self.peer_buf[1][p], self.peer_ctrl[p] = self.buf_ptr[0], self.buf_ptr[1], self.ctrl_ptr
                    continue
                opened = []
                for i in range(3):
                    h = (ctypes.c_ubyte * 64).from_buffer_copy(bytes(gathered[p][i].tolist()))
                    ptr = ctypes.c_void_p(0)
                    _lib.check(self.lib.gnna_ipc_open(h, ctypes.byref(ptr)), "ipc_open")
                    opened.append(ptr.value)
                    self._opened.append(ptr.value)
                self.peer_buf[0][p], self.peer_buf[1][p], self.peer_ctrl[p] = opened
        # where my block of rows starts inside peer p's buffer: after p's own rows and the blocks of lower ranks
        self.dst_row0 = [0] * world
        for p in range(world):
            n_local_p = int(metas[p][0])
            recv_p = [int(v) for v in metas[p][1:]]
            self.dst_row0[p] = n_local_p + sum(recv_p[:rank])
        sb = [0]
        for c in sg.send_counts:
            sb.append(sb[-1] + c)
        self.send_begin = (ctypes.c_int32 * (world + 1))(*sb)
        self.c_dst_row0 = (ctypes.c_int64 * world)(*self.dst_row0)
        self.c_peer_ctrl = (ctypes.c_void_p * world)(*self.peer_ctrl)
        self.c_peer_buf = [(ctypes.c_void_p * world)(*self.peer_buf[b]) for b in range(2)]
        # peers that asked for my whole range are served by the copy engines (GNNA_HALO_CE=0: always the push kernel)
        self.dense_mask = 0
        if os.environ.get("GNNA_HALO_CE", "0") == "1":
            for p in range(world):
                if sg.dense_to[p]:
                    self.dense_mask |= 1 << p
        dist.barrier(group=group)

    def _view(self, parity, dim):
        key = (parity, int(dim))
        if key not in self._views:
            if dim > self.dim:
                raise ValueError("PeerHalo was mapped for rows of <= %d floats, asked for %d" % (self.dim, dim))
            self._views[key] = torch.as_tensor(_RawCuda(self.buf_ptr[parity], (self.sg.n_ext, int(dim)), "<f4"), device=self.sg.device)
        return self._views[key]

    def features(self, step=None, dim=None):
        """The [n_ext, dim] buffer step `step` (default: the next one) gathers from.  Asking for the NEXT step's buffer
        also fixes the width that step exchanges."""
        d = self.cur_dim if dim is None else int(dim)
        if step is None:
            self.cur_dim = d
        s = self.step + 1 if step is None else step
        return self._view(s & 1, d)

    stage = lambda self, dim: self.features(None, dim)   # noqa: E731  (reads better at the call sites of the layers)

    def begin_step(self):
        """First call of a step, on the compute (current) stream: advances the step counter on the device.
        Returns the step's feature buffer."""
        from . import _lib
        self.step += 1
        with torch.cuda.device(self.sg.device):
            st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(self.lib.gnna_halo_begin_step(ctypes.c_void_p(self.ctrl_ptr), st), "halo_begin_step")
        return self._view(self.step & 1, self.cur_dim)

    def push(self, stream=None):
        """Store my rows into my peers' halo rows (ring order) on `stream`, which must be ordered after begin_step."""
        from . import _lib
        sg = self.sg
        b = self.step & 1
        if self.dense_mask:
            with torch.cuda.device(sg.device):
                st = ctypes.c_void_p((stream or torch.cuda.current_stream()).cuda_stream)
                _lib.check(self.lib.gnna_halo_push_ce(ctypes.c_void_p(self.buf_ptr[b]), sg.n_local,
                                                      ctypes.c_void_p(sg.send_idx.data_ptr() if sg.send_idx.numel() else 0),
                                                      self.send_begin, self.c_peer_buf[b], self.c_peer_ctrl, self.c_dst_row0,
                                                      ctypes.c_void_p(self.ctrl_ptr), sg.world, sg.rank, self.cur_dim,
                                                      self.dense_mask, st), "halo_push_ce")
            return self._view(b, self.cur_dim)
        with torch.cuda.device(sg.device):
            st = ctypes.c_void_p((stream or torch.cuda.current_stream()).cuda_stream)
            _lib.check(self.lib.gnna_halo_push_f32(ctypes.c_void_p(self.buf_ptr[b]),
                                                   ctypes.c_void_p(sg.send_idx.data_ptr() if sg.send_idx.numel() else 0),
                                                   self.send_begin, self.c_peer_buf[b], self.c_peer_ctrl, self.c_dst_row0,
                                                   ctypes.c_void_p(self.ctrl_ptr), sg.world, sg.rank, self.cur_dim, st), "halo_push")
        return self._view(b, self.cur_dim)

    def wait(self, peers=None, stream=None):
        """Make `stream` wait until the rows of `peers` (ranks; None = all) for the current step have landed."""
        from . import _lib
        mask = 0
        for q in (peers or []):
            mask |= 1 << q
        with torch.cuda.device(self.sg.device):
            st = ctypes.c_void_p((stream or torch.cuda.current_stream()).cuda_stream)
            _lib.check(self.lib.gnna_halo_wait(ctypes.c_void_p(self.ctrl_ptr), self.sg.world, self.sg.rank, mask, st), "halo_wait")

    def exchange(self):
        """Push my rows into my peers' halo rows for the next step and wait for theirs (on the current stream)."""
        self.begin_step()
        buf = self.push()
        self.wait()
        return buf

    def ack(self):
        """After the aggregation of the current step: producers may overwrite this parity again."""
        from . import _lib
        with torch.cuda.device(self.sg.device):
            st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(self.lib.gnna_halo_ack(self.c_peer_ctrl, ctypes.c_void_p(self.ctrl_ptr), self.sg.world, self.sg.rank, st), "halo_ack")

    def error_word(self):
        """The control block's error word as a 1-element device tensor (no synchronisation): 0 = fine, 1 = a producer
        here gave up waiting for a peer's acknowledgement, 2 = a flag wait here timed out, 3 = a PRODUCER could not
        deliver this rank's halo rows (they are stale).  Read it together with something the caller synchronises on
        anyway (the loss of an epoch)."""
        return torch.as_tensor(_RawCuda(self.ctrl_ptr + 48 * 4, (1,), "<i4"), device=self.sg.device)

    def error(self):
        """error_word() on the host.  Synchronises."""
        torch.cuda.synchronize(self.sg.device)
        return int(self.error_word().item())

    def check(self):
        """Raise if any bounded wait of the exchange timed out since the buffers were mapped.  Synchronises."""
        e = self.error()
        if e:
            raise RuntimeError("halo exchange failed on rank %d (error word %d: %s)" % (
                self.sg.rank, e, {1: "no acknowledgement from a peer", 2: "a peer's rows never arrived",
                                  3: "a peer could not deliver this rank's halo rows"}.get(e, "?")))

    def close(self):
        torch.cuda.synchronize(self.sg.device)
        dist.barrier(group=self.sg.group)
        self.bufs, self._views = [], {}
        with torch.cuda.device(self.sg.device):
            for p in self._opened:
                self.lib.gnna_ipc_close(ctypes.c_void_p(p))
            dist.barrier(group=self.sg.group)
            for p in self._own:
                self.lib.gnna_ipc_free(ctypes.c_void_p(p))
        self._opened, self._own = [], []


def allreduce_weight_grad(d_weight, group=None):
    """dW = X^T G is a sum over rows => sum over ranks (tiny: din x dout)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(d_weight, group=group)
    return d_weight
