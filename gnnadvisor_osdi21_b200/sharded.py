"""Sharded GCN / GIN layers and the multi-GPU training epoch: the distributed counterpart of
GNNAdvisor/gnn_conv.py:31-147 and of the epoch loop in GNNAdvisor/GNNA_main.py:142-187.

The reference is single-GPU; BASELINE.json's metric asks for the GCN epoch at 1/2/4/8 GPUs and for graphs
that exceed one GPU.  One process per GPU; rank r owns a contiguous vertex range (dist.ShardedGraph) and with it the
rows of every activation and gradient matrix; weights are replicated.

  GCN forward   T = X_r W            (row-local product)          kernel.cu:280
                halo exchange of n_j*T_j                           -- the one collective of the aggregation
                out_i = n_i * sum_j (n_j T_j)                      kernel.cu:383-413 on [own rows | halo rows]
  GCN backward  halo exchange of n_j*dOut_j, G = Ahat dOut         kernel.cu:436-463 (same CSR, same exchange: F4)
                dX = G W^T (row-local), dW_r = X_r^T G_r           kernel.cu:472-473
  GIN forward   halo exchange of X, S = eps * A X, out = S W       kernel.cu:572-605
  GIN backward  dW_r = S_r^T dOut_r, Pm = dOut W^T, exchange, dX = eps * A Pm     kernel.cu:710-738
  after backward: ONE all-reduce of all the weight gradients (dW = sum over ranks of dW_r; they are tiny), Adam replicated.

The exchange is NCCL all_to_all_single ("nccl"), the NVLink push kernel over CUDA-IPC mapped buffers ("peer",
csrc/halo.cu) or the push hidden behind per-owner sub-shard aggregation ("overlap"), see dist.py.  Row widths that are not
a multiple of four floats (41 or 47 classes) are zero-padded through the weight matrix, so every exchanged row is whole
16-byte chunks and the padded columns cost nothing but their bytes.

`compute` is the object that runs the three device operations (dense product, row pre-scale, aggregation): the product's is
CudaCompute = libgnna_b200.so through the C ABI, nothing else ships.  The CPU tests (gloo) pass their own.
"""
import ctypes
import math

import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None and t.numel() > 0 else ctypes.c_void_p(0)


class CudaCompute:
    """The device operations of a sharded layer on libgnna_b200.so.  Raises if the library is missing: no fallback."""

    def __init__(self):
        self.lib = _lib.load()

    @staticmethod
    def _stream():
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def mm(self, A, B, ta=False, tb=False, out=None):
        """Row-major op(A) @ op(B) in fp32 (cuBLAS SGEMM on the library's handle, TF32 off: what torch::mm is in the reference)."""
        m = A.shape[1] if ta else A.shape[0]
        k = A.shape[0] if ta else A.shape[1]
        n = B.shape[0] if tb else B.shape[1]
        assert (B.shape[1] if tb else B.shape[0]) == k and A.is_contiguous() and B.is_contiguous()
        if out is None:
            out = torch.empty(m, n, dtype=torch.float32, device=A.device)
        assert out.is_contiguous() and out.shape == (m, n)
        with torch.cuda.device(A.device):
            _lib.check(self.lib.gnna_sgemm_f32(int(ta), int(tb), m, n, k, _p(A), _p(B), _p(out), self._stream()), "sgemm")
        return out

    def prescale(self, X, degrees, out):
        """out[i,:] = degrees[i] * X[i,:] (out may be X)."""
        with torch.cuda.device(X.device):
            _lib.check(self.lib.gnna_prescale_rows_f32(_p(X), _p(out), _p(degrees), X.shape[0], X.shape[1], self._stream()), "prescale")
        return out

    def aggregate(self, sg, mode, x_ext, out, eps, dim_worker, warp_per_block):
        """out[n_local, D] over the local CSR of `sg` and x_ext [n_ext, D]; mode 0 SAG, 2 GIN, 3 GCN on pre-scaled rows."""
        with torch.cuda.device(x_ext.device):
            _lib.check(self.lib.gnna_aggregate_part_f32_ex(
                int(mode), 0, _p(x_ext), sg.n_ext, _p(out), sg.n_local, _p(sg.row_ptr), _p(sg.col_idx),
                _p(sg.degrees_ext) if mode == 3 else ctypes.c_void_p(0), float(eps), _p(sg.part_ptr), _p(sg.part2node),
                x_ext.shape[1], sg.part2node.numel(), sg.part_size, int(dim_worker), int(warp_per_block), self._stream()),
                "sharded aggregate")
        return out


class ShardedInputInfo:
    """What a sharded layer needs besides its input: the rank's shard, the kernel parameters (the reference's
    inputProperty fields partSize / dimWorker / warpPerBlock, param.py:19-29) and how halo rows travel.

    exchange: "nccl"    gather + ONE all_to_all_single per aggregation (works on every backend; gloo in the CPU tests)
              "peer"    NVLink push kernel over CUDA-IPC mapped buffers, then the aggregation
              "overlap" the push hidden behind per-owner sub-shard aggregation (rows of width % 4 == 0 only)
    max_dim: widest matrix any layer will exchange (peer / overlap map their buffers once, for that width)."""

    def __init__(self, sg, dimWorker=32, warpPerBlock=4, exchange="nccl", max_dim=None, compute=None):
        assert exchange in ("nccl", "peer", "overlap")
        if not sg._tables_built:
            sg.build_tables()
        self.sg = sg
        self.partSize, self.dimWorker, self.warpPerBlock = sg.part_size, int(dimWorker), int(warpPerBlock)
        self.exchange = exchange if sg.world > 1 else "nccl"
        self.compute = compute if compute is not None else CudaCompute()
        self.peer = None
        self._bufs = {}
        if self.exchange in ("peer", "overlap"):
            from .dist import PeerHalo
            assert max_dim, "peer / overlap exchange needs max_dim"
            self.peer = PeerHalo(sg, (int(max_dim) + 3) // 4 * 4)
            if self.exchange == "overlap":
                sg.build_owner_shards()

    # ---- one aggregation = stage (producer fills the first n_local rows) -> gather (exchange + kernel)
    def stage(self, dim):
        """[n_ext, dim] buffer of the next aggregation; the caller fills rows [0, n_local)."""
        if self.peer is not None:
            return self.peer.stage(dim)
        if dim not in self._bufs:
            self._bufs[dim] = torch.empty(self.sg.n_ext, dim, dtype=torch.float32, device=self.sg.device)
        return self._bufs[dim]

    def gather(self, mode, x_ext, eps=0.5):
        """Exchange the halo rows of the staged buffer and aggregate: returns out [n_local, dim].
        mode 0 SAG, 2 GIN (eps * sum), 3 GCN on rows pre-scaled by degrees."""
        sg = self.sg
        out = torch.empty(sg.n_local, x_ext.shape[1], dtype=torch.float32, device=sg.device)
        if self.exchange == "overlap":
            sg.aggregate_overlapped(1 if mode == 3 else mode, self.peer, out, eps=eps,
                                    dim_worker=self.dimWorker, warp_per_block=self.warpPerBlock)
            return out
        if self.peer is not None:
            x_ext = self.peer.exchange()
        else:
            sg.exchange(x_ext)
        self.compute.aggregate(sg, mode, x_ext, out, eps, self.dimWorker, self.warpPerBlock)
        if self.peer is not None:
            self.peer.ack()
        return out

    def check(self):
        if self.peer is not None:
            self.peer.check()

    def close(self):
        if self.peer is not None:
            self.peer.close()
            self.peer = None


def _pad4(d):
    return (d + 3) // 4 * 4


class ShardedGNNAFunction(torch.autograd.Function):
    """GCN layer on a shard (gnn_conv.py:31-78 / kernel.cu:267-322, 422-476).  weight's width is a multiple of 4."""

    @staticmethod
    def forward(ctx, X, weight, info):
        sg, c = info.sg, info.compute
        X, weight = X.contiguous(), weight.contiguous()
        buf = info.stage(weight.shape[1])
        local = buf[:sg.n_local]
        c.mm(X, weight, out=local)                                       # T = X W                      kernel.cu:280
        c.prescale(local, sg.degrees_ext, local)                         # n_j * T_j, in place
        ctx.save_for_backward(X, weight)
        ctx.info = info
        return info.gather(3, buf)                                       # n_i * sum_j (...)            :383-413

    @staticmethod
    def backward(ctx, d_output):
        X, weight = ctx.saved_tensors
        info = ctx.info
        sg, c = info.sg, info.compute
        d_output = d_output.contiguous()
        buf = info.stage(d_output.shape[1])
        c.prescale(d_output, sg.degrees_ext, buf[:sg.n_local])
        G = info.gather(3, buf)                                          # G = Ahat dOut                :436-463
        d_input = c.mm(G, weight, tb=True) if ctx.needs_input_grad[0] else None       # :472 (skipped when unused)
        d_weight = c.mm(X, G, ta=True)                                   # this rank's share of X^T G   :473
        return d_input, d_weight, None


class ShardedGNNAFunction_GIN(torch.autograd.Function):
    """GIN layer on a shard (gnn_conv.py:101-126 / kernel.cu:559-617, 696-747).  X's width is a multiple of 4."""

    @staticmethod
    def forward(ctx, X, weight, info, eplison):
        sg, c = info.sg, info.compute
        weight = weight.contiguous()
        buf = info.stage(X.shape[1])
        buf[:sg.n_local].copy_(X)
        S = info.gather(2, buf, eps=eplison)                             # S = eps * A X                :572-603
        ctx.save_for_backward(S, weight)
        ctx.info, ctx.eplison = info, eplison
        return c.mm(S, weight)                                           # :605

    @staticmethod
    def backward(ctx, d_output):
        S, weight = ctx.saved_tensors
        info = ctx.info
        sg, c = info.sg, info.compute
        d_output = d_output.contiguous()
        d_weight = c.mm(S, d_output, ta=True)                            # :710
        d_input = None
        if ctx.needs_input_grad[0]:
            buf = info.stage(weight.shape[0])
            c.mm(d_output, weight, tb=True, out=buf[:sg.n_local])        # Pm = dOut W^T                :711
            d_input = info.gather(2, buf, eps=ctx.eplison)               # :712-738
        return d_input, d_weight, None, None


class _ShardedConv(torch.nn.Module):
    def __init__(self, input_dim, output_dim):
        super().__init__()
        self.weights = torch.nn.Parameter(torch.randn(input_dim, output_dim))      # drawn like gnn_conv.py:83 (same seed, same weights)
        bound = 1.0 / math.sqrt(output_dim)                              # gnn_conv.py:86-88, 136-138
        with torch.no_grad():
            self.weights.uniform_(-bound, bound)


class ShardedGCNConv(_ShardedConv):
    """GCNConv (gnn_conv.py:80-98) on this rank's rows: forward(X_local, info) -> out_local."""

    def forward(self, X, info):
        dout = self.weights.shape[1]
        ld = _pad4(dout)
        W = F.pad(self.weights, (0, ld - dout)) if ld != dout else self.weights
        out = ShardedGNNAFunction.apply(X, W, info)
        return out[:, :dout] if ld != dout else out


class ShardedGINConv(_ShardedConv):
    """GINConv (gnn_conv.py:128-147) on this rank's rows; eps fixed at 0.5 as in the reference."""

    def __init__(self, input_dim, output_dim):
        super().__init__(input_dim, output_dim)
        self.eplison = 0.5

    def forward(self, X, info):
        din = self.weights.shape[0]
        ld = _pad4(din)
        if ld != din:
            X = F.pad(X, (0, ld - din))
            W = F.pad(self.weights, (0, 0, 0, ld - din))
        else:
            W = self.weights
        return ShardedGNNAFunction_GIN.apply(X, W, info, self.eplison)


# ------------------------------------------------------------------------------------------ model + epoch
class ShardedNet(torch.nn.Module):
    """The two models of GNNA_main.py:142-171 on a shard: GCN = conv(in,hid) relu conv(hid,cls); GIN = 5 convs
    (in,hid) (hid,hid)x3 (hid,cls) with relu between; log_softmax at the end."""

    def __init__(self, model, in_dim, hidden, classes):
        super().__init__()
        conv = ShardedGCNConv if model == "gcn" else ShardedGINConv
        dims = [in_dim, hidden, classes] if model == "gcn" else [in_dim] + [hidden] * 4 + [classes]
        self.convs = torch.nn.ModuleList([conv(a, b) for a, b in zip(dims[:-1], dims[1:])])

    def forward(self, x, info):
        for i, c in enumerate(self.convs):
            x = c(x, info)
            if i < len(self.convs) - 1:
                x = F.relu(x)
        return F.log_softmax(x, dim=1)


def broadcast_parameters(module, group=None, src=0):
    """Every rank starts from rank `src`'s weights (they are replicated, each rank draws its own otherwise)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        for prm in module.parameters():
            dist.broadcast(prm.data, dist.get_global_rank(group, src) if group is not None else src, group=group)


def allreduce_gradients(params, group=None):
    """dW = sum over ranks of the per-shard X_r^T G_r: ONE all-reduce over all weight gradients of the model (flattened;
    they are din x dout each, tiny next to the activations)."""
    params = [q for q in params if q.grad is not None]
    if not params or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    flat = torch.cat([q.grad.reshape(-1) for q in params])
    dist.all_reduce(flat, group=group)
    off = 0
    for q in params:
        n = q.grad.numel()
        q.grad.copy_(flat[off:off + n].view_as(q.grad))
        off += n


def train_epoch(model, optimizer, x_local, y_local, info, group=None):
    """One epoch as GNNA_main.py:182-187 (forward, nll_loss, backward, Adam step) on a sharded graph.  The loss is the
    mean over ALL nodes: each rank contributes sum_over_its_rows / N_global, the gradients add up in the all-reduce.
    Returns this rank's share of the loss as a device tensor (no synchronisation)."""
    optimizer.zero_grad()
    out = model(x_local, info)
    loss = F.nll_loss(out, y_local, reduction="sum") / info.sg.num_nodes_global
    loss.backward()
    allreduce_gradients(list(model.parameters()), group)
    optimizer.step()
    return loss.detach()
