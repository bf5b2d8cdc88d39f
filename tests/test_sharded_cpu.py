"""world_size-2 gloo tests (CPU) of the sharded layers and the multi-GPU epoch (gnnadvisor_osdi21_b200/sharded.py):
the gradients of a sharded epoch equal those of the unsharded model, for GCN and GIN, including class counts that are
not a multiple of four (padded through the weights) and a vertex range that owns no halo.

The device operations are injected (`compute=`): here they are the CPU checker (oracle/), on the GPU box they are
libgnna_b200.so (tests/test_sharded_gpu.py runs the same comparison with the CUDA kernels)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from helpers import make_graph
from gnnadvisor_osdi21_b200 import dist as gdist, sharded


class OracleCompute:
    """sharded.CudaCompute's three operations on the CPU checker (test infrastructure)."""

    def mm(self, A, B, ta=False, tb=False, out=None):
        C = torch.from_numpy(oracle.mm(A.numpy(), B.numpy(), trans_a=ta, trans_b=tb))
        if out is None:
            return C
        out.copy_(C)
        return out

    def prescale(self, X, degrees, out):
        out.copy_(X * degrees[:X.shape[0], None])
        return out

    def aggregate(self, sg, mode, x_ext, out, eps, dim_worker, warp_per_block):
        X = np.ascontiguousarray(x_ext.numpy())
        cols, pp, pn = sg.col_idx.numpy(), sg.part_ptr.numpy(), sg.part2node.numpy()
        if mode == 3:       # rows already scaled by n_j: plain sum, then n_i
            got = oracle.aggregate(0, X, cols, None, 1.0, pp, pn)[:sg.n_local] * sg.degrees_ext.numpy()[:sg.n_local, None]
        else:
            got = oracle.aggregate(mode, X, cols, None, eps, pp, pn)[:sg.n_local]
        out.copy_(torch.from_numpy(got))
        return out


def dense_reference(model, rp, ci, X, y, weights, dims):
    """The unsharded model in float64 torch autograd on a dense adjacency: loss and weight gradients."""
    n = len(rp) - 1
    A = torch.zeros(n, n, dtype=torch.float64)
    rows = np.repeat(np.arange(n), np.diff(rp))
    A[rows, ci] = 1.0
    nrm = torch.from_numpy(oracle.degrees(rp)).double()
    Ahat = nrm[:, None] * A * nrm[None, :]
    Ws = [w.clone().double().requires_grad_(True) for w in weights]
    h = torch.from_numpy(X).double()
    for i, W in enumerate(Ws):
        h = Ahat @ (h @ W) if model == "gcn" else (0.5 * (A @ h)) @ W
        if i < len(Ws) - 1:
            h = torch.relu(h)
    loss = torch.nn.functional.nll_loss(torch.log_softmax(h, dim=1), torch.from_numpy(y))
    loss.backward()
    return float(loss), [W.grad for W in Ws]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


CASES = {"gcn": (12, 8, 5), "gin": (10, 8, 3)}     # (in, hidden, classes): 5 / 3 classes and 10 inputs exercise the padding


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, ps = 600, 8
        rp, ci = make_graph("rmat", n, 9000, 81)
        for model, (din, hid, cls) in CASES.items():
            gen = torch.Generator().manual_seed(7)
            X = torch.randn(n, din, generator=gen).numpy() * 0.1          # weights n_i*n_j grow layer over layer (F1)
            y = torch.randint(0, cls, (n,), generator=gen).numpy()
            for ranges in (None, [0, 1, n] if world == 2 else None):      # default cut, and a one-row shard
                sg = gdist.ShardedGraph(torch.from_numpy(rp), torch.from_numpy(ci), ps, device="cpu", ranges=ranges)
                info = sharded.ShardedInputInfo(sg, exchange="nccl", compute=OracleCompute())
                torch.manual_seed(100 + rank)                              # ranks draw different weights ...
                net = sharded.ShardedNet(model, din, hid, cls)
                sharded.broadcast_parameters(net)                          # ... and start from rank 0's
                w0 = [p.detach().clone() for p in net.parameters()]
                opt = torch.optim.Adam(net.parameters(), lr=0.01)
                v0, v1 = sg.ranges[rank], sg.ranges[rank + 1]
                loss = sharded.train_epoch(net, opt, torch.from_numpy(X[v0:v1]), torch.from_numpy(y[v0:v1]), info)
                total = loss.clone()
                dist.all_reduce(total)
                ref_loss, ref_grads = dense_reference(model, rp, ci, X, y, w0, None)
                assert abs(float(total) - ref_loss) <= 1e-4 * abs(ref_loss), (model, float(total), ref_loss)
                for p, g in zip(net.parameters(), ref_grads):
                    err = (p.grad.double() - g).abs().max() / g.abs().max().clamp_min(1e-30)
                    assert float(err) < 2e-4, "%s grad rel err %g (ranges %s)" % (model, float(err), ranges)
                # replicated Adam: every rank holds the same weights after the step
                for p in net.parameters():
                    both = [torch.empty_like(p.data) for _ in range(world)]
                    dist.all_gather(both, p.data)
                    assert all(torch.equal(both[0], b) for b in both)
        ret[rank] = "ok"
    except Exception as e:   # noqa: BLE001
        import traceback
        ret[rank] = "FAIL: %s\n%s" % (e, traceback.format_exc())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [1, 2, 3])
def test_sharded_epoch_gradients_equal_unsharded(world):
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret.get(r) == "ok" for r in range(world)), dict(ret)
