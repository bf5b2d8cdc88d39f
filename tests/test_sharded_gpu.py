"""The sharded layers and epoch on the CUDA kernels (libgnna_b200.so): gradients of a 2-rank epoch equal the unsharded
float64 model, the sharded aggregation of a generated-per-shard graph equals the oracle on the whole graph.

On a box with >= 2 GPUs: one rank per GPU over NCCL, all three exchanges (all_to_all, NVLink push, overlapped push).
On a ONE-GPU box (the driver's GPUTEST box): both ranks run their kernels on cuda:0 and exchange over gloo (staged
through the host), so the multi-rank path is still exercised with the real kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from helpers import assert_close, make_graph
from test_sharded_cpu import CASES, dense_reference

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ngpu, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    multi = ngpu >= world
    dev = torch.device("cuda", rank if multi else 0)
    torch.cuda.set_device(dev)
    if multi:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gnnadvisor_osdi21_b200 import dist as gdist, graph, sharded
        n, ps = 3000, 16
        rp, ci = make_graph("rmat", n, 90000, 81)
        exchanges = ("nccl", "peer", "overlap") if multi else ("nccl",)
        for model, (din, hid, cls) in CASES.items():
            gen = torch.Generator().manual_seed(7)
            X = torch.randn(n, din, generator=gen).numpy() * 0.05
            y = torch.randint(0, cls, (n,), generator=gen).numpy()
            for ex in exchanges:
                sg = gdist.ShardedGraph(torch.from_numpy(rp).to(dev), torch.from_numpy(ci).to(dev), ps, device=dev)
                info = sharded.ShardedInputInfo(sg, exchange=ex, max_dim=max(din, hid, cls))
                torch.manual_seed(100 + rank)
                net = sharded.ShardedNet(model, din, hid, cls).to(dev)
                sharded.broadcast_parameters(net)
                w0 = [p.detach().cpu().clone() for p in net.parameters()]
                opt = torch.optim.Adam(net.parameters(), lr=0.01)
                v0, v1 = sg.ranges[rank], sg.ranges[rank + 1]
                xl, yl = torch.from_numpy(X[v0:v1]).to(dev), torch.from_numpy(y[v0:v1]).to(dev)
                loss = sharded.train_epoch(net, opt, xl, yl, info)
                total = loss.clone()
                dist.all_reduce(total)
                ref_loss, ref_grads = dense_reference(model, rp, ci, X, y, w0, None)
                assert abs(float(total) - ref_loss) <= 1e-4 * abs(ref_loss), (model, ex, float(total), ref_loss)
                for p, g in zip(net.parameters(), ref_grads):
                    err = (p.grad.double().cpu() - g).abs().max() / g.abs().max().clamp_min(1e-30)
                    assert float(err) < 2e-4, "%s/%s grad rel err %g" % (model, ex, float(err))
                for _ in range(3):                                   # more steps: both buffer parities, acks
                    sharded.train_epoch(net, opt, xl, yl, info)
                info.check()
                info.close()
        # a graph generated shard by shard (no rank holds it) against the oracle on the whole graph
        n2, e2, dim = 20000, 400000, 32
        est = graph.stream_degree_estimate(n2, e2 // 2, device=dev, chunk=1 << 16)
        ranges = gdist.partition_ranges(torch.cat([est.new_zeros(1), torch.cumsum(est, 0)]), world)
        r, c = graph.synth_graph_shard(n2, e2, ranges[rank], ranges[rank + 1], device=dev, chunk=1 << 16)
        sg = gdist.ShardedGraph.from_rows(ranges, r, c, ps, device=dev).build_tables()
        wrp, wci = graph.synth_graph(n2, e2, exact=False)
        assert sg.num_edges_global == int(wrp[-1]) and sg.num_nodes_global == n2
        wrp, wci = wrp.numpy(), wci.numpy()
        Xw = torch.randn(n2, dim, generator=torch.Generator().manual_seed(3)).numpy()
        deg = oracle.degrees(wrp)
        pp, pn = oracle.build_part(ps, wrp, exact=True)
        v0, v1 = ranges[rank], ranges[rank + 1]
        x_ext = sg.new_features(dim)
        sg.local(x_ext).copy_(torch.from_numpy(Xw[v0:v1]).to(dev))
        for mode in (0, 1, 2):
            got = sg.aggregate(mode, x_ext, eps=0.5).cpu().numpy()
            ref = oracle.aggregate(mode, Xw, wci, deg, 0.5, pp, pn)[v0:v1]
            terms = oracle.aggregate(mode, np.abs(Xw), wci, deg, 0.5, pp, pn)[v0:v1]
            assert_close(got, ref, what="per-shard graph rank %d mode %d" % (rank, mode), terms=terms)
        ret[rank] = "ok"
    except Exception as e:   # noqa: BLE001
        import traceback
        ret[rank] = "FAIL: %s\n%s" % (e, traceback.format_exc())
    finally:
        dist.destroy_process_group()


def test_sharded_epoch_and_per_shard_graph_on_cuda():
    ngpu, world = torch.cuda.device_count(), 2
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, ngpu, ret), nprocs=world, join=True)
    assert all(ret.get(r) == "ok" for r in range(world)), dict(ret)
