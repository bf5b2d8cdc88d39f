"""NCCL test of the sharded path on >= 2 GPUs: the sharded aggregation equals the single-GPU product
and the oracle.  Skipped on boxes with one GPU (the host logic is covered by tests/test_dist_cpu.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from helpers import assert_close, make_graph, rand_features

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from gnnadvisor_osdi21_b200 import dist as gdist
        n, ps = 3000, 16
        rp, ci = make_graph("rmat", n, 90000, 71)
        deg = oracle.degrees(rp)
        pp, pn = oracle.build_part(ps, rp, exact=True)
        sg = gdist.ShardedGraph(torch.from_numpy(rp).to(dev), torch.from_numpy(ci).to(dev), ps, device=dev).build_tables()
        v0, v1 = sg.ranges[rank], sg.ranges[rank + 1]
        for dim in (64, 41):
            X = rand_features(n, dim, 72 + dim)
            x_ext = sg.new_features(dim)
            sg.local(x_ext).copy_(torch.from_numpy(X[v0:v1]).to(dev))
            for mode in (0, 1, 2):
                got = sg.aggregate(mode, x_ext, eps=0.5).cpu().numpy()
                ref = oracle.aggregate(mode, X, ci, deg, 0.5, pp, pn)[v0:v1]
                terms = oracle.aggregate(mode, np.abs(X), ci, deg, 0.5, pp, pn)[v0:v1]
                assert_close(got, ref, what="rank %d dim %d mode %d" % (rank, dim, mode), terms=terms)
        # NVLink-native exchange (CUDA IPC + push kernel): same results, several steps so that both buffer
        # parities, the flags and the ack handshake are exercised
        dim = 64
        peer = gdist.PeerHalo(sg, dim)
        for step in range(1, 7):
            X = rand_features(n, dim, 500 + step)
            sg.local(peer.features()).copy_(torch.from_numpy(X[v0:v1]).to(dev))
            got = sg.aggregate(1, None, peer=peer).cpu().numpy()
            ref = oracle.aggregate(1, X, ci, deg, 0.5, pp, pn)[v0:v1]
            terms = oracle.aggregate(1, np.abs(X), ci, deg, 0.5, pp, pn)[v0:v1]
            assert_close(got, ref, what="peer halo rank %d step %d" % (rank, step), terms=terms)
        assert peer.error() == 0
        # overlapped step: per-owner sub-shards consumed as their rows land, push on a second stream
        sg.build_owner_shards()
        total = sum(int(s[1].numel()) for s in sg.owner_shards)
        assert total == sg.col_idx.numel() and sg.owner_order[0] == rank
        out = torch.empty(sg.n_local, dim, device=dev)
        # GNNA_GATED=1 (default): ONE kernel whose CTAs wait for a peer's flag inside the kernel (exchange fused into the
        # aggregation); 0: one kernel per owner sub-shard with wait kernels in between
        for gated in ("1", "0", "1"):
            os.environ["GNNA_GATED"] = gated
            for step in range(1, 6):
                X = rand_features(n, dim, 700 + step)
                for mode in (1, 2):
                    sg.write_local(peer, torch.from_numpy(X[v0:v1]).to(dev), prescale=(mode == 1))
                    out.fill_(float("nan"))
                    got = sg.aggregate_overlapped(mode, peer, out).cpu().numpy()
                    ref = oracle.aggregate(mode, X, ci, deg, 0.5, pp, pn)[v0:v1]
                    terms = oracle.aggregate(mode, np.abs(X), ci, deg, 0.5, pp, pn)[v0:v1]
                    assert_close(got, ref, what="overlapped (gated=%s) rank %d step %d mode %d" % (gated, rank, step, mode), terms=terms)
        assert peer.error() == 0
        peer.close()
        ret[rank] = "ok"
    except Exception as e:   # noqa: BLE001
        import traceback
        ret[rank] = "FAIL: %s\n%s" % (e, traceback.format_exc())
    finally:
        dist.destroy_process_group()


def test_sharded_aggregation_nccl():
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret.get(r) == "ok" for r in range(world)), dict(ret)
