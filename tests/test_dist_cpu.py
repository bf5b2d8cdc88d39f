"""world_size-2/3 gloo tests (CPU) of the sharding host logic: vertex ranges, local/halo index space,
the halo exchange, and that aggregating the shards reproduces the unsharded result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from helpers import make_graph, rand_features
from gnnadvisor_osdi21_b200 import dist as gdist


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, dim, ps = 1200, 24, 8
        rp, ci = make_graph("rmat", n, 30000, 61)
        X = rand_features(n, dim, 62)
        deg = oracle.degrees(rp)
        sg = gdist.ShardedGraph(torch.from_numpy(rp), torch.from_numpy(ci), ps, device="cpu", dense_halo=0).build_tables()
        v = sg.ranges
        assert v[0] == 0 and v[-1] == n and all(a <= b for a, b in zip(v, v[1:]))
        assert sg.n_local == v[rank + 1] - v[rank]
        # edge balance: no shard holds more than its share + the heaviest row
        share = sg.num_edges_global / world
        assert sg.num_edges_local <= share + (rp[1:] - rp[:-1]).max() + 1
        # halo = exactly the remote neighbours, sorted, each once
        cols = ci[rp[v[rank]]:rp[v[rank + 1]]]
        remote = np.unique(cols[(cols < v[rank]) | (cols >= v[rank + 1])])
        assert np.array_equal(sg.halo_ids.numpy(), remote)
        # local index space round-trips to the global ids
        l2g = np.concatenate([np.arange(v[rank], v[rank + 1]), remote])
        assert np.array_equal(l2g[sg.col_idx.numpy()], cols)
        # exchange: X_ext rows == the global rows they stand for; halo degrees are the GLOBAL degrees
        x_ext = sg.new_features(dim)
        sg.local(x_ext).copy_(torch.from_numpy(X[v[rank]:v[rank + 1]]))
        sg.exchange(x_ext)
        assert np.array_equal(x_ext.numpy(), X[l2g])
        assert np.array_equal(sg.degrees_ext.numpy(), deg[l2g])
        # sharded aggregation (oracle on the local CSR over X_ext) == rows of the unsharded one
        pp, pn = oracle.build_part(ps, rp, exact=True)
        for mode in (0, 1, 2):
            full = oracle.aggregate(mode, X, ci, deg, 0.5, pp, pn)
            loc = oracle.aggregate(mode, x_ext.numpy(), sg.col_idx.numpy(), sg.degrees_ext.numpy(), 0.5,
                                   sg.part_ptr.numpy(), sg.part2node.numpy())[:sg.n_local]
            assert np.array_equal(loc, full[v[rank]:v[rank + 1]]), "mode %d" % mode
        # dense halos: an owner most of whose rows are needed is asked for its whole range (one contiguous block);
        # the aggregation is unchanged, the owner sees "this peer wants all my rows, in order"
        sd = gdist.ShardedGraph(torch.from_numpy(rp), torch.from_numpy(ci), ps, device="cpu", dense_halo=0.05).build_tables()
        assert any(sd.dense_from) and not sd.dense_from[rank]
        hd = sd.halo_ids.numpy()
        assert set(remote).issubset(set(hd)) and np.array_equal(hd, np.unique(hd))
        for q in range(world):
            if sd.dense_from[q]:
                assert sd.recv_counts[q] == v[q + 1] - v[q] and sd.halo_rows_needed[q] <= sd.recv_counts[q]
        flags = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(flags, torch.tensor([int(f) for f in sd.dense_from]))
        for p in range(world):                                   # what I see as a sender == what the receivers decided
            assert bool(flags[p][rank]) == (sd.dense_to[p] and p != rank)
            if sd.dense_to[p]:
                lo = sum(sd.send_counts[:p])
                assert torch.equal(sd.send_idx[lo:lo + sd.n_local], torch.arange(sd.n_local))
        xd = sd.new_features(dim)
        sd.local(xd).copy_(torch.from_numpy(X[v[rank]:v[rank + 1]]))
        sd.exchange(xd)
        l2g_d = np.concatenate([np.arange(v[rank], v[rank + 1]), hd])
        assert np.array_equal(xd.numpy(), X[l2g_d])
        for mode in (0, 1, 2):
            full = oracle.aggregate(mode, X, ci, deg, 0.5, pp, pn)
            loc = oracle.aggregate(mode, xd.numpy(), sd.col_idx.numpy(), sd.degrees_ext.numpy(), 0.5,
                                   sd.part_ptr.numpy(), sd.part2node.numpy())[:sd.n_local]
            assert np.array_equal(loc, full[v[rank]:v[rank + 1]]), "dense halo, mode %d" % mode
        # dW reduction helper
        w = torch.full((3, 2), float(rank + 1))
        gdist.allreduce_weight_grad(w)
        assert float(w[0, 0]) == sum(range(1, world + 1))
        ret[rank] = "ok"
    except Exception as e:   # noqa: BLE001
        import traceback
        ret[rank] = "FAIL: %s\n%s" % (e, traceback.format_exc())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_graph_gloo(world):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret.get(r) == "ok" for r in range(world)), dict(ret)


def test_partition_ranges_balance_edges_not_nodes():
    # one hub holding half the edges: the hub's rank gets few nodes
    deg = np.full(1000, 10); deg[0] = 10000
    rp = torch.from_numpy(np.concatenate([[0], np.cumsum(deg)]).astype(np.int32))
    v = gdist.partition_ranges(rp, 4)
    assert v[0] == 0 and v[-1] == 1000 and v == sorted(v)
    e = [int(rp[v[i + 1]] - rp[v[i]]) for i in range(4)]
    assert max(e) <= 10000 + 10 and v[1] - v[0] < 250
    assert gdist.partition_ranges(rp, 1) == [0, 1000]
    # a per-row cost moves rows away from the ranks that hold many light rows
    w = gdist.partition_ranges(rp, 4, row_weight=40)
    assert w[0] == 0 and w[-1] == 1000 and w == sorted(w)
    rows_e = [v[i + 1] - v[i] for i in range(4)]
    rows_w = [w[i + 1] - w[i] for i in range(4)]
    assert max(rows_w) < max(rows_e) and gdist.default_row_weight(1) == 0 and gdist.default_row_weight(8) == 280
    assert gdist.default_row_weight(8, avg_degree=492) == 70 and gdist.default_row_weight(8, avg_degree=14.5) == 280
