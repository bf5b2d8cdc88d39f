"""N > 1 arm of bench.py: the same aggregation step on a graph sharded over N GPUs of one node.

One rank per GPU (torchrun), NCCL.  The graph is FIXED (the Reddit look-alike of the N=1 run), so this
is strong scaling: rank r owns a contiguous vertex range holding ~E/N edges (dist.ShardedGraph), one
step = halo exchange of the remote neighbour rows (gather + ONE all_to_all_single over NVLink) + the
local aggregation kernel.  value = E_global * D / max-over-ranks device time.
"""
import json
import os
import time

import torch
import torch.distributed as dist


def run(args, bench):
    from . import _lib, graph, sharded, dist as gdist
    # stdout carries exactly ONE JSON line (rank 0): anything libraries print there (NCCL's version
    # banner does) is sent to stderr instead
    import sys
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)

    D = args.dim
    n_all, e_all, in_dim, hidden, classes, kind = graph.LOOKALIKES[args.workload]
    n_all, e_all = max(2, int(n_all * args.scale)), max(2, int(e_all * args.scale))
    # the config line (and the reference arm's) computes the same weight from the same two numbers
    row_weight = int(os.environ.get("GNNA_ROW_WEIGHT", gdist.default_row_weight(world, e_all / n_all)))
    # graphs that do not fit one GPU are generated SHARD BY SHARD: every rank walks the same counter-based pair stream and
    # keeps the rows it owns (graph.synth_graph_shard); nobody ever holds the whole graph.  Smaller ones (Reddit) keep the
    # exact edge count of the dataset, which needs a global duplicate count: built whole on every GPU, then cut.
    sharded_gen = os.environ.get("GNNA_SHARDED_GEN", "1" if e_all > 400_000_000 else "0") == "1"
    t_gen = time.perf_counter()
    if sharded_gen:
        est = graph.stream_degree_estimate(n_all, e_all // 2, kind=kind, device=device, every=4)
        rp_est = torch.cat([est.new_zeros(1), torch.cumsum(est, 0)])
        ranges = gdist.partition_ranges(rp_est, world, row_weight)
        del est, rp_est
        r, c = graph.synth_graph_shard(n_all, e_all, ranges[rank], ranges[rank + 1], kind=kind, device=device)
        sg = gdist.ShardedGraph.from_rows(ranges, r, c, args.part_size, device=device).build_tables()
        del r, c
        N, E = n_all, sg.num_edges_global
    else:
        gr = graph.lookalike(args.workload, device=device, scale=args.scale)
        rp, ci = gr["row_ptr"], gr["col_idx"]
        N, E = gr["num_nodes"], ci.numel()
        sg = gdist.ShardedGraph(rp, ci, args.part_size, device=device, row_weight=row_weight).build_tables()
        del rp, ci, gr["row_ptr"], gr["col_idx"]
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t_gen
    # features are a function of the GLOBAL row id (graph.stream_features): a rank makes its own rows, the checker any row
    own_ids = torch.arange(sg.v0, sg.v0 + sg.n_local, device=device)
    X_local = torch.cat([graph.stream_features(own_ids[i:i + (1 << 20)], D) for i in range(0, max(sg.n_local, 1), 1 << 20)])
    del own_ids
    torch.cuda.empty_cache()
    out = torch.empty(sg.n_local, D, device=device)
    P_local = sg.part2node.numel()
    compute = sharded.CudaCompute()
    tune = dict(dim_worker=args.dim_worker, warp_per_block=args.warp_per_block)

    # halo exchange: NVLink push kernel over CUDA-IPC mapped peer buffers (csrc/halo.cu); NCCL all_to_all_single if the
    # box does not allow IPC mappings or the double-buffered mapped buffers would not fit beside the graph
    peer, halo_mode = None, "nccl all_to_all_single"
    free_b, _ = torch.cuda.mem_get_info(device)
    ext_bytes = max(sg.n_ext, 1) * D * 4
    big = ext_bytes > 0.12 * free_b                    # papers100M-size shards: no second copy of anything n_ext-sized
    want_peer = os.environ.get("GNNA_HALO", "peer") == "peer"
    if want_peer and 2 * ext_bytes > 0.6 * free_b:
        want_peer, halo_mode = False, "nccl all_to_all_single (mapped peer buffers would need %.0f GB)" % (2 * ext_bytes / 1e9)
    if want_peer:
        try:
            peer = gdist.PeerHalo(sg, D)
            n_dense = bin(peer.dense_mask).count("1")
            halo_mode = ("CUDA-IPC mapped peer buffers over NVLink: copy-engine memcpy of the whole range to %d dense peer(s) "
                         "(gnna_halo_push_ce), push kernel for %d sparse one(s) (gnna_halo_push_f32)" % (n_dense, world - 1 - n_dense))
        except Exception as e:   # noqa: BLE001
            peer, halo_mode = None, "nccl all_to_all_single (peer mapping failed: %s)" % str(e)[:80]
    ok = torch.tensor([1 if peer is not None else 0], device=device)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if int(ok.item()) == 0 and peer is not None:
        peer.close()
        peer = None
        halo_mode = "nccl all_to_all_single (peer mapping failed on another rank)"
    x_ext = sg.new_features(D) if (peer is None or not big) else None       # the all_to_all path's [own | halo] buffer

    def step_nccl():       # pre-scale own rows (the X*W epilogue of a layer), gather + ONE all_to_all, aggregate
        compute.prescale(X_local, sg.degrees_ext, sg.local(x_ext))
        sg.exchange(x_ext)
        compute.aggregate(sg, 3, x_ext, out, 0.5, args.dim_worker, args.warp_per_block)

    def step_peer():       # the same with the NVLink push kernel, no overlap
        compute.prescale(X_local, sg.degrees_ext, sg.local(peer.stage(D)))
        xb = peer.exchange()
        compute.aggregate(sg, 3, xb, out, 0.5, args.dim_worker, args.warp_per_block)
        peer.ack()

    # overlapped step: the push runs on a second stream, per-owner sub-shards are aggregated as their rows land
    overlap = peer is not None and D % 4 == 0 and os.environ.get("GNNA_OVERLAP", "1") == "1"
    if overlap:
        sg.build_owner_shards()

    def step_overlap():
        sg.write_local(peer, X_local, prescale=True)
        sg.aggregate_overlapped(1, peer, out, **tune)

    step_serial = step_peer if peer is not None else step_nccl
    step = step_overlap if overlap else step_serial

    # the overlapped step is ~12 small launches on two streams: replay it from CUDA graphs (one per buffer
    # parity; the step counter lives in device memory) so the host is not the bottleneck at 8 GPUs
    graphs, launches_per_step = None, None
    if overlap and os.environ.get("GNNA_GRAPH", "1") == "1":
        try:
            for _ in range(3):
                step_overlap()
            torch.cuda.synchronize()
            dist.barrier(device_ids=[local])
            base_step = peer.step
            graphs = []
            for i in range(2):
                g = torch.cuda.CUDAGraph()
                c0 = _lib.launch_count()
                with torch.cuda.graph(g):
                    step_overlap()
                launches_per_step = _lib.launch_count() - c0
                graphs.append(g)
            peer.step = base_step                       # nothing ran during capture
            state = {"i": 0}

            def step_graph():
                graphs[state["i"] & 1].replay()
                state["i"] += 1
                peer.step += 1
            step = step_graph
        except Exception as e:   # noqa: BLE001
            graphs = None
            halo_mode += " (graph capture failed: %s)" % str(e)[:60]
            step = step_overlap

    # ---- correctness inside the bench: (1) the exchange implementations agree, (2) sampled rows against the CPU oracle
    halo_check, halo_diff_elems = None, None
    if peer is not None and x_ext is not None:
        step_nccl()
        ref_out = out.clone()
        step_serial()
        step()
        step()
        step()
        halo_check = ((out - ref_out).abs().max() / ref_out.abs().max().clamp_min(1e-30)).item()
        halo_diff_elems = int((out != ref_out).sum().item())
        out.zero_()                                    # and once more from a zeroed output: the step must rewrite it
        step()
        torch.cuda.synchronize()
        halo_check = max(halo_check, ((out - ref_out).abs().max() / ref_out.abs().max().clamp_min(1e-30)).item())
        assert halo_check < 1e-4, "overlapped / peer exchange disagrees with the NCCL path: %g" % halo_check
        del ref_out
    step()
    torch.cuda.synchronize()
    oracle_err = oracle_check_rows(bench, graph, sg, out, D, args.part_size)
    errs = [torch.zeros(1, dtype=torch.float64, device=device) for _ in range(world)]
    dist.all_gather(errs, torch.tensor([oracle_err], dtype=torch.float64, device=device))
    oracle_err = max(float(e.item()) for e in errs)
    assert oracle_err <= 1e-4, "sharded aggregation disagrees with the CPU oracle: %g" % oracle_err

    def kernel_only():     # the local kernel on the step's buffer as it stands (own + halo rows, pre-scaled)
        xb = peer.features(peer.step) if peer is not None else x_ext
        compute.aggregate(sg, 3, xb, out, 0.5, args.dim_worker, args.warp_per_block)

    def exchange_only():
        if peer is not None:
            peer.stage(D)
            peer.exchange()
            peer.ack()
        else:
            sg.exchange(x_ext)

    def reduce_max(ms):
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    barrier = lambda: dist.barrier(device_ids=[local])   # noqa: E731
    sampler = bench.ClockSampler(local) if rank == 0 else None
    step(); torch.cuda.synchronize()
    _lib.launch_count(reset=True)
    if sampler:
        sampler.start()
    ms = reduce_max(bench.timed(step, args.steps, args.warmup, barrier)) / args.steps
    clocks = sampler.stop() if sampler else None
    launches = _lib.launch_count() * args.steps // (args.steps + args.warmup)
    if graphs is not None:
        launches = launches_per_step * args.steps
    k = max(3, args.steps // 4)
    ms_kernel = reduce_max(bench.timed(kernel_only, k, 3, barrier)) / k
    ms_exch = reduce_max(bench.timed(exchange_only, k, 3, barrier)) / k
    ms_serial = reduce_max(bench.timed(step_serial, k, 3, barrier)) / k if overlap else ms

    # ---- end to end: this rank's features start and end in pinned host memory.  One step = H2D of its input, the sharded
    # aggregation, D2H of its result.  The three stages of consecutive steps run side by side (copy engines + SMs): replay k
    # uploads the input of step k+1, aggregates step k and downloads the result of step k-1.  With the peer exchange the
    # whole replay is ONE CUDA graph per buffer parity, so the host issues one launch per step.
    # (papers100M-size shards: 7 GB of features per rank -- three pinned copies of that on each of 8 ranks would not fit the
    # host; the end-to-end leg then moves the first 1/16 of the rank's rows per step and says so)
    e2e_rows = sg.n_local if not big else max(1, sg.n_local // 16)
    x_host = X_local[:e2e_rows].cpu().pin_memory()
    out_host = [torch.empty(e2e_rows, D).pin_memory() for _ in range(2)]
    x_stage = [torch.empty(sg.n_local, D, device=device) for _ in range(2)] if not big else [X_local, X_local]
    o_stage = [torch.empty(sg.n_local, D, device=device) for _ in range(2)] if not big else [out, torch.empty_like(out)]
    s_in, s_out = torch.cuda.Stream(device), torch.cuda.Stream(device)

    def e2e_body(i):
        """One pipelined replay with staging parity i; forks to the copy streams and joins back."""
        cur = torch.cuda.current_stream()
        ev0 = torch.cuda.Event(); ev0.record(cur)
        s_in.wait_event(ev0); s_out.wait_event(ev0)
        with torch.cuda.stream(s_in):                          # input of the NEXT step
            x_stage[i ^ 1][:e2e_rows].copy_(x_host, non_blocking=True)
        with torch.cuda.stream(s_out):                         # result of the PREVIOUS step
            out_host[i ^ 1].copy_(o_stage[i ^ 1][:e2e_rows], non_blocking=True)
        if overlap:
            sg.write_local(peer, x_stage[i], prescale=True)
            sg.aggregate_overlapped(1, peer, o_stage[i], **tune)
        elif peer is not None:
            compute.prescale(x_stage[i], sg.degrees_ext, sg.local(peer.stage(D)))
            xb = peer.exchange()
            compute.aggregate(sg, 3, xb, o_stage[i], 0.5, args.dim_worker, args.warp_per_block)
            peer.ack()
        else:
            compute.prescale(x_stage[i], sg.degrees_ext, sg.local(x_ext))
            sg.exchange(x_ext)
            compute.aggregate(sg, 3, x_ext, o_stage[i], 0.5, args.dim_worker, args.warp_per_block)
        e1, e2 = torch.cuda.Event(), torch.cuda.Event()
        e1.record(s_in); e2.record(s_out)
        cur.wait_event(e1); cur.wait_event(e2)

    if not big:
        for b in range(2):
            x_stage[b].copy_(X_local)
    e2e_graphs, e2e_state = None, {"i": 0}
    if overlap and graphs is not None:
        try:
            for _ in range(2):
                e2e_body(e2e_state["i"] & 1); e2e_state["i"] += 1
            torch.cuda.synchronize()
            dist.barrier(device_ids=[local])
            base_step = peer.step
            e2e_graphs = []
            for i in range(2):                                 # staging parity follows the exchange-buffer parity
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    e2e_body((e2e_state["i"] + i) & 1)
                e2e_graphs.append(g)
            peer.step = base_step
        except Exception as e:   # noqa: BLE001
            e2e_graphs = None
            halo_mode += " (e2e graph capture failed: %s)" % str(e)[:60]

    def e2e_step():
        if e2e_graphs is not None:
            e2e_graphs[e2e_state["k"] & 1].replay()
            peer.step += 1
        else:
            e2e_body(e2e_state["i"] & 1)
        e2e_state["i"] += 1
        e2e_state["k"] = e2e_state.get("k", 0) + 1
    e2e_state["k"] = 0
    ke = max(4, min(args.steps, 30))
    ms_e2e = reduce_max(bench.timed(e2e_step, ke, 4, barrier)) / ke
    # the pipeline delivers step k's result during replay k+1: drain it and compare with the device-resident run
    e2e_step()
    torch.cuda.synchronize()
    last = (e2e_state["i"] - 2) & 1
    e2e_diff = ((out_host[last].to(device) - out[:e2e_rows]).abs().max() / out.abs().max().clamp_min(1e-30)).item()

    # ---- GCN epoch (BASELINE.json's second metric) on the sharded layers: forward + backward + Adam, GNNA_main.py:142-202
    epoch = None
    if not args.no_extras and not big:
        try:
            epoch = sharded_epoch_ms(args, sg, in_dim, hidden, classes, device, local, bench,
                                     "overlap" if overlap else ("peer" if peer is not None else "nccl"), reduce_max, barrier)
        except Exception as e:   # noqa: BLE001
            epoch = {"error": str(e)[:200]}

    halo = sg.halo_bytes(D)
    stats = torch.tensor([sg.num_edges_local, sg.n_local, sg.n_halo, P_local, halo["recv"], halo["send"], sum(sg.halo_rows_needed)],
                         device=device, dtype=torch.float64)
    allstats = [torch.zeros_like(stats) for _ in range(world)]
    dist.all_gather(allstats, stats)
    if rank == 0:
        peak, peak_src = bench.peak_gbs()
        # roofline of the local kernel on the most loaded rank
        per_rank = [[float(v) for v in s.cpu()] for s in allstats]
        worst = max(per_rank, key=lambda s: s[0])
        B = bench.alg_bytes(int(worst[0]), int(worst[1]), D, int(worst[3]))
        achieved = B / (ms_kernel * 1e-3) / 1e9
        line = {"metric": "aggregation throughput (GCN SpMM), edges*dim/s", "value": E * D / (ms * 1e-3), "unit": "edge*dim/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": bench.config_of(args, N, E, int(sum(s[3] for s in per_rank)), world=world),
                "roofline": bench.roofline_of(B, ms_kernel, int(worst[2] + worst[1]) * D * 4,
                                              kernel="gnna::aggregate_kernel<float,4,16,1,false> on the most loaded shard",
                                              note="per-GPU kernel, exchange excluded; max over ranks of the kernel-only time"),
                "e2e": {"value": E * D / (ms_e2e * 1e-3), "unit": "edge*dim/s", "ms_per_step": ms_e2e,
                        "h2d_bytes_per_step": int(sum(s[1] for s in per_rank) * D * 4) // (16 if big else 1),
                        "d2h_bytes_per_step": int(sum(s[1] for s in per_rank) * D * 4) // (16 if big else 1),
                        "rows_moved": "all" if not big else "the first 1/16 of every rank's rows (host memory)",
                        "cuda_graph": e2e_graphs is not None, "max_rel_diff_vs_device_run": e2e_diff,
                        "note": "per rank: pinned host rows -> H2D -> pre-scale + halo exchange + aggregation -> D2H, every step; "
                                "the copies of neighbouring steps overlap the aggregation (two copy streams)"},
                "gpu_launches": int(launches), "clocks": clocks, "impl": "ours",
                "extras": {"ms_kernel_only": ms_kernel, "ms_exchange_only": ms_exch, "ms_step_without_overlap": ms_serial,
                           "overlap": bool(overlap), "cuda_graph": graphs is not None,
                           "halo_exchange": halo_mode + ("; per-owner sub-shards aggregated while the rest is in flight" if overlap else ""),
                           "graph_built": ("shard by shard from the counter-based pair stream (no rank holds the graph)" if sharded_gen
                                           else "whole on every GPU, then cut (exact dataset edge count)"),
                           "graph_build_s": t_gen,
                           "peer_vs_nccl_max_rel_diff": halo_check, "peer_vs_nccl_differing_elements": halo_diff_elems,
                           "oracle_check_max_rel_err": oracle_err,
                           "oracle_check": "up to 2048 rows per rank, chosen at random, recomputed by the CPU oracle from features "
                                           "regenerated from their GLOBAL ids (max over ranks, tolerance 1e-4)",
                           "gcn_epoch_ms": epoch,
                           "shards": [{"edges": int(s[0]), "rows": int(s[1]), "halo_rows": int(s[2]), "halo_rows_needed": int(s[6]),
                                       "halo_recv_bytes": int(s[4]), "halo_send_bytes": int(s[5])} for s in per_rank]}}
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if peer is not None:
        err = peer.error()
        peer.close()
        assert err == 0, "halo wait timed out (%d)" % err
    dist.barrier(device_ids=[local])
    dist.destroy_process_group()


def oracle_check_rows(bench, graph, sg, out, D, part_size, rows=2048, seed=5):
    """Recompute up to `rows` randomly chosen rows of this rank's GCN aggregation with the CPU oracle (the checker, never
    the thing measured) from RAW features regenerated from the neighbours' GLOBAL ids, and return the largest relative
    error.  Checks the kernel, the local/halo index space, the exchanged degrees and that the halo rows that arrived are
    the rows that were asked for -- at every graph size, because only the sampled rows' neighbourhoods go to the host."""
    import numpy as np
    oracle = bench._oracle()
    dev = sg.device
    n_local = sg.n_local
    if n_local == 0:
        return 0.0
    g = torch.Generator().manual_seed(seed + sg.rank)
    pick = torch.sort(torch.randperm(n_local, generator=g)[:rows])[0].to(dev)
    rp = sg.row_ptr.to(torch.int64)
    beg, end = rp[pick], rp[pick + 1]
    lens = end - beg
    off = torch.zeros(pick.numel() + 1, dtype=torch.int64, device=dev)
    off[1:] = torch.cumsum(lens, 0)
    total = int(off[-1])
    pos = torch.arange(total, device=dev) - torch.repeat_interleave(off[:-1], lens) + torch.repeat_interleave(beg, lens)
    cols_ext = sg.col_idx[pos].to(torch.int64)                                   # ids in [own rows | halo rows]
    uniq, inv = torch.unique(cols_ext, return_inverse=True)
    glob = torch.where(uniq < n_local, uniq + sg.v0, sg.halo_ids[(uniq - n_local).clamp_min(0)])
    k = pick.numel()
    X = torch.cat([graph.stream_features(pick + sg.v0, D), graph.stream_features(glob, D)]).cpu().numpy()
    deg = torch.cat([sg.degrees_ext[pick], sg.degrees_ext[uniq]]).cpu().numpy()
    sub_rp = np.concatenate([off.cpu().numpy(), np.full(len(uniq), total)]).astype(np.int32)   # neighbour rows have no edges
    sub_ci = (inv + k).cpu().numpy().astype(np.int32)
    pp, pn = oracle.build_part(part_size, sub_rp, exact=True)
    ref = oracle.aggregate(1, X, sub_ci, deg, 1.0, pp, pn, threads=bench.host_threads())[:k].astype(np.float64)
    terms = oracle.aggregate(1, np.abs(X), sub_ci, deg, 1.0, pp, pn, threads=bench.host_threads())[:k].astype(np.float64)
    got = out[pick].cpu().numpy().astype(np.float64)
    scale = np.maximum(np.maximum(np.abs(ref), 1e-3 * np.abs(ref).max()), 0.1 * terms)
    return float((np.abs(got - ref) / np.maximum(scale, 1e-30)).max())


def sharded_epoch_ms(args, sg, in_dim, hidden, classes, device, local, bench, exchange, reduce_max, barrier):
    """2-layer GCN epoch (forward + backward + ONE gradient all-reduce + Adam) on the sharded layers."""
    from . import graph, sharded
    info = sharded.ShardedInputInfo(sg, dimWorker=args.dim_worker, warpPerBlock=args.warp_per_block,
                                    exchange=exchange, max_dim=max(hidden, classes))
    own = torch.arange(sg.v0, sg.v0 + sg.n_local, device=device)
    x = torch.cat([graph.stream_features(own[i:i + (1 << 20)], in_dim, seed=77) for i in range(0, max(sg.n_local, 1), 1 << 20)])
    y = torch.ones(sg.n_local, dtype=torch.long, device=device)                   # dataset.py:136
    torch.manual_seed(20213)
    net = sharded.ShardedNet("gcn", in_dim, hidden, classes).to(device)
    sharded.broadcast_parameters(net)
    opt = torch.optim.Adam(net.parameters(), lr=0.01)

    def train():
        sharded.train_epoch(net, opt, x, y, info)
    k = max(3, min(args.steps // 5, 20))
    ms = reduce_max(bench.timed(train, k, 3, barrier)) / k
    info.check()
    info.close()
    return {"ms": ms, "epochs_timed": k, "exchange": exchange,
            "model": "GCN %d-%d-%d sharded over %d GPUs: fwd+bwd (4 halo exchanges) + gradient all-reduce + Adam (GNNA_main.py:142-202)"
                     % (in_dim, hidden, classes, sg.world)}
