"""Exhaustive interleaving check of the halo-exchange protocol of csrc/halo.cu (flags, acknowledgements, buffers
double-buffered by step parity).  The reference has nothing of the kind (single GPU, SURVEY.md 5: "race detection: none");
on the GPU the protocol is exercised by tests/test_dist_gpu.py and the multi-GPU bench, here its DESIGN is checked:
every interleaving of the ranks' stream-ordered actions, for 2 and 3 ranks over several steps, must

  * never let a producer overwrite rows a consumer is still aggregating (the race the acknowledgements exist for),
  * never let a consumer aggregate rows of another step (the race the flags exist for),
  * never deadlock, and never let a rank run more than one step ahead of a peer.

The guards below are the ones in the kernels:
  push, step s, to peer q   waits until  ack[me <- q] + 2 >= s        halo_push_kernel:  while (ack + 2 < step) spin
  aggregate, step s         waits until  flag[me <- p] >= s  for all p   halo_wait_kernel:  while (flag < step) spin
  after aggregating step s  ack[p <- me] = s for all p                   halo_ack_kernel
and a rank issues, per step: begin_step; push to every peer (comm stream); wait + aggregate + ack (compute stream);
the next step's rows are written into the other parity only after this step's push has read them (ev_pushed).
CPU only, pure Python."""
import pytest


def explore(world, steps, ack_gate=True, flag_gate=True, needs=None, max_lead=1):
    """Breadth-first over all reachable states.  Returns (states, violation or None).
    needs[r] = the peers whose rows rank r consumes (default: all of them); a rank waits only for those flags -- as the
    CTAs of the fused kernel do, which wait per owner SEGMENT and not at all for an owner they have no neighbours at.

    Per rank the program is, for s = 1..steps:   PUSH(s, q) for each peer q   then   AGG_BEGIN(s)  AGG_END(s) [= ack]
    The pushes of step s+1 may start as soon as AGG of step s has BEGUN?  No: a rank's next begin_step is issued after
    its aggregation was issued on the same (compute) stream, and the push of step s+1 waits for that begin_step, so
    program order per rank is push(s, *) ; agg(s) ; push(s+1, *) ...  -- except that in the overlapped step the push runs on
    the communication stream CONCURRENTLY with the aggregation of the same step; that is modelled by letting the pushes of
    step s and AGG_BEGIN(s) interleave freely (the aggregation only needs the peers' rows, not this rank's own pushes)."""
    needs = needs or {r: [q for q in range(world) if q != r] for r in range(world)}
    peers = {r: [q for q in range(world) if r in needs[q]] for r in range(world)}      # whom r pushes rows to
    # state: per rank (step being worked on, set of peers already pushed to this step, agg phase 0/1/2);
    #        flag[q][p], ack[p][q] (ack[p][q] = last step q acknowledged to p); buf[q][parity][p] = step whose rows lie there
    init_rank = (1, frozenset(), 0)
    init = (tuple(init_rank for _ in range(world)),
            tuple(tuple(0 for _ in range(world)) for _ in range(world)),       # flag[q][p]
            tuple(tuple(0 for _ in range(world)) for _ in range(world)),       # ack[p][q]
            tuple(tuple(tuple(0 for _ in range(world)) for _ in range(2)) for _ in range(world)))   # buf[q][parity][p]
    seen, frontier = {init}, [init]
    while frontier:
        nxt = []
        for st in frontier:
            ranks, flag, ack, buf = st
            moves = 0
            for r in range(world):
                s, pushed, phase = ranks[r]
                if s > steps:
                    continue
                # -- push of step s to a peer q not yet served
                for q in peers[r]:
                    if q in pushed:
                        continue
                    if ack_gate and not (ack[r][q] + 2 >= s):
                        continue                                   # spinning on the acknowledgement of step s-2
                    qs, _, qphase = ranks[q]
                    if qphase == 1 and (qs & 1) == (s & 1) and qs != s:
                        return len(seen), "rank %d overwrites parity %d of rank %d (step %d) while it aggregates step %d" % (r, s & 1, q, s, qs)
                    nb = [[list(x) for x in b] for b in buf]
                    nb[q][s & 1][r] = s
                    nf = [list(x) for x in flag]
                    nf[q][r] = s
                    nr = list(ranks)
                    nr[r] = (s, pushed | {q}, phase)
                    new = (tuple(nr), tuple(map(tuple, nf)), ack, tuple(tuple(tuple(x) for x in b) for b in nb))
                    moves += 1
                    if new not in seen:
                        seen.add(new)
                        nxt.append(new)
                # -- aggregation of step s begins: needs every peer's flag
                if phase == 0 and (not flag_gate or all(flag[r][p] >= s for p in needs[r])):
                    for p in needs[r]:
                        if buf[r][s & 1][p] != s:
                            return len(seen), "rank %d aggregates step %d over rows of step %d from rank %d" % (r, s, buf[r][s & 1][p], p)
                    nr = list(ranks)
                    nr[r] = (s, pushed, 1)
                    new = (tuple(nr), flag, ack, buf)
                    moves += 1
                    if new not in seen:
                        seen.add(new)
                        nxt.append(new)
                # -- aggregation ends: acknowledge; the step is over once the pushes are out too (ev_pushed)
                if phase == 1:
                    na = [list(x) for x in ack]
                    for p in needs[r]:
                        na[p][r] = s
                    nr = list(ranks)
                    nr[r] = (s, pushed, 2)
                    new = (tuple(nr), flag, tuple(map(tuple, na)), buf)
                    moves += 1
                    if new not in seen:
                        seen.add(new)
                        nxt.append(new)
                if phase == 2 and len(pushed) == len(peers[r]):
                    nr = list(ranks)
                    nr[r] = (s + 1, frozenset(), 0)
                    new = (tuple(nr), flag, ack, buf)
                    moves += 1
                    if new not in seen:
                        seen.add(new)
                        nxt.append(new)
            if moves == 0 and any(rk[0] <= steps for rk in ranks):
                return len(seen), "deadlock at %r" % (ranks,)
            lead = [rk[0] for rk in ranks]
            if max(lead) - min(lead) > max_lead:   # all-to-all: finishing step s needs every peer's rows of step s
                return len(seen), "rank steps drifted apart: %r" % (lead,)
        frontier = nxt
    return len(seen), None


@pytest.mark.parametrize("world,steps", [(2, 6), (3, 4)])
def test_protocol_has_no_race_and_no_deadlock(world, steps):
    states, bad = explore(world, steps)
    assert bad is None, bad
    assert states > 25 * steps             # the search did look at interleavings (163 states for 2 ranks, thousands for 3)


ONE_WAY = {0: [], 1: [0]}                   # rank 1 consumes rank 0's rows, rank 0 consumes nothing: it never waits for a flag
CHAIN = {0: [], 1: [0], 2: [1]}


@pytest.mark.parametrize("needs,world", [(ONE_WAY, 2), (CHAIN, 3)])
def test_one_way_halos_are_held_back_by_the_acknowledgements(needs, world):
    """When the dependence is not all-to-all the flags alone order nothing for the producer: the acknowledgement is what
    keeps it from lapping its consumer (two steps of rows in flight: the two buffer parities; as step counters, which
    advance after the acknowledgement, three per hop)."""
    states, bad = explore(world, 6, needs=needs, max_lead=3 * (world - 1))
    assert bad is None, bad
    _, bad = explore(world, 6, needs=needs, ack_gate=False, max_lead=99)
    # the checker has teeth: without the gate the producer laps its consumer -- seen either as an overwrite under a running
    # aggregation or as an aggregation over rows of a later step
    assert bad is not None and ("overwrites" in bad or "over rows of step" in bad)


def test_the_checker_finds_stale_reads_without_the_flag_gate():
    _, bad = explore(2, 3, flag_gate=False)
    assert bad is not None and "aggregates step" in bad     # no flag wait: rows of another step are consumed
