"""Rabbit-Order replacement on the host (csrc/reorder.cu): time and ordering quality against the
window length, on the planted-community graph of tools/locality_experiment.py (communities of 512 vertices hidden by a
random relabelling) and on an R-MAT look-alike.  window 1 = the sequential algorithm.  CPU only.

    OMP_WAIT_POLICY=passive GNNA_RABBIT_VERBOSE=1 python tools/reorder_study.py [nodes] [windows ...]
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnnadvisor_osdi21_b200 import graph, reorder  # noqa: E402


def planted(n, comm=512, deg_in=20, deg_out=5, seed=1):
    rng = np.random.default_rng(seed)
    ids = np.arange(n, dtype=np.int64)
    src = np.concatenate([np.repeat(ids, deg_in), np.repeat(ids, deg_out)])
    dst_in = np.minimum((np.repeat(ids, deg_in) // comm) * comm + rng.integers(0, comm, n * deg_in), n - 1)
    dst = np.concatenate([dst_in, rng.integers(0, n, n * deg_out)])
    keep = src != dst
    hide = rng.permutation(n)
    return hide[src[keep]], hide[dst[keep]]


def span(perm, s, d):
    return float(np.mean(np.abs(perm[s] - perm[d])))


def in_block_fraction(perm, s, d, block=512):
    """share of the edges whose endpoints fall into the same aligned block of `block` new ids (what one L2-resident tile holds)"""
    return float(np.mean(perm[s] // block == perm[d] // block))


def halo_rows(perm, s, d, n, world=8):
    """Rows a rank must receive per aggregation when the relabelled graph is cut into `world` edge-balanced vertex ranges
    (dist.partition_ranges with row weight 0): (mean, max) over the ranks of the number of distinct remote neighbours."""
    ps, pd = perm[np.concatenate([s, d])], perm[np.concatenate([d, s])]          # symmetric
    deg = np.bincount(ps, minlength=n)
    cum = np.concatenate([[0], np.cumsum(deg)])
    cuts = np.searchsorted(cum, [cum[-1] * g // world for g in range(world + 1)])
    cuts[0], cuts[-1] = 0, n
    owner = np.searchsorted(cuts, ps, side="right") - 1
    remote = (pd < cuts[owner]) | (pd >= cuts[owner + 1])
    key = np.unique(owner[remote].astype(np.int64) * n + pd[remote])
    per_rank = np.bincount(key // n, minlength=world)
    return float(per_rank.mean()), int(per_rank.max())


def main():
    args = [int(a) for a in sys.argv[1:]]
    n = args[0] if args else 600000
    windows = args[1:] or [1, 256, 4096, 16384, 65536]
    threads = len(os.sched_getaffinity(0))
    for name, (s, d) in (("planted communities of 512", planted(n)),
                         ("R-MAT look-alike", tuple(t.numpy() for t in graph.stream_pairs(n, 0, 12 * n, kind="rmat", seed=3)))):
        e = torch.from_numpy(np.stack([s, d]).astype(np.int32))
        ident = np.arange(n)
        print("\n%s: %d vertices, %d edge pairs, %d host threads; as given: avg edge span %.0f, same-block share %.3f, "
              "halo rows per rank at 8 ranks mean %.0f / max %d"
              % ((name, n, len(s), threads, span(ident, s, d), in_block_fraction(ident, s, d)) + halo_rows(ident, s, d, n)), flush=True)
        for w in windows:
            t = time.perf_counter()
            perm = reorder.permutation(e, n, window=w).numpy().astype(np.int64)
            dt = time.perf_counter() - t
            assert np.array_equal(np.sort(perm), ident)
            print("  window %6d: %7.2f s   avg edge span %9.0f   same-block share %.3f   halo rows per rank at 8 ranks mean %.0f / max %d"
                  % ((w, dt, span(perm, s, d), in_block_fraction(perm, s, d)) + halo_rows(perm, s, d, n)), flush=True)


if __name__ == "__main__":
    main()
