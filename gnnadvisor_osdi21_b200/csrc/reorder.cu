// reorder.cu -- community-based vertex renumbering ("Rabbit Order", Arai et al., IPDPS'16), host code.
//
// Replaces the python module `rabbit` of the reference (rabbit_module/src/reorder.cpp:235-295 on top of
// rabbit_order.hpp: aggregate :554-621, merge :477-526, unite :393-449, find_best :455-467,
// compute_perm :623-673), which cannot be built in this image (boost, libnuma, tcmalloc are absent,
// SURVEY.md F12).  Same algorithm, own implementation, no dependencies:
//   1. symmetrise the edge list, drop self loops, merge duplicates into weights;
//   2. visit vertices in ascending degree order; `v` is merged into the neighbouring community `u` with
//      the largest modularity gain  w(v,u) - str(v)*str(u)/(2m)  if that gain is positive, otherwise it
//      becomes a top-level community; merged vertices hang under `u` (child / sibling links), their
//      edges are folded into `u`'s lazily, the next time `u` itself is visited;
//   3. new ids = depth-first walk of every top-level community's merge tree, communities laid out one
//      after the other, so vertices merged together get adjacent ids.
//
// Parallelism.  The reference runs step 2 with optimistic merges under compare-and-swap (:477-526): fast, but
// its permutation differs from run to run and no test can pin it.  Here step 2 is parallel AND
// deterministic: the degree-ordered vertex list is cut into WINDOWS; all host threads fold the edges of a
// window's vertices and pick their best neighbour against the state at the start of the window (the expensive
// part: every edge is traced to its community, sorted and summed), then the window's merges are applied one
// after the other in list order (a few stores per vertex).  A vertex whose choice meanwhile joined the vertex
// itself is carried into the next window (the counterpart of the reference's `pends`).  The result depends on
// the window length only -- never on the number of threads or their timing -- and a window of 1 IS the
// sequential algorithm (GNNA_RABBIT_WINDOW=1).  Step 1 is parallel sorts plus per-vertex folding, step 3 a
// linear walk.
// What is checked: the result is a permutation, applied consistently, identical from run to run and for any
// thread count, and it shortens the average edge span of graphs with community structure as well as the
// sequential order does (tests/test_reorder.py).
#include <algorithm>
#include <parallel/algorithm>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <memory>
#include <numeric>
#include <utility>
#include <vector>

#include <omp.h>

#include "common.h"

namespace {

typedef std::pair<int32_t, float> WEdge;   // (neighbour community, weight)

struct Dendrogram {
    std::vector<std::vector<WEdge>> es;    // aggregated edges of the community rooted at v
    std::vector<int32_t> com;              // parent community (== v for a root), path-compressed
    std::vector<int32_t> child, sibling, united_child;
    std::vector<double> str;               // total weighted degree of the community
    double tot_wgt = 0.0;
};

// Root of v's community.  Called concurrently while a window is evaluated: `com` only changes between windows, so every
// thread finds the same root; the compressing stores write ancestors of the entry they replace (relaxed atomics: any
// interleaving leaves a valid, shorter path).
inline int32_t trace_com(Dendrogram &g, int32_t v)
{
    int32_t r = v, p;
    while ((p = __atomic_load_n(&g.com[r], __ATOMIC_RELAXED)) != r) r = p;
    while ((p = __atomic_load_n(&g.com[v], __ATOMIC_RELAXED)) != r) {
        __atomic_store_n(&g.com[v], r, __ATOMIC_RELAXED);
        v = p;
    }
    return r;
}

// sort by neighbour and add up the weights of equal neighbours
void compact(std::vector<WEdge> &e)
{
    if (e.empty()) return;
    std::sort(e.begin(), e.end(), [](const WEdge &a, const WEdge &b) { return a.first < b.first; });
    size_t o = 0;
    for (size_t i = 1; i < e.size(); i++) {
        if (e[i].first == e[o].first) e[o].second += e[i].second;
        else e[++o] = e[i];
    }
    e.resize(o + 1);
}

// fold the edges of v and of every community merged into v since the last visit into es[v]
// (only v's own data and its children's edge lists are written: vertices of one window never share those)
void unite(Dendrogram &g, int32_t v, std::vector<WEdge> &buf)
{
    buf.clear();
    auto push = [&](int32_t u) {
        for (const WEdge &e : g.es[u]) {
            int32_t c = trace_com(g, e.first);
            if (c != v) buf.push_back(WEdge(c, e.second));   // edges inside the community vanish
        }
        if (u != v) std::vector<WEdge>().swap(g.es[u]);
    };
    push(v);
    for (int32_t w = g.child[v]; w != -1 && w != g.united_child[v]; w = g.sibling[w]) push(w);
    g.united_child[v] = g.child[v];
    compact(buf);
    g.es[v].assign(buf.begin(), buf.end());
}

// neighbour community with the largest positive modularity gain, or v itself (rabbit_order.hpp:455-467)
inline int32_t find_best(const Dendrogram &g, int32_t v)
{
    const double vstr = g.str[v];
    double best_gain = 0.0;
    int32_t best = v;
    for (const WEdge &e : g.es[v]) {
        const double gain = (double)e.second - vstr * g.str[e.first] / g.tot_wgt;
        if (gain > best_gain) { best_gain = gain; best = e.first; }
    }
    return best;
}

double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

// window <= 0: GNNA_RABBIT_WINDOW from the environment, else the built-in choice (a function of num_nodes only)
extern "C" int gnna_rabbit_reorder_host_ex(const int32_t *src, const int32_t *dst, int64_t num_edges,
                                           int64_t num_nodes, int32_t *perm_old_to_new, int64_t window)
{
    GNNA_REQUIRE(num_nodes >= 0 && num_edges >= 0, "rabbit_reorder: negative size");
    GNNA_REQUIRE(num_nodes < 0x7fffffffLL, "rabbit_reorder: too many vertices");
    if (num_nodes == 0) return GNNA_OK;
    GNNA_REQUIRE(perm_old_to_new && (num_edges == 0 || (src && dst)), "rabbit_reorder: null pointer");
    const int32_t n = (int32_t)num_nodes;
    const bool verbose = getenv("GNNA_RABBIT_VERBOSE") != nullptr;
    double t_mark = now_s();
    auto lap = [&](const char *what) {
        if (!verbose) return;
        const double t = now_s();
        fprintf(stderr, "[rabbit] %-28s %8.3f s\n", what, t - t_mark);
        t_mark = t;
    };
    if (window <= 0) {
        const char *w = getenv("GNNA_RABBIT_WINDOW");
        window = w ? atoll(w) : 0;
    }
    if (window <= 0) window = std::min<int64_t>(16384, std::max<int64_t>(1, num_nodes / 64));

    // 1. symmetric weighted adjacency, no self loops, duplicates merged.  Both directions of every edge become a packed
    // key (a << 32 | b); the keys are brought into order by a two-level counting sort like csrc/dataset.cu's: slices of the
    // edge list are histogrammed over row-range buckets, a prefix sum over (bucket, slice) gives every slice private slots
    // (no atomics, deterministic), and the buckets -- ~16 K keys, cache-resident -- are sorted independently; buckets are
    // ascending row ranges, so the array is then sorted as a whole.
    int shift = 0;
    {
        const int64_t want_buckets = std::max<int64_t>(1, std::min<int64_t>(num_edges / 8192 + 1, 1 << 16));
        while ((((int64_t)n - 1) >> shift) + 1 > want_buckets) shift++;
    }
    const int64_t B = (((int64_t)n - 1) >> shift) + 1;
    const int T = std::max(1, omp_get_max_threads());          // slices (handed to whatever team the runtime grants)
    std::vector<int64_t> hist((size_t)T * B, 0);
    long long bad = -1;
#pragma omp parallel for schedule(static, 1) num_threads(T)
    for (int t = 0; t < T; t++) {
        int64_t *h = hist.data() + (size_t)t * B;
        long long my_bad = -1;
        for (int64_t i = num_edges * t / T; i < num_edges * (t + 1) / T; i++) {
            const int32_t a = src[i], b = dst[i];
            if (a < 0 || a >= n || b < 0 || b >= n) { if (my_bad < 0) my_bad = i; continue; }
            if (a == b) continue;                              // self loop: dropped
            h[a >> shift]++;
            h[b >> shift]++;
        }
        if (my_bad >= 0) {
#pragma omp critical
            if (bad < 0 || my_bad < bad) bad = my_bad;
        }
    }
    GNNA_REQUIRE(bad < 0, "rabbit_reorder: vertex id out of range at edge %lld", bad);
    std::vector<int64_t> bucket_begin((size_t)B + 1, 0);
    {
        int64_t run = 0;
        for (int64_t b = 0; b < B; b++) {
            bucket_begin[b] = run;
            for (int t = 0; t < T; t++) {
                const int64_t c = hist[(size_t)t * B + b];
                hist[(size_t)t * B + b] = run;                 // becomes slice t's write cursor in bucket b
                run += c;
            }
        }
        bucket_begin[B] = run;
    }
    const size_t nkeys = (size_t)bucket_begin[B];
    std::unique_ptr<uint64_t[]> keys_mem(new uint64_t[nkeys ? nkeys : 1]);   // first touched by the scatter's threads
    uint64_t *const keys = keys_mem.get();
#pragma omp parallel for schedule(static, 1) num_threads(T)
    for (int t = 0; t < T; t++) {
        int64_t *cur = hist.data() + (size_t)t * B;
        for (int64_t i = num_edges * t / T; i < num_edges * (t + 1) / T; i++) {
            const int32_t a = src[i], b = dst[i];
            if (a == b) continue;
            keys[cur[a >> shift]++] = ((uint64_t)(uint32_t)a << 32) | (uint32_t)b;
            keys[cur[b >> shift]++] = ((uint64_t)(uint32_t)b << 32) | (uint32_t)a;
        }
    }
#pragma omp parallel for schedule(dynamic, 4) num_threads(T)
    for (int64_t b = 0; b < B; b++) std::sort(keys + bucket_begin[b], keys + bucket_begin[b + 1]);
    std::vector<int64_t>().swap(hist);
    lap("edge keys + sort");
    Dendrogram g;
    g.es.resize(n);
    g.com.resize(n);
    g.child.assign(n, -1);
    g.sibling.assign(n, -1);
    g.united_child.assign(n, -1);
    g.str.assign(n, 0.0);
    // first[v] = index of the first key whose source is >= v; every key that starts a new source fills its gap
    std::vector<size_t> first((size_t)n + 1, 0);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)nkeys; i++) {
        const int32_t a = (int32_t)(keys[i] >> 32);
        const int32_t prev = i ? (int32_t)(keys[i - 1] >> 32) : -1;
        for (int32_t v = prev + 1; v <= a; v++) first[v] = (size_t)i;
    }
    {
        const int32_t last = nkeys ? (int32_t)(keys[nkeys - 1] >> 32) : -1;
        for (int32_t v = last + 1; v <= n; v++) first[v] = nkeys;
    }
    double tot = 0.0;
#pragma omp parallel for schedule(dynamic, 1024) reduction(+ : tot)
    for (int32_t v = 0; v < n; v++) {
        g.com[v] = v;
        std::vector<WEdge> &e = g.es[v];
        double s = 0.0;
        for (size_t i = first[v]; i < first[v + 1];) {
            size_t j = i;
            while (j < first[v + 1] && keys[j] == keys[i]) j++;
            e.push_back(WEdge((int32_t)(keys[i] & 0xffffffffu), (float)(j - i)));
            s += (double)(j - i);
            i = j;
        }
        g.str[v] = s;
        tot += s;
    }
    g.tot_wgt = tot;
    keys_mem.reset();
    std::vector<size_t>().swap(first);
    lap("adjacency");

    // 2. incremental aggregation in ascending (unweighted) degree order, window by window
    std::vector<int32_t> order(n);
    std::iota(order.begin(), order.end(), 0);
    __gnu_parallel::stable_sort(order.begin(), order.end(),
                                [&](int32_t a, int32_t b) { return g.es[a].size() < g.es[b].size(); });
    lap("degree order");
    std::vector<int32_t> tops;
    std::vector<int32_t> win, carried, cand;
    win.reserve((size_t)window * 2);
    long long n_windows = 0, n_carried = 0;
    if (window == 1) {                     // the sequential algorithm, without the window machinery
        std::vector<WEdge> buf;
        for (int32_t v : order) {
            unite(g, v, buf);
            const int32_t best = find_best(g, v);
            if (best == v) { tops.push_back(v); continue; }
            g.str[best] += g.str[v];       // v joins best: hang it under best, newest child first
            g.sibling[v] = g.child[best];
            g.child[best] = v;
            g.com[v] = best;
        }
    } else {
        const int threads = std::max(1, omp_get_max_threads());
        std::vector<std::vector<WEdge>> bufs(threads);
        size_t next = 0;
        while (next < (size_t)n || !carried.empty()) {
            win.assign(carried.begin(), carried.end());       // carried vertices keep their place at the front
            carried.clear();
            while (next < (size_t)n && (int64_t)win.size() < window) win.push_back(order[next++]);
            const int64_t wn = (int64_t)win.size();
            cand.resize((size_t)wn);
            // (a) every thread: fold and choose against the state at the start of the window
            // (a short window is not worth waking the team for: same result, the calling thread does it)
#pragma omp parallel for schedule(dynamic, 16) num_threads(threads) if (wn >= 256)
            for (int64_t i = 0; i < wn; i++) {
                unite(g, win[i], bufs[omp_get_thread_num()]);
                cand[i] = find_best(g, win[i]);
            }
            // (b) in list order: apply
            for (int64_t i = 0; i < wn; i++) {
                const int32_t v = win[i];
                if (cand[i] == v) { tops.push_back(v); continue; }
                const int32_t r = trace_com(g, cand[i]);       // the choice may have joined a community in this window
                if (r == v) {                                   // ... v's own: look again with its edges folded in
                    carried.push_back(v);
                    n_carried++;
                    continue;
                }
                g.str[r] += g.str[v];
                g.sibling[v] = g.child[r];
                g.child[r] = v;
                g.com[v] = r;
            }
            n_windows++;
        }
    }
    if (verbose) fprintf(stderr, "[rabbit] window %lld: %lld windows, %lld vertices carried over, %zu top-level communities\n",
                         (long long)window, n_windows, n_carried, tops.size());
    lap("aggregation");

    // 3. depth-first numbering of the merge trees, one top-level community after the other
    int32_t next_id = 0;
    std::vector<int32_t> stack;
    auto push_chain = [&](int32_t v) {     // v, its first child, that child's first child, ...
        for (; v != -1; v = g.child[v]) stack.push_back(v);
    };
    for (int32_t t : tops) {
        push_chain(t);
        while (!stack.empty()) {
            const int32_t v = stack.back();
            stack.pop_back();
            perm_old_to_new[v] = next_id++;
            if (g.sibling[v] != -1) push_chain(g.sibling[v]);
        }
    }
    lap("numbering");
    GNNA_REQUIRE(next_id == n, "rabbit_reorder: internal error, numbered %d of %d vertices", next_id, n);
    return GNNA_OK;
}

extern "C" int gnna_rabbit_reorder_host(const int32_t *src, const int32_t *dst, int64_t num_edges,
                                        int64_t num_nodes, int32_t *perm_old_to_new)
{
    return gnna_rabbit_reorder_host_ex(src, dst, num_edges, num_nodes, perm_old_to_new, 0);
}
