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
// The reference runs step 2 with optimistic parallel merges, so its permutation differs from run to run
// and no test pins it; this implementation is sequential and deterministic.  What is
// checked: the result is a permutation, applied consistently, and it shortens the average edge span
// of graphs with community structure (tests/test_reorder.py).
#include <algorithm>
#include <parallel/algorithm>
#include <cstdint>
#include <numeric>
#include <utility>
#include <vector>

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

inline int32_t trace_com(Dendrogram &g, int32_t v)
{
    int32_t r = v;
    while (g.com[r] != r) r = g.com[r];
    while (g.com[v] != r) {                // path compression
        int32_t nx = g.com[v];
        g.com[v] = r;
        v = nx;
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

}  // namespace

extern "C" int gnna_rabbit_reorder_host(const int32_t *src, const int32_t *dst, int64_t num_edges,
                                        int64_t num_nodes, int32_t *perm_old_to_new)
{
    GNNA_REQUIRE(num_nodes >= 0 && num_edges >= 0, "rabbit_reorder: negative size");
    GNNA_REQUIRE(num_nodes < 0x7fffffffLL, "rabbit_reorder: too many vertices");
    if (num_nodes == 0) return GNNA_OK;
    GNNA_REQUIRE(perm_old_to_new && (num_edges == 0 || (src && dst)), "rabbit_reorder: null pointer");
    const int32_t n = (int32_t)num_nodes;

    // 1. symmetric weighted adjacency, no self loops, duplicates merged
    // (the edge list passes are the bulk of the time at 10^8 edges: all host threads, as the reference's OpenMP build)
    std::vector<uint64_t> keys((size_t)num_edges * 2);
    long long bad = -1;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < num_edges; i++) {
        const int32_t a = src[i], b = dst[i];
        if (a < 0 || a >= n || b < 0 || b >= n) {
#pragma omp critical
            if (bad < 0 || i < bad) bad = i;
            keys[2 * i] = keys[2 * i + 1] = ~0ull;
            continue;
        }
        if (a == b) { keys[2 * i] = keys[2 * i + 1] = ~0ull; continue; }     // self loop: dropped below
        keys[2 * i] = ((uint64_t)(uint32_t)a << 32) | (uint32_t)b;
        keys[2 * i + 1] = ((uint64_t)(uint32_t)b << 32) | (uint32_t)a;
    }
    GNNA_REQUIRE(bad < 0, "rabbit_reorder: vertex id out of range at edge %lld", bad);
    __gnu_parallel::sort(keys.begin(), keys.end());
    while (!keys.empty() && keys.back() == ~0ull) keys.pop_back();
    Dendrogram g;
    g.es.resize(n);
    g.com.resize(n);
    std::iota(g.com.begin(), g.com.end(), 0);
    g.child.assign(n, -1);
    g.sibling.assign(n, -1);
    g.united_child.assign(n, -1);
    g.str.assign(n, 0.0);
    for (size_t i = 0; i < keys.size();) {
        size_t j = i;
        while (j < keys.size() && keys[j] == keys[i]) j++;
        const int32_t a = (int32_t)(keys[i] >> 32), b = (int32_t)(keys[i] & 0xffffffffu);
        const float w = (float)(j - i);
        g.es[a].push_back(WEdge(b, w));
        g.str[a] += w;
        g.tot_wgt += w;
        i = j;
    }
    std::vector<uint64_t>().swap(keys);

    // 2. incremental aggregation in ascending (unweighted) degree order
    std::vector<int32_t> order(n);
    std::iota(order.begin(), order.end(), 0);
    __gnu_parallel::stable_sort(order.begin(), order.end(),
                                [&](int32_t a, int32_t b) { return g.es[a].size() < g.es[b].size(); });
    std::vector<int32_t> tops;
    std::vector<WEdge> buf;
    for (int32_t v : order) {
        unite(g, v, buf);
        const double vstr = g.str[v];
        double best_gain = 0.0;
        int32_t best = v;
        for (const WEdge &e : g.es[v]) {
            const double gain = (double)e.second - vstr * g.str[e.first] / g.tot_wgt;
            if (gain > best_gain) { best_gain = gain; best = e.first; }
        }
        if (best == v) {
            tops.push_back(v);
        } else {                           // v joins best: hang it under best, newest child first
            g.str[best] += vstr;
            g.sibling[v] = g.child[best];
            g.child[best] = v;
            g.com[v] = best;
        }
    }

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
    GNNA_REQUIRE(next_id == n, "rabbit_reorder: internal error, numbered %d of %d vertices", next_id, n);
    return GNNA_OK;
}
