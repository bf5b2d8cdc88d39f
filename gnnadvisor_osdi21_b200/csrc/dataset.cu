// dataset.cu -- the input side of the hot path on the host: edge-list files and COO -> CSR, on all host threads.
//
// Replaces the loader of the reference (GNNAdvisor/dataset.py:55-122): a Python loop over the lines of a
// whitespace `src dst` text file (:62-70: two list appends and two set inserts per edge), `[1] * E` value lists
// and scipy's single-threaded coo_matrix(...).tocsr() (:108-111).  At 10^8 edges that is minutes before the
// first kernel runs.  Same results, bit for bit:
//   * the text parser returns the edges in file order, num_nodes = largest id + 1 (:72);
//   * the CSR has duplicate edges merged and the columns of a row ascending -- what scipy's coo -> csr followed
//     by sum_duplicates gives and what the kernels assume (self loops stay in, like scipy keeps them).
// Host code only (no kernel in this file); the GPU builder for graphs that are generated on the device is
// graph.csr_from_edges on CUDA tensors.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#include <omp.h>

#include "common.h"

namespace {

struct MappedFile {
    const char *data = nullptr;
    size_t size = 0;
    int fd = -1;
    ~MappedFile()
    {
        if (data && size) munmap(const_cast<char *>(data), size);
        if (fd >= 0) close(fd);
    }
    // 0 ok, -1 cannot open, -2 cannot map
    int open_ro(const char *path)
    {
        fd = ::open(path, O_RDONLY);
        if (fd < 0) return -1;
        struct stat st;
        if (fstat(fd, &st) != 0) return -1;
        size = (size_t)st.st_size;
        if (size == 0) return 0;
        void *p = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (p == MAP_FAILED) { data = nullptr; return -2; }
        data = static_cast<const char *>(p);
        madvise(p, size, MADV_SEQUENTIAL);
        return 0;
    }
};

inline bool is_blank(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

// One decimal integer with an optional sign at p (python's int() of one token); false if the token is anything else.
inline bool parse_int(const char *&p, const char *end, int64_t &v)
{
    bool neg = false;
    if (p < end && (*p == '-' || *p == '+')) { neg = (*p == '-'); p++; }
    if (p >= end || *p < '0' || *p > '9') return false;
    uint64_t a = 0;
    int digits = 0;
    while (p < end && *p >= '0' && *p <= '9') {
        a = a * 10 + (uint64_t)(*p - '0');
        if (++digits > 18) return false;                   // beyond int64: not a vertex id
        p++;
    }
    if (p < end && !is_blank(*p) && *p != '\n') return false;   // "12abc", "1.5"
    v = neg ? -(int64_t)a : (int64_t)a;
    return true;
}

// Lines [begin, end) of the file (begin is the first byte of a line): appends (src, dst) per edge line.
// Lines that are empty, or start with '#' or '%' (SNAP / MatrixMarket comments), are skipped; tokens after the second
// are ignored (weights).  Returns the byte offset of the first malformed line, or -1.
long long parse_range(const char *base, size_t begin, size_t end, std::vector<int64_t> &out)
{
    const char *p = base + begin, *stop = base + end;
    while (p < stop) {
        const char *line = p;
        const char *nl = static_cast<const char *>(memchr(p, '\n', (size_t)(stop - p)));
        const char *eol = nl ? nl : stop;
        while (p < eol && is_blank(*p)) p++;
        if (p < eol && *p != '#' && *p != '%') {
            int64_t a, b;
            if (!parse_int(p, eol, a)) return (long long)(line - base);
            while (p < eol && is_blank(*p)) p++;
            if (!parse_int(p, eol, b)) return (long long)(line - base);
            out.push_back(a);
            out.push_back(b);
        }
        p = nl ? nl + 1 : stop;
    }
    return -1;
}

// first byte of the first line that starts at or after `pos`
size_t line_start_at_or_after(const char *base, size_t size, size_t pos)
{
    if (pos == 0) return 0;
    if (pos >= size) return size;
    const char *nl = static_cast<const char *>(memchr(base + pos - 1, '\n', size - (pos - 1)));
    return nl ? (size_t)(nl - base) + 1 : size;
}

}  // namespace

extern "C" int gnna_edge_text_scan(const char *path_host, int64_t *max_edges_host)
{
    GNNA_REQUIRE(path_host && max_edges_host, "edge_text_scan: null pointer");
    MappedFile f;
    const int rc = f.open_ro(path_host);
    GNNA_REQUIRE(rc == 0, "edge_text_scan: cannot %s %s", rc == -1 ? "open" : "map", path_host);
    int64_t lines = 0;
    const int64_t size = (int64_t)f.size;
    const int64_t block = 1 << 22;
#pragma omp parallel for schedule(dynamic) reduction(+ : lines)
    for (int64_t b = 0; b < size; b += block) {
        const char *p = f.data + b, *stop = f.data + std::min(size, b + block);
        int64_t n = 0;
        while (p < stop) {
            const char *nl = static_cast<const char *>(memchr(p, '\n', (size_t)(stop - p)));
            if (!nl) break;
            n++;
            p = nl + 1;
        }
        lines += n;
    }
    if (size > 0 && f.data[size - 1] != '\n') lines++;     // last line without a newline
    *max_edges_host = lines;
    return GNNA_OK;
}

extern "C" int gnna_edge_text_parse(const char *path_host, int64_t *src_host, int64_t *dst_host, int64_t capacity,
                                    int64_t *num_edges_host, int64_t *num_nodes_host)
{
    GNNA_REQUIRE(path_host && num_edges_host && num_nodes_host, "edge_text_parse: null pointer");
    GNNA_REQUIRE(capacity >= 0 && (capacity == 0 || (src_host && dst_host)), "edge_text_parse: bad output buffers");
    MappedFile f;
    const int rc = f.open_ro(path_host);
    GNNA_REQUIRE(rc == 0, "edge_text_parse: cannot %s %s", rc == -1 ? "open" : "map", path_host);
    const int threads = std::max(1, omp_get_max_threads());
    // more pieces than threads: lines differ in length, the pieces do not
    const int pieces = (int)std::min<size_t>((size_t)threads * 8, f.size / (1 << 16) + 1);
    std::vector<std::vector<int64_t>> part(pieces);
    std::vector<long long> bad(pieces, -1);
#pragma omp parallel for schedule(dynamic)
    for (int i = 0; i < pieces; i++) {
        const size_t b = line_start_at_or_after(f.data, f.size, f.size * (size_t)i / pieces);
        const size_t e = line_start_at_or_after(f.data, f.size, f.size * (size_t)(i + 1) / pieces);
        if (b < e) {
            part[i].reserve((e - b) / 6);
            bad[i] = parse_range(f.data, b, e, part[i]);
        }
    }
    std::vector<int64_t> first(pieces + 1, 0);
    for (int i = 0; i < pieces; i++) {
        if (bad[i] >= 0) {
            // the line number is only wanted on this path: count the newlines before the offending line
            long long line = 1;
            for (long long k = 0; k < bad[i]; k++) line += (f.data[k] == '\n');
            return gnna::fail(GNNA_ERR_INVALID, "edge_text_parse: %s line %lld is not `src dst` (two integers)", path_host, line);
        }
        first[i + 1] = first[i] + (int64_t)(part[i].size() / 2);
    }
    const int64_t total = first[pieces];
    GNNA_REQUIRE(total <= capacity, "edge_text_parse: %lld edges, room for %lld", (long long)total, (long long)capacity);
    int64_t max_id = -1;
#pragma omp parallel for schedule(dynamic) reduction(max : max_id)
    for (int i = 0; i < pieces; i++) {
        const std::vector<int64_t> &v = part[i];
        int64_t *s = src_host + first[i], *d = dst_host + first[i];
        for (size_t k = 0; k < v.size() / 2; k++) {
            s[k] = v[2 * k];
            d[k] = v[2 * k + 1];
            max_id = std::max(max_id, std::max(s[k], d[k]));
        }
    }
    *num_edges_host = total;
    *num_nodes_host = max_id + 1;                          // dataset.py:72  max(self.nodes) + 1
    return GNNA_OK;
}

// COO -> CSR with duplicates merged and columns ascending.  Row ranges of equal width are the buckets of a two-level
// counting sort: (1) the edge list is cut into slices, each histogrammed over the buckets, (2) a prefix sum over
// (bucket, slice) gives each slice a private range inside each bucket, so the scatter needs no atomics and is
// deterministic, (3) buckets are sorted as packed (row << 32 | col) keys and made unique independently -- a bucket is a few
// 10^4 keys, cache-resident -- and (4) the unique keys are copied to their final offsets.
extern "C" int gnna_csr_from_edges_host(const int64_t *src_host, const int64_t *dst_host, int64_t num_edges,
                                        int64_t num_nodes, int32_t *row_ptr_host, int32_t *col_idx_host,
                                        int64_t *nnz_host)
{
    GNNA_REQUIRE(num_edges >= 0 && num_nodes >= 0, "csr_from_edges: negative size");
    GNNA_REQUIRE(num_nodes < 0x7fffffffLL, "csr_from_edges: %lld nodes exceed the int32 CSR contract", (long long)num_nodes);
    GNNA_REQUIRE(row_ptr_host && nnz_host && (num_edges == 0 || (src_host && dst_host && col_idx_host)),
                 "csr_from_edges: null pointer");
    const int64_t N = num_nodes, E = num_edges;
    if (E == 0) {
        for (int64_t i = 0; i <= N; i++) row_ptr_host[i] = 0;
        *nnz_host = 0;
        return GNNA_OK;
    }
    GNNA_REQUIRE(N > 0, "csr_from_edges: edge endpoint outside [0, 0)");
    int shift = 0;                                         // rows per bucket = 1 << shift; about 16 K edges per bucket
    {
        const int64_t want_buckets = std::max<int64_t>(1, std::min<int64_t>(E / 16384 + 1, 1 << 16));
        while (((N - 1) >> shift) + 1 > want_buckets) shift++;
    }
    const int64_t B = ((N - 1) >> shift) + 1;
    // The edge list is cut into T slices, one histogram each.  T is a number of SLICES, not a promise about the team: the
    // loops below hand slices to whatever threads the runtime grants (all regions ask for the same team size: libgomp
    // re-docks its pool, a spin-wait of ~100 ms in a sandboxed container, whenever the size changes).
    const int T = std::max(1, omp_get_max_threads());
    std::vector<int64_t> hist((size_t)T * B, 0);
    long long bad = -1;
#pragma omp parallel for schedule(static, 1) num_threads(T)
    for (int t = 0; t < T; t++) {
        const int64_t e0 = E * t / T, e1 = E * (t + 1) / T;
        int64_t *h = hist.data() + (size_t)t * B;
        long long my_bad = -1;
        for (int64_t i = e0; i < e1; i++) {
            const int64_t r = src_host[i], c = dst_host[i];
            if (r < 0 || r >= N || c < 0 || c >= N) { if (my_bad < 0) my_bad = i; continue; }
            h[r >> shift]++;
        }
        if (my_bad >= 0) {
#pragma omp critical
            if (bad < 0 || my_bad < bad) bad = my_bad;
        }
    }
    // scipy's coo_matrix raises here too ("row index exceeds matrix dimensions")
    GNNA_REQUIRE(bad < 0, "csr_from_edges: edge endpoint outside [0, %lld) at edge %lld", (long long)N, bad);
    std::vector<int64_t> bucket_begin((size_t)B + 1, 0);
    {
        int64_t run = 0;
        for (int64_t b = 0; b < B; b++) {
            bucket_begin[b] = run;
            for (int t = 0; t < T; t++) {
                const int64_t c = hist[(size_t)t * B + b];
                hist[(size_t)t * B + b] = run;             // becomes thread t's write cursor in bucket b
                run += c;
            }
        }
        bucket_begin[B] = run;
    }
    std::unique_ptr<uint64_t[]> keys_mem(new uint64_t[(size_t)E]);   // not value-initialised: first touched by the scatter's threads
    uint64_t *const keys = keys_mem.get();
#pragma omp parallel for schedule(static, 1) num_threads(T)
    for (int t = 0; t < T; t++) {
        const int64_t e0 = E * t / T, e1 = E * (t + 1) / T;
        int64_t *cur = hist.data() + (size_t)t * B;
        for (int64_t i = e0; i < e1; i++) {
            const uint64_t r = (uint64_t)src_host[i], c = (uint64_t)dst_host[i];
            keys[(size_t)cur[r >> shift]++] = (r << 32) | c;
        }
    }
    std::vector<int64_t> uniq((size_t)B + 1, 0);           // unique keys per bucket, then their exclusive prefix sum
#pragma omp parallel for schedule(dynamic, 4) num_threads(T)
    for (int64_t b = 0; b < B; b++) {
        uint64_t *k0 = keys + bucket_begin[b], *k1 = keys + bucket_begin[b + 1];
        std::sort(k0, k1);
        uniq[b] = std::unique(k0, k1) - k0;
    }
    {
        int64_t run = 0;
        for (int64_t b = 0; b <= B; b++) {
            const int64_t c = b < B ? uniq[b] : 0;
            uniq[b] = run;
            run += c;
        }
    }
    const int64_t nnz = uniq[B];
    GNNA_REQUIRE(nnz < 0x80000000LL, "csr_from_edges: graph has %lld edges; the int32 CSR contract of the reference stops at 2^31-1",
                 (long long)nnz);
#pragma omp parallel for schedule(dynamic, 4) num_threads(T)
    for (int64_t b = 0; b < B; b++) {
        const uint64_t *k = keys + bucket_begin[b];
        const int64_t cnt = uniq[b + 1] - uniq[b], out0 = uniq[b];
        const int64_t r_first = b << shift, r_last = std::min<int64_t>(N, (b + 1) << shift);
        int64_t j = 0;
        for (int64_t r = r_first; r < r_last; r++) {       // rows of the bucket: offset of the first key of row r
            row_ptr_host[r] = (int32_t)(out0 + j);
            while (j < cnt && (int64_t)(k[j] >> 32) == r) {
                col_idx_host[out0 + j] = (int32_t)(k[j] & 0xffffffffu);
                j++;
            }
        }
    }
    row_ptr_host[N] = (int32_t)nnz;
    *nnz_host = nnz;
    return GNNA_OK;
}
