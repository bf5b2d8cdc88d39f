// build_part.cu -- the neighbour-group table (replaces build_part, GNNAdvisor.cpp:210-251) and
// the degree vector (GNNAdvisor/dataset.py:11-18,121-122).
//
// The reference builds the table with a single-threaded loop of per-element Tensor::operator[]
// stores (about 3 us per group) into FLOAT32 tensors (F5).  Here:
//   * host version: two passes over node ranges on all host threads, int32 output; `compat`
//     reproduces the reference's float32 round trip and its terminal-entry rule (F6) exactly;
//   * device version: count -> exclusive scan -> one thread per GROUP finds its node by binary
//     search in the scanned offsets (balanced no matter how skewed the degrees are).
#include <algorithm>
#include <thread>
#include <vector>

#include "common.h"

namespace gnna {

static inline int parts_of(int degree, int ps) { return degree <= 0 ? 0 : (degree + ps - 1) / ps; }

static int host_threads(int64_t n)
{
    unsigned hw = std::thread::hardware_concurrency();
    int t = hw ? (int)hw : 1;
    if (t > 64) t = 64;
    int64_t by_work = n / 65536 + 1;
    return (int)std::min<int64_t>(t, by_work);
}

}  // namespace gnna

extern "C" int64_t gnna_count_parts_host(int part_size, const int32_t *indptr, int64_t num_nodes)
{
    if (part_size <= 0 || !indptr || num_nodes < 0) {
        gnna::fail(GNNA_ERR_INVALID, "gnna_count_parts_host: bad argument");
        return -1;
    }
    const int T = gnna::host_threads(num_nodes);
    std::vector<int64_t> partial(T, 0);
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++)
        th.emplace_back([&, t] {
            int64_t a = num_nodes * t / T, b = num_nodes * (t + 1) / T, s = 0;
            for (int64_t i = a; i < b; i++) s += gnna::parts_of(indptr[i + 1] - indptr[i], part_size);
            partial[t] = s;
        });
    for (auto &x : th) x.join();
    int64_t total = 0;
    for (int64_t s : partial) total += s;
    return total;
}

extern "C" int gnna_build_part_host(int part_size, const int32_t *indptr, int64_t num_nodes,
                                    int32_t *part_ptr, int32_t *part2node, int64_t num_parts, int compat)
{
    GNNA_REQUIRE(part_size > 0, "build_part: partSize must be positive (got %d)", part_size);
    GNNA_REQUIRE(indptr && part_ptr && num_nodes >= 0, "build_part: null pointer");
    GNNA_REQUIRE(part2node || num_parts == 0, "build_part: null part2node");
    const int T = gnna::host_threads(num_nodes);
    std::vector<int64_t> start(T + 1, 0);
    {
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++)
            th.emplace_back([&, t] {
                int64_t a = num_nodes * t / T, b = num_nodes * (t + 1) / T, s = 0;
                for (int64_t i = a; i < b; i++) s += gnna::parts_of(indptr[i + 1] - indptr[i], part_size);
                start[t + 1] = s;
            });
        for (auto &x : th) x.join();
    }
    for (int t = 0; t < T; t++) start[t + 1] += start[t];
    GNNA_REQUIRE(start[T] == num_parts, "build_part: num_parts %lld does not match the graph (%lld)",
                 (long long)num_parts, (long long)start[T]);
    {
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++)
            th.emplace_back([&, t] {
                int64_t a = num_nodes * t / T, b = num_nodes * (t + 1) / T, c = start[t];
                for (int64_t i = a; i < b; i++) {
                    const int32_t lo = indptr[i];
                    const int np = gnna::parts_of(indptr[i + 1] - lo, part_size);
                    for (int pid = 0; pid < np; pid++) {
                        int32_t v = lo + pid * part_size;
                        // compat: the reference stores into a float32 tensor and the caller casts back
                        part_ptr[c] = compat ? (int32_t)(float)v : v;
                        part2node[c] = compat ? (int32_t)(float)(int32_t)i : (int32_t)i;
                        c++;
                    }
                }
            });
        for (auto &x : th) x.join();
    }
    const int32_t last = num_nodes > 0 ? indptr[num_nodes] : 0;
    if (compat) {
        // GNNAdvisor.cpp:246-247: the terminal is written only while visiting the LAST node's last group
        const bool last_has_nbr = num_nodes > 0 && indptr[num_nodes] > indptr[num_nodes - 1];
        part_ptr[num_parts] = last_has_nbr ? (int32_t)(float)last : 0;
    } else {
        part_ptr[num_parts] = last;
    }
    return GNNA_OK;
}

// ------------------------------------------------------------------------------------------
// device version
// ------------------------------------------------------------------------------------------
namespace gnna {

constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int dev_parts_of(const int32_t *indptr, long long i, long long n, int ps)
{
    if (i >= n) return 0;
    int d = indptr[i + 1] - indptr[i];
    return d <= 0 ? 0 : (d + ps - 1) / ps;
}

// block-wide exclusive scan of one int per thread; returns the block total to every thread
__device__ __forceinline__ int block_exclusive_scan(int v, int &total, int *smem /* >= 32 */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = lane < (blockDim.x >> 5) ? smem[lane] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        smem[lane] = winc - w;                    // exclusive warp offsets
        if (lane == 31) smem[32] = winc;          // block total
    }
    __syncthreads();
    total = smem[32];
    int r = smem[warp] + inc - v;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_THREADS)
part_tile_sums(const int32_t *__restrict__ indptr, long long n, int ps, long long *__restrict__ tile_sums)
{
    __shared__ int smem[33];
    const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) s += dev_parts_of(indptr, base + k, n, ps);
    int total;
    block_exclusive_scan(s, total, smem);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// one block: exclusive scan of the tile sums in place, grand total in tile_sums[num_tiles]
__global__ void __launch_bounds__(1024)
scan_tile_sums(long long *tile_sums, long long num_tiles)
{
    __shared__ long long carry;
    __shared__ long long wsum[33];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long b = 0; b < num_tiles; b += blockDim.x) {
        long long i = b + threadIdx.x;
        long long v = i < num_tiles ? tile_sums[i] : 0, inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            long long t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            long long w = wsum[lane], winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                long long t = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += t;
            }
            wsum[lane] = winc - w;
            if (lane == 31) wsum[32] = winc;
        }
        __syncthreads();
        if (i < num_tiles) tile_sums[i] = carry + wsum[warp] + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += wsum[32];
        __syncthreads();
    }
    if (threadIdx.x == 0) tile_sums[num_tiles] = carry;
}

// offsets[i] = number of groups before node i  (int32; P < 2^31 is checked on the host)
__global__ void __launch_bounds__(SCAN_THREADS)
part_offsets(const int32_t *__restrict__ indptr, long long n, int ps, const long long *__restrict__ tile_sums,
             int32_t *__restrict__ offsets)
{
    __shared__ int smem[33];
    const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
    int c[SCAN_ITEMS], s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { c[k] = dev_parts_of(indptr, base + k, n, ps); s += c[k]; }
    int total;
    int ex = block_exclusive_scan(s, total, smem);
    long long run = tile_sums[blockIdx.x] + ex;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        if (base + k <= n) offsets[base + k] = (int32_t)run;   // includes offsets[n] = P
        run += c[k];
    }
}

// one thread per group: node = last i with offsets[i] <= g
__global__ void __launch_bounds__(256)
part_expand(const int32_t *__restrict__ indptr, const int32_t *__restrict__ offsets, long long n, int ps,
            long long num_parts, int32_t *__restrict__ part_ptr, int32_t *__restrict__ part2node)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g == 0) part_ptr[num_parts] = indptr[n];
    if (g >= num_parts) return;
    long long lo = 0, hi = n;                      // offsets[lo] <= g < offsets[hi]
    while (hi - lo > 1) {
        long long mid = (lo + hi) >> 1;
        if ((long long)offsets[mid] <= g) lo = mid; else hi = mid;
    }
    part2node[g] = (int32_t)lo;
    part_ptr[g] = indptr[lo] + (int32_t)(g - offsets[lo]) * ps;
}

__global__ void __launch_bounds__(256)
degrees_kernel(const int32_t *__restrict__ indptr, long long n, float *__restrict__ degrees)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        int d = indptr[i + 1] - indptr[i];
        degrees[i] = sqrtf((float)(d > 0 ? d : 1));
    }
}

static inline long long num_tiles_of(long long n) { return (n + 1 + SCAN_TILE - 1) / SCAN_TILE; }

}  // namespace gnna

extern "C" int64_t gnna_build_part_workspace_bytes(int64_t num_nodes)
{
    if (num_nodes < 0) return -1;
    const long long tiles = gnna::num_tiles_of(num_nodes);
    // [tile sums: (tiles+1) int64][offsets: (n+1) int32], 256-byte aligned pieces
    long long a = ((tiles + 1) * 8 + 255) / 256 * 256;
    long long b = ((num_nodes + 1) * 4 + 255) / 256 * 256;
    return a + b;
}

extern "C" int gnna_build_part_device(int part_size, const int32_t *indptr, int64_t num_nodes,
                                      int32_t *part_ptr, int32_t *part2node, int64_t *num_parts_out_host,
                                      void *workspace, int64_t workspace_bytes, void *stream_)
{
    using namespace gnna;
    cudaStream_t stream = (cudaStream_t)stream_;
    GNNA_REQUIRE(part_size > 0, "build_part: partSize must be positive (got %d)", part_size);
    GNNA_REQUIRE(indptr && workspace && num_nodes >= 0, "build_part_device: null pointer");
    GNNA_REQUIRE(workspace_bytes >= gnna_build_part_workspace_bytes(num_nodes), "build_part_device: workspace too small");
    const long long tiles = num_tiles_of(num_nodes);
    long long *tile_sums = (long long *)workspace;
    int32_t *offsets = (int32_t *)((char *)workspace + ((tiles + 1) * 8 + 255) / 256 * 256);

    part_tile_sums<<<(unsigned)tiles, SCAN_THREADS, 0, stream>>>(indptr, num_nodes, part_size, tile_sums);
    scan_tile_sums<<<1, 1024, 0, stream>>>(tile_sums, tiles);
    part_offsets<<<(unsigned)tiles, SCAN_THREADS, 0, stream>>>(indptr, num_nodes, part_size, tile_sums, offsets);
    GNNA_CUDA_CHECK(cudaGetLastError());
    count_launch(3);
    long long P = 0;
    GNNA_CUDA_CHECK(cudaMemcpyAsync(&P, tile_sums + tiles, sizeof(long long), cudaMemcpyDeviceToHost, stream));
    GNNA_CUDA_CHECK(cudaStreamSynchronize(stream));
    if (num_parts_out_host) *num_parts_out_host = P;
    GNNA_REQUIRE(P < 0x7fffffffLL, "build_part_device: %lld groups do not fit int32 tables", P);
    if (!part_ptr) return GNNA_OK;                 // sizing call
    GNNA_REQUIRE(part2node || P == 0, "build_part_device: null part2node");
    const long long threads = P > 0 ? P : 1;
    part_expand<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(indptr, offsets, num_nodes, part_size, P,
                                                                      part_ptr, part2node);
    GNNA_CUDA_CHECK(cudaGetLastError());
    count_launch(1);
    return GNNA_OK;
}

extern "C" int gnna_degrees(const int32_t *indptr, int64_t num_nodes, float *degrees, void *stream)
{
    GNNA_REQUIRE(num_nodes >= 0, "gnna_degrees: negative size");
    if (num_nodes == 0) return GNNA_OK;
    GNNA_REQUIRE(indptr && degrees, "gnna_degrees: null pointer");
    gnna::degrees_kernel<<<(unsigned)((num_nodes + 255) / 256), 256, 0, (cudaStream_t)stream>>>(indptr, num_nodes, degrees);
    GNNA_CUDA_CHECK(cudaGetLastError());
    gnna::count_launch(1);
    return GNNA_OK;
}
