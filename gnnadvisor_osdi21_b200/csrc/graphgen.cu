// graphgen.cu -- the counter-based pair stream of gnnadvisor_osdi21_b200/graph.py (stream_pairs) as ONE kernel.
//
// graph.py defines the synthetic look-alike graphs with torch integer ops so that the same graph comes out on the CPU (the
// --impl reference arm, the tests) and on any GPU.  At ogbn-papers100M size every rank walks 8*10^8 pairs of the stream to
// pick the rows it owns (graph.synth_graph_shard); as ~450 elementwise torch kernels per 32 M-pair chunk that is minutes
// of an 8-GPU lease, as one fused kernel it is seconds.  Bit-for-bit the same pairs (tests/test_parity_gpu.py checks).
// Workload generation, no reference counterpart (the reference reads dataset files, dataset.py:58-94).
#include "common.h"

namespace gnna {

__device__ __forceinline__ unsigned long long mix64(unsigned long long x)
{
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL;
    x ^= x >> 27; x *= 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

// kind 0: R-MAT with thresholds (a, a+b, a+b+c) * 2^32, `bits` levels; kind 1: uniform.  Dropped pairs (self loop or id >= N)
// are written as src = dst = -1.
__global__ void __launch_bounds__(256)
stream_pairs_kernel(long long start, long long count, unsigned long long seed_mix, long long N, int kind, int bits,
                    unsigned long long t_a, unsigned long long t_ab, unsigned long long t_abc,
                    long long *__restrict__ src_out, long long *__restrict__ dst_out)
{
    const unsigned long long GOLD = 0x9E3779B97F4A7C15ULL;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
        const unsigned long long base = mix64((unsigned long long)(start + i) ^ seed_mix);
        unsigned long long s = 0, d = 0;
        if (kind == 1) {
            const unsigned long long h = mix64(base + GOLD);
            s = ((h >> 32) * (unsigned long long)N) >> 32;
            d = ((h & 0xFFFFFFFFULL) * (unsigned long long)N) >> 32;
        } else {
            unsigned long long h = 0;
            for (int lvl = 0; lvl < bits; lvl++) {
                unsigned long long u;
                if ((lvl & 1) == 0) {
                    h = mix64(base + (unsigned long long)(lvl / 2 + 1) * GOLD);
                    u = h >> 32;
                } else {
                    u = h & 0xFFFFFFFFULL;
                }
                s = s * 2 + (u >= t_ab ? 1 : 0);
                d = d * 2 + (((u >= t_a) && (u < t_ab)) || (u >= t_abc) ? 1 : 0);
            }
        }
        const bool keep = (long long)s < N && (long long)d < N && s != d;
        src_out[i] = keep ? (long long)s : -1;
        dst_out[i] = keep ? (long long)d : -1;
    }
}

}  // namespace gnna

extern "C" int gnna_stream_pairs(int64_t start, int64_t count, uint64_t seed_mix, int64_t num_nodes, int kind, int bits,
                                 uint64_t t_a, uint64_t t_ab, uint64_t t_abc, int64_t *src_out, int64_t *dst_out, void *stream)
{
    using namespace gnna;
    GNNA_REQUIRE(count >= 0 && num_nodes > 0 && (kind == 0 || kind == 1) && bits >= 1 && bits <= 62, "gnna_stream_pairs: bad argument");
    if (count == 0) return GNNA_OK;
    GNNA_REQUIRE(src_out && dst_out, "gnna_stream_pairs: null pointer");
    long long blocks = (count + 255) / 256;
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    stream_pairs_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(start, count, seed_mix, num_nodes, kind, bits, t_a, t_ab, t_abc,
                                                                          (long long *)src_out, (long long *)dst_out);
    GNNA_CUDA_CHECK(cudaGetLastError());
    count_launch(1);
    return GNNA_OK;
}
