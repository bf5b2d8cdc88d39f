// probe.cu -- bandwidth micro-benchmarks that give bench.py its roofline denominators on the box it runs on.
//
// The aggregation's feature matrix is L2-resident on the Reddit-size workload (60 MB of 126 MB), so its roof is the
// L2 -> SM path, not HBM (VERDICT r1 "roofline labelling").  gnna_probe_l2_read streams a buffer that fits in L2 through
// the same path with the same instruction the gather uses (ld.global.nc.v4; the buffer is far larger than L1): mode 0 reads it front to back, fully coalesced (the
// highest rate the path delivers), mode 1 reads it as randomly ordered 256-byte rows by 16-lane sub-warps -- the access
// pattern of the D=64 fp32 gather.  The caller times the launch with CUDA events: bytes = buffer bytes * passes.
// No reference counterpart (measurement infrastructure).
#include "common.h"

namespace gnna {

// the gather's own load instruction (gather.cuh: ld.global.nc.v4, volatile so a batch is issued before it is consumed)
__device__ __forceinline__ uint4 ld_nc_v4(const uint4 *p)
{
    uint4 v;
    asm volatile("ld.global.nc.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

constexpr int PROBE_UNROLL = 8;

// n16 (16-byte chunks of the buffer) is a multiple of gridDim.x * blockDim.x * PROBE_UNROLL (the host wrapper rounds the
// size down), so every thread issues only full batches of PROBE_UNROLL independent loads -- no serialised tail -- and in
// mode 0 every chunk of the buffer is loaded exactly once per pass.  The buffer is ~100x an SM's L1 and a CTA reads a
// different slice in every pass, so every load is served by L2.
__global__ void __launch_bounds__(512, 3)   // <= 42 registers: 1536 threads per SM at any block size
l2_read_probe_kernel(const uint4 *__restrict__ buf, long long n16, int passes, int mode, unsigned *sink)
{
    const long long T = (long long)gridDim.x * blockDim.x;
    unsigned acc = 0;
    for (int p = 0; p < passes; p++) {
        const long long tid = (long long)((blockIdx.x + 41u * (unsigned)p) % gridDim.x) * blockDim.x + threadIdx.x;
        if (mode == 0) {
            for (long long i = tid; i < n16; i += T * PROBE_UNROLL) {
                uint4 v[PROBE_UNROLL];
#pragma unroll
                for (int u = 0; u < PROBE_UNROLL; u++) v[u] = ld_nc_v4(buf + i + u * T);
#pragma unroll
                for (int u = 0; u < PROBE_UNROLL; u++) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
            }
        } else {
            // 256-byte rows (16 chunks, 16 lanes each) in a scrambled order: the D=64 fp32 gather's access pattern
            const unsigned rows = (unsigned)(n16 / 16);
            const long long subs = T / 16, sub = tid / 16;
            const int lane = (int)(tid & 15);
            for (long long r = sub; r < rows; r += subs * PROBE_UNROLL) {
                uint4 v[PROBE_UNROLL];
#pragma unroll
                for (int u = 0; u < PROBE_UNROLL; u++) {
                    unsigned h = (unsigned)(r + u * subs) * 0x9E3779B1u + (unsigned)p * 0x85EBCA6Bu;
                    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12;
                    const unsigned row = (unsigned)(((unsigned long long)h * rows) >> 32);
                    v[u] = ld_nc_v4(buf + (long long)row * 16 + lane);
                }
#pragma unroll
                for (int u = 0; u < PROBE_UNROLL; u++) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
            }
        }
    }
    if (acc == 0x9e3779b9u) *sink = acc;     // never true in practice: keeps the loads alive
}

}  // namespace gnna

// Reads `buf` `passes` times.  mode 0: coalesced stream, 1: random 256-byte rows.  `bytes` is rounded DOWN to a whole number
// of batches (SMs * ctas_per_sm * 512 threads * 8 loads * 16 bytes); the bytes one pass really reads come back in
// *bytes_per_pass.  threads_per_sm <= 0 picks 1536 (8 loads of 16 bytes in flight each: 196 KB per SM, what the gather keeps in
// flight with 48 resident warps); block_threads (128, 256 or 512; <= 0 picks 512) is the CTA size they come in -- the gather
// runs 128-thread CTAs, and small CTAs drift apart instead of issuing their batches in lockstep.
// `sink` is any 4 writable bytes of device memory.
extern "C" int gnna_probe_l2_read(const void *buf, int64_t bytes, int passes, int mode, int threads_per_sm, int block_threads,
                                  void *sink, int64_t *bytes_per_pass, void *stream)
{
    using namespace gnna;
    GNNA_REQUIRE(buf && sink && passes > 0 && (mode == 0 || mode == 1), "gnna_probe_l2_read: bad argument");
    GNNA_REQUIRE((((uintptr_t)buf) & 15) == 0, "gnna_probe_l2_read: buffer not 16-byte aligned");
    int dev = 0, sms = 148;
    GNNA_CUDA_CHECK(cudaGetDevice(&dev));
    GNNA_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (threads_per_sm <= 0) threads_per_sm = 1536;
    if (block_threads <= 0) block_threads = 512;
    GNNA_REQUIRE(block_threads == 128 || block_threads == 256 || block_threads == 512, "gnna_probe_l2_read: block_threads must be 128, 256 or 512");
    const int ctas_per_sm = threads_per_sm / block_threads > 0 ? threads_per_sm / block_threads : 1;
    const long long batch16 = (long long)sms * ctas_per_sm * block_threads * PROBE_UNROLL;      // chunks per full batch of the grid
    const long long n16 = (bytes / 16) / batch16 * batch16;
    GNNA_REQUIRE(n16 > 0, "gnna_probe_l2_read: buffer smaller than one batch (%lld bytes)", batch16 * 16);
    if (bytes_per_pass) *bytes_per_pass = n16 * 16;
    l2_read_probe_kernel<<<sms * ctas_per_sm, block_threads, 0, (cudaStream_t)stream>>>((const uint4 *)buf, n16, passes, mode, (unsigned *)sink);
    GNNA_CUDA_CHECK(cudaGetLastError());
    count_launch(1);
    return GNNA_OK;
}
