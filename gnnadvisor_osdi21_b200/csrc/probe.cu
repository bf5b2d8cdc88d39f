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

// Every 16-byte chunk of the buffer is loaded EXACTLY once per pass (full batches of PROBE_UNROLL loads per thread, then
// a one-load-at-a-time tail), so bytes moved = buffer bytes * passes.  The buffer (32 MB) is far larger than an SM's L1, and
// a thread never re-reads a line within a pass, so every load is served by L2.
__global__ void __launch_bounds__(512, 2)
l2_read_probe_kernel(const uint4 *__restrict__ buf, long long n16, int passes, int mode, unsigned *sink)
{
    const long long T = (long long)gridDim.x * blockDim.x;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned acc = 0;
    if (mode == 0) {
        for (int p = 0; p < passes; p++) {
            // a CTA reads a different slice of the buffer in every pass (rotated by 41 CTAs), so what its SM's L1 kept
            // from the previous pass is never what it asks for next: every load goes to L2
            long long i = (long long)((blockIdx.x + 41u * (unsigned)p) % gridDim.x) * blockDim.x + threadIdx.x;
            for (; i + (PROBE_UNROLL - 1) * T < n16; i += T * PROBE_UNROLL) {
                uint4 v[PROBE_UNROLL];
#pragma unroll
                for (int u = 0; u < PROBE_UNROLL; u++) v[u] = ld_nc_v4(buf + i + u * T);
#pragma unroll
                for (int u = 0; u < PROBE_UNROLL; u++) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
            }
            for (; i < n16; i += T) {
                const uint4 v = ld_nc_v4(buf + i);
                acc ^= v.x ^ v.y ^ v.z ^ v.w;
            }
        }
    } else {
        // rows of 16 chunks (256 bytes) visited in a scrambled order: row index = bit-mixed counter (a bijection on
        // [0, rows) because rows is a power of two and the multiplier is odd); 16 lanes read one row
        const long long rows = n16 / 16, subs = T / 16, sub = tid / 16;
        const unsigned long long mask = (unsigned long long)rows - 1;
        const int lane = (int)(tid & 15);
        for (int p = 0; p < passes; p++) {
            long long r = sub;
            for (; r + (PROBE_UNROLL - 1) * subs < rows; r += subs * PROBE_UNROLL) {
                uint4 v[PROBE_UNROLL];
#pragma unroll
                for (int u = 0; u < PROBE_UNROLL; u++) {
                    const unsigned long long k = ((unsigned long long)(r + u * subs) * 0x9E3779B97F4A7C15ULL + (unsigned)p) & mask;
                    v[u] = ld_nc_v4(buf + (long long)k * 16 + lane);
                }
#pragma unroll
                for (int u = 0; u < PROBE_UNROLL; u++) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
            }
            for (; r < rows; r += subs) {
                const unsigned long long k = ((unsigned long long)r * 0x9E3779B97F4A7C15ULL + (unsigned)p) & mask;
                const uint4 v = ld_nc_v4(buf + (long long)k * 16 + lane);
                acc ^= v.x ^ v.y ^ v.z ^ v.w;
            }
        }
    }
    if (acc == 0x9e3779b9u) *sink = acc;     // never true in practice: keeps the loads alive
}

}  // namespace gnna

// Reads `bytes` (rounded down to 256) of `buf` `passes` times.  mode 0: coalesced stream, 1: random 256-byte rows.
// ctas_per_sm <= 0 picks 2 (1024 threads per SM, 8 loads of 16 bytes in flight each).  `sink` is any 4 writable bytes of device memory.
extern "C" int gnna_probe_l2_read(const void *buf, int64_t bytes, int passes, int mode, int ctas_per_sm, void *sink, void *stream)
{
    using namespace gnna;
    GNNA_REQUIRE(buf && sink && bytes >= 256 && passes > 0 && (mode == 0 || mode == 1), "gnna_probe_l2_read: bad argument");
    GNNA_REQUIRE((((uintptr_t)buf) & 15) == 0, "gnna_probe_l2_read: buffer not 16-byte aligned");
    GNNA_REQUIRE(mode == 0 || ((bytes / 256) & (bytes / 256 - 1)) == 0, "gnna_probe_l2_read: mode 1 needs a power-of-two number of 256-byte rows");
    int dev = 0, sms = 148;
    GNNA_CUDA_CHECK(cudaGetDevice(&dev));
    GNNA_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (ctas_per_sm <= 0) ctas_per_sm = 2;
    l2_read_probe_kernel<<<sms * ctas_per_sm, 512, 0, (cudaStream_t)stream>>>((const uint4 *)buf, (bytes / 256) * 16, passes, mode,
                                                                             (unsigned *)sink);
    GNNA_CUDA_CHECK(cudaGetLastError());
    count_launch(1);
    return GNNA_OK;
}
