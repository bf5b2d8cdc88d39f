// gemm_tf32x3.cu -- the layer's dense products on the 5th-generation tensor cores at fp32-grade accuracy.
//
// Reference: the six torch::mm call sites of GNNAdvisor/GNNConv/GNNAdvisor_kernel.cu (:280 X*W, :472, :473 X^T*G, :605,
// :710, :711) -- cuBLAS SGEMM there, and until round 2 here too (cuBLAS picks a SIMT kernel, `cutlass_80_simt_sgemm`: 12 % of
// the fp32 GCN epoch and 20 % of the bf16-row epoch on the Reddit look-alike).  The two products that are real
// contractions are tall and skinny:
//     NN   C[M, N] = A[M, K] * B[K, N]            X*W:    M = 232 965 nodes, K = 602, N = 64     (kernel.cu:280)
//     TN   C[Mc, N] = A[Kn, Mc]^T * B[Kn, N]      X^T*G:  Mc = 602, N = 64, reduced over Kn = 232 965 nodes  (kernel.cu:473)
// Both are bound by reading X once (561 MB: 0.09 ms of HBM time against 0.2-0.4 ms of SIMT SGEMM).  Here:
//   * tcgen05.mma.cta_group::1.kind::tf32, M = 128 x N = round_up(n, 16) x K = 8 per instruction, accumulator in TMEM;
//   * fp32 accuracy by the 3xTF32 split: x = hi + lo with hi = x truncated to TF32's 10 mantissa bits (exactly
//     representable, so the result does not depend on how the tensor core rounds a 32-bit container) and lo = x - hi (exact in
//     fp32); D += A_hi*B_hi + A_lo*B_hi + A_hi*B_lo -- the dropped lo*lo term and the truncation of lo are O(2^-22) relative,
//     i.e. SGEMM-grade (the parity bar is 1e-4);
//   * operand tiles are brought in with cp.async (LDGSTS) straight into the canonical K-major, no-swizzle UMMA layout
//     (8-row x 16-byte core matrices): a warp fills whole core matrices, so shared-memory writes are conflict-free and every
//     32-byte sector fetched is used.  TMA cannot serve this path: X's row pitch (602 * 4 = 2408 bytes) is not a multiple of
//     16 bytes, which cuTensorMapEncodeTiled requires.  The TN product transposes while it loads (4-byte copies);
//   * R raw stages (7 for N <= 64, 5 up to N = 128: what 227 KB of shared memory hold), R - 2 k-blocks in flight while one is
//     split and one multiplied -- 80 KB of loads per SM in flight, what it takes to pull HBM bandwidth through one CTA per
//     SM (with 2 in flight the first version ran at cuBLAS-SIMT speed) -- split tiles double buffered; the elected thread
//     issues 12 MMAs per k-block and `tcgen05.commit`s to the mbarrier that frees the buffers;
//   * persistent CTAs (one per SM) over row tiles (NN) or over a split of the node range (TN, merged with red.global.add);
//     the loads of the next tile are already in flight while TMEM is drained by tcgen05.ld.
#include <string.h>

#include "common.h"

namespace gnna {

namespace {

constexpr int GT_M = 128;          // rows of the accumulator tile (UMMA M)
constexpr int GT_BK = 32;          // fp32 elements of K per k-block: 128 bytes = 8 core matrices per row
constexpr int GT_THREADS = 256;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle, 4-byte elements: element (r, k) of an [R x 32] tile lives at
//   (r/8)*1024 + (k/4)*128 + (r%8)*16 + (k%4)*4          (core matrix = 8 rows x 16 bytes; 8 of them per 8-row group)
__device__ __forceinline__ uint32_t tile_offset(int r, int k) { return (uint32_t)((r >> 3) * 1024 + (k >> 2) * 128 + (r & 7) * 16 + (k & 3) * 4); }

// cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version = 1 [46,48), SWIZZLE_NONE
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr)
{
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46);
}

// cute::UMMA::InstrDescriptor, kind::tf32: D = f32 [4,6) = 1, A = B = TF32 [7,10) = [10,13) = 2, both K-major, N>>3 [17,23), M>>4 [24,29)
__device__ __forceinline__ uint32_t idesc_tf32(int M, int N)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "GT_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra GT_DONE;\n\t"
        "bra GT_WAIT;\n\t"
        "GT_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}

// bounded wait (a protocol bug traps instead of hanging the GPU): ~2^28 polls are seconds
__device__ __forceinline__ void mbar_wait_b(uint64_t *bar, uint32_t parity)
{
    uint32_t done = 0;
    for (unsigned spins = 0; !done; spins++) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (spins > (1u << 28)) asm volatile("trap;");
    }
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
// the barrier receives one arrival from this thread when all cp.async it issued so far have landed
__device__ __forceinline__ void cp_async_arrive(uint64_t *bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// cp.async of CP bytes with zero fill of the bytes beyond `valid` (0 <= valid <= CP); src must be a valid address even
// when valid == 0 (nothing is read then)
template <int CP>
__device__ __forceinline__ void cp_async(uint32_t dst, const void *src, int valid)
{
    if constexpr (CP == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(valid) : "memory");
    else if constexpr (CP == 8)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(valid) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(valid) : "memory");
}

struct GemmArgs {
    const float *A, *B;
    float *C;
    long long M;        // NN: rows of A and C.  TN: rows of A and B (the reduced dimension)
    int K;              // NN: columns of A = rows of B.  TN: columns of A = rows of C
    int N;              // columns of B and C
    int npad;           // N rounded up to 16
    long long work;     // NN: number of row tiles.  TN: k-blocks of the reduced dimension per split
    int splits;         // TN: CTAs per row tile of C
    const float *row_scale;   // NN epilogue: C[i,:] *= row_scale[i] (may be null)
};

// ---- loaders: one k-block of the A operand ([128 x 32]) and of the B operand ([npad x 32]) into K-major tiles -----------
// NN A operand: rows of A, K contiguous in global memory.  A warp fills whole core matrices: lane -> (row % 8 = lane % 8,
// 16-byte chunk = lane / 8), i.e. 8 rows x 64 contiguous bytes per request on the global side, 512 contiguous bytes of
// shared memory.  CP = copy size the alignment of A allows (16, 8 or 4 bytes).
template <int CP>
__device__ __forceinline__ void load_rows_kmajor(uint32_t tile, const float *base, long long row0, long long rows_total, int ld, int k0,
                                                 int k_total, int tile_rows, int t = threadIdx.x, int nthreads = GT_THREADS)
{
    // The tile is filled in SHARED-memory order, CP bytes per lane: piece i lives at byte i * CP of the tile, i.e. in 16-byte
    // chunk c = i * CP / 16 = (group of 8 rows, k / 4, row % 8).  A warp request therefore writes 32 * CP contiguous bytes
    // (conflict-free) and reads, per row it touches, (16 / CP lanes) x CP = 16 contiguous bytes per chunk: with 8-byte copies
    // the two halves of a chunk sit on neighbouring lanes, so every 32-byte sector a request touches is used completely
    // (the first version put them in different instructions: half-used sectors, X*W at 1.0 TB/s).
    constexpr int PER = 16 / CP;                            // pieces per 16-byte chunk
    const int pieces = tile_rows * 8 * PER;
    for (int i = t; i < pieces; i += nthreads) {
        const int c = i / PER, h = i % PER;
        const int rg = c >> 6, kc = (c >> 3) & 7, r8 = c & 7;
        const int r = rg * 8 + r8, k = kc * 4 + h * (CP / 4);
        const long long grow = row0 + r;
        const int kk = k0 + k;
        int valid = 0;
        if (grow < rows_total && kk < k_total) valid = min(CP, (k_total - kk) * 4);
        cp_async<CP>(tile + (uint32_t)(c * 16 + h * CP), valid ? (const void *)(base + grow * ld + kk) : (const void *)base, valid);
    }
}

// Transposing loader: operand element (r, k) = base[(k0 + k) * ld + col0 + r] (r contiguous in global memory): 4-byte copies,
// lane -> (k % 4 = lane % 4, row % 8 = lane / 4): one core matrix (128 contiguous bytes of shared memory) per warp request,
// 4 global rows x 32 contiguous bytes on the global side.
__device__ __forceinline__ void load_cols_kmajor(uint32_t tile, const float *base, int col0, int cols_total, int ld, long long k0,
                                                 long long k_total, int tile_rows, int t = threadIdx.x, int nthreads = GT_THREADS)
{
    const int lane = t & 31, warp = t >> 5;
    const int cms = tile_rows;                              // core matrices in the tile: (tile_rows / 8) * 8
    for (int cm = warp; cm < cms; cm += nthreads / 32) {
        const int rg = cm >> 3, kc = cm & 7;
        const int r = rg * 8 + (lane >> 2), k = kc * 4 + (lane & 3);
        const long long gk = k0 + k;
        const int gc = col0 + r;
        const bool ok = gk < k_total && gc < cols_total;
        cp_async<4>(tile + (uint32_t)(rg * 1024 + kc * 128 + (lane >> 2) * 16 + (lane & 3) * 4),
                    ok ? (const void *)(base + gk * ld + gc) : (const void *)base, ok ? 4 : 0);
    }
}

// hi/lo split of a tile in place: raw -> hi (same buffer), lo (second buffer).  Element-wise, layout-agnostic, 128-bit.
__device__ __forceinline__ void split_tile(float *raw, float *lo, int bytes, int t = threadIdx.x, int nthreads = GT_THREADS)
{
    const int n4 = bytes >> 4;
    for (int i = t; i < n4; i += nthreads) {
        float4 v = reinterpret_cast<float4 *>(raw)[i];
        float4 h, l;
        h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
        h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
        h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
        h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
        reinterpret_cast<float4 *>(raw)[i] = h;
        reinterpret_cast<float4 *>(lo)[i] = l;
    }
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

}  // namespace

// TRANS = false: NN (A rows K-major in memory, B transposed while loading).  TRANS = true: TN (both operands transposed
// while loading, reduced dimension split over gridDim.y CTAs, result merged with reductions).  CP: copy size for A (NN only).
template <bool TRANS, int CP, int GT_RAW>
__global__ void __launch_bounds__(GT_THREADS, 1)
gemm_tf32x3_kernel(const GemmArgs g)
{
    constexpr int DIST = GT_RAW - 2;               // k-blocks in flight
    extern __shared__ __align__(1024) unsigned char smem[];
    const int npad = g.npad;
    const int a_bytes = GT_M * GT_BK * 4, b_bytes = npad * GT_BK * 4;
    unsigned char *raw_a = smem;                                    // [GT_RAW][a_bytes]
    unsigned char *raw_b = raw_a + GT_RAW * a_bytes;                // [GT_RAW][b_bytes]
    unsigned char *lo_a = raw_b + GT_RAW * b_bytes;                 // [2][a_bytes]
    unsigned char *lo_b = lo_a + 2 * a_bytes;                       // [2][b_bytes]
    uint64_t *bars = reinterpret_cast<uint64_t *>(lo_b + 2 * b_bytes);   // [2] "MMAs of this lo buffer are done"
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bars + 2);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int tmem_cols = 32;
    while (tmem_cols < npad) tmem_cols *= 2;
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *s_tmem;
    const uint32_t idesc = idesc_tf32(GT_M, npad);

    // ---- the flat sequence of k-blocks this CTA walks: (tile, kb) pairs
    //   NN: tiles blockIdx.x, blockIdx.x + gridDim.x, ... of `work` row tiles; KB = ceil(K / 32) k-blocks each
    //   TN: ONE tile of C (row tile blockIdx.x of the K columns of A); k-blocks [blockIdx.y * work, ...) of the reduced dimension
    const int KB = TRANS ? 0 : (g.K + GT_BK - 1) / GT_BK;
    long long n_tiles, kb_per_tile;
    if (!TRANS) {
        n_tiles = ((long long)blockIdx.x < g.work) ? (g.work - 1 - blockIdx.x) / gridDim.x + 1 : 0;
        kb_per_tile = KB;
    } else {
        const long long total_kb = (g.M + GT_BK - 1) / GT_BK;
        const long long first = (long long)blockIdx.y * g.work;
        kb_per_tile = first < total_kb ? min(g.work, total_kb - first) : 0;
        n_tiles = kb_per_tile > 0 ? 1 : 0;
    }
    const long long total = n_tiles * kb_per_tile;

    auto issue_loads = [&](long long q) {          // k-block q of the flat sequence -> raw stage q % GT_RAW
        const int st = (int)(q % GT_RAW);
        const uint32_t ta = smem_u32(raw_a + st * a_bytes), tb = smem_u32(raw_b + st * b_bytes);
        if (!TRANS) {
            const long long tile = blockIdx.x + (q / kb_per_tile) * gridDim.x;
            const int k0 = (int)(q % kb_per_tile) * GT_BK;
            load_rows_kmajor<CP>(ta, g.A, tile * GT_M, g.M, g.K, k0, g.K, GT_M);
            load_cols_kmajor(tb, g.B, 0, g.N, g.N, k0, g.K, npad);                 // B[k, n] -> operand (n, k)
        } else {
            const long long k0 = ((long long)blockIdx.y * g.work + q) * GT_BK;
            load_cols_kmajor(ta, g.A, blockIdx.x * GT_M, g.K, g.K, k0, g.M, GT_M);  // A[k, m] -> operand (m, k)
            load_cols_kmajor(tb, g.B, 0, g.N, g.N, k0, g.M, npad);
        }
    };

    // prologue: DIST k-blocks in flight
    for (int q = 0; q < DIST; q++) {
        if (q < total) issue_loads(q);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }

    for (long long q = 0; q < total; q++) {
        const int st = (int)(q % GT_RAW), lb = (int)(q & 1);
        asm volatile("cp.async.wait_group %0;" ::"n"(DIST - 1) : "memory");   // this thread's copies of k-block q have landed
        __syncthreads();                                           // ... and everybody else's
        if (q >= 2) mbar_wait(&bars[lb], (uint32_t)(((q - 2) >> 1) & 1));   // MMAs of k-block q-2 done: lo[lb] and raw stage
        if (q + DIST < total) issue_loads(q + DIST);                        // (q + DIST) % GT_RAW == (q - 2) % GT_RAW are free
        asm volatile("cp.async.commit_group;" ::: "memory");
        split_tile(reinterpret_cast<float *>(raw_a + st * a_bytes), reinterpret_cast<float *>(lo_a + lb * a_bytes), a_bytes);
        split_tile(reinterpret_cast<float *>(raw_b + st * b_bytes), reinterpret_cast<float *>(lo_b + lb * b_bytes), b_bytes);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy (UMMA)
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        const long long kb_in_tile = q % kb_per_tile;
        if (warp == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                const uint32_t a_hi = smem_u32(raw_a + st * a_bytes), a_lo = smem_u32(lo_a + lb * a_bytes);
                const uint32_t b_hi = smem_u32(raw_b + st * b_bytes), b_lo = smem_u32(lo_b + lb * b_bytes);
#pragma unroll
                for (int ks = 0; ks < GT_BK / 8; ks++) {           // K = 8 per instruction = 32 bytes = 2 core matrices
                    const uint32_t o = ks * 256;
                    mma_tf32(tmem_base, umma_desc(a_lo + o), umma_desc(b_hi + o), idesc, (kb_in_tile > 0 || ks > 0) ? 1u : 0u);
                    mma_tf32(tmem_base, umma_desc(a_hi + o), umma_desc(b_lo + o), idesc, 1u);
                    mma_tf32(tmem_base, umma_desc(a_hi + o), umma_desc(b_hi + o), idesc, 1u);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[lb]))
                             : "memory");
            }
            __syncwarp();
        }
        if (kb_in_tile == kb_per_tile - 1) {
            // ---- epilogue of this tile: wait for its last MMAs, drain TMEM (the next tile's loads are already in flight)
            mbar_wait(&bars[lb], (uint32_t)((q >> 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int qd = warp & 3;                               // TMEM lanes 32*qd .. +31 are the ones this warp may read
            const int half = warp >> 2;                            // the two warps with the same qd split the columns
            const int r_in_tile = qd * 32 + lane;
            long long row;
            if (!TRANS) row = (blockIdx.x + (q / kb_per_tile) * gridDim.x) * GT_M + r_in_tile;
            else row = (long long)blockIdx.x * GT_M + r_in_tile;
            const long long rows_total = TRANS ? g.K : g.M;
            float rs = 1.f;
            if (!TRANS && g.row_scale && row < rows_total) rs = __ldg(g.row_scale + row);
            for (int c = half * 16; c < npad; c += 32) {
                uint32_t r[16];
                const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)c;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (row < rows_total) {
                    float *o = g.C + row * g.N + c;
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        if (c + j < g.N) {
                            const float v = __uint_as_float(r[j]);
                            if (!TRANS) o[j] = g.row_scale ? __fmul_rn(rs, v) : v;
                            else asm volatile("red.global.add.f32 [%0], %1;" ::"l"(o + j), "f"(v) : "memory");
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();                                       // TMEM drained before the next tile's first MMA overwrites it
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
}


// ---- warp-specialised variant ------------------------------------------------------------------------------------
// The lockstep kernel above spends ~700 cycles of SOFTWARE per k-block (every warp issues copies, splits, then waits while
// one thread issues the MMAs, with two CTA-wide barriers in between) -- as long as the k-block's HBM time, so it ran at
// cuBLAS-SIMT speed (measured: GCN epoch 8.27 vs 7.78 ms).  Here the three jobs run side by side, coupled by mbarriers only:
//   warps 0-3  PRODUCERS   cp.async the raw A/B tiles of k-block q into stage q % R as soon as the MMAs that last read the
//                          stage have retired (empty[stage], arrived by tcgen05.commit); completion -> full[stage]
//                          (cp.async.mbarrier.arrive.noinc, one arrival per producer thread);
//   warps 4-7  CONVERTERS  wait full[stage] and the lo buffer (done[q & 1] of k-block q-2), split raw -> hi (in place) + lo,
//                          fence to the async proxy, arrive ready[q & 1]; at the end of a tile they wait for its last MMAs and
//                          drain TMEM (they are the four warps whose TMEM lane quarter is warp % 4);
//   warp 8     MMA         one thread: wait ready[q & 1], issue the 12 MMAs, commit to empty[stage] and done[q & 1].
// The next tile's first MMA is ordered after the drain of TMEM because its ready[] arrival comes from the same converter
// threads, after their epilogue.
constexpr int WS_PROD = 128, WS_CONV = 128, WS_THREADS = WS_PROD + WS_CONV + 32;

template <bool TRANS, int CP, int GT_RAW>
__global__ void __launch_bounds__(WS_THREADS, 1)
gemm_tf32x3_ws_kernel(const GemmArgs g)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    const int npad = g.npad;
    const int a_bytes = GT_M * GT_BK * 4, b_bytes = npad * GT_BK * 4;
    unsigned char *raw_a = smem;
    unsigned char *raw_b = raw_a + GT_RAW * a_bytes;
    unsigned char *lo_a = raw_b + GT_RAW * b_bytes;
    unsigned char *lo_b = lo_a + 2 * a_bytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(lo_b + 2 * b_bytes);   // [GT_RAW] raw tiles landed         (128 cp.async arrivals)
    uint64_t *empty = full + GT_RAW;                                      // [GT_RAW] raw stage may be refilled (1 commit)
    uint64_t *ready = empty + GT_RAW;                                     // [2] hi/lo tiles are in place      (128 converter arrivals)
    uint64_t *done = ready + 2;                                           // [2] MMAs of this lo buffer retired (1 commit)
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(done + 2);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int tmem_cols = 32;
    while (tmem_cols < npad) tmem_cols *= 2;
    if (tid == 0) {
        for (int i = 0; i < GT_RAW; i++) { mbar_init(&full[i], WS_PROD); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; i++) { mbar_init(&ready[i], WS_CONV); mbar_init(&done[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *s_tmem;

    long long n_tiles, kb_per_tile;
    if (!TRANS) {
        n_tiles = ((long long)blockIdx.x < g.work) ? (g.work - 1 - blockIdx.x) / gridDim.x + 1 : 0;
        kb_per_tile = (g.K + GT_BK - 1) / GT_BK;
    } else {
        const long long total_kb = (g.M + GT_BK - 1) / GT_BK;
        const long long first = (long long)blockIdx.y * g.work;
        kb_per_tile = first < total_kb ? min(g.work, total_kb - first) : 0;
        n_tiles = kb_per_tile > 0 ? 1 : 0;
    }
    const long long total = n_tiles * kb_per_tile;

    if (warp < 4) {
        // ===== PRODUCERS
        for (long long q = 0; q < total; q++) {
            const int st = (int)(q % GT_RAW);
            const uint32_t fill = (uint32_t)(q / GT_RAW);
            mbar_wait_b(&empty[st], (fill & 1) ^ 1);                // passes at once on a fresh barrier
            const uint32_t ta = smem_u32(raw_a + st * a_bytes), tb = smem_u32(raw_b + st * b_bytes);
            if (!TRANS) {
                const long long tile = blockIdx.x + (q / kb_per_tile) * gridDim.x;
                const int k0 = (int)(q % kb_per_tile) * GT_BK;
                load_rows_kmajor<CP>(ta, g.A, tile * GT_M, g.M, g.K, k0, g.K, GT_M, tid, WS_PROD);
                load_cols_kmajor(tb, g.B, 0, g.N, g.N, k0, g.K, npad, tid, WS_PROD);
            } else {
                const long long k0 = ((long long)blockIdx.y * g.work + q) * GT_BK;
                load_cols_kmajor(ta, g.A, blockIdx.x * GT_M, g.K, g.K, k0, g.M, GT_M, tid, WS_PROD);
                load_cols_kmajor(tb, g.B, 0, g.N, g.N, k0, g.M, npad, tid, WS_PROD);
            }
            cp_async_arrive(&full[st]);
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    } else if (warp < 8) {
        // ===== CONVERTERS (+ epilogue)
        const int ct = tid - WS_PROD;
        for (long long q = 0; q < total; q++) {
            const int st = (int)(q % GT_RAW), lb = (int)(q & 1);
            const uint32_t fill = (uint32_t)(q / GT_RAW), use = (uint32_t)(q >> 1);
            mbar_wait_b(&full[st], fill & 1);
            mbar_wait_b(&done[lb], (use & 1) ^ 1);                  // MMAs of k-block q-2 retired: lo[lb] is free
            split_tile(reinterpret_cast<float *>(raw_a + st * a_bytes), reinterpret_cast<float *>(lo_a + lb * a_bytes), a_bytes, ct, WS_CONV);
            split_tile(reinterpret_cast<float *>(raw_b + st * b_bytes), reinterpret_cast<float *>(lo_b + lb * b_bytes), b_bytes, ct, WS_CONV);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(&ready[lb]);
            const long long kb_in_tile = q % kb_per_tile;
            if (kb_in_tile == kb_per_tile - 1) {
                mbar_wait_b(&done[lb], use & 1);                    // the tile's last MMAs retired
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const int qd = warp & 3;
                const int r_in_tile = qd * 32 + lane;
                long long row;
                if (!TRANS) row = (blockIdx.x + (q / kb_per_tile) * gridDim.x) * GT_M + r_in_tile;
                else row = (long long)blockIdx.x * GT_M + r_in_tile;
                const long long rows_total = TRANS ? g.K : g.M;
                float rs = 1.f;
                if (!TRANS && g.row_scale && row < rows_total) rs = __ldg(g.row_scale + row);
                for (int c = 0; c < npad; c += 16) {
                    uint32_t r[16];
                    const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)c;
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                        : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (row < rows_total) {
                        float *o = g.C + row * g.N + c;
#pragma unroll
                        for (int j = 0; j < 16; j++) {
                            if (c + j < g.N) {
                                const float v = __uint_as_float(r[j]);
                                if (!TRANS) o[j] = g.row_scale ? __fmul_rn(rs, v) : v;
                                else asm volatile("red.global.add.f32 [%0], %1;" ::"l"(o + j), "f"(v) : "memory");
                            }
                        }
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");   // TMEM reads ordered before the next ready[] arrival
            }
        }
    } else if (lane == 0) {
        // ===== MMA issuer (one thread)
        const uint32_t idesc = idesc_tf32(GT_M, npad);
        for (long long q = 0; q < total; q++) {
            const int st = (int)(q % GT_RAW), lb = (int)(q & 1);
            const uint32_t use = (uint32_t)(q >> 1);
            mbar_wait_b(&ready[lb], use & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const long long kb_in_tile = q % kb_per_tile;
            const uint32_t a_hi = smem_u32(raw_a + st * a_bytes), a_lo = smem_u32(lo_a + lb * a_bytes);
            const uint32_t b_hi = smem_u32(raw_b + st * b_bytes), b_lo = smem_u32(lo_b + lb * b_bytes);
#pragma unroll
            for (int ks = 0; ks < GT_BK / 8; ks++) {
                const uint32_t o = ks * 256;
                mma_tf32(tmem_base, umma_desc(a_lo + o), umma_desc(b_hi + o), idesc, (kb_in_tile > 0 || ks > 0) ? 1u : 0u);
                mma_tf32(tmem_base, umma_desc(a_hi + o), umma_desc(b_lo + o), idesc, 1u);
                mma_tf32(tmem_base, umma_desc(a_hi + o), umma_desc(b_hi + o), idesc, 1u);
            }
            umma_commit(&empty[st]);
            umma_commit(&done[lb]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
}

static int gemm_smem_bytes(int npad, int raw) { return (raw + 2) * (GT_M * GT_BK * 4 + npad * GT_BK * 4) + 8 * (2 * raw + 4) + 64; }

// -1: not decided (environment GNNA_TC_GEMM, else 3); 0 cuBLAS only; 1 lockstep kernel for both products; 2 warp-specialised
// kernel for both; 3 (default) = what the measurements say (tools/gemm_bench.py, profiles/r02_gemm_bench.txt, Reddit layer 1
// on B200): X^T*G on the warp-specialised kernel (0.33 vs 0.42 ms for cuBLAS' SIMT SGEMM), X*W on cuBLAS (0.36 ms; the
// kernel here needs 0.53 ms: with N = 64 the three MMAs of the split re-read a 16 KB A tile from shared memory per 8 KB of
// B, and the split itself moves another 72 KB per k-block -- shared-memory bandwidth, not HBM, bounds it).
static int g_tc_gemm = -1;

template <bool TRANS, int CP, int RAW>
static int launch_gemm(const GemmArgs &g, dim3 grid, cudaStream_t st)
{
    const int smem = gemm_smem_bytes(g.npad, RAW);
    if (g_tc_gemm >= 2) {
        GNNA_CUDA_CHECK(cudaFuncSetAttribute(gemm_tf32x3_ws_kernel<TRANS, CP, RAW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        gemm_tf32x3_ws_kernel<TRANS, CP, RAW><<<grid, WS_THREADS, smem, st>>>(g);
    } else {
        GNNA_CUDA_CHECK(cudaFuncSetAttribute(gemm_tf32x3_kernel<TRANS, CP, RAW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        gemm_tf32x3_kernel<TRANS, CP, RAW><<<grid, GT_THREADS, smem, st>>>(g);
    }
    GNNA_CUDA_CHECK(cudaGetLastError());
    return GNNA_OK;
}

template <bool TRANS, int CP>
static int launch_gemm_raw(const GemmArgs &g, dim3 grid, cudaStream_t st)
{
    return g.npad <= 64 ? launch_gemm<TRANS, CP, 7>(g, grid, st) : launch_gemm<TRANS, CP, 5>(g, grid, st);
}

bool tc_gemm_enabled()
{
    if (g_tc_gemm < 0) {
        const char *e = getenv("GNNA_TC_GEMM");
        g_tc_gemm = (e && e[0] >= '0' && e[0] <= '3') ? e[0] - '0' : 3;
    }
    return g_tc_gemm >= 1;
}

// C[m, n] = op(A) * B on the tensor cores where the shape is one of the two tall-skinny contractions of a layer; returns
// GNNA_ERR_UNSUPPORTED otherwise (the caller then uses cuBLAS).  row_scale (NN only, may be null): C[i, :] *= row_scale[i].
int gemm_tf32x3(cudaStream_t st, bool ta, bool tb, int64_t m, int64_t n, int64_t k, const float *A, const float *B, float *C,
                const float *row_scale)
{
    if (tb || n < 1 || n > 128 || m < 1 || k < 1) return GNNA_ERR_UNSUPPORTED;
    if ((((uintptr_t)A | (uintptr_t)B | (uintptr_t)C) & 3) != 0) return GNNA_ERR_UNSUPPORTED;
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.A = A; g.B = B; g.C = C;
    g.N = (int)n;
    g.npad = ((int)n + 15) / 16 * 16;
    g.row_scale = row_scale;
    int dev = 0, sms = 148;
    GNNA_CUDA_CHECK(cudaGetDevice(&dev));
    GNNA_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (!ta) {
        // NN: worth it when A is tall (many row tiles) and K is a real contraction (for K < 256 the per-tile epilogue
        // dominates and cuBLAS' SIMT kernel is as fast); in the default mode cuBLAS keeps this product (see g_tc_gemm)
        if (g_tc_gemm == 3 || m < 8192 || k < 256 || k > 0x7fffffff) return GNNA_ERR_UNSUPPORTED;
        g.M = m; g.K = (int)k;
        g.work = (m + GT_M - 1) / GT_M;
        const dim3 grid((unsigned)(g.work < sms ? g.work : sms));
        const bool a16 = (k % 4 == 0) && (((uintptr_t)A & 15) == 0), a8 = (k % 2 == 0) && (((uintptr_t)A & 7) == 0);
        int rc;
        if (a16) rc = launch_gemm_raw<false, 16>(g, grid, st);
        else if (a8) rc = launch_gemm_raw<false, 8>(g, grid, st);
        else rc = launch_gemm_raw<false, 4>(g, grid, st);
        if (rc != GNNA_OK) return rc;
    } else {
        // TN: C[m, n] = A[k, m]^T B[k, n], reduced over k (the node dimension): split over the SMs, merged with reductions
        if (row_scale || k < 8192 || m < 256 || m > 0x7fffffff) return GNNA_ERR_UNSUPPORTED;
        g.M = k; g.K = (int)m;
        const int m_tiles = (int)((m + GT_M - 1) / GT_M);
        const long long total_kb = (k + GT_BK - 1) / GT_BK;
        int splits = sms / m_tiles;
        if (splits < 1) splits = 1;
        if (splits > total_kb) splits = (int)total_kb;
        g.work = (total_kb + splits - 1) / splits;
        g.splits = splits;
        GNNA_CUDA_CHECK(cudaMemsetAsync(C, 0, sizeof(float) * (size_t)m * (size_t)n, st));
        const int rc = launch_gemm_raw<true, 4>(g, dim3(m_tiles, splits), st);
        if (rc != GNNA_OK) return rc;
    }
    GNNA_CUDA_CHECK(cudaGetLastError());
    count_launch(1);
    return GNNA_OK;
}

}  // namespace gnna

extern "C" int gnna_set_tc_gemm(int mode)
{
    gnna::tc_gemm_enabled();
    const int prev = gnna::g_tc_gemm;
    gnna::g_tc_gemm = mode < 0 ? 0 : (mode > 3 ? 3 : mode);
    return prev;
}
