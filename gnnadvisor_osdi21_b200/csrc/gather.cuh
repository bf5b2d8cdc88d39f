// gather.cuh -- device building blocks shared by the aggregation kernels (aggregate.cu, fused_gemm.cu):
// batched, order-preserving 128-bit row gathers with the ids broadcast inside a sub-warp.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace gnna {

// ------------------------------------------------------------------------------------------
// vector loads of VEC elements of T through the read-only path, unpacked to fp32
// ------------------------------------------------------------------------------------------
template <int BYTES> struct Raw;
template <> struct Raw<16> { uint4 v; };
template <> struct Raw<8> { uint2 v; };
template <> struct Raw<4> { uint32_t v; };
template <> struct Raw<2> { uint16_t v; };

// Predicated read-only loads as volatile asm: the U*KCH loads of one batch must be ISSUED before
// the first is consumed (that is the memory-level parallelism this kernel lives on).  Written as
// `if (ok) x = __ldg(p); else x = 0;` the compiler sinks every load next to its use and keeps a
// single row in flight (seen in SASS: the same destination registers reused back to back).
__device__ __forceinline__ void ldg_raw(Raw<16> &r, const void *p, bool ok) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t"
                 "mov.b32 %0, 0;\n\tmov.b32 %1, 0;\n\tmov.b32 %2, 0;\n\tmov.b32 %3, 0;\n\t"
                 "@p ld.global.nc.v4.b32 {%0, %1, %2, %3}, [%4];\n\t}"
                 : "=r"(r.v.x), "=r"(r.v.y), "=r"(r.v.z), "=r"(r.v.w) : "l"(p), "r"((int)ok));
}
__device__ __forceinline__ void ldg_raw(Raw<8> &r, const void *p, bool ok) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %3, 0;\n\t"
                 "mov.b32 %0, 0;\n\tmov.b32 %1, 0;\n\t"
                 "@p ld.global.nc.v2.b32 {%0, %1}, [%2];\n\t}"
                 : "=r"(r.v.x), "=r"(r.v.y) : "l"(p), "r"((int)ok));
}
__device__ __forceinline__ void ldg_raw(Raw<4> &r, const void *p, bool ok) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t"
                 "mov.b32 %0, 0;\n\t"
                 "@p ld.global.nc.b32 %0, [%1];\n\t}"
                 : "=r"(r.v) : "l"(p), "r"((int)ok));
}
__device__ __forceinline__ void ldg_raw(Raw<2> &r, const void *p, bool ok) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t"
                 "mov.b16 %0, 0;\n\t"
                 "@p ld.global.nc.b16 %0, [%1];\n\t}"
                 : "=h"(r.v) : "l"(p), "r"((int)ok));
}

// unpredicated variants (fast path: whole batch valid) -- still volatile asm so they stay batched
__device__ __forceinline__ void ldg_raw(Raw<16> &r, const void *p) {
    asm volatile("ld.global.nc.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.v.x), "=r"(r.v.y), "=r"(r.v.z), "=r"(r.v.w) : "l"(p));
}
__device__ __forceinline__ void ldg_raw(Raw<8> &r, const void *p) {
    asm volatile("ld.global.nc.v2.b32 {%0, %1}, [%2];" : "=r"(r.v.x), "=r"(r.v.y) : "l"(p));
}
__device__ __forceinline__ void ldg_raw(Raw<4> &r, const void *p) {
    asm volatile("ld.global.nc.b32 %0, [%1];" : "=r"(r.v) : "l"(p));
}
__device__ __forceinline__ void ldg_raw(Raw<2> &r, const void *p) {
    asm volatile("ld.global.nc.b16 %0, [%1];" : "=h"(r.v) : "l"(p));
}
// index streams (col_idx, group table) are read exactly once: keep them out of L1
__device__ __forceinline__ int ldg_stream(const int32_t *p) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// small per-node tables read at the START of a group for its flush (row_ptr, degrees): volatile so they are issued
// where they are written, not sunk next to their use after the gather
__device__ __forceinline__ int ldg_keep(const int32_t *p) {
    int v;
    asm volatile("ld.global.nc.b32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float ldg_keep_f(const float *p) {
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ float bf16lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// fp32
__device__ __forceinline__ void unpack(const Raw<16> &r, float (&f)[4], float) {
    f[0] = __uint_as_float(r.v.x); f[1] = __uint_as_float(r.v.y); f[2] = __uint_as_float(r.v.z); f[3] = __uint_as_float(r.v.w);
}
__device__ __forceinline__ void unpack(const Raw<8> &r, float (&f)[2], float) {
    f[0] = __uint_as_float(r.v.x); f[1] = __uint_as_float(r.v.y);
}
__device__ __forceinline__ void unpack(const Raw<4> &r, float (&f)[1], float) { f[0] = __uint_as_float(r.v); }
// bf16
__device__ __forceinline__ void unpack(const Raw<16> &r, float (&f)[8], __nv_bfloat16) {
    f[0] = bf16lo(r.v.x); f[1] = bf16hi(r.v.x); f[2] = bf16lo(r.v.y); f[3] = bf16hi(r.v.y);
    f[4] = bf16lo(r.v.z); f[5] = bf16hi(r.v.z); f[6] = bf16lo(r.v.w); f[7] = bf16hi(r.v.w);
}
__device__ __forceinline__ void unpack(const Raw<8> &r, float (&f)[4], __nv_bfloat16) {
    f[0] = bf16lo(r.v.x); f[1] = bf16hi(r.v.x); f[2] = bf16lo(r.v.y); f[3] = bf16hi(r.v.y);
}
__device__ __forceinline__ void unpack(const Raw<4> &r, float (&f)[2], __nv_bfloat16) { f[0] = bf16lo(r.v); f[1] = bf16hi(r.v); }
__device__ __forceinline__ void unpack(const Raw<2> &r, float (&f)[1], __nv_bfloat16) { f[0] = bf16lo(r.v); }

// acc += bf16 elements, fp32 accumulation, with the mixed-precision add of PTX ISA 8.6 (sm_100+):
// `add.rn.f32.bf16 d, a, c` = fl(float(a) + c) in ONE instruction (SASS FHADD.BF16 with an .H0/.H1 selector),
// where the conversion is exact -- the same result as unpack + __fadd_rn at half the instruction count.
__device__ __forceinline__ void add_bf16x2(uint32_t w, float &lo_acc, float &hi_acc) {
    asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\t"
        "add.rn.f32.bf16 %0, lo, %0;\n\tadd.rn.f32.bf16 %1, hi, %1;\n\t}"
        : "+f"(lo_acc), "+f"(hi_acc) : "r"(w));
}
__device__ __forceinline__ void add_bf16(const Raw<16> &r, float (&a)[8]) {
    add_bf16x2(r.v.x, a[0], a[1]); add_bf16x2(r.v.y, a[2], a[3]); add_bf16x2(r.v.z, a[4], a[5]); add_bf16x2(r.v.w, a[6], a[7]);
}
__device__ __forceinline__ void add_bf16(const Raw<8> &r, float (&a)[4]) {
    add_bf16x2(r.v.x, a[0], a[1]); add_bf16x2(r.v.y, a[2], a[3]);
}
__device__ __forceinline__ void add_bf16(const Raw<4> &r, float (&a)[2]) { add_bf16x2(r.v, a[0], a[1]); }
__device__ __forceinline__ void add_bf16(const Raw<2> &r, float (&a)[1]) {
    asm("add.rn.f32.bf16 %0, %1, %0;" : "+f"(a[0]) : "h"(r.v));
}
// fp32 element types never take this path; the overloads only have to exist for the template to compile
__device__ __forceinline__ void add_bf16(const Raw<16> &, float (&)[4]) {}
__device__ __forceinline__ void add_bf16(const Raw<8> &, float (&)[2]) {}
__device__ __forceinline__ void add_bf16(const Raw<4> &, float (&)[1]) {}

// base + (int64)a * b as ONE IMAD.WIDE (the compiler emits IMAD.WIDE + a 64-bit add otherwise)
__device__ __forceinline__ const char *mad_wide(int a, int b, const char *base) {
    const char *r;
    asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(base));
    return r;
}

// One batch step: U neighbour rows of this sub-warp are loaded (all loads issued first), then summed in
// neighbour order.  PRED=false is the fast path (whole warp has U more neighbours, all chunks in range).
template <typename T, int VEC, int LPR, int KCH, int U, int IPL, bool WEIGHTED, bool PRED>
__device__ __forceinline__ void batch_step(const char *const (&lane_base)[KCH], int row_bytes, int nchunks, int chunk0, int j0,
                                           const int (&nid)[IPL], const float (&wgt)[IPL], float (&acc)[KCH][VEC])
{
    using RawT = Raw<VEC * (int)sizeof(T)>;
    constexpr unsigned FULL = 0xffffffffu;
    RawT raw[U][KCH];
    float w[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
        const int j = j0 + u;                        // compile-time after unrolling
        const int nj = __shfl_sync(FULL, nid[j / LPR], j % LPR, LPR);
        if (WEIGHTED) w[u] = __shfl_sync(FULL, wgt[j / LPR], j % LPR, LPR);
#pragma unroll
        for (int k = 0; k < KCH; k++) {
            // one IMAD.WIDE per load: lane_base[k] (64-bit, this lane's chunk inside row 0) + nj * row_bytes
            const char *p = mad_wide(nj, row_bytes, lane_base[k]);
            if (PRED) ldg_raw(raw[u][k], p, nj >= 0);
            else ldg_raw(raw[u][k], p);
        }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
#pragma unroll
        for (int k = 0; k < KCH; k++) {
            if constexpr (!WEIGHTED && sizeof(T) == 2) {
                add_bf16(raw[u][k], acc[k]);         // one FHADD.BF16 per element instead of unpack + FADD
            } else {
                float f[VEC];
                unpack(raw[u][k], f, T());
#pragma unroll
                for (int v = 0; v < VEC; v++) {
                    if (WEIGHTED) acc[k][v] = __fadd_rn(acc[k][v], __fmul_rn(w[u], f[v]));
                    else acc[k][v] = __fadd_rn(acc[k][v], f[v]);
                }
            }
        }
    }
}

}  // namespace gnna
