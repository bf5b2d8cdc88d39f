// halo.cu -- NVLink-native halo exchange for the sharded aggregation path (no reference counterpart:
// the reference is single-GPU).
//
// The NCCL formulation is gather(send rows) -> all_to_all_single -> (receiver) aggregate: two extra
// passes over the halo bytes plus a collective launch per aggregation.  Here every rank maps its peers'
// feature buffers (CUDA IPC) and ONE kernel writes each row a peer needs straight into that peer's halo
// rows over NVLink (128-bit remote stores), then raises a per-source flag in the peer's memory
// (release at system scope).  The consumer waits for its world-1 flags on its compute stream and runs
// the aggregation kernel on [own rows | halo rows] unchanged.  Halo buffers are double-buffered by step
// parity; a producer may overwrite parity b only after the consumer acknowledged the step that last
// read it (ack counters, also in peer memory), so no rank can run more than one step ahead.
// All waits are bounded (about 4 s of GPU clock) and report through an error word instead of hanging.
#include <cuda_runtime.h>

#include "common.h"

namespace gnna {

constexpr int MAX_PEERS = 16;
constexpr int PUSH_WARPS = 8;
constexpr int PUSH_ROWS_PER_CTA = 64;
constexpr long long SPIN_LIMIT_CYCLES = 8000000000LL;   // ~4 s at 2 GHz

struct HaloPushParams {
    float *peer_base[MAX_PEERS];          // peer p's feature buffer for this step parity (mapped)
    unsigned *peer_flag[MAX_PEERS];       // &flags[my_rank] inside peer p's control block
    const unsigned *ack_from[MAX_PEERS];  // &acks[p] inside MY control block (written by peer p)
    long long dst_row0[MAX_PEERS];        // row of peer p's buffer where my block of rows starts
    int send_begin[MAX_PEERS + 1];        // my send list is send_idx[send_begin[p] .. send_begin[p+1])
    int cta_begin[MAX_PEERS + 1];         // CTAs [cta_begin[p], cta_begin[p+1]) serve peer p
    unsigned *done_counter;               // [MAX_PEERS] local scratch, zero between launches
    unsigned *error_word;                 // local: set to non-zero when a wait timed out
    int world, my_rank, dim;
    unsigned step;                        // 1, 2, 3, ...
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(PUSH_WARPS * 32)
halo_push_kernel(const float *__restrict__ x_local, const long long *__restrict__ send_idx, HaloPushParams prm)
{
    __shared__ int s_ok;
    int p = 0;
    while (p + 1 < prm.world && (int)blockIdx.x >= prm.cta_begin[p + 1]) p++;
    while (p < prm.world && prm.cta_begin[p + 1] == prm.cta_begin[p]) p++;   // skip peers without CTAs (myself)
    const int chunk = blockIdx.x - prm.cta_begin[p];
    const int ctas_p = prm.cta_begin[p + 1] - prm.cta_begin[p];

    // the buffer of this parity was last read by peer p at step-2: wait for its acknowledgement
    if (threadIdx.x == 0) {
        int ok = 1;
        if (prm.step > 2) {
            const long long t0 = clock64();
            while (ld_acquire_sys(prm.ack_from[p]) + 2 < prm.step) {
                if (clock64() - t0 > SPIN_LIMIT_CYCLES) { ok = 0; atomicExch(prm.error_word, 1u); break; }
            }
        }
        s_ok = ok;
    }
    __syncthreads();

    if (s_ok) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int r0 = prm.send_begin[p] + chunk * PUSH_ROWS_PER_CTA;
        const int r1 = min(r0 + PUSH_ROWS_PER_CTA, prm.send_begin[p + 1]);
        float *dst_base = prm.peer_base[p] + (prm.dst_row0[p] - prm.send_begin[p]) * (long long)prm.dim;
        if ((prm.dim & 3) == 0) {
            const int c4 = prm.dim >> 2;
            for (int r = r0 + warp; r < r1; r += PUSH_WARPS) {
                const float4 *src = reinterpret_cast<const float4 *>(x_local + send_idx[r] * prm.dim);
                float4 *dst = reinterpret_cast<float4 *>(dst_base + (long long)r * prm.dim);
                for (int c = lane; c < c4; c += 32) dst[c] = __ldg(src + c);
            }
        } else {
            for (int r = r0 + warp; r < r1; r += PUSH_WARPS) {
                const float *src = x_local + send_idx[r] * prm.dim;
                float *dst = dst_base + (long long)r * prm.dim;
                for (int c = lane; c < prm.dim; c += 32) dst[c] = __ldg(src + c);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();                              // my CTA's remote stores are visible system-wide
        const unsigned done = atomicAdd(prm.done_counter + p, 1u);
        if (done == (unsigned)ctas_p - 1) {                  // last CTA serving peer p: raise the flag
            prm.done_counter[p] = 0;
            __threadfence_system();
            st_release_sys(prm.peer_flag[p], prm.step);
        }
    }
}

// one thread per peer: wait until flags[q] reaches `step`
__global__ void halo_wait_kernel(const unsigned *flags, int world, int my_rank, unsigned step, unsigned *error_word)
{
    const int q = threadIdx.x;
    if (q >= world || q == my_rank) return;
    const long long t0 = clock64();
    while (ld_acquire_sys(flags + q) < step) {
        if (clock64() - t0 > SPIN_LIMIT_CYCLES) { atomicExch(error_word, 2u); break; }
    }
}

struct HaloAckParams {
    unsigned *peer_ack[MAX_PEERS];   // &acks[my_rank] inside peer p's control block
    int world, my_rank;
    unsigned step;
};

// after the aggregation that consumed step `step`: tell every producer its buffer may be reused
__global__ void halo_ack_kernel(HaloAckParams prm)
{
    const int p = threadIdx.x;
    if (p >= prm.world || p == prm.my_rank) return;
    __threadfence_system();
    st_release_sys(prm.peer_ack[p], prm.step);
}

}  // namespace gnna

using namespace gnna;

// ---- CUDA IPC plumbing (device memory that peers on the same node can map) ---------------------
extern "C" int gnna_ipc_alloc(int64_t bytes, void **ptr, unsigned char *handle64)
{
    GNNA_REQUIRE(bytes > 0 && ptr && handle64, "gnna_ipc_alloc: bad argument");
    GNNA_CUDA_CHECK(cudaMalloc(ptr, (size_t)bytes));
    GNNA_CUDA_CHECK(cudaMemset(*ptr, 0, (size_t)bytes));
    cudaIpcMemHandle_t h;
    GNNA_CUDA_CHECK(cudaIpcGetMemHandle(&h, *ptr));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle64, &h, 64);
    return GNNA_OK;
}

extern "C" int gnna_ipc_open(const unsigned char *handle64, void **ptr)
{
    GNNA_REQUIRE(handle64 && ptr, "gnna_ipc_open: bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    GNNA_CUDA_CHECK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return GNNA_OK;
}

extern "C" int gnna_ipc_close(void *ptr)
{
    if (ptr) GNNA_CUDA_CHECK(cudaIpcCloseMemHandle(ptr));
    return GNNA_OK;
}

extern "C" int gnna_ipc_free(void *ptr)
{
    if (ptr) GNNA_CUDA_CHECK(cudaFree(ptr));
    return GNNA_OK;
}

// ---- one halo exchange step ------------------------------------------------------------------
// control block layout (uint32, in IPC memory of every rank): [0..15] flags (written by producers),
// [16..31] acks (written by consumers), [32..47] done counters (local), [48] error word (local)
extern "C" int gnna_halo_push_f32(const float *x_local, const int64_t *send_idx, const int32_t *send_begin_host,
                                  void *const *peer_feature_base_host, void *const *peer_ctrl_host,
                                  const int64_t *peer_dst_row0_host, void *my_ctrl,
                                  int world, int my_rank, int dim, uint32_t step, void *stream)
{
    GNNA_REQUIRE(world >= 1 && world <= MAX_PEERS && my_rank >= 0 && my_rank < world, "halo_push: bad world/rank");
    GNNA_REQUIRE(dim > 0 && step >= 1, "halo_push: bad dim/step");
    if (world == 1) return GNNA_OK;
    GNNA_REQUIRE(x_local && send_begin_host && peer_feature_base_host && peer_ctrl_host && peer_dst_row0_host && my_ctrl,
                 "halo_push: null pointer");
    HaloPushParams prm;
    memset(&prm, 0, sizeof(prm));
    unsigned *ctrl = (unsigned *)my_ctrl;
    int ctas = 0;
    for (int p = 0; p < world; p++) {
        prm.send_begin[p] = send_begin_host[p];
        prm.cta_begin[p] = ctas;
        if (p != my_rank) {
            const int rows = send_begin_host[p + 1] - send_begin_host[p];
            ctas += rows > 0 ? (rows + PUSH_ROWS_PER_CTA - 1) / PUSH_ROWS_PER_CTA : 1;   // >= 1: it raises the flag
            prm.peer_base[p] = (float *)peer_feature_base_host[p];
            prm.peer_flag[p] = (unsigned *)peer_ctrl_host[p] + my_rank;
            prm.ack_from[p] = ctrl + 16 + p;
            prm.dst_row0[p] = peer_dst_row0_host[p];
        }
    }
    prm.send_begin[world] = send_begin_host[world];
    prm.cta_begin[world] = ctas;
    GNNA_REQUIRE(send_idx || send_begin_host[world] == 0, "halo_push: null send_idx");
    prm.done_counter = ctrl + 32;
    prm.error_word = ctrl + 48;
    prm.world = world;
    prm.my_rank = my_rank;
    prm.dim = dim;
    prm.step = step;
    halo_push_kernel<<<ctas, PUSH_WARPS * 32, 0, (cudaStream_t)stream>>>(x_local, (const long long *)send_idx, prm);
    GNNA_CUDA_CHECK(cudaGetLastError());
    count_launch(1);
    return GNNA_OK;
}

extern "C" int gnna_halo_wait(void *my_ctrl, int world, int my_rank, uint32_t step, void *stream)
{
    GNNA_REQUIRE(my_ctrl && world >= 1 && world <= MAX_PEERS, "halo_wait: bad argument");
    if (world == 1) return GNNA_OK;
    unsigned *ctrl = (unsigned *)my_ctrl;
    halo_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(ctrl, world, my_rank, step, ctrl + 48);
    GNNA_CUDA_CHECK(cudaGetLastError());
    count_launch(1);
    return GNNA_OK;
}

extern "C" int gnna_halo_ack(void *const *peer_ctrl_host, int world, int my_rank, uint32_t step, void *stream)
{
    GNNA_REQUIRE(peer_ctrl_host && world >= 1 && world <= MAX_PEERS, "halo_ack: bad argument");
    if (world == 1) return GNNA_OK;
    HaloAckParams prm;
    memset(&prm, 0, sizeof(prm));
    for (int p = 0; p < world; p++)
        if (p != my_rank) prm.peer_ack[p] = (unsigned *)peer_ctrl_host[p] + 16 + my_rank;
    prm.world = world;
    prm.my_rank = my_rank;
    prm.step = step;
    halo_ack_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(prm);
    GNNA_CUDA_CHECK(cudaGetLastError());
    count_launch(1);
    return GNNA_OK;
}
