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
// All waits are bounded (about 4 s of GPU clock) and report through an error word instead of hanging; a producer that
// gave up poisons the consumer's error word too, so stale halo rows are never consumed silently.
#include <cuda_runtime.h>
#include <stdlib.h>

#include "common.h"

namespace gnna {

constexpr int MAX_PEERS = 16;
constexpr int PUSH_WARPS = 8;
constexpr int PUSH_ROWS_PER_CTA = 128;
constexpr long long SPIN_LIMIT_CYCLES = 8000000000LL;   // ~4 s at 2 GHz

// Indexed by SLOT, not rank: slot s serves peer (my_rank + 1 + s) % world; the kernel feeds my peers in ring
// order, so every receiver sees its chunks arrive one after the other -- which is what lets it start
// aggregating a peer's rows while the others are still in flight.
struct HaloPushParams {
    float *peer_base[MAX_PEERS];          // the peer's feature buffer for this step parity (mapped)
    unsigned *peer_flag[MAX_PEERS];       // &flags[my_rank] inside the peer's control block
    unsigned *peer_err[MAX_PEERS];        // the peer's error word: poisoned when I could not deliver its rows
    const unsigned *ack_from[MAX_PEERS];  // &acks[peer] inside MY control block (written by the peer)
    long long dst_row0[MAX_PEERS];        // row of the peer's buffer where my block of rows starts
    int send_lo[MAX_PEERS];               // my send list for the peer is send_idx[send_lo .. send_hi)
    int send_hi[MAX_PEERS];
    unsigned *done_counter;               // [MAX_PEERS] local scratch, zero between launches
    unsigned *error_word;                 // local: set to non-zero when a wait timed out
    int slots, dim;
    unsigned skip_mask;                   // slots (bit s) served by the copy engines (gnna_halo_push_ce): not this kernel's
    int interleave;                       // 1: CTA c starts at slot c % slots, so all peers are fed at once (see the kernel)
    const unsigned *step_ptr;             // local control word holding the current step (1, 2, 3, ...): read on the
                                          // device so that a captured CUDA graph can be replayed step after step
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

// Persistent: a fixed, small number of CTAs (prm.ctas, about one per SM or fewer) so that the push leaves most
// thread slots to the aggregation kernel running next to it.  All CTAs serve slot 0 first, then slot 1, ...:
// the peers' flags go up one after the other.
__global__ void __launch_bounds__(PUSH_WARPS * 32)
halo_push_kernel(const float *__restrict__ x_local, const long long *__restrict__ send_idx, HaloPushParams prm)
{
    __shared__ int s_ok;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned step = *reinterpret_cast<const volatile unsigned *>(prm.step_ptr);
    // interleave = 0 (default): all CTAs serve slot 0, then slot 1, ... -- blocks arrive one after the other, the first one
    // soonest, which is what the consumer's segment-by-segment aggregation wants.  1 (GNNA_PUSH_INTERLEAVE=1): the CTAs start
    // at different slots and walk the ring from there, so every peer is written at any moment; measured on 4xB200 the whole
    // exchange is then 5 % shorter (0.130 vs 0.136 ms) and the overlapped step no faster (profiles/r02_v4_sweep_4gpu.txt).
    const int first = prm.interleave ? (int)(blockIdx.x % (unsigned)prm.slots) : 0;
    for (int pi = 0; pi < prm.slots; pi++) {                 // slot
        const int p = (pi + first) % prm.slots;
        if ((prm.skip_mask >> p) & 1u) continue;
        // the buffer of this parity was last read by this peer at step-2: wait for its acknowledgement
        if (threadIdx.x == 0) {
            int ok = 1;
            if (step > 2) {
                const long long t0 = clock64();
                while (ld_acquire_sys(prm.ack_from[p]) + 2 < step) {
                    if (clock64() - t0 > SPIN_LIMIT_CYCLES) {
                        // the rows are NOT delivered: say so on both sides.  The flag below still goes up so the
                        // consumer does not spin as well, but its error word now reads 3 ("halo rows of this step are
                        // stale") -- PeerHalo.error() / ShardedGraph.check() turn that into an exception on the host.
                        ok = 0;
                        atomicExch(prm.error_word, 1u);
                        st_release_sys(prm.peer_err[p], 3u);
                        break;
                    }
                }
            }
            s_ok = ok;
        }
        __syncthreads();
        if (s_ok) {
            float *dst_base = prm.peer_base[p] + (prm.dst_row0[p] - prm.send_lo[p]) * (long long)prm.dim;
            const int rows = prm.send_hi[p] - prm.send_lo[p];
            const int chunks = (rows + PUSH_ROWS_PER_CTA - 1) / PUSH_ROWS_PER_CTA;
            for (int chunk = blockIdx.x; chunk < chunks; chunk += gridDim.x) {
                const int r0 = prm.send_lo[p] + chunk * PUSH_ROWS_PER_CTA;
                const int r1 = min(r0 + PUSH_ROWS_PER_CTA, prm.send_hi[p]);
                if ((prm.dim & 3) == 0) {
                    // 128-bit copies; `lpr` lanes per row (power of two), 32/lpr rows per warp pass, 4 passes in flight
                    const int c4 = prm.dim >> 2;
                    int lpr = 1;
                    while (lpr < c4 && lpr < 32) lpr <<= 1;
                    const int rpw = 32 / lpr, sub = lane / lpr, l = lane % lpr;
                    constexpr int UNROLL = 4;
                    for (int r = r0 + warp * rpw + sub; r < r1; r += PUSH_WARPS * rpw * UNROLL) {
                        for (int c = l; c < c4; c += lpr) {
                            float4 v[UNROLL];
#pragma unroll
                            for (int u = 0; u < UNROLL; u++) {
                                const int rr = r + u * PUSH_WARPS * rpw;
                                if (rr < r1) v[u] = __ldg(reinterpret_cast<const float4 *>(x_local + send_idx[rr] * prm.dim) + c);
                            }
#pragma unroll
                            for (int u = 0; u < UNROLL; u++) {
                                const int rr = r + u * PUSH_WARPS * rpw;
                                if (rr < r1) reinterpret_cast<float4 *>(dst_base + (long long)rr * prm.dim)[c] = v[u];
                            }
                        }
                    }
                } else {
                    for (int r = r0 + warp; r < r1; r += PUSH_WARPS) {
                        const float *src = x_local + send_idx[r] * prm.dim;
                        float *dst = dst_base + (long long)r * prm.dim;
                        for (int c = lane; c < prm.dim; c += 32) dst[c] = __ldg(src + c);
                    }
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();                          // my CTA's remote stores are visible system-wide
            const unsigned done = atomicAdd(prm.done_counter + p, 1u);
            if (done == gridDim.x - 1) {                     // last CTA done with this peer: raise its flag
                prm.done_counter[p] = 0;
                __threadfence_system();
                st_release_sys(prm.peer_flag[p], step);
            }
        }
    }
}

// one thread per peer: wait until flags[q] reaches `step` (only the peers in `peer_mask`)
__global__ void halo_wait_kernel(const unsigned *flags, int world, int my_rank, unsigned peer_mask, const unsigned *step_ptr,
                                 unsigned *error_word)
{
    const int q = threadIdx.x;
    if (q >= world || q == my_rank || !((peer_mask >> q) & 1u)) return;
    const unsigned step = *reinterpret_cast<const volatile unsigned *>(step_ptr);
    const long long t0 = clock64();
    while (ld_acquire_sys(flags + q) < step) {
        if (clock64() - t0 > SPIN_LIMIT_CYCLES) { atomicExch(error_word, 2u); break; }
    }
}

struct HaloAckParams {
    unsigned *peer_ack[MAX_PEERS];   // &acks[my_rank] inside peer p's control block
    int world, my_rank;
    const unsigned *step_ptr;
};

// after the aggregation that consumed step `step`: tell every producer its buffer may be reused
__global__ void halo_ack_kernel(HaloAckParams prm)
{
    const int p = threadIdx.x;
    if (p >= prm.world || p == prm.my_rank) return;
    __threadfence_system();
    st_release_sys(prm.peer_ack[p], *reinterpret_cast<const volatile unsigned *>(prm.step_ptr));
}

// ---- copy-engine exchange (dense halos) ---------------------------------------------------------------------------
// When a peer needs (nearly) ALL of this rank's rows -- every peer does on a dense graph: 92 % of the rows on the Reddit
// look-alike at 8 GPUs -- the "gather the wanted rows" kernel is the wrong tool: it occupies SMs the aggregation wants
// and tops out near 350 GB/s per rank.  The receiver then simply asks for the whole vertex range
// (dist.ShardedGraph: dense halo), the block a peer wants is this rank's local rows as they lie in memory, and ONE
// cudaMemcpyAsync per peer moves it over NVLink on the copy engines: no SM, no index list.  Three tiny kernels keep the
// protocol of the push kernel: wait for the peers' acknowledgements of step-2 (the buffer parity about to be overwritten),
// then after each copy raise that peer's flag.
struct HaloCeParams {
    const unsigned *ack_from[MAX_PEERS];
    unsigned *peer_err[MAX_PEERS];
    int n;
    const unsigned *step_ptr;
    unsigned *error_word;
};

__global__ void halo_wait_acks_kernel(HaloCeParams prm)
{
    const int p = threadIdx.x;
    if (p >= prm.n) return;
    const unsigned step = *reinterpret_cast<const volatile unsigned *>(prm.step_ptr);
    if (step <= 2) return;
    const long long t0 = clock64();
    while (ld_acquire_sys(prm.ack_from[p]) + 2 < step) {
        if (clock64() - t0 > SPIN_LIMIT_CYCLES) {
            atomicExch(prm.error_word, 1u);
            st_release_sys(prm.peer_err[p], 3u);     // the copy that follows lands in a buffer the peer may still read
            break;
        }
    }
}

__global__ void halo_raise_flag_kernel(unsigned *peer_flag, const unsigned *step_ptr)
{
    __threadfence_system();
    st_release_sys(peer_flag, *reinterpret_cast<const volatile unsigned *>(step_ptr));
}

static int push_interleave()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("GNNA_PUSH_INTERLEAVE");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v;
}

// first kernel of a step on the compute stream: step += 1 (everything else of the step is ordered after it)
__global__ void halo_bump_kernel(unsigned *step_ptr) { *step_ptr = *step_ptr + 1; }

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
// [16..31] acks (written by consumers), [32..47] done counters (local), [48] error word (local),
// [49] current step (local; incremented by gnna_halo_begin_step, read by push / wait / ack on the device)
extern "C" int gnna_halo_push_f32(const float *x_local, const int64_t *send_idx, const int32_t *send_begin_host,
                                  void *const *peer_feature_base_host, void *const *peer_ctrl_host,
                                  const int64_t *peer_dst_row0_host, void *my_ctrl,
                                  int world, int my_rank, int dim, void *stream)
{
    GNNA_REQUIRE(world >= 1 && world <= MAX_PEERS && my_rank >= 0 && my_rank < world, "halo_push: bad world/rank");
    GNNA_REQUIRE(dim > 0, "halo_push: bad dim");
    if (world == 1) return GNNA_OK;
    GNNA_REQUIRE(x_local && send_begin_host && peer_feature_base_host && peer_ctrl_host && peer_dst_row0_host && my_ctrl,
                 "halo_push: null pointer");
    HaloPushParams prm;
    memset(&prm, 0, sizeof(prm));
    unsigned *ctrl = (unsigned *)my_ctrl;
    int max_chunks = 1;
    for (int sl = 0; sl < world - 1; sl++) {
        const int p = (my_rank + 1 + sl) % world;
        const int rows = send_begin_host[p + 1] - send_begin_host[p];
        const int chunks = (rows + PUSH_ROWS_PER_CTA - 1) / PUSH_ROWS_PER_CTA;
        if (chunks > max_chunks) max_chunks = chunks;
        prm.send_lo[sl] = send_begin_host[p];
        prm.send_hi[sl] = send_begin_host[p + 1];
        prm.peer_base[sl] = (float *)peer_feature_base_host[p];
        prm.peer_flag[sl] = (unsigned *)peer_ctrl_host[p] + my_rank;
        prm.peer_err[sl] = (unsigned *)peer_ctrl_host[p] + 48;
        prm.ack_from[sl] = ctrl + 16 + p;
        prm.dst_row0[sl] = peer_dst_row0_host[p];
    }
    static int push_ctas = 0;
    if (push_ctas == 0) {
        const char *e = getenv("GNNA_PUSH_CTAS");
        push_ctas = e ? atoi(e) : 0;
        if (push_ctas <= 0) push_ctas = 96;        // < 148 SMs x 8 resident: most thread slots stay free for the aggregation
    }
    const int ctas = push_ctas < max_chunks ? push_ctas : max_chunks;
    GNNA_REQUIRE(send_idx || send_begin_host[world] == 0, "halo_push: null send_idx");
    prm.done_counter = ctrl + 32;
    prm.error_word = ctrl + 48;
    prm.slots = world - 1;
    prm.dim = dim;
    prm.step_ptr = ctrl + 49;
    prm.interleave = push_interleave();
    halo_push_kernel<<<ctas, PUSH_WARPS * 32, 0, (cudaStream_t)stream>>>(x_local, (const long long *)send_idx, prm);
    GNNA_CUDA_CHECK(cudaGetLastError());
    count_launch(1);
    return GNNA_OK;
}

// Same contract as gnna_halo_push_f32.  Peers in `dense_mask` (bit = RANK) receive this rank's n_local rows as ONE
// device-to-device copy (they asked for the whole range: their block for this rank is n_local rows long); the others are
// served by the push kernel from their send lists.
extern "C" int gnna_halo_push_ce(const float *x_local, int64_t n_local, const int64_t *send_idx, const int32_t *send_begin_host,
                                 void *const *peer_feature_base_host, void *const *peer_ctrl_host,
                                 const int64_t *peer_dst_row0_host, void *my_ctrl,
                                 int world, int my_rank, int dim, uint32_t dense_mask, void *stream)
{
    GNNA_REQUIRE(world >= 1 && world <= MAX_PEERS && my_rank >= 0 && my_rank < world, "halo_push_ce: bad world/rank");
    GNNA_REQUIRE(dim > 0 && n_local >= 0, "halo_push_ce: bad size");
    if (world == 1) return GNNA_OK;
    GNNA_REQUIRE(x_local && send_begin_host && peer_feature_base_host && peer_ctrl_host && peer_dst_row0_host && my_ctrl,
                 "halo_push_ce: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned *ctrl = (unsigned *)my_ctrl;
    HaloCeParams ce;
    memset(&ce, 0, sizeof(ce));
    HaloPushParams prm;
    memset(&prm, 0, sizeof(prm));
    int max_chunks = 0, sparse = 0;
    for (int sl = 0; sl < world - 1; sl++) {
        const int p = (my_rank + 1 + sl) % world;
        const int rows = send_begin_host[p + 1] - send_begin_host[p];
        prm.send_lo[sl] = send_begin_host[p];
        prm.send_hi[sl] = send_begin_host[p + 1];
        prm.peer_base[sl] = (float *)peer_feature_base_host[p];
        prm.peer_flag[sl] = (unsigned *)peer_ctrl_host[p] + my_rank;
        prm.peer_err[sl] = (unsigned *)peer_ctrl_host[p] + 48;
        prm.ack_from[sl] = ctrl + 16 + p;
        prm.dst_row0[sl] = peer_dst_row0_host[p];
        ce.ack_from[sl] = ctrl + 16 + p;
        ce.peer_err[sl] = prm.peer_err[sl];
        if ((dense_mask >> p) & 1u) {
            GNNA_REQUIRE(rows == n_local, "halo_push_ce: peer %d is marked dense but wants %d of %lld rows", p, rows, (long long)n_local);
            prm.skip_mask |= 1u << sl;
        } else {
            const int chunks = (rows + PUSH_ROWS_PER_CTA - 1) / PUSH_ROWS_PER_CTA;
            if (chunks > max_chunks) max_chunks = chunks;
            sparse++;
        }
    }
    ce.n = world - 1;
    ce.step_ptr = ctrl + 49;
    ce.error_word = ctrl + 48;
    halo_wait_acks_kernel<<<1, 32, 0, st>>>(ce);
    GNNA_CUDA_CHECK(cudaGetLastError());
    count_launch(1);
    // One point-to-point copy of a 7.5-15 MB block runs at ~270 GB/s (measured on 4 and 8 B200s); the copies are issued one
    // after the other on the caller's stream, in ring order, so every receiver sees its blocks arrive in turn.  (Running
    // them concurrently on side streams forked inside a stream capture hung the 4-GPU bench in round 2 and was removed.)
    const size_t bytes = (size_t)n_local * (size_t)dim * sizeof(float);
    for (int sl = 0; sl < world - 1; sl++) {
        if (!((prm.skip_mask >> sl) & 1u)) continue;
        if (bytes)
            GNNA_CUDA_CHECK(cudaMemcpyAsync(prm.peer_base[sl] + prm.dst_row0[sl] * (long long)dim, x_local, bytes,
                                            cudaMemcpyDeviceToDevice, st));
        halo_raise_flag_kernel<<<1, 1, 0, st>>>(prm.peer_flag[sl], ctrl + 49);
        GNNA_CUDA_CHECK(cudaGetLastError());
        count_launch(1);
    }
    if (sparse > 0) {
        GNNA_REQUIRE(send_idx || max_chunks == 0, "halo_push_ce: null send_idx");
        int ctas = 96;
        const char *e = getenv("GNNA_PUSH_CTAS");
        if (e && atoi(e) > 0) ctas = atoi(e);
        if (ctas > max_chunks) ctas = max_chunks > 0 ? max_chunks : 1;
        prm.done_counter = ctrl + 32;
        prm.error_word = ctrl + 48;
        prm.slots = world - 1;
        prm.dim = dim;
        prm.step_ptr = ctrl + 49;
        prm.interleave = push_interleave();
        halo_push_kernel<<<ctas, PUSH_WARPS * 32, 0, st>>>(x_local, (const long long *)send_idx, prm);
        GNNA_CUDA_CHECK(cudaGetLastError());
        count_launch(1);
    }
    return GNNA_OK;
}

extern "C" int gnna_halo_begin_step(void *my_ctrl, void *stream)
{
    GNNA_REQUIRE(my_ctrl, "halo_begin_step: null control block");
    halo_bump_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((unsigned *)my_ctrl + 49);
    GNNA_CUDA_CHECK(cudaGetLastError());
    count_launch(1);
    return GNNA_OK;
}

extern "C" int gnna_halo_wait(void *my_ctrl, int world, int my_rank, uint32_t peer_mask, void *stream)
{
    GNNA_REQUIRE(my_ctrl && world >= 1 && world <= MAX_PEERS, "halo_wait: bad argument");
    if (world == 1) return GNNA_OK;
    unsigned *ctrl = (unsigned *)my_ctrl;
    if (peer_mask == 0) peer_mask = 0xffffffffu;
    halo_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(ctrl, world, my_rank, peer_mask, ctrl + 49, ctrl + 48);
    GNNA_CUDA_CHECK(cudaGetLastError());
    count_launch(1);
    return GNNA_OK;
}

extern "C" int gnna_halo_ack(void *const *peer_ctrl_host, void *my_ctrl, int world, int my_rank, void *stream)
{
    GNNA_REQUIRE(peer_ctrl_host && my_ctrl && world >= 1 && world <= MAX_PEERS, "halo_ack: bad argument");
    if (world == 1) return GNNA_OK;
    HaloAckParams prm;
    memset(&prm, 0, sizeof(prm));
    for (int p = 0; p < world; p++)
        if (p != my_rank) prm.peer_ack[p] = (unsigned *)peer_ctrl_host[p] + 16 + my_rank;
    prm.world = world;
    prm.my_rank = my_rank;
    prm.step_ptr = (const unsigned *)my_ctrl + 49;
    halo_ack_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(prm);
    GNNA_CUDA_CHECK(cudaGetLastError());
    count_launch(1);
    return GNNA_OK;
}
