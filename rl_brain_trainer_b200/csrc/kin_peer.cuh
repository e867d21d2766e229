// kin_peer.cuh -- receive-buffer layout of the NVLink peer-memory gradient exchange and the in-kernel ("fused") form of it:
// the tail every CTA of the gradient kernel runs after its partial row is written (csrc/kin_ppo_tc.cu), replacing the separate
// kin_peer_push_kernel + kin_peer_gather_kernel launches between the gradient kernel and Adam.
//
//   receive buffer of a rank (cudaMalloc'ed by kin_peer_buffer_create, opened by the peers through CUDA IPC):
//     [0]      unsigned arrived[8]            two-kernel form: arrived[s] = pushes of sender s that have fully landed here
//     [64]     unsigned local_count           two-kernel form: CTAs of this rank's push kernel that have finished
//     [68]     unsigned grid_count            fused form: CTAs of this rank's gradient kernel whose partial row is complete (monotonic)
//     [128]    unsigned slice_flag[8][512]    fused form: slice_flag[s][c] = last exchange whose slice c sender s has delivered here
//     [16512]  float    slot[2][world][row]   row = n_params + 8 (gradient, then the 5 loss statistics); slot = exchange parity
#pragma once

#include "kin_internal.h"
#include "kin_ppo_layout.cuh"

namespace kin {

constexpr int PEER_MAX = 8;
constexpr int PEER_MAX_CTA = 512;
constexpr size_t PEER_FLAGS = 128;
constexpr size_t PEER_HEADER = PEER_FLAGS + sizeof(unsigned) * PEER_MAX * PEER_MAX_CTA;      // 16 512 bytes

struct PeerTable {
    unsigned char* base[PEER_MAX];
};

// arguments of the fused exchange (world == 0: no exchange, the caller reduces `partials` itself)
struct PeerFused {
    PeerTable peers;
    int rank, world;
    unsigned epoch;
    float* grad;                      // [P] this rank's copy of the summed gradient
    float* stats;                     // [KIN_PPO_STATS] (5 summed loss statistics, slot KIN_PPO_STAT_SKIP on a timeout)
    int* timed_out;                   // sticky device flag
    unsigned long long timeout_cycles;
};

unsigned long long kin_peer_timeout_cycles();      // kin_peer.cu

__host__ __device__ inline int peer_row(int P) { return (P + KIN_PPO_STATS + 3) & ~3; }

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Tail of the gradient kernel, called by ALL threads of EVERY CTA (blockDim.x >= 256) after the CTA's part of its partial row is in
// global memory.  Needs every CTA of the grid co-resident (it contains a grid barrier): the gradient kernel's grid is <= 2 CTAs / SM.
//   1. grid barrier: all partial rows of this rank are complete;
//   2. CTA c reduces its slice of the P + 5 columns over the rows -- in the order kin_peer_push_kernel / kin_ppo_reduce_kernel use
//      (8 interleaved row groups, folded in warp order), so the result is bitwise the two-kernel path's -- and stores it into
//      slot[epoch & 1][rank] of every peer (posted NVLink stores), then publishes slice_flag[rank][c] = epoch at every peer;
//   3. waits for slice_flag[r][c] >= epoch from every rank r and writes grad / stats = sum over ranks in RANK ORDER (bitwise
//      identical on every rank).  A peer that never delivers sets the sticky timeout flag and marks the minibatch to be skipped.
__device__ __forceinline__ void peer_exchange_tail(const PeerFused& px, const float* __restrict__ partials, int n_rows, int P, float inv_global_batch,
                                                   float (*part)[32], int* flag_smem) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int n_cta = (int)(gridDim.x * gridDim.y), cta = (int)(blockIdx.y * gridDim.x + blockIdx.x);
    unsigned char* own = px.peers.base[px.rank];
    // ---- 1. grid barrier (monotonic counter: exchange e completes at e * n_cta)
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        unsigned* cnt = reinterpret_cast<unsigned*>(own + 68);
        atomicAdd(cnt, 1u);
        const unsigned target = px.epoch * (unsigned)n_cta;
        const long long t0 = clock64();
        while ((int)(ld_acquire_gpu(cnt) - target) < 0) {
            if ((unsigned long long)(clock64() - t0) > px.timeout_cycles) { atomicExch(px.timed_out, 1); break; }
        }
    }
    __syncthreads();
    // ---- 2. this CTA's column slice: reduce over the rows, push to every peer
    const int cols = P + 5, prow = P + KIN_PPO_STATS + 8;
    const int per = ((cols + n_cta - 1) / n_cta + 31) & ~31;      // whole 32-column blocks per CTA
    const int p0 = cta * per, p1 = min(p0 + per, cols);
    for (int base = p0; base < p1; base += 32) {
        const int p = base + lane;
        float a0 = 0.0f, a1 = 0.0f;
        if (w < 8 && p < cols) {
            // rows w, w + 16, ... into a0 and w + 8, w + 24, ... into a1, added in that order (kin_ppo_reduce_kernel's order); the loads
            // of a batch are all issued before the first add, so the L2 latency is paid once per batch, not once per row
            constexpr int B = 6;
            int c = w;
            while (c < n_rows) {
                float v0[B], v1[B];
#pragma unroll
                for (int k = 0; k < B; ++k) {
                    const int r0 = c + 16 * k, r1 = r0 + 8;
                    v0[k] = r0 < n_rows ? __ldcg(partials + (size_t)r0 * prow + p) : 0.0f;
                    v1[k] = (r1 < n_rows && r0 + 8 < n_rows) ? __ldcg(partials + (size_t)r1 * prow + p) : 0.0f;
                }
#pragma unroll
                for (int k = 0; k < B; ++k) {
                    const int r0 = c + 16 * k;
                    if (r0 + 8 < n_rows) { a0 += v0[k]; a1 += v1[k]; }
                    else if (r0 < n_rows) a0 += v0[k];
                }
                c += 16 * B;
            }
        }
        if (w < 8) part[w][lane] = a0 + a1;
        __syncthreads();
        if (w < px.world && p < cols) {
            float a = part[0][lane];
#pragma unroll
            for (int k = 1; k < 8; ++k) a += part[k][lane];
            if (p >= P) a *= inv_global_batch;
            float* slot = reinterpret_cast<float*>(px.peers.base[w] + PEER_HEADER) + ((size_t)(px.epoch & 1u) * px.world + px.rank) * peer_row(P);
            slot[p] = a;
        }
        __syncthreads();
    }
    // the release stores below are the only system-scope fences of the CTA: release is cumulative, so coming after the CTA barrier it
    // also orders the other threads' slot stores before the flag
    if (tid < px.world) {
        unsigned* flag = reinterpret_cast<unsigned*>(px.peers.base[tid] + PEER_FLAGS) + px.rank * PEER_MAX_CTA + cta;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(px.epoch) : "memory");
    }
    // ---- 3. wait for the same slice from every rank, rank-ordered sum
    if (tid == 0) *flag_smem = 0;
    __syncthreads();
    if (tid < px.world && p0 < p1 && *reinterpret_cast<volatile int*>(px.timed_out) == 0) {
        const unsigned* flag = reinterpret_cast<const unsigned*>(own + PEER_FLAGS) + tid * PEER_MAX_CTA + cta;
        const long long t0 = clock64();
        while ((int)(ld_acquire_sys(flag) - px.epoch) < 0) {
            if ((unsigned long long)(clock64() - t0) > px.timeout_cycles) { atomicExch(px.timed_out, 1); break; }
        }
    }
    __syncthreads();
    const float* slot = reinterpret_cast<const float*>(own + PEER_HEADER) + (size_t)(px.epoch & 1u) * px.world * peer_row(P);
    for (int p = p0 + tid; p < p1; p += (int)blockDim.x) {
        float a = 0.0f;
        for (int r = 0; r < px.world; ++r) a += __ldcg(slot + (size_t)r * peer_row(P) + p);      // L2 is where the peers' stores land
        if (p < P) px.grad[p] = a;
        else if (px.stats) px.stats[p - P] = a;
    }
    if (cta == 0 && tid == 0 && px.stats) px.stats[KIN_PPO_STAT_SKIP] = *reinterpret_cast<volatile int*>(px.timed_out) ? 1.0f : 0.0f;
}

}  // namespace kin
