// kin_peer.cuh -- receive-buffer layout of the NVLink peer-memory gradient exchange and the in-kernel ("fused") form of it:
// the tail every CTA of the gradient kernel runs after its partial row is written (csrc/kin_ppo_tc.cu), replacing the separate
// kin_peer_push_kernel + kin_peer_gather_kernel launches between the gradient kernel and Adam.
//
//   receive buffer of a rank (cudaMalloc'ed by kin_peer_buffer_create, opened by the peers through CUDA IPC):
//     [0]      unsigned arrived[8]            two-kernel form: arrived[s] = pushes of sender s that have fully landed here
//     [64]     unsigned local_count           two-kernel form: CTAs of this rank's push kernel that have finished
//     [68]     unsigned grid_count            fused form: CTAs of this rank's gradient kernel whose partial row is complete (monotonic)
//     [72]     unsigned adam_count            fused update: CTAs whose slice of the summed gradient and its sum of squares are written (monotonic)
//     [128]    unsigned slice_flag[8][512]    fused form: slice_flag[s][c] = last exchange whose slice c sender s has delivered here
//     [16512]  float    slot[2][world][row]   row = n_params + 8 (gradient, then the 5 loss statistics); slot = exchange parity
#pragma once

#include <cuda_bf16.h>

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

// optional clip + Adam step in the same tail (params == nullptr: the caller launches kin_ppo_adam)
struct AdamFused {
    float* params;                    // [P] flat parameters, updated in place
    float* m;                         // [P] Adam first moments
    float* v;                         // [P] Adam second moments
    unsigned short* wimg;             // bf16 operand image of the 56-input weights kept in step (nullable)
    float* stats_accum;               // [KIN_PPO_STATS] running sums over the minibatches of an update (nullable)
    float* norm_part;                 // [n_cta] scratch: sum of squares of each CTA's gradient slice
    float bc1, bc2;                   // 1 - beta^step
    int in_dim;
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
    AdamFused adam;
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

// Second half of the fused update, called by ALL threads of EVERY CTA right after peer_exchange_tail when px.adam.params is set:
// clip_grad_norm_ + Adam (kin_ppo_adam's arithmetic) on the slice of the summed gradient this CTA has just written, so a minibatch is ONE
// launch (gradient, reduction, exchange, optimiser step) instead of three.
//   1. every CTA writes the sum of squares of its slice (fixed order) to norm_part[cta]; grid barrier (adam_count);
//   2. every CTA adds the n_cta partial sums in the same fixed order -> the same clip coefficient everywhere (and on every rank: the
//      summed gradient is bitwise identical on all ranks);
//   3. Adam on the slice: m, v, params, and the bf16 weight image the next launch's TMA copies read.
// The parameters are only read in the gradient kernel's prologue and every CTA is past the first grid barrier, so updating them here
// races with nothing.  A timed-out exchange leaves them untouched (same rule as kin_ppo_adam with KIN_PPO_STAT_SKIP).
__device__ __forceinline__ void peer_adam_tail(const PeerFused& px, const KinPpoHyper& hp, int P, float* red /* >= 40 floats of smem */) {
    const AdamFused& A = px.adam;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = (int)(blockDim.x + 31) >> 5;
    const int n_cta = (int)(gridDim.x * gridDim.y), cta = (int)(blockIdx.y * gridDim.x + blockIdx.x);
    unsigned char* own = px.peers.base[px.rank];
    const int cols = P + 5;
    const int per = ((cols + n_cta - 1) / n_cta + 31) & ~31;
    const int p0 = min(cta * per, P), p1 = min(cta * per + per, P);       // the gradient part of this CTA's slice
    __syncthreads();                 // px.grad[p0 .. p1) was written by this CTA's threads
    float ss = 0.0f;
    for (int p = p0 + tid; p < p1; p += (int)blockDim.x) {
        const float g = px.grad[p];
        ss = fmaf(g, g, ss);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
    if (lane == 0) red[w] = ss;
    __syncthreads();
    if (tid == 0) {
        float t = 0.0f;
        for (int k = 0; k < nw; ++k) t += red[k];
        A.norm_part[cta] = t;
        __threadfence();
        unsigned* cnt = reinterpret_cast<unsigned*>(own + 72);
        atomicAdd(cnt, 1u);
        const unsigned target = px.epoch * (unsigned)n_cta;
        const long long t0 = clock64();
        while ((int)(ld_acquire_gpu(cnt) - target) < 0) {
            if ((unsigned long long)(clock64() - t0) > px.timeout_cycles) { atomicExch(px.timed_out, 1); break; }
        }
    }
    __syncthreads();
    if (w == 0) {
        float t = 0.0f;
        for (int k = lane; k < n_cta; k += 32) t += __ldcg(A.norm_part + k);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
        if (lane == 0) {
            const float norm = sqrtf(t);
            const float c = hp.max_grad_norm > 0.0f ? hp.max_grad_norm / (norm + 1e-6f) : 1.0f;
            red[32] = c < 1.0f ? c : 1.0f;
            red[33] = norm;
        }
    }
    __syncthreads();
    const bool skip = *reinterpret_cast<volatile int*>(px.timed_out) != 0;
    if (cta == 0 && tid == 0 && px.stats) {
        px.stats[KIN_PPO_STAT_GRAD_NORM] = red[33];
        px.stats[KIN_PPO_STAT_SKIP] = skip ? 1.0f : 0.0f;
        if (A.stats_accum && !skip) {
#pragma unroll
            for (int q = 0; q < 5; ++q) A.stats_accum[q] += __ldcg(px.stats + q);
            A.stats_accum[KIN_PPO_STAT_GRAD_NORM] += red[33];
            A.stats_accum[7] += 1.0f;
        }
    }
    if (skip) return;
    const float cf = red[32];
    const PpoOffsets O = ppo_offsets(A.in_dim);
    for (int p = p0 + tid; p < p1; p += (int)blockDim.x) {
        const float g = px.grad[p] * cf;
        const float mm = fmaf(hp.adam_beta1, A.m[p], (1.0f - hp.adam_beta1) * g);
        const float vv = fmaf(hp.adam_beta2, A.v[p], (1.0f - hp.adam_beta2) * g * g);
        A.m[p] = mm;
        A.v[p] = vv;
        const float denom = sqrtf(vv) / sqrtf(A.bc2) + hp.adam_eps;
        const float np = A.params[p] - (hp.learning_rate / A.bc1) * (mm / denom);
        A.params[p] = np;
        if (A.wimg) {
            const int off = wimg_offset(O, A.in_dim, p);
            if (off >= 0) A.wimg[off >> 1] = __bfloat16_as_ushort(__float2bfloat16_rn(np));
        }
    }
}

}  // namespace kin
