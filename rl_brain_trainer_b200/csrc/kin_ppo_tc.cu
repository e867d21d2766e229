// kin_ppo_tc.cu -- K3-TC: the PPO minibatch gradient with EVERY GEMM on the 5th-generation tensor cores
// (tcgen05.mma kind::f16, bf16 operands, fp32 accumulation in TMEM).
//
// Same contract as kin_ppo_grad (kin_ppo.cu; SB3 2.8.0 ppo.py train()), different arithmetic engine:
//   * a GEMM tile is 128 samples (two 64-sample minibatch tiles); sample row r <-> TMEM lane r.  A CTA has 256 threads:
//     warps 0-3 own the actor half of every activation row, warps 4-7 the critic half (thread = (row, net)).
//   * every operand lives in shared memory as a [rows][64 bf16] SWIZZLE_128B tile (kin_umma.cuh).  The tiles written for
//     the forward pass (X, H1, H2 as K-major A operands with M = sample) are re-read UNCHANGED as MN-major operands
//     (K = sample) by the weight-gradient GEMMs  dW = G^T H, so nothing is ever transposed; W1 is likewise read K-major by
//     the forward pass and MN-major by the data-gradient GEMM  dH1 = G2 W1.
//   * GEMMs per 128-sample tile (M x N x K):
//       forward   Z  = X [W0a;W0c]^T            128 x 128 x 64     (bias b0 rides on the constant-one column 56 of X)
//                 Z  = H1a W1a^T | H1c W1c^T    2 x (128 x 64 x 64)
//                 O  = H2a WOa^T + H2c WOc^T    128 x 16 x 64 (x2, accumulated: cols 0-6 action means, col 7 value)
//       backward  Z  = dO WOa | dO WOc          2 x (128 x 64 x 16)
//                 Z  = G2a W1a | G2c W1c        2 x (128 x 64 x 64)
//       weights   dWO += H2^T dO, db1 += G2^T dO(ones col), dW1 += G2^T H1, dW0|db0 += G1^T X, dbo += X(ones row)^T dO
//                                               M = 64, K = 128 samples, accumulated in TMEM across all tiles of the CTA
//     The weight-gradient accumulators (and the bias gradients, which fall out of the constant-one columns) stay in TMEM
//     for the whole kernel: 512 columns = Z 128 | O 16 | - | dbo 16 | dWO 32 | db1 32 | dW1 128 | dW0 128.
//   * one thread issues the MMAs; completion is tracked with two mbarriers (the forward/backward chain, and the
//     weight-gradient batch that must drain before the next tile overwrites the operand tiles).
//   * the elementwise work (tanh, 1 - h^2, the loss and its derivative, log_std gradient, statistics) is fp32 in registers;
//     each thread keeps packed bf16 copies of its H1 / H2 half-rows in registers for the backward pass.
//
// Numerics: bf16 operands (8-bit mantissa) + fp32 accumulation + tanh.approx -> gradients agree with the strict-fp32 kernel
// to ~1e-2 relative per tensor (tests/test_gpu_ppo.py::test_minibatch_gradient_tc_*); because log-probs move by O(1e-3) the
// trainer refreshes old_logp with THIS kernel's forward (forward_only) so the first-epoch ratio is exactly 1.
#include "kin_ppo_layout.cuh"
#include "kin_umma.cuh"

namespace kin {

using namespace umma;

constexpr int TCG_THREADS = 256;
constexpr int TCG_ROWS = 128;
constexpr int TILE_BYTES = TCG_ROWS * 128;        // [128][64 bf16]
constexpr float kHalfLog2PiTc = 0.91893853320467274178f;

// TMEM column map (fp32 columns)
constexpr unsigned COL_Z = 0;         // 128: actor 0..63 | critic 64..127
constexpr unsigned COL_O = 128;       // 16
constexpr unsigned COL_BO = 160;      // 16   (M = 64 rows = X columns; row 56 = sum over samples of dO)
constexpr unsigned COL_WO = 192;      // 2 x 16
constexpr unsigned COL_B1 = 224;      // 2 x 16 (column 8 = db1)
constexpr unsigned COL_W1 = 256;      // 2 x 64
constexpr unsigned COL_W0 = 384;      // 2 x 64 (column 56 = db0)
constexpr unsigned TMEM_COLS_G = 512;

struct __align__(1024) TcGradSmem {
    unsigned char X[2][TILE_BYTES];      // double-buffered in image mode (the next tile's image is prefetched by the TMA engine)
    unsigned char H1[2][TILE_BYTES];
    unsigned char H2[2][TILE_BYTES];
    unsigned char G2[2][TILE_BYTES];
    unsigned char G1[2][TILE_BYTES];
    unsigned char DO[TILE_BYTES];        // cols 0..7 dL/d(mean, value), col 8 = 1, rest 0
    unsigned char W0[TILE_BYTES];        // rows 0..63 actor, 64..127 critic; col 56 = b0
    unsigned char W1[2][64 * 128];
    unsigned char WO[2][16 * 128];       // actor: rows 0..6 = act_w; critic: row 7 = val_w
    float b1[128];
    float bo[8];
    float ls[8];
    float inv_sig[8];
    float scal[32];                      // 0 adv mean, 1 1/(std+eps), 2..6 statistics, 8..14 d log_std
    unsigned long long mbar[4];          // 0 forward/backward chain, 1 weight-gradient batch, 2/3 X image buffers
    unsigned tmem_base;
};

__device__ __forceinline__ void bulk_load_tile(unsigned dst_saddr, const void* src, unsigned mbar_saddr) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_saddr), "r"(TILE_BYTES) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_saddr), "l"(src), "r"(TILE_BYTES), "r"(mbar_saddr) : "memory");
}

__device__ __forceinline__ void st_bf16(unsigned char* tile, int row, int col, float v) {
    *reinterpret_cast<unsigned short*>(tile + sw_elem(row, col)) = (unsigned short)(pack_bf16(v, 0.0f) & 0xffffu);
}

// 64 accumulator columns of (row, net) -> f -> bf16 row of `tile`.  MODE 0: tanh; 1: tanh(x + b1); 2: x * (1 - h^2), h = keep[]
template <int MODE>
__device__ __forceinline__ void epilogue64(unsigned tz, unsigned char* tile, int row, const float* bias, unsigned* keep) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        float v[32];
        tmem_ld32(tz + half * 32, v);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            unsigned p[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int c = 8 * j + 2 * e;
                float a = v[c], b = v[c + 1];
                if (MODE == 0) {
                    a = tanh_fast(a);
                    b = tanh_fast(b);
                } else if (MODE == 1) {
                    a = tanh_fast(a + bias[half * 32 + c]);
                    b = tanh_fast(b + bias[half * 32 + c + 1]);
                } else {
                    const unsigned h = keep[half * 16 + 4 * j + e];
                    const float hl = bf16_lo(h), hh = bf16_hi(h);
                    a *= fmaf(-hl, hl, 1.0f);
                    b *= fmaf(-hh, hh, 1.0f);
                }
                p[e] = pack_bf16(a, b);
                if (MODE != 2) keep[half * 16 + 4 * j + e] = p[e];
            }
            *reinterpret_cast<uint4*>(tile + sw_chunk(row, half * 4 + j)) = make_uint4(p[0], p[1], p[2], p[3]);
        }
    }
}

// IMG: obs is the rollout buffer of bf16 operand images written by kin_ppo_collect (one 16 KB image per 128 consecutive samples)
template <bool IMG>
__global__ void __launch_bounds__(TCG_THREADS, 1)
kin_ppo_grad_tc_kernel(const float* __restrict__ params, KinPpoHyper hp, const float* __restrict__ obs, const float* __restrict__ action,
                       const float* __restrict__ old_logp, const float* __restrict__ advantage, const float* __restrict__ returns,
                       const double* __restrict__ tile_sums, const int* __restrict__ tile_ids, int n_pairs, float inv_global_batch,
                       float* __restrict__ partials, float* __restrict__ logp_out, float* __restrict__ value_out, int forward_only) {
    constexpr int IN = 56;
    extern __shared__ unsigned char smem_raw[];
    TcGradSmem& S = *reinterpret_cast<TcGradSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const PpoOffsets O = ppo_offsets(IN);
    const int P = O.total;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = tid & 127, net = tid >> 7;

    // ---- prologue: weights -> bf16 operand tiles --------------------------------------------------------------------
    {
        uint4* z = reinterpret_cast<uint4*>(S.DO);
        for (int i = tid; i < TILE_BYTES / 16; i += TCG_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
        uint4* zw = reinterpret_cast<uint4*>(S.WO);
        for (int i = tid; i < 2 * 16 * 128 / 16; i += TCG_THREADS) zw[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    for (int i = tid; i < 128 * 64; i += TCG_THREADS) {
        const int n = i >> 6, k = i & 63, u = n & 63;
        const int wbase = (n < 64) ? O.pi_w0 : O.vf_w0, bbase = (n < 64) ? O.pi_b0 : O.vf_b0;
        const float v = k < IN ? __ldg(params + wbase + u * IN + k) : (k == IN ? __ldg(params + bbase + u) : 0.0f);
        st_bf16(S.W0, n, k, v);
    }
    for (int i = tid; i < 2 * 4096; i += TCG_THREADS) {
        const int nt = i >> 12, u = (i >> 6) & 63, k = i & 63;
        st_bf16(S.W1[nt], u, k, __ldg(params + (nt ? O.vf_w1 : O.pi_w1) + u * 64 + k));
    }
    __syncthreads();   // WO / DO zero fill is complete before the real rows go in
    for (int i = tid; i < 7 * 64; i += TCG_THREADS) st_bf16(S.WO[0], i >> 6, i & 63, __ldg(params + O.act_w + i));
    if (tid < 64) st_bf16(S.WO[1], 7, tid, __ldg(params + O.val_w + tid));
    if (tid < 128) {
        S.b1[tid] = __ldg(params + (tid < 64 ? O.pi_b1 + tid : O.vf_b1 + tid - 64));
        *reinterpret_cast<uint4*>(S.DO + sw_chunk(tid, 1)) = make_uint4(0x00003F80u, 0u, 0u, 0u);   // col 8 = 1.0 (bf16)
    }
    if (tid < 8) {
        const float ls = tid < 7 ? __ldg(params + O.log_std + tid) : 0.0f;
        S.bo[tid] = tid < 7 ? __ldg(params + O.act_b + tid) : __ldg(params + O.val_b);
        S.ls[tid] = ls;
        S.inv_sig[tid] = expf(-ls);
    }
    if (tid < 32) S.scal[tid] = 0.0f;
    __syncwarp();
    if (tid < 32 && !forward_only) {
        double s1 = 0.0, s2 = 0.0;
        for (int j = tid; j < 2 * n_pairs; j += 32) {
            const int t = tile_ids[j];
            s1 += tile_sums[2 * t];
            s2 += tile_sums[2 * t + 1];
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, off);
            s2 += __shfl_xor_sync(0xffffffffu, s2, off);
        }
        if (tid == 0) {
            const double nsamp = (double)n_pairs * TCG_ROWS;
            const double mean = s1 / nsamp;
            const double var = nsamp > 1.0 ? fmax((s2 - nsamp * mean * mean) / (nsamp - 1.0), 0.0) : 0.0;
            S.scal[0] = hp.normalize_advantage ? (float)mean : 0.0f;
            S.scal[1] = hp.normalize_advantage ? (float)(1.0 / (sqrt(var) + 1e-8)) : 1.0f;
        }
    }
    if (warp == 0) tmem_alloc(smem_u32(&S.tmem_base), TMEM_COLS_G);
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&S.mbar[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();

    const unsigned tb = S.tmem_base;
    const unsigned tlane = tb + ((unsigned)((warp & 3) * 32) << 16);     // this warp's lane quadrant, column 0
    const unsigned tz = tlane + COL_Z + net * 64;
    const unsigned mb_main = smem_u32(&S.mbar[0]), mb_wg = smem_u32(&S.mbar[1]);
    const unsigned aXb[2] = {smem_u32(S.X[0]), smem_u32(S.X[1])}, aDO = smem_u32(S.DO), aW0 = smem_u32(S.W0);
    const unsigned mb_x[2] = {smem_u32(&S.mbar[2]), smem_u32(&S.mbar[3])};
    unsigned par_x[2] = {0u, 0u};
    const unsigned char* img = reinterpret_cast<const unsigned char*>(obs);
    if (IMG && tid == 0 && (int)blockIdx.x < n_pairs)
        bulk_load_tile(aXb[0], img + (size_t)(tile_ids[2 * blockIdx.x] >> 1) * TILE_BYTES, mb_x[0]);
    const unsigned aH1[2] = {smem_u32(S.H1[0]), smem_u32(S.H1[1])}, aH2[2] = {smem_u32(S.H2[0]), smem_u32(S.H2[1])};
    const unsigned aG2[2] = {smem_u32(S.G2[0]), smem_u32(S.G2[1])}, aG1[2] = {smem_u32(S.G1[0]), smem_u32(S.G1[1])};
    const unsigned aW1[2] = {smem_u32(S.W1[0]), smem_u32(S.W1[1])}, aWO[2] = {smem_u32(S.WO[0]), smem_u32(S.WO[1])};
    unsigned par_main = 0u, par_wg = 0u;
    unsigned h1p[32], h2p[32];
    float dls[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, st[5] = {0.f, 0.f, 0.f, 0.f, 0.f};   // per-thread partial sums (net 0 threads)

    int it = 0;
    for (int j = blockIdx.x; j < n_pairs; j += gridDim.x, ++it) {
        const int t0 = tile_ids[2 * j], t1 = tile_ids[2 * j + 1];
        const int xb = IMG ? (it & 1) : 0;
        const unsigned aX = aXb[xb];
        // loss inputs of this thread's sample: issue the loads now, consume them after layer 3
        const size_t g = (size_t)(row < 64 ? t0 : t1) * 64 + (row & 63);
        float act_r[7], adv_r = 0.0f, olp_r = 0.0f, ret_r = 0.0f;
        if (net == 0) {
#pragma unroll
            for (int d = 0; d < 7; ++d) act_r[d] = __ldg(action + g * 7 + d);
            if (!forward_only) {
                adv_r = __ldg(advantage + g);
                olp_r = __ldg(old_logp + g);
                ret_r = __ldg(returns + g);
            }
        }
        if (it > 0 && !forward_only) {      // the previous tile's weight-gradient GEMMs still read X / H / G / dO
            mbar_wait(mb_wg, par_wg);
            par_wg ^= 1u;
        }
        if (IMG) {
            // the other buffer is free (its last readers were the previous tile's GEMMs): prefetch the next tile's image into it
            const int jn = j + gridDim.x;
            if (tid == 0 && jn < n_pairs) bulk_load_tile(aXb[xb ^ 1], img + (size_t)(tile_ids[2 * jn] >> 1) * TILE_BYTES, mb_x[xb ^ 1]);
            mbar_wait(mb_x[xb], par_x[xb]);
            par_x[xb] ^= 1u;
            fence_before();
            __syncthreads();
        } else {
            // ---- X tile: obs fp32 -> bf16, coalesced float4 reads; column 56 = 1 carries the layer-1 bias --------------
            unsigned char* X = S.X[0];
#pragma unroll
            for (int i = 0; i < 7; ++i) {
                const int idx = tid + TCG_THREADS * i;           // 0 .. 1791
                const int half = idx >= 896, rem = idx - half * 896;
                const int r = half * 64 + rem / 14, q = rem % 14;
                const float4 v = __ldg(reinterpret_cast<const float4*>(obs + (size_t)(half ? t1 : t0) * 64 * IN) + rem);
                *reinterpret_cast<uint2*>(X + sw_chunk(r, q >> 1) + ((q & 1) << 3)) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
            }
            if (tid < 128) *reinterpret_cast<uint4*>(X + sw_chunk(tid, 7)) = make_uint4(0x00003F80u, 0u, 0u, 0u);
            fence_async_smem();
            fence_before();
            __syncthreads();
        }
        // ---- layer 1 ----------------------------------------------------------------------------------------------------
        if (tid == 0) {
            fence_after();
            constexpr unsigned id = idesc_bf16(128, 128, false, false);
#pragma unroll
            for (int k = 0; k < 4; ++k) mma_bf16(tb + COL_Z, desc_k(aX + k * 32), desc_k(aW0 + k * 32), id, k > 0);
            commit(mb_main);
        }
        mbar_wait(mb_main, par_main);
        par_main ^= 1u;
        fence_after();
        epilogue64<0>(tz, S.H1[net], row, nullptr, h1p);
        fence_async_smem();
        fence_before();
        __syncthreads();
        // ---- layer 2 ----------------------------------------------------------------------------------------------------
        if (tid == 0) {
            fence_after();
            constexpr unsigned id = idesc_bf16(128, 64, false, false);
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int k = 0; k < 4; ++k) mma_bf16(tb + COL_Z + nt * 64, desc_k(aH1[nt] + k * 32), desc_k(aW1[nt] + k * 32), id, k > 0);
            commit(mb_main);
        }
        mbar_wait(mb_main, par_main);
        par_main ^= 1u;
        fence_after();
        epilogue64<1>(tz, S.H2[net], row, S.b1 + net * 64, h2p);
        fence_async_smem();
        fence_before();
        __syncthreads();
        // ---- layer 3: means (cols 0..6) and value (col 7) -----------------------------------------------------------------
        if (tid == 0) {
            fence_after();
            constexpr unsigned id = idesc_bf16(128, 16, false, false);
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int k = 0; k < 4; ++k) mma_bf16(tb + COL_O, desc_k(aH2[nt] + k * 32), desc_k(aWO[nt] + k * 32), id, (nt | k) > 0);
            commit(mb_main);
        }
        mbar_wait(mb_main, par_main);
        par_main ^= 1u;
        fence_after();
        // ---- loss and d(loss)/d(outputs): one thread per sample (the actor half of the CTA) ---------------------------------
        if (net == 0) {
            float o[16];
            tmem_ld16(tlane + COL_O, o);
            float lp = 0.0f, z[7];
#pragma unroll
            for (int d = 0; d < 7; ++d) {
                z[d] = (act_r[d] - (o[d] + S.bo[d])) * S.inv_sig[d];
                lp += -0.5f * z[d] * z[d] - S.ls[d] - kHalfLog2PiTc;
            }
            const float v = o[7] + S.bo[7];
            if (logp_out) logp_out[g] = lp;
            if (value_out) value_out[g] = v;
            if (!forward_only) {
                const float adv_n = (adv_r - S.scal[0]) * S.scal[1];
                const float log_ratio = lp - olp_r;
                const float ratio = expf(log_ratio);
                const float pl1 = adv_n * ratio, pl2 = adv_n * fminf(fmaxf(ratio, 1.0f - hp.clip_range), 1.0f + hp.clip_range);
                const float dpl_dlp = (pl1 <= pl2) ? -adv_n * ratio : 0.0f;
                const float rt = ret_r;
                float dm[8];
                float ent = 0.0f;
#pragma unroll
                for (int d = 0; d < 7; ++d) {
                    dm[d] = inv_global_batch * dpl_dlp * z[d] * S.inv_sig[d];
                    dls[d] += inv_global_batch * dpl_dlp * (z[d] * z[d] - 1.0f) - inv_global_batch * hp.ent_coef;
                    ent += 0.5f + kHalfLog2PiTc + S.ls[d];
                }
                dm[7] = inv_global_batch * hp.vf_coef * 2.0f * (v - rt);
                *reinterpret_cast<uint4*>(S.DO + sw_chunk(row, 0)) =
                    make_uint4(pack_bf16(dm[0], dm[1]), pack_bf16(dm[2], dm[3]), pack_bf16(dm[4], dm[5]), pack_bf16(dm[6], dm[7]));
                st[0] += -fminf(pl1, pl2);
                st[1] += (rt - v) * (rt - v);
                st[2] += ent;
                st[3] += (ratio - 1.0f) - log_ratio;
                st[4] += fabsf(ratio - 1.0f) > hp.clip_range ? 1.0f : 0.0f;
            }
        }
        if (forward_only) {
            fence_before();
            __syncthreads();     // O is re-written by the next tile's layer 3 only after everyone has read it
            continue;
        }
        fence_async_smem();
        fence_before();
        __syncthreads();
        // ---- dZ2 = (dO WO) * (1 - H2^2) -------------------------------------------------------------------------------------
        if (tid == 0) {
            fence_after();
            constexpr unsigned id = idesc_bf16(128, 64, false, true);
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) mma_bf16(tb + COL_Z + nt * 64, desc_k(aDO), desc_mn(aWO[nt]), id, 0u);
            commit(mb_main);
        }
        mbar_wait(mb_main, par_main);
        par_main ^= 1u;
        fence_after();
        epilogue64<2>(tz, S.G2[net], row, nullptr, h2p);
        fence_async_smem();
        fence_before();
        __syncthreads();
        // ---- dZ1 = (G2 W1) * (1 - H1^2); the weight-gradient GEMMs that need only G2 follow it on the tensor pipe -------------
        if (tid == 0) {
            fence_after();
            constexpr unsigned id = idesc_bf16(128, 64, false, true);
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int k = 0; k < 4; ++k) mma_bf16(tb + COL_Z + nt * 64, desc_k(aG2[nt] + k * 32), desc_mn(aW1[nt] + k * 2048), id, k > 0);
            commit(mb_main);
            const unsigned acc0 = it > 0;
            constexpr unsigned id16 = idesc_bf16(64, 16, true, true), id64 = idesc_bf16(64, 64, true, true);
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
                for (int k = 0; k < 8; ++k) mma_bf16(tb + COL_WO + nt * 16, desc_mn(aH2[nt] + k * 2048), desc_mn(aDO + k * 2048), id16, acc0 | (k > 0));
#pragma unroll
                for (int k = 0; k < 8; ++k) mma_bf16(tb + COL_B1 + nt * 16, desc_mn(aG2[nt] + k * 2048), desc_mn(aDO + k * 2048), id16, acc0 | (k > 0));
#pragma unroll
                for (int k = 0; k < 8; ++k) mma_bf16(tb + COL_W1 + nt * 64, desc_mn(aG2[nt] + k * 2048), desc_mn(aH1[nt] + k * 2048), id64, acc0 | (k > 0));
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) mma_bf16(tb + COL_BO, desc_mn(aX + k * 2048), desc_mn(aDO + k * 2048), id16, acc0 | (k > 0));
        }
        mbar_wait(mb_main, par_main);
        par_main ^= 1u;
        fence_after();
        epilogue64<2>(tz, S.G1[net], row, nullptr, h1p);
        fence_async_smem();
        fence_before();
        __syncthreads();
        // ---- dW0 | db0 += G1^T X ---------------------------------------------------------------------------------------------
        if (tid == 0) {
            fence_after();
            const unsigned acc0 = it > 0;
            constexpr unsigned id64 = idesc_bf16(64, 64, true, true);
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int k = 0; k < 8; ++k) mma_bf16(tb + COL_W0 + nt * 64, desc_mn(aG1[nt] + k * 2048), desc_mn(aX + k * 2048), id64, acc0 | (k > 0));
            commit(mb_wg);
        }
    }

    if (!forward_only) {
        if (it > 0) mbar_wait(mb_wg, par_wg);
        fence_after();
        // log_std gradient and statistics: warp shuffle, then shared atomics (4 warps)
        if (net == 0) {
#pragma unroll
            for (int d = 0; d < 7; ++d) {
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) dls[d] += __shfl_xor_sync(0xffffffffu, dls[d], off);
            }
#pragma unroll
            for (int q = 0; q < 5; ++q) {
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) st[q] += __shfl_xor_sync(0xffffffffu, st[q], off);
            }
            if (lane == 0) {
#pragma unroll
                for (int q = 0; q < 5; ++q) atomicAdd(&S.scal[2 + q], st[q]);
#pragma unroll
                for (int d = 0; d < 7; ++d) atomicAdd(&S.scal[8 + d], dls[d]);
            }
        }
        __syncthreads();
        // ---- accumulators (M = 64: row m lives in lane m % 16 + 32 * (m / 16)) -> this CTA's partial gradient ---------------
        float* out = partials + (size_t)blockIdx.x * (P + KIN_PPO_STATS + 8);
        const int u = (warp & 3) * 16 + lane;       // valid for lane < 16
        const bool rowok = lane < 16;
        const int w1_base = net ? O.vf_w1 : O.pi_w1, w0_base = net ? O.vf_w0 : O.pi_w0;
        const int b1_base = net ? O.vf_b1 : O.pi_b1, b0_base = net ? O.vf_b0 : O.pi_b0;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float v[32];
            tmem_ld32(tlane + COL_W1 + net * 64 + half * 32, v);
            if (rowok) {
#pragma unroll
                for (int c = 0; c < 32; ++c) out[w1_base + u * 64 + half * 32 + c] = v[c];   // flat offsets are not 16-byte aligned
            }
            tmem_ld32(tlane + COL_W0 + net * 64 + half * 32, v);
            if (rowok) {
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const int k = half * 32 + c;
                    if (k < IN) out[w0_base + u * IN + k] = v[c];
                    else if (k == IN) out[b0_base + u] = v[c];
                }
            }
        }
        {
            float o[16];
            tmem_ld16(tlane + COL_WO + net * 16, o);
            if (rowok) {
                if (net == 0) {
#pragma unroll
                    for (int d = 0; d < 7; ++d) out[O.act_w + d * 64 + u] = o[d];
                } else {
                    out[O.val_w + u] = o[7];
                }
            }
            tmem_ld16(tlane + COL_B1 + net * 16, o);
            if (rowok) out[b1_base + u] = o[8];
            if (warp == 3) {                        // X column 56 (the ones column) = row 56 -> lane 96 + 8
                tmem_ld16(tlane + COL_BO, o);
                if (lane == 8) {
#pragma unroll
                    for (int d = 0; d < 7; ++d) out[O.act_b + d] = o[d];
                    out[O.val_b] = o[7];
                }
            }
        }
        if (tid < 7) out[O.log_std + tid] = S.scal[8 + tid];
        if (tid < 5) out[P + tid] = S.scal[2 + tid];
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, TMEM_COLS_G);
}

}  // namespace kin

using namespace kin;

extern "C" int kin_ppo_grad_tc(const float* params, int in_dim, const KinPpoHyper* hp, const void* obs_any, const float* action, const float* old_logp,
                               const float* advantage, const float* returns, const double* tile_sums, const int* tile_ids, int n_tiles,
                               long long global_batch, float* partials, int grid, float* grad, float* stats, float* logp_out, float* value_out,
                               int forward_only, int obs_is_image, void* stream) {
    const float* obs = static_cast<const float*>(obs_any);
    if (!params || !hp || !obs || !action || !tile_ids || n_tiles <= 0 || grid <= 0)
        return kin_fail(KIN_ERR_INVALID_ARG, "kin_ppo_grad_tc: bad arguments");
    if (!forward_only && (!old_logp || !advantage || !returns || !tile_sums || !partials || !grad || global_batch <= 0))
        return kin_fail(KIN_ERR_INVALID_ARG, "kin_ppo_grad_tc: the gradient pass needs old_logp, advantage, returns, tile_sums, partials and grad");
    if (forward_only && !logp_out && !value_out) return kin_fail(KIN_ERR_INVALID_ARG, "kin_ppo_grad_tc: forward_only without an output");
    if (in_dim != 56) return kin_fail(KIN_ERR_UNSUPPORTED, "kin_ppo_grad_tc: in_dim must be 56");
    if (n_tiles & 1) return kin_fail(KIN_ERR_INVALID_ARG, "kin_ppo_grad_tc: n_tiles must be even (two 64-sample tiles per 128-row GEMM tile)");
    if (((uintptr_t)obs & 15u)) return kin_fail(KIN_ERR_INVALID_ARG, "kin_ppo_grad_tc: obs must be 16-byte aligned");
    const int P = ppo_offsets(in_dim).total;
    const size_t smem = sizeof(TcGradSmem) + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kin_ppo_grad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kin_ppo_grad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return kin_fail_cuda(e, "kin_ppo_grad_tc: smem attribute");
        attr_set = true;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int n_pairs = n_tiles / 2;
    const int g = grid < n_pairs ? grid : n_pairs;
    const float inv = forward_only ? 0.0f : 1.0f / (float)global_batch;
    if (obs_is_image)
        kin_ppo_grad_tc_kernel<true><<<g, TCG_THREADS, smem, st>>>(params, *hp, obs, action, old_logp, advantage, returns, tile_sums, tile_ids, n_pairs,
                                                                   inv, partials, logp_out, value_out, forward_only);
    else
        kin_ppo_grad_tc_kernel<false><<<g, TCG_THREADS, smem, st>>>(params, *hp, obs, action, old_logp, advantage, returns, tile_sums, tile_ids, n_pairs,
                                                                    inv, partials, logp_out, value_out, forward_only);
    if (!forward_only) kin_ppo_reduce_launch(partials, g, P, grad, stats, inv, st);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? KIN_OK : kin_fail_cuda(e, "kin_ppo_grad_tc");
}
