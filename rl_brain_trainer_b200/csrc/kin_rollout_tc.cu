// kin_rollout_tc.cu -- K2-TC: the fused policy-in-loop Approach -> Finisher rollout with the actor MLP on the
// 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM).
//
// Mapping: one episode <-> one thread <-> one row of the A operand tile <-> one TMEM lane.
//   * A CTA has 256 threads = two independent 128-episode tiles (warps 0-3 / 4-7) that share one copy of the
//     policy weights in shared memory and ping-pong: while one tile's threads run the env arithmetic on the
//     FP32 pipe, the other tile's GEMM runs on the tensor pipe.
//   * Per env-step and tile: every thread writes its 56-float observation (+ a constant 1 that carries the
//     bias, + zero pad to K = 64) straight into the 128-byte-swizzled K-major UMMA layout (its own row),
//     fence.proxy.async + a 128-thread named barrier, ONE thread issues 8 tcgen05.mma (M128 N64 K8) and
//     commits to an mbarrier; every thread then pulls its own accumulator row out of TMEM with
//     tcgen05.ld.32x32b, applies tanh and writes the hidden row back into the same A tile for the next layer
//     (layer 2: M128 N64, layer 3: M128 N8).  The 7 outputs go to the in-register env step (kin_core.cuh).
//   * The weights are B operands, K-major == torch's [out][in] row-major, pre-rounded to TF32
//     (cvt.rna) and laid out swizzled once per phase.  The A rows are rounded with cvt.rna as well
//     (the tensor core would otherwise truncate the low 13 mantissa bits).
//   * HBM traffic: 27 floats in, 24 words out per EPISODE; the kernel is bound by the FP32 / XU pipes (env
//     arithmetic, tanh, TMEM loads), the tensor pipe carries the 16,256 FLOP of the MLP per env-step.
//
// Numerics: env arithmetic fp32 (identical to the FFMA variant), MLP operands TF32 with fp32 accumulation and
// tanh.approx -> actions differ from strict fp32 by O(1e-3); closed-loop success rates agree statistically
// (tests/test_gpu_rollout.py::test_fused_rollout_tc_*).  The FFMA variant stays the strict-parity path.
//
// Replaces the same reference loops as kin_rollout.cu (eval_workspace_expansion.py:126-147 etc.).
#include <cstdlib>

#include "kin_internal.h"
#include "kin_state.cuh"

namespace kin {

constexpr int TC_TILE = 128;               // episodes per tile == UMMA M == TMEM lanes
constexpr int TC_TILES = 4;                // tiles per CTA (the last one may be partial: 1..4 warps)
constexpr int TC_MAX_THREADS = TC_TILE * TC_TILES;
constexpr int TC_K = 64;                   // padded reduction width of every layer
constexpr int TC_HID = 64;
constexpr int CHUNK_FLOATS_A = TC_TILE * 32;   // one 128-byte-wide K chunk of an A tile: 128 rows x 32 floats
constexpr int A_TILE_FLOATS = 2 * CHUNK_FLOATS_A;
constexpr int CHUNK_FLOATS_W = TC_HID * 32;    // 64 rows x 32 floats
constexpr int W_FLOATS = 2 * CHUNK_FLOATS_W;
constexpr int CHUNK_FLOATS_WO = 8 * 32;        // output layer: 8 rows (7 actions + zero row)
constexpr int WO_FLOATS = 2 * CHUNK_FLOATS_WO;

struct TcSmem {
    float W0[W_FLOATS];                 // 16 KB  [64][64]: 56 inputs | bias column | zero pad
    float W1[W_FLOATS];                 // 16 KB
    float WO[WO_FLOATS];                // 2 KB   [8][64]
    float b1[TC_HID];
    float bo[8];
    unsigned long long mbar[TC_TILES];
    unsigned tmem_base;
    int run_flags[2][TC_TILES][4];   // double-buffered by step parity: written before, read after the layer-1 barrier
    alignas(1024) float A[1][A_TILE_FLOATS];   // one 32 KB A tile per tile of the CTA (1..4, sized at launch), 1024-byte aligned
};

struct DevPolicyTc {
    const float *w0, *b0, *w1, *b1, *wo, *bo;
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// Round-to-nearest (ties away) TF32: cvt.rna.tf32 is "add half an ulp of the 10-bit mantissa, clear the low 13 bits"; the
// tensor core ignores those 13 bits of a kind::tf32 operand, so the add alone feeds it the same operand (values here are finite
// and far from overflow: observations in [-1, 1], tanh outputs, trained weights).
__device__ __forceinline__ float to_tf32(float x) { return __uint_as_float(__float_as_uint(x) + 0x1000u); }
__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// float offset of element (row, k) inside a K-major SWIZZLE_128B operand whose K chunks hold `rows` rows each
__device__ __forceinline__ int sw128_offset(int row, int k, int chunk_floats) {
    const int chunk = k >> 5, kk = k & 31;
    return chunk * chunk_floats + row * 32 + ((((kk >> 2) ^ (row & 7)) << 2) | (kk & 3));
}

// UMMA shared-memory descriptor, K-major, SWIZZLE_128B: start>>4 | LBO(=1)<<16 | SBO(=1024 B>>4)<<32 | version 1<<46 | layout 2<<61
__device__ __forceinline__ unsigned long long umma_desc(unsigned saddr) {
    return (unsigned long long)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D fp32, A/B tf32, both K-major, N>>3 at bit 17, M>>4 at bit 24
__host__ __device__ constexpr unsigned umma_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned mbar_saddr) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar_saddr) : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned saddr, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(saddr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned saddr, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}"
        ::"r"(saddr), "r"(parity) : "memory");
}
// one lane of a converged warp (the MMA issue is then compiled with uniform-register descriptors instead of per-MMA R2UR chains)
__device__ __forceinline__ bool elect_one_tc() {
    unsigned pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0u;
}
__device__ __forceinline__ void tile_barrier(int tile, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(tile + 1), "r"(threads) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// this thread's 32 consecutive accumulator columns [col0, col0+32) of its own TMEM lane
__device__ __forceinline__ void tmem_ld32(unsigned taddr, float* v) {
    unsigned r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(unsigned taddr, float* v) {
    unsigned r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// stage one policy's actor into the swizzled B-operand images (all threads of the CTA)
__device__ void load_weights_tc(TcSmem& S, const DevPolicyTc& p, int tid, int nthreads) {
    for (int i = tid; i < TC_HID * TC_K; i += nthreads) {
        const int n = i >> 6, k = i & 63;
        float v0 = k < KIN_OBS_DIM ? __ldg(p.w0 + n * KIN_OBS_DIM + k) : (k == KIN_OBS_DIM ? __ldg(p.b0 + n) : 0.0f);
        S.W0[sw128_offset(n, k, CHUNK_FLOATS_W)] = to_tf32(v0);
        S.W1[sw128_offset(n, k, CHUNK_FLOATS_W)] = to_tf32(__ldg(p.w1 + n * TC_HID + k));
    }
    for (int i = tid; i < 8 * TC_K; i += nthreads) {
        const int n = i >> 6, k = i & 63;
        S.WO[sw128_offset(n, k, CHUNK_FLOATS_WO)] = n < KIN_NJ ? to_tf32(__ldg(p.wo + n * TC_HID + k)) : 0.0f;
    }
    for (int i = tid; i < TC_HID; i += nthreads) S.b1[i] = __ldg(p.b1 + i);     // a CTA may be a single warp
    if (tid < 8) S.bo[tid] = tid < KIN_NJ ? __ldg(p.bo + tid) : 0.0f;
}

// write 4 consecutive K elements [k4*4, k4*4+4) of this thread's A row (already TF32-rounded)
__device__ __forceinline__ void a_store4(float* A, int row, int k4, float x0, float x1, float x2, float x3) {
    const int chunk = k4 >> 3, j = k4 & 7;
    float4* dst = reinterpret_cast<float4*>(A + chunk * CHUNK_FLOATS_A + row * 32 + ((j ^ (row & 7)) << 2));
    *dst = make_float4(x0, x1, x2, x3);
}

// one GEMM of the tile: D[128 x N] (TMEM) = A[128 x 64] (smem) * W[N x 64]^T (smem); issued by one thread
__device__ __forceinline__ void issue_layer(unsigned a_saddr, unsigned w_saddr, int w_chunk_bytes, unsigned tmem_d, unsigned idesc, unsigned mbar_saddr) {
    tc_fence_after();
#pragma unroll
    for (int k = 0; k < TC_K / 8; ++k) {
        const unsigned a_off = (k >> 2) * (CHUNK_FLOATS_A * 4) + (k & 3) * 32;
        const unsigned w_off = (k >> 2) * w_chunk_bytes + (k & 3) * 32;
        umma_tf32(tmem_d, umma_desc(a_saddr + a_off), umma_desc(w_saddr + w_off), idesc, k > 0 ? 1u : 0u);
    }
    umma_commit(mbar_saddr);
}

struct TileCtx {
    float* A;
    unsigned a_saddr, w0_saddr, w1_saddr, wo_saddr, mbar_saddr, tmem_d, tmem_row;
    int tile, row, flag_buf, tile_threads;
    unsigned parity;
    bool issuer;
};

// obs[56] -> act[7]; collective over the tile's 128 threads.  Returns false (without running the MLP) once no episode
// of the tile is still running: the vote rides on the layer-1 barrier, flags double-buffered by step parity.
__device__ __forceinline__ bool mlp_tc(TcSmem& S, TileCtx& c, const float* o, float* act, bool running) {
    const int fb = c.flag_buf;
    c.flag_buf ^= 1;
    const unsigned any = __any_sync(0xffffffffu, running);
    if ((threadIdx.x & 31) == 0) S.run_flags[fb][c.tile][(threadIdx.x >> 5) & 3] = (int)any;
    // ---- layer 1: A = [obs | 1 | 0...]
#pragma unroll
    for (int k4 = 0; k4 < KIN_OBS_DIM / 4; ++k4) a_store4(c.A, c.row, k4, to_tf32(o[4 * k4]), to_tf32(o[4 * k4 + 1]), to_tf32(o[4 * k4 + 2]), to_tf32(o[4 * k4 + 3]));
    a_store4(c.A, c.row, 14, 1.0f, 0.0f, 0.0f, 0.0f);
    a_store4(c.A, c.row, 15, 0.0f, 0.0f, 0.0f, 0.0f);
    fence_async_smem();
    tc_fence_before();
    tile_barrier(c.tile, c.tile_threads);
    if (!(S.run_flags[fb][c.tile][0] | S.run_flags[fb][c.tile][1] | S.run_flags[fb][c.tile][2] | S.run_flags[fb][c.tile][3])) return false;
    if (c.row < 32 && elect_one_tc()) issue_layer(c.a_saddr, c.w0_saddr, CHUNK_FLOATS_W * 4, c.tmem_d, umma_idesc(TC_TILE, TC_HID), c.mbar_saddr);
    mbar_wait(c.mbar_saddr, c.parity);
    c.parity ^= 1u;
    tc_fence_after();
    float v[32];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        tmem_ld32(c.tmem_row + half * 32, v);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            a_store4(c.A, c.row, half * 8 + j, to_tf32(tanh_approx(v[4 * j])), to_tf32(tanh_approx(v[4 * j + 1])),
                     to_tf32(tanh_approx(v[4 * j + 2])), to_tf32(tanh_approx(v[4 * j + 3])));
    }
    // ---- layer 2
    fence_async_smem();
    tc_fence_before();
    tile_barrier(c.tile, c.tile_threads);
    if (c.row < 32 && elect_one_tc()) issue_layer(c.a_saddr, c.w1_saddr, CHUNK_FLOATS_W * 4, c.tmem_d, umma_idesc(TC_TILE, TC_HID), c.mbar_saddr);
    mbar_wait(c.mbar_saddr, c.parity);
    c.parity ^= 1u;
    tc_fence_after();
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        tmem_ld32(c.tmem_row + half * 32, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float* b = S.b1 + half * 32 + 4 * j;
            a_store4(c.A, c.row, half * 8 + j, to_tf32(tanh_approx(v[4 * j] + b[0])), to_tf32(tanh_approx(v[4 * j + 1] + b[1])),
                     to_tf32(tanh_approx(v[4 * j + 2] + b[2])), to_tf32(tanh_approx(v[4 * j + 3] + b[3])));
        }
    }
    // ---- layer 3 (N = 8)
    fence_async_smem();
    tc_fence_before();
    tile_barrier(c.tile, c.tile_threads);
    if (c.row < 32 && elect_one_tc()) issue_layer(c.a_saddr, c.wo_saddr, CHUNK_FLOATS_WO * 4, c.tmem_d, umma_idesc(TC_TILE, 8), c.mbar_saddr);
    mbar_wait(c.mbar_saddr, c.parity);
    c.parity ^= 1u;
    tc_fence_after();
    float a8[8];
    tmem_ld8(c.tmem_row, a8);
    tc_fence_before();
#pragma unroll
    for (int i = 0; i < KIN_NJ; ++i) act[i] = clampf(a8[i] + S.bo[i], -1.0f, 1.0f);
    return true;
}

__device__ __forceinline__ bool ready_pred_tc(float pos_thr, float ori_thr, float a_thr, float dq_thr, float pos, float ori, float an, float dqn) {
    return pos_thr > 0.0f && ori_thr > 0.0f && pos <= pos_thr && ori <= ori_thr && (a_thr <= 0.0f || an <= a_thr) && (dq_thr <= 0.0f || dqn <= dq_thr);
}

// Launch shape: blockDim.x = 32 * W threads, W = 1..16 warps = up to 4 tiles of 4 warps (the last tile may have fewer warps: its
// GEMMs are still M = 128, the rows of the missing warps are simply nobody's); shared memory and TMEM are sized by the tile count.
// A batch that fits one wave gets ONE CTA per SM with W = ceil(warps / SMs), so every SM carries the same number of episodes
// (65 536 episodes -> 148 CTAs x 14 warps); larger batches run 16-warp CTAs wave after wave.
__global__ void __launch_bounds__(TC_MAX_THREADS, 1)
kin_rollout_tc_kernel(const __grid_constant__ KinEnvParams PA, const __grid_constant__ KinEnvParams PF, DevPolicyTc pol_a, DevPolicyTc pol_f,
                      int has_finisher, const float* __restrict__ iq, const float* __restrict__ idq, const float* __restrict__ ipa,
                      const float* __restrict__ gq, const float* __restrict__ gpose, int n, int stride, int confirm,
                      uint32_t* __restrict__ result, unsigned long long* __restrict__ env_steps) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // round up to 1024 bytes WITHOUT leaving the shared address space (pointer + offset, not an integer round trip), so every
    // access below compiles to LDS / STS rather than generic LD / ST with 64-bit address arithmetic
    TcSmem& S = *reinterpret_cast<TcSmem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    TileCtx c;
    c.tile = tid >> 7;
    c.row = tid & (TC_TILE - 1);
    c.tile_threads = min(TC_TILE, (int)blockDim.x - c.tile * TC_TILE);
    c.A = &S.A[0][0] + (size_t)c.tile * A_TILE_FLOATS;
    c.issuer = c.row == 0;
    c.parity = 0u;
    c.flag_buf = 0;

    const int n_tiles_cta = ((int)blockDim.x + TC_TILE - 1) / TC_TILE;
    const unsigned tmem_cols = n_tiles_cta == 1 ? 64u : (n_tiles_cta == 2 ? 128u : 256u);   // 64 accumulator columns per tile, power of two
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < TC_TILES; ++i) mbar_init(smem_u32(&S.mbar[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 2 * TC_TILES * 4) (&S.run_flags[0][0][0])[tid] = 0;   // warps a partial tile does not have never vote
    load_weights_tc(S, pol_a, tid, (int)blockDim.x);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem_base = S.tmem_base;
    c.a_saddr = smem_u32(c.A);
    c.w0_saddr = smem_u32(S.W0);
    c.w1_saddr = smem_u32(S.W1);
    c.wo_saddr = smem_u32(S.WO);
    c.mbar_saddr = smem_u32(&S.mbar[c.tile]);
    c.tmem_d = tmem_base + c.tile * TC_HID;                                   // lane 0, this tile's 64 columns
    c.tmem_row = c.tmem_d + ((unsigned)((warp & 3) * 32) << 16);              // this warp's 32-lane slice

    const int ep = blockIdx.x * (int)blockDim.x + tid;
    const bool active = ep < n;
    const int epc = active ? ep : n - 1;

    EnvRegs s;
    s.flags = 0u;
    float goal_q[NJ];
    {
        float r_iq[NJ], r_idq[NJ], r_ipa[NJ], r_gq[NJ], r_gp[6];
#pragma unroll
        for (int k = 0; k < NJ; ++k) {
            r_iq[k] = iq[(size_t)epc * NJ + k];
            r_idq[k] = idq ? idq[(size_t)epc * NJ + k] : 0.0f;
            r_ipa[k] = ipa ? ipa[(size_t)epc * NJ + k] : 0.0f;
            r_gq[k] = gq ? gq[(size_t)epc * NJ + k] : 0.0f;
        }
        if (gpose) {
#pragma unroll
            for (int k = 0; k < 6; ++k) r_gp[k] = gpose[(size_t)epc * 6 + k];
        }
        reset_core(PA, s, KIN_MODE_APPROACH, r_iq, r_idq, r_ipa, r_gq, gpose ? r_gp : nullptr, goal_q);
    }

    // ---- approach phase ------------------------------------------------------------------------------
    float min_pos = s.entry[0], min_ori = s.entry[1];
    int steps = 0, streak = 0, max_streak = 0, first_ready = -1;
    bool ready_hit = false, have_snap = false;
    float snap[3 * NJ];
    int snap_step = -1;
    float last_an = 0.0f, last_dqn = 0.0f;
    StepOut so;
    so.done = 0u; so.pos = s.entry[0]; so.ori = s.entry[1]; so.dq_l2 = 0.0f;
    bool running = active;
    pose_error(s.ee, s.goal, so.pe, so.oe);
#pragma unroll
    for (int i = 0; i < NJ; ++i) so.margin[i] = joint_margin(PA, s.q[i], i);
    while (true) {
        float o[OBS], act[NJ];
        build_obs_from(PA, s, KIN_MODE_APPROACH, so.pe, so.oe, so.margin, o);
        if (!mlp_tc(S, c, o, act, running)) break;
        if (running) {
            float an2 = 0.0f;
#pragma unroll
            for (int i = 0; i < NJ; ++i) an2 = fmaf(act[i], act[i], an2);
            const float an = sqrtf(an2);
            step_core<KIN_MODE_APPROACH, false>(PA, s, act, so, nullptr);
            steps += 1;
            min_pos = fminf(min_pos, so.pos);
            min_ori = fminf(min_ori, so.ori);
            if (ready_pred_tc(PA.ar_dock_coarse_ready_pos_threshold_m, PA.ar_dock_coarse_ready_ori_threshold_rad,
                              PA.ar_dock_coarse_ready_action_threshold, PA.ar_dock_coarse_ready_dq_threshold, so.pos, so.ori, an, so.dq_l2)) {
                ready_hit = true;
                if (first_ready < 0) first_ready = steps;
                streak += 1;
            } else {
                streak = 0;
            }
            max_streak = max(max_streak, streak);
            if (!have_snap && streak >= confirm) {
                have_snap = true;
                snap_step = steps;
#pragma unroll
                for (int i = 0; i < NJ; ++i) { snap[i] = s.q[i]; snap[NJ + i] = s.dq[i]; snap[2 * NJ + i] = s.pa[i]; }
            }
            last_an = an;
            last_dqn = so.dq_l2;
            running = !(so.done & (KIN_DONE_TERMINATED | KIN_DONE_TRUNCATED));
        }
    }
    if (active) {   // approach-phase motion metrics (eval_workspace_expansion.py:47-66 classifies failures with them)
        result[(size_t)KIN_RES_APPROACH_ACTION * stride + ep] = __float_as_uint(last_an);
        result[(size_t)KIN_RES_APPROACH_DQ * stride + ep] = __float_as_uint(last_dqn);
    }
    const int approach_steps = steps;
    const bool approach_success = (so.done & KIN_DONE_SUCCESS) != 0;
    const float approach_pos = so.pos, approach_ori = so.ori;
    const bool final_ready = ready_pred_tc(PA.ar_finisher_ready_pos_threshold_m, PA.ar_finisher_ready_ori_threshold_rad,
                                           PA.ar_finisher_ready_action_threshold, PA.ar_finisher_ready_dq_threshold, so.pos, so.ori, last_an, last_dqn);
    const int handoff_kind = final_ready ? 2 : (have_snap ? 1 : 0);
    const int handoff_step = final_ready ? steps : (have_snap ? snap_step : -1);
    bool success = approach_success;
    float final_pos = so.pos, final_ori = so.ori, final_an = last_an, final_dqn = last_dqn;
    int finisher_steps = 0;

    // ---- finisher phase ------------------------------------------------------------------------------
    __syncthreads();   // both tiles are out of the approach loop: nobody reads the approach weights any more
    if (has_finisher) {
        load_weights_tc(S, pol_f, tid, (int)blockDim.x);
        fence_async_smem();
        __syncthreads();
        running = active && handoff_kind != 0;
        if (running) {
            float r_iq[NJ], r_idq[NJ], r_ipa[NJ], gq_out[NJ], r_gp[6];
#pragma unroll
            for (int i = 0; i < NJ; ++i) {
                r_iq[i] = (handoff_kind == 2) ? s.q[i] : snap[i];
                r_idq[i] = (handoff_kind == 2) ? s.dq[i] : snap[NJ + i];
                r_ipa[i] = (handoff_kind == 2) ? s.pa[i] : snap[2 * NJ + i];
            }
#pragma unroll
            for (int k = 0; k < 6; ++k) r_gp[k] = s.goal[k];
            reset_core(PF, s, KIN_MODE_DOCK, r_iq, r_idq, r_ipa, goal_q, r_gp, gq_out);
        }
        steps = 0;
        pose_error(s.ee, s.goal, so.pe, so.oe);
#pragma unroll
        for (int i = 0; i < NJ; ++i) so.margin[i] = joint_margin(PF, s.q[i], i);
        while (true) {
            float o[OBS], act[NJ];
            build_obs_from(PF, s, KIN_MODE_DOCK, so.pe, so.oe, so.margin, o);
            if (!mlp_tc(S, c, o, act, running)) break;
            if (running) {
                float an2 = 0.0f;
#pragma unroll
                for (int i = 0; i < NJ; ++i) an2 = fmaf(act[i], act[i], an2);
                step_core<KIN_MODE_DOCK, false>(PF, s, act, so, nullptr);
                steps += 1;
                final_an = sqrtf(an2);
                running = !(so.done & (KIN_DONE_TERMINATED | KIN_DONE_TRUNCATED));
            }
        }
        if (active && handoff_kind != 0) {
            finisher_steps = steps;
            success = (so.done & KIN_DONE_SUCCESS) != 0;
            final_pos = so.pos; final_ori = so.ori; final_dqn = so.dq_l2;
        }
    }

    if (active) {
        auto put_u = [&](int row, uint32_t v) { result[(size_t)row * stride + ep] = v; };
        auto put_f = [&](int row, float v) { result[(size_t)row * stride + ep] = __float_as_uint(v); };
        put_u(KIN_RES_SUCCESS, success ? 1u : 0u);
        put_u(KIN_RES_FLAGS, (approach_success ? 1u : 0u) | ((ready_hit || final_ready) ? 2u : 0u) |
                                 ((max_streak >= confirm || final_ready) ? 4u : 0u) | (final_ready ? 8u : 0u) | ((uint32_t)handoff_kind << 4));
        put_u(KIN_RES_HANDOFF_STEP, (uint32_t)handoff_step);
        put_u(KIN_RES_FIRST_READY_STEP, (uint32_t)first_ready);
        put_u(KIN_RES_MAX_READY_STREAK, (uint32_t)max_streak);
        put_u(KIN_RES_STEPS, (uint32_t)approach_steps | ((uint32_t)finisher_steps << 16));
        put_f(KIN_RES_FINAL_POS, final_pos); put_f(KIN_RES_FINAL_ORI, final_ori);
        put_f(KIN_RES_APPROACH_POS, approach_pos); put_f(KIN_RES_APPROACH_ORI, approach_ori);
        put_f(KIN_RES_MIN_POS, min_pos); put_f(KIN_RES_MIN_ORI, min_ori);
        put_f(KIN_RES_FINAL_ACTION, final_an); put_f(KIN_RES_FINAL_DQ, final_dqn);
#pragma unroll
        for (int i = 0; i < NJ; ++i) put_f(KIN_RES_FINAL_Q + i, s.q[i]);
    }
    if (env_steps) {
        unsigned long long mine = active ? (unsigned long long)(approach_steps + finisher_steps) : 0ull;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, off);
        if ((tid & 31) == 0 && mine) atomicAdd(env_steps, mine);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
}

}  // namespace kin

using namespace kin;

int kin_rollout_tc_launch(const KinHandle* ha, const KinHandle* hf, const KinPolicyWeights* pa, const KinPolicyWeights* pf,
                          const float* iq, const float* idq, const float* ipa, const float* gq, const float* gpose, int n, int stride,
                          int confirm, int variant, uint32_t* result, unsigned long long* env_steps, cudaStream_t st) {
    (void)variant;
    static bool attr_set = false;
    if (!attr_set) {
        const size_t smem_max = sizeof(TcSmem) + (TC_TILES - 1) * A_TILE_FLOATS * sizeof(float) + 1024;
        cudaError_t e = cudaFuncSetAttribute(kin_rollout_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
        if (e != cudaSuccess) return kin_fail_cuda(e, "kin_rollout_approach_finisher(tc): smem attribute");
        attr_set = true;
    }
    DevPolicyTc da{pa->pi_w0, pa->pi_b0, pa->pi_w1, pa->pi_b1, pa->act_w, pa->act_b};
    DevPolicyTc df = pf ? DevPolicyTc{pf->pi_w0, pf->pi_b0, pf->pi_w1, pf->pi_b1, pf->act_w, pf->act_b} : da;
    const KinEnvParams& PF = hf ? hf->params : ha->params;
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0) n_sm = 148;
    }
    // one wave: spread the warps evenly over the SMs; several waves: full 16-warp CTAs, one per SM (measured on the 1 M-pair
    // random-start sweep: 8.9 G env-steps/s vs 8.1 G with two 8-warp CTAs per SM)
    const int warps = (n + 31) / 32;
    int w_per_cta = (warps + n_sm - 1) / n_sm;
    if (w_per_cta > 16) w_per_cta = 16;
    if (const char* v = getenv("KIN_TC_WARPS")) { const int f = atoi(v); if (f >= 1 && f <= 16) w_per_cta = f; }
    const int threads = 32 * w_per_cta;
    const size_t smem = sizeof(TcSmem) + ((threads + TC_TILE - 1) / TC_TILE - 1) * A_TILE_FLOATS * sizeof(float) + 1024;
    kin_rollout_tc_kernel<<<(n + threads - 1) / threads, threads, smem, st>>>(ha->params, PF, da, df, (hf && pf) ? 1 : 0, iq, idq, ipa, gq,
                                                                                       gpose, n, stride, confirm, result, env_steps);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? KIN_OK : kin_fail_cuda(e, "kin_rollout_approach_finisher(tc)");
}
