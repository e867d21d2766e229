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

#include "kin_tc_mlp.cuh"

namespace kin {

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
    static bool attr_set[KIN_MAX_DEVICES] = {};
    const int dev_slot = kin_device_slot();
    if (!attr_set[dev_slot]) {
        const size_t smem_max = sizeof(TcSmem) + (TC_TILES - 1) * A_TILE_FLOATS * sizeof(float) + 1024;
        cudaError_t e = cudaFuncSetAttribute(kin_rollout_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
        if (e != cudaSuccess) return kin_fail_cuda(e, "kin_rollout_approach_finisher(tc): smem attribute");
        attr_set[dev_slot] = true;
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
