// kin_rollout.cu -- K2 (fp32 FFMA variant): the fused policy-in-loop Approach -> Finisher rollout, and the
// standalone batched policy forward.
//
// One thread per episode.  The env state never leaves registers for the whole episode (128 approach steps +
// 36 finisher steps in the official configs); both policies' weights are staged in shared memory; the
// Approach -> Finisher handoff (ready streak, first-confirmed snapshot, final-settled override, hand-over of
// q / dq / prev_action into a dock-mode env) is per-lane predicated logic -- lanes that finish early idle
// until their warp is done.  HBM traffic is the episode inputs (27 floats) and the result rows, i.e. < 1 B
// per env-step; the kernel is bound by the FP32 pipe (MLP) and this variant exists for strict-fp32 parity.
//
// Replaces the per-episode Python loops of eval/eval_workspace_expansion.py:126-147 and
// eval/eval_full_workspace_coverage.py:120-164: _run_approach_with_handoff
// (eval_pipeline_ablation.py:60-147), _dock_coarse_ready (eval_three_stage.py:41-56), _finisher_ready
// (eval_approach_finisher.py:24-32), _state_reset_options + _run_policy (eval_three_stage.py:30-125).
#include "kin_internal.h"
#include "kin_mlp.cuh"
#include "kin_state.cuh"

namespace kin {

constexpr int RO_THREADS = 128;

struct DevPolicy {
    const float *w0, *b0, *w1, *b1, *wo, *bo;
};

__device__ __forceinline__ bool ready_pred(float pos_thr, float ori_thr, float a_thr, float dq_thr, float pos, float ori, float an, float dqn) {
    return pos_thr > 0.0f && ori_thr > 0.0f && pos <= pos_thr && ori <= ori_thr && (a_thr <= 0.0f || an <= a_thr) && (dq_thr <= 0.0f || dqn <= dq_thr);
}

__global__ void __launch_bounds__(RO_THREADS)
kin_rollout_ffma_kernel(const __grid_constant__ KinEnvParams PA, const __grid_constant__ KinEnvParams PF, DevPolicy pol_a, DevPolicy pol_f,
                        int has_finisher, const float* __restrict__ iq, const float* __restrict__ idq, const float* __restrict__ ipa,
                        const float* __restrict__ gq, const float* __restrict__ gpose, int n, int stride, int confirm,
                        uint32_t* __restrict__ result, unsigned long long* __restrict__ env_steps, float* __restrict__ handoff_out) {
    extern __shared__ __align__(16) float smem[];
    float* sw = smem;                                   // weights of the policy currently in the loop
    float* scratch_all = smem + MlpSmem<OBS>::FLOATS;   // [64][RO_THREADS]
    const int tid = threadIdx.x;
    const int ep = blockIdx.x * RO_THREADS + tid;
    const bool active = ep < n;
    const int epc = active ? ep : n - 1;
    float* scratch = scratch_all + tid;

    mlp_load_smem<OBS>(sw, pol_a.w0, pol_a.b0, pol_a.w1, pol_a.b1, pol_a.wo, pol_a.bo, ACT, tid, RO_THREADS);

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
    __syncthreads();

    // ---- approach phase: _run_approach_with_handoff --------------------------------------------------
    float min_pos = s.entry[0], min_ori = s.entry[1];
    int steps = 0, streak = 0, max_streak = 0, first_ready = -1;
    bool ready_hit = false, have_snap = false;
    float snap[3 * NJ];   // q, dq, prev_action at the first confirmed handoff
    float snap_m[4] = {0.0f, 0.0f, 0.0f, 0.0f};   // its position / orientation error, action magnitude, dq norm
    int snap_step = -1;
    float last_an = 0.0f, last_dqn = 0.0f;
    StepOut so;
    so.done = 0u; so.pos = s.entry[0]; so.ori = s.entry[1];
    bool running = active;
    while (__any_sync(0xffffffffu, running)) {
        if (running) {
            float o[OBS], act[ACT];
            build_obs(PA, s, KIN_MODE_APPROACH, o);
            mlp_forward<OBS, ACT, RO_THREADS>(sw, o, act, scratch);
            float an2 = 0.0f;
#pragma unroll
            for (int i = 0; i < ACT; ++i) {
                act[i] = clampf(act[i], -1.0f, 1.0f);  // predict() clips to the Box bounds
                an2 = fmaf(act[i], act[i], an2);
            }
            const float an = sqrtf(an2);
            step_core<KIN_MODE_APPROACH, false>(PA, s, act, so, nullptr);
            steps += 1;
            min_pos = fminf(min_pos, so.pos);
            min_ori = fminf(min_ori, so.ori);
            if (ready_pred(PA.ar_dock_coarse_ready_pos_threshold_m, PA.ar_dock_coarse_ready_ori_threshold_rad,
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
                snap_m[0] = so.pos; snap_m[1] = so.ori; snap_m[2] = an; snap_m[3] = so.dq_l2;
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
    const bool final_ready = ready_pred(PA.ar_finisher_ready_pos_threshold_m, PA.ar_finisher_ready_ori_threshold_rad,
                                        PA.ar_finisher_ready_action_threshold, PA.ar_finisher_ready_dq_threshold, so.pos, so.ori, last_an, last_dqn);
    int handoff_kind = final_ready ? 2 : (have_snap ? 1 : 0);
    int handoff_step = final_ready ? steps : (have_snap ? snap_step : -1);
    bool success = approach_success;
    float final_pos = so.pos, final_ori = so.ori, final_an = last_an, final_dqn = last_dqn;
    int finisher_steps = 0;

    // ---- handoff states for the Finisher's reset buffer (training/build_finisher_handoff_state_buffer.py:73-112) -----
    if (handoff_out && active) {
        auto put = [&](int row, float v) { handoff_out[(size_t)row * stride + ep] = v; };
#pragma unroll
        for (int i = 0; i < NJ; ++i) {
            put(KIN_HO_FINAL_Q + i, s.q[i]); put(KIN_HO_FINAL_DQ + i, s.dq[i]); put(KIN_HO_FINAL_PA + i, s.pa[i]);
            put(KIN_HO_SNAP_Q + i, have_snap ? snap[i] : 0.0f); put(KIN_HO_SNAP_DQ + i, have_snap ? snap[NJ + i] : 0.0f);
            put(KIN_HO_SNAP_PA + i, have_snap ? snap[2 * NJ + i] : 0.0f);
            put(KIN_HO_GOAL_Q + i, goal_q[i]);
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) put(KIN_HO_GOAL_POSE + k, s.goal[k]);
        put(KIN_HO_FINAL_METRICS + 0, so.pos); put(KIN_HO_FINAL_METRICS + 1, so.ori);
        put(KIN_HO_FINAL_METRICS + 2, last_an); put(KIN_HO_FINAL_METRICS + 3, last_dqn);
#pragma unroll
        for (int k = 0; k < 4; ++k) put(KIN_HO_SNAP_METRICS + k, snap_m[k]);
        put(KIN_HO_FINAL_STEP, (float)steps);
        put(KIN_HO_SNAP_STEP, (float)snap_step);
    }

    // ---- finisher phase: _run_policy(dock) from the handed-over state ----------------------------------
    __syncthreads();  // everyone is done reading the approach weights
    if (has_finisher) {
        mlp_load_smem<OBS>(sw, pol_f.w0, pol_f.b0, pol_f.w1, pol_f.b1, pol_f.wo, pol_f.bo, ACT, tid, RO_THREADS);
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
        while (__any_sync(0xffffffffu, running)) {
            if (running) {
                float o[OBS], act[ACT];
                build_obs(PF, s, KIN_MODE_DOCK, o);
                mlp_forward<OBS, ACT, RO_THREADS>(sw, o, act, scratch);
                float an2 = 0.0f;
#pragma unroll
                for (int i = 0; i < ACT; ++i) {
                    act[i] = clampf(act[i], -1.0f, 1.0f);
                    an2 = fmaf(act[i], act[i], an2);
                }
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
}

// standalone batched policy forward: obs [n,IN] -> action [n,7] (clipped), value [n] (optional)
template <int IN>
__global__ void __launch_bounds__(RO_THREADS)
kin_policy_forward_kernel(DevPolicy pi, DevPolicy vf, int has_value, const float* __restrict__ obs, float* __restrict__ action,
                          float* __restrict__ value, int n) {
    extern __shared__ __align__(16) float smem[];
    float* sw = smem;
    float* scratch = smem + MlpSmem<IN>::FLOATS + threadIdx.x;
    const int tid = threadIdx.x;
    const int i = blockIdx.x * RO_THREADS + tid;
    const int ic = min(i, n - 1);
    float x[IN];
    const float4* src = reinterpret_cast<const float4*>(obs + (size_t)ic * IN);
#pragma unroll
    for (int k = 0; k < IN / 4; ++k) {
        float4 v = __ldg(src + k);
        x[4 * k] = v.x; x[4 * k + 1] = v.y; x[4 * k + 2] = v.z; x[4 * k + 3] = v.w;
    }
    mlp_load_smem<IN>(sw, pi.w0, pi.b0, pi.w1, pi.b1, pi.wo, pi.bo, ACT, tid, RO_THREADS);
    __syncthreads();
    float act[ACT];
    mlp_forward<IN, ACT, RO_THREADS>(sw, x, act, scratch);
    if (i < n) {
#pragma unroll
        for (int k = 0; k < ACT; ++k) action[(size_t)i * ACT + k] = clampf(act[k], -1.0f, 1.0f);
    }
    if (has_value && value) {
        __syncthreads();
        mlp_load_smem<IN>(sw, vf.w0, vf.b0, vf.w1, vf.b1, vf.wo, vf.bo, 1, tid, RO_THREADS);
        __syncthreads();
        float v[1];
        mlp_forward<IN, 1, RO_THREADS>(sw, x, v, scratch);
        if (i < n) value[i] = v[0];
    }
}

static DevPolicy actor_of(const KinPolicyWeights* w) { return DevPolicy{w->pi_w0, w->pi_b0, w->pi_w1, w->pi_b1, w->act_w, w->act_b}; }
static DevPolicy critic_of(const KinPolicyWeights* w) { return DevPolicy{w->vf_w0, w->vf_b0, w->vf_w1, w->vf_b1, w->val_w, w->val_b}; }
static bool actor_ok(const KinPolicyWeights* w) { return w && w->pi_w0 && w->pi_b0 && w->pi_w1 && w->pi_b1 && w->act_w && w->act_b; }

}  // namespace kin

using namespace kin;

// implemented in kin_rollout_tc16.cu (tcgen05 kind::f16 / TMEM variant)
int kin_rollout_tc16_launch(const KinHandle* ha, const KinHandle* hf, const KinPolicyWeights* pa, const KinPolicyWeights* pf,
                            const float* iq, const float* idq, const float* ipa, const float* gq, const float* gpose, int n, int stride,
                            int confirm, uint32_t* result, unsigned long long* env_steps, cudaStream_t st);

extern "C" int kin_rollout_approach_finisher(void* approach_handle, void* finisher_handle, const KinPolicyWeights* host_approach,
                                             const KinPolicyWeights* host_finisher, const float* initial_q, const float* initial_dq,
                                             const float* initial_prev_action, const float* goal_q, const float* goal_pose6, int n,
                                             int stride, int handoff_confirm_steps, int variant, uint32_t* result,
                                             unsigned long long* env_steps, void* stream) {
    KinHandle* ha = kin_handle(approach_handle);
    KinHandle* hf = finisher_handle ? kin_handle(finisher_handle) : nullptr;
    if (!ha || (finisher_handle && !hf)) return kin_fail(KIN_ERR_INVALID_ARG, "kin_rollout_approach_finisher: bad handle");
    if (!actor_ok(host_approach) || host_approach->in_dim != OBS) return kin_fail(KIN_ERR_INVALID_ARG, "kin_rollout_approach_finisher: approach policy must be a 56-input actor");
    const bool has_f = hf != nullptr && host_finisher != nullptr;
    if (has_f && (!actor_ok(host_finisher) || host_finisher->in_dim != OBS)) return kin_fail(KIN_ERR_INVALID_ARG, "kin_rollout_approach_finisher: finisher policy must be a 56-input actor");
    if (!initial_q || (!goal_q && !goal_pose6) || !result || n <= 0 || stride < n) return kin_fail(KIN_ERR_INVALID_ARG, "kin_rollout_approach_finisher: bad buffers / sizes");
    if (handoff_confirm_steps < 1) return kin_fail(KIN_ERR_INVALID_ARG, "kin_rollout_approach_finisher: handoff_confirm_steps >= 1");
    cudaStream_t st = (cudaStream_t)stream;
    if (variant != 0)
        return kin_rollout_tc16_launch(ha, has_f ? hf : nullptr, host_approach, has_f ? host_finisher : nullptr, initial_q, initial_dq,
                                       initial_prev_action, goal_q, goal_pose6, n, stride, handoff_confirm_steps, result, env_steps, st);
    const size_t smem = (size_t)(MlpSmem<OBS>::FLOATS + HID * RO_THREADS) * sizeof(float);
    static bool attr_set[KIN_MAX_DEVICES] = {};
    const int dev_slot = kin_device_slot();
    if (!attr_set[dev_slot]) {
        cudaError_t e = cudaFuncSetAttribute(kin_rollout_ffma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return kin_fail_cuda(e, "kin_rollout_approach_finisher: smem attribute");
        attr_set[dev_slot] = true;
    }
    const KinEnvParams& PF = has_f ? hf->params : ha->params;
    DevPolicy da = actor_of(host_approach), df = has_f ? actor_of(host_finisher) : actor_of(host_approach);
    kin_rollout_ffma_kernel<<<(n + RO_THREADS - 1) / RO_THREADS, RO_THREADS, smem, st>>>(ha->params, PF, da, df, has_f ? 1 : 0, initial_q, initial_dq,
                                                                                           initial_prev_action, goal_q, goal_pose6, n, stride,
                                                                                           handoff_confirm_steps, result, env_steps, nullptr);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? KIN_OK : kin_fail_cuda(e, "kin_rollout_approach_finisher");
}

extern "C" int kin_rollout_handoff_states(void* approach_handle, const KinPolicyWeights* host_approach, const float* initial_q, const float* initial_dq,
                                          const float* initial_prev_action, const float* goal_q, const float* goal_pose6, int n, int stride,
                                          int handoff_confirm_steps, uint32_t* result, float* handoff_out, unsigned long long* env_steps, void* stream) {
    KinHandle* ha = kin_handle(approach_handle);
    if (!ha) return kin_fail(KIN_ERR_INVALID_ARG, "kin_rollout_handoff_states: bad handle");
    if (!actor_ok(host_approach) || host_approach->in_dim != OBS) return kin_fail(KIN_ERR_INVALID_ARG, "kin_rollout_handoff_states: approach policy must be a 56-input actor");
    if (!initial_q || (!goal_q && !goal_pose6) || !result || !handoff_out || n <= 0 || stride < n) return kin_fail(KIN_ERR_INVALID_ARG, "kin_rollout_handoff_states: bad buffers / sizes");
    if (handoff_confirm_steps < 1) return kin_fail(KIN_ERR_INVALID_ARG, "kin_rollout_handoff_states: handoff_confirm_steps >= 1");
    const size_t smem = (size_t)(MlpSmem<OBS>::FLOATS + HID * RO_THREADS) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(kin_rollout_ffma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return kin_fail_cuda(e, "kin_rollout_handoff_states: smem attribute");
    DevPolicy da = actor_of(host_approach);
    kin_rollout_ffma_kernel<<<(n + RO_THREADS - 1) / RO_THREADS, RO_THREADS, smem, (cudaStream_t)stream>>>(
        ha->params, ha->params, da, da, 0, initial_q, initial_dq, initial_prev_action, goal_q, goal_pose6, n, stride, handoff_confirm_steps, result, env_steps,
        handoff_out);
    e = cudaGetLastError();
    return e == cudaSuccess ? KIN_OK : kin_fail_cuda(e, "kin_rollout_handoff_states");
}

extern "C" int kin_policy_forward(const KinPolicyWeights* w, const float* obs, float* action, float* value, int n, void* stream) {
    if (!actor_ok(w) || !obs || !action || n < 0) return kin_fail(KIN_ERR_INVALID_ARG, "kin_policy_forward: bad arguments");
    if (n == 0) return KIN_OK;
    if (((uintptr_t)obs & 15u)) return kin_fail(KIN_ERR_INVALID_ARG, "kin_policy_forward: obs must be 16-byte aligned");
    const int has_value = (value && w->has_value && w->vf_w0) ? 1 : 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = (n + RO_THREADS - 1) / RO_THREADS;
    cudaError_t e;
    if (w->in_dim == KIN_OBS_DIM) {
        const size_t smem = (size_t)(MlpSmem<KIN_OBS_DIM>::FLOATS + HID * RO_THREADS) * sizeof(float);
        e = cudaFuncSetAttribute(kin_policy_forward_kernel<KIN_OBS_DIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return kin_fail_cuda(e, "kin_policy_forward: smem attribute");
        kin_policy_forward_kernel<KIN_OBS_DIM><<<blocks, RO_THREADS, smem, st>>>(actor_of(w), critic_of(w), has_value, obs, action, value, n);
    } else if (w->in_dim == KIN_ROUTE_OBS_DIM) {
        const size_t smem = (size_t)(MlpSmem<KIN_ROUTE_OBS_DIM>::FLOATS + HID * RO_THREADS) * sizeof(float);
        e = cudaFuncSetAttribute(kin_policy_forward_kernel<KIN_ROUTE_OBS_DIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return kin_fail_cuda(e, "kin_policy_forward: smem attribute");
        kin_policy_forward_kernel<KIN_ROUTE_OBS_DIM><<<blocks, RO_THREADS, smem, st>>>(actor_of(w), critic_of(w), has_value, obs, action, value, n);
    } else {
        return kin_fail(KIN_ERR_UNSUPPORTED, "kin_policy_forward: in_dim must be 56 or 80");
    }
    e = cudaGetLastError();
    return e == cudaSuccess ? KIN_OK : kin_fail_cuda(e, "kin_policy_forward");
}
