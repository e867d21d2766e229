// kin_route.cu -- the dense holder-route wrappers on the GPU: RouteKinematicEnv / RouteSequenceKinematicEnv step with
// the in-episode waypoint advance, and the fused sequential route probe with the 80-input route policy in the loop.
//
// The route wrapper's bookkeeping is per-lane predicated logic on top of step_core: route-ready (5 thresholds), ready
// streak, success, target advance (goal swap + entry metrics), nearest-waypoint scan over the whole route (table of
// waypoint joint vectors staged in shared memory, warp-uniform reads), 13-term route reward, +24 observation floats.
// Two things the reference recomputes every step are reused instead: FK(prev_q) is the cached ee pose and FK(curr_q)
// is the pose the base step just produced (route/route_env.py:127,137 call FK a second and third time).
//
// Replaces route/route_env.py:49-212, route/route_sequence_env.py:96-278, route/reward_route.py:36-143,
// route/route_observation.py:31-61, eval/eval_route_curriculum.py:55-136,188-218 (paths under kinematic_phase1/).
#include "kin_internal.h"
#include "kin_mlp.cuh"
#include "kin_state.cuh"

namespace kin {

constexpr int ROBS = KIN_ROUTE_OBS_DIM;
constexpr int ROBS_TILE_FLOATS = WARP * ROBS;   // 2560 floats = 10 240 B per warp
constexpr int RT_THREADS = 128;
constexpr int RT_WARPS = RT_THREADS / WARP;
constexpr int ROUTE_MAX_SMEM_WP = 1024;         // waypoint joint vectors staged in smem for the nearest-waypoint scan

struct RouteView {
    int n;
    const float* q;      // [n][7]
    const float* pose;   // [n][6]
    const float* tan;    // [n][7] next_q_delta
    const float* prog;   // [n]
};

struct RouteRegs {
    int index, streak, last, completed;
};

struct RouteOut {
    float reward, q_err, nearest;
    unsigned flags;      // bit0 ready, bit1 regression, bit2 orientation hit, bit3 waypoint success
    unsigned done;       // KIN_DONE_* with route semantics
};

__device__ __forceinline__ int wp_clamp(const RouteView& R, int i) { return min(max(i, 0), R.n - 1); }

__device__ __forceinline__ float dist7(const float* a, const float* b) {
    float acc = 0.0f;
#pragma unroll
    for (int i = 0; i < NJ; ++i) acc = fmaf(a[i] - b[i], a[i] - b[i], acc);
    return sqrtf(acc);
}

// route/route_observation.py:31-61 spliced into the alphabetical 80-vector
__device__ __forceinline__ void build_route_obs(const KinEnvParams& P, const RouteView& R, const EnvRegs& s, int route_index, const float* base56, float* o) {
#pragma unroll
    for (int k = 0; k < 47; ++k) o[k] = base56[k];
    const float* goal = R.q + (size_t)wp_clamp(R, route_index) * NJ;
    const float* tan = R.tan + (size_t)wp_clamp(R, max(route_index - 1, 0)) * NJ;
#pragma unroll
    for (int i = 0; i < NJ; ++i) {
        const float g = __ldg(goal + i);
        o[47 + i] = clampf((g - s.q[i]) * P.k_inv_delta_limit[i], -1.0f, 1.0f);
        o[54 + i] = clampf(fmaf(2.0f * P.k_inv_span[i], g - P.joint_lower[i], -1.0f), -1.0f, 1.0f);
        o[64 + i] = clampf(__ldg(tan + i) * P.k_inv_delta_limit[i], -1.0f, 1.0f);
    }
    const int max_idx = R.n - 1;
    o[61] = clampf((float)route_index / (float)max(max_idx, 1), 0.0f, 1.0f);
    o[62] = clampf(__ldg(R.prog + wp_clamp(R, route_index)) / fmaxf(__ldg(R.prog + max_idx), 1e-9f), 0.0f, 1.0f);
    o[63] = 0.0f;
    o[71] = base56[47]; o[72] = base56[48]; o[73] = base56[49];
#pragma unroll
    for (int k = 0; k < 6; ++k) o[74 + k] = base56[50 + k];
}

__device__ __forceinline__ bool route_ready(const KinEnvParams& P, float qe, float pos, float ori, float an, float dqn) {
    return qe <= P.rr_route_ready_q_threshold && pos <= P.rr_route_ready_pos_threshold_m && ori <= P.rr_route_ready_ori_threshold_rad &&
           an <= P.rr_route_ready_action_threshold && dqn <= P.rr_route_ready_dq_threshold;
}

// One wrapper step.  q_table: waypoint joint vectors for the nearest scan (smem or global), nullptr -> scan skipped.
template <bool SEQ, bool COMP>
__device__ __forceinline__ void route_step_core(const KinEnvParams& P, const RouteView& R, const float* q_table, EnvRegs& s, RouteRegs& rr,
                                                const float* action, bool reset_streak_on_advance, StepOut& so, RouteOut& ro, float* rc) {
    float prev_q[NJ], prev_pa[NJ], prev_ee[6];
#pragma unroll
    for (int i = 0; i < NJ; ++i) { prev_q[i] = s.q[i]; prev_pa[i] = s.pa[i]; }
#pragma unroll
    for (int k = 0; k < 6; ++k) prev_ee[k] = s.ee[k];
    const int target = rr.index;
    float goal_q[NJ], goal_pose[6], tangent[NJ];
    {
        const float* gq = R.q + (size_t)wp_clamp(R, target) * NJ;
        const float* gp = R.pose + (size_t)wp_clamp(R, target) * 6;
        const float* tn = R.tan + (size_t)wp_clamp(R, max(target - 1, 0)) * NJ;
#pragma unroll
        for (int i = 0; i < NJ; ++i) { goal_q[i] = __ldg(gq + i); tangent[i] = __ldg(tn + i); }
#pragma unroll
        for (int k = 0; k < 6; ++k) goal_pose[k] = __ldg(gp + k);
    }
    step_core<KIN_MODE_APPROACH, false>(P, s, action, so, nullptr);

    const float q_err = dist7(goal_q, s.q), prev_q_err = dist7(goal_q, prev_q);
    // the wrapper norms the RAW action (route_env.py:140); a policy's action is already inside [-1, 1]
    float an2 = 0.0f, msq = 0.0f, dmsq = 0.0f, dot = 0.0f, tn2 = 0.0f;
#pragma unroll
    for (int i = 0; i < NJ; ++i) {
        an2 = fmaf(action[i], action[i], an2);
        const float da = action[i] - prev_pa[i];
        dmsq = fmaf(da, da, dmsq);
        dot = fmaf(s.q[i] - prev_q[i], tangent[i], dot);
        tn2 = fmaf(tangent[i], tangent[i], tn2);
    }
    msq = an2 * (1.0f / NJ);
    dmsq *= (1.0f / NJ);
    const float an = sqrtf(an2), dqn = so.dq_l2, tn = sqrtf(tn2);
    float nearest = 0.0f;
    if (q_table) {
        nearest = CUDART_INF_F;
        for (int w = 0; w < R.n; ++w) {
            const float* qw = q_table + w * NJ;
            float acc = 0.0f;
#pragma unroll
            for (int i = 0; i < NJ; ++i) acc = fmaf(qw[i] - s.q[i], qw[i] - s.q[i], acc);
            nearest = fminf(nearest, acc);
        }
        nearest = sqrtf(nearest);
    }
    const bool ready = route_ready(P, q_err, so.pos, so.ori, an, dqn);
    rr.streak = ready ? rr.streak + 1 : 0;

    // route reward (reward_route.py:54-143): pose errors against the waypoint's FK pose, prev pose = cached FK(prev_q)
    float pe[3], oe[3];
    pose_error(prev_ee, goal_pose, pe, oe);
    const float prev_pos = norm3(pe[0], pe[1], pe[2]), prev_ori = norm3(oe[0], oe[1], oe[2]);
    pose_error(s.ee, goal_pose, pe, oe);
    const float curr_pos = norm3(pe[0], pe[1], pe[2]), curr_ori = norm3(oe[0], oe[1], oe[2]);
    const float tangent_progress = tn > 0.0f ? dot / fmaxf(tn, 1e-9f) : 0.0f;
    const bool ready_r = route_ready(P, q_err, curr_pos, curr_ori, an, dqn);
    float low_motion = 0.0f;
    if (curr_pos <= 2.0f * P.rr_route_ready_pos_threshold_m && curr_ori <= 2.0f * P.rr_route_ready_ori_threshold_rad) {
        const float a_clean = fmaxf(1.0f - an / fmaxf(P.rr_route_ready_action_threshold, 1e-9f), 0.0f);
        const float d_clean = fmaxf(1.0f - dqn / fmaxf(P.rr_route_ready_dq_threshold, 1e-9f), 0.0f);
        low_motion = P.rr_low_motion_near_waypoint_bonus * 0.5f * (a_clean + d_clean);
    }
    float c[13];
    c[0] = P.rr_q_goal_progress_weight * (prev_q_err - q_err);
    c[1] = P.rr_ee_position_progress_weight * (prev_pos - curr_pos);
    c[2] = P.rr_ee_orientation_progress_weight * (prev_ori - curr_ori);
    c[3] = P.rr_route_tangent_progress_weight * fmaxf(tangent_progress, 0.0f);
    c[4] = ready_r ? P.rr_same_step_route_ready_bonus : 0.0f;
    c[5] = (ready_r && rr.streak >= 1) ? P.rr_route_ready_dwell_bonus : 0.0f;
    c[6] = low_motion;
    c[7] = -P.rr_orientation_regression_penalty_weight * fmaxf(curr_ori - prev_ori, 0.0f);
    c[8] = -P.rr_q_route_regression_penalty_weight * fmaxf(q_err - prev_q_err, 0.0f);
    c[9] = -P.rr_off_route_penalty_weight * fmaxf(nearest, 0.0f);
    float smooth = -P.rr_action_magnitude_weight * msq;
    smooth += -P.rr_action_delta_weight * dmsq;
    c[10] = smooth;
    c[11] = -P.rr_dq_penalty_weight * dqn;
    c[12] = (q_err >= prev_q_err && curr_pos >= prev_pos && curr_ori >= prev_ori) ? -P.rr_no_progress_penalty : 0.0f;
    float reward = 0.0f;
#pragma unroll
    for (int k = 0; k < 13; ++k) reward += c[k];
    if (COMP) {
#pragma unroll
        for (int k = 0; k < 13; ++k) rc[k] = c[k];
        rc[13] = q_err; rc[14] = curr_pos; rc[15] = curr_ori; rc[16] = ready_r ? 1.0f : 0.0f;
    }

    const bool wp_success = ready && rr.streak >= P.term_success_dwell_steps;
    const bool base_term = (so.done & KIN_DONE_TERMINATED) != 0;
    const unsigned base_reason = (so.done >> KIN_DONE_REASON_SHIFT) & 3u;
    bool success, terminated;
    if (!SEQ) {
        success = wp_success;
        terminated = base_term;
        if (base_term && base_reason == 1u && !success) terminated = false;
        if (success && P.term_terminate_on_success) terminated = true;
    } else {
        success = false;
        terminated = false;
        if (wp_success) {
            rr.completed += 1;
            if (target >= rr.last) {
                success = true;
                terminated = true;
            } else {   // _advance_target (route_sequence_env.py:253-257): swap the goal in place, re-capture entry metrics
                rr.index = target + 1;
                const float* gp = R.pose + (size_t)wp_clamp(R, rr.index) * 6;
#pragma unroll
                for (int k = 0; k < 6; ++k) s.goal[k] = __ldg(gp + k);
                capture_entry_metrics(s);
                if (reset_streak_on_advance) rr.streak = 0;
            }
        }
        if (base_term && !terminated && base_reason != 1u) terminated = true;
    }
    ro.reward = reward;
    ro.q_err = q_err;
    ro.nearest = nearest;
    ro.flags = (ready ? 1u : 0u) | (q_err > prev_q_err ? 2u : 0u) | (so.ori <= P.rr_route_ready_ori_threshold_rad ? 4u : 0u) | (wp_success ? 8u : 0u);
    ro.done = (terminated ? KIN_DONE_TERMINATED : 0u) | (so.done & KIN_DONE_TRUNCATED) | (success ? KIN_DONE_SUCCESS : 0u) |
              (so.done & (KIN_DONE_PRE_NEAR | KIN_DONE_NEAR)) | (base_reason << KIN_DONE_REASON_SHIFT);
}

__device__ __forceinline__ void stage_route_obs_row(float* tile, int lane, const float* o) {
    float4* dst = reinterpret_cast<float4*>(tile + lane * ROBS);
#pragma unroll
    for (int k = 0; k < ROBS / 4; ++k) dst[k] = make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
}

__device__ __forceinline__ void load_q_table(float* dst, const RouteView& R, int tid, int nthreads) {
    const int n = min(R.n, ROUTE_MAX_SMEM_WP) * NJ;
    for (int i = tid; i < n; i += nthreads) dst[i] = __ldg(R.q + i);
}

template <bool SEQ, bool COMP>
__global__ void __launch_bounds__(RT_THREADS)
kin_route_step_kernel(const __grid_constant__ KinEnvParams P, RouteView R, float* __restrict__ state, int stride, int n,
                      const float* __restrict__ action, float* __restrict__ obs, float* __restrict__ reward, uint8_t* __restrict__ done,
                      float* __restrict__ raux, float* __restrict__ rcomp, int reset_streak) {
    extern __shared__ __align__(128) float smem[];
    float* tiles = smem;                                        // [RT_WARPS][ROBS_TILE_FLOATS]
    float* q_table = smem + RT_WARPS * ROBS_TILE_FLOATS;        // [min(n_wp, 1024)][7]
    load_q_table(q_table, R, threadIdx.x, RT_THREADS);
    __syncthreads();
    const int lane = threadIdx.x & (WARP - 1), warp = threadIdx.x >> 5;
    const int env0 = (blockIdx.x * RT_WARPS + warp) * WARP;
    if (env0 >= n) return;
    const int env = env0 + lane;
    const bool active = env < n;
    const int envc = active ? env : n - 1;
    float* tile = tiles + warp * ROBS_TILE_FLOATS;
    float a[NJ];
    load_action_tile(action, env0, n, tile, lane, a);
    EnvRegs s;
    load_env<true>(state, stride, envc, s);
    RouteRegs rr;
    {
        const unsigned r0 = ld_row_u(state, stride, KIN_ROW_ROUTE, envc), r1 = ld_row_u(state, stride, KIN_ROW_ROUTE2, envc);
        rr.index = (int)(r0 & 0xffffu); rr.streak = (int)(r0 >> 16); rr.last = (int)(r1 & 0xffffu); rr.completed = (int)(r1 >> 16);
    }
    StepOut so;
    RouteOut ro;
    float rc[COMP ? 17 : 1];
    const float* table = R.n <= ROUTE_MAX_SMEM_WP ? q_table : R.q;
    route_step_core<SEQ, COMP>(P, R, table, s, rr, a, reset_streak != 0, so, ro, rc);
    if (active) {
        store_env_step(state, stride, env, s);
        if (SEQ) {   // the advance rewrote goal pose + entry metrics
#pragma unroll
            for (int k = 0; k < 6; ++k) st_row(state, stride, KIN_ROW_GOAL_POSE + k, env, s.goal[k]);
#pragma unroll
            for (int k = 0; k < 4; ++k) st_row(state, stride, KIN_ROW_ENTRY + k, env, s.entry[k]);
            const float* gq = R.q + (size_t)wp_clamp(R, rr.index) * NJ;
#pragma unroll
            for (int i = 0; i < NJ; ++i) st_row(state, stride, KIN_ROW_GOAL_Q + i, env, __ldg(gq + i));
        }
        st_row_u(state, stride, KIN_ROW_ROUTE, env, (unsigned)rr.index | ((unsigned)min(rr.streak, 0xffff) << 16));
        st_row_u(state, stride, KIN_ROW_ROUTE2, env, (unsigned)rr.last | ((unsigned)min(rr.completed, 0xffff) << 16));
        reward[env] = ro.reward;
        done[env] = (uint8_t)ro.done;
        if (raux) {
            raux[(size_t)KIN_RAUX_Q_ERR * stride + env] = ro.q_err;
            raux[(size_t)KIN_RAUX_NEAREST * stride + env] = ro.nearest;
            raux[(size_t)KIN_RAUX_POS_ERR * stride + env] = so.pos;
            raux[(size_t)KIN_RAUX_ORI_ERR * stride + env] = so.ori;
            raux[(size_t)KIN_RAUX_FLAGS * stride + env] = __uint_as_float(ro.flags);
            raux[(size_t)KIN_RAUX_ROUTE_INDEX * stride + env] = __uint_as_float((unsigned)rr.index);
            raux[(size_t)KIN_RAUX_STREAK * stride + env] = __uint_as_float((unsigned)rr.streak);
            raux[(size_t)KIN_RAUX_COMPLETED * stride + env] = __uint_as_float((unsigned)rr.completed);
        }
        if (COMP) {
#pragma unroll
            for (int k = 0; k < 17; ++k) rcomp[(size_t)k * stride + env] = rc[k];
        }
    }
    float o56[OBS], o[ROBS];
    build_obs(P, s, KIN_MODE_APPROACH, o56);   // after an advance the goal changed, so recompute the goal errors
    build_route_obs(P, R, s, rr.index, o56, o);
    stage_route_obs_row(tile, lane, o);
    bulk_store_tile(obs + (size_t)env0 * ROBS, tile, min(WARP, n - env0) * ROBS * 4, lane);
    bulk_store_wait_read(lane);
}

__global__ void __launch_bounds__(128)
kin_route_reset_kernel(const __grid_constant__ KinEnvParams P, RouteView R, float* __restrict__ state, int stride, int n_envs,
                       const int* __restrict__ env_ids, int n_reset, const int* __restrict__ route_index, const int* __restrict__ start_index,
                       const int* __restrict__ last_index, const float* __restrict__ iq, const float* __restrict__ idq,
                       const float* __restrict__ ipa, float* __restrict__ obs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_reset) return;
    const int env = env_ids ? env_ids[i] : i;
    if (env < 0 || env >= n_envs) return;
    const int ri = route_index[i];
    const int st = start_index ? start_index[i] : max(ri - 1, 0);
    float r_iq[NJ], r_idq[NJ], r_ipa[NJ], r_gq[NJ], gq_out[NJ];
    const float* sq = R.q + (size_t)wp_clamp(R, st) * NJ;
    const float* gq = R.q + (size_t)wp_clamp(R, ri) * NJ;
#pragma unroll
    for (int k = 0; k < NJ; ++k) {
        r_iq[k] = iq ? iq[(size_t)i * NJ + k] : __ldg(sq + k);
        r_idq[k] = idq ? idq[(size_t)i * NJ + k] : 0.0f;
        r_ipa[k] = ipa ? ipa[(size_t)i * NJ + k] : 0.0f;
        r_gq[k] = __ldg(gq + k);
    }
    EnvRegs s;
    s.flags = 0u;
    reset_core(P, s, KIN_MODE_APPROACH, r_iq, r_idq, r_ipa, r_gq, nullptr, gq_out);
    store_env_reset(state, stride, env, s, gq_out);
    st_row_u(state, stride, KIN_ROW_ROUTE, env, (unsigned)ri);
    st_row_u(state, stride, KIN_ROW_ROUTE2, env, (unsigned)(last_index ? last_index[i] : ri));
    if (obs) {
        float o56[OBS], o[ROBS];
        build_obs(P, s, KIN_MODE_APPROACH, o56);
        build_route_obs(P, R, s, ri, o56, o);
        float4* dst = reinterpret_cast<float4*>(obs + (size_t)i * ROBS);
#pragma unroll
        for (int k = 0; k < ROBS / 4; ++k) dst[k] = make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
    }
}

struct DevPolicyR {
    const float *w0, *b0, *w1, *b1, *wo, *bo;
};

// evaluate_sequential_route for one replica per thread, policy in the loop (strict fp32 MLP)
__global__ void __launch_bounds__(RT_THREADS)
kin_route_probe_kernel(const __grid_constant__ KinEnvParams P, RouteView R, DevPolicyR pol, const float* __restrict__ start_q, int start_index,
                       int end_index, int n, int* __restrict__ prefix_out, uint32_t* __restrict__ success_bits, int words,
                       unsigned long long* __restrict__ env_steps) {
    extern __shared__ __align__(16) float smem[];
    float* sw = smem;
    float* scratch = smem + MlpSmem<ROBS>::FLOATS + threadIdx.x;
    const int tid = threadIdx.x;
    const int rep = blockIdx.x * RT_THREADS + tid;
    const bool active = rep < n;
    const int repc = active ? rep : n - 1;
    mlp_load_smem<ROBS>(sw, pol.w0, pol.b0, pol.w1, pol.b1, pol.wo, pol.bo, ACT, tid, RT_THREADS);
    __syncthreads();
    float cq[NJ], cdq[NJ], cpa[NJ];
    {
        const float* q0 = start_q ? start_q + (size_t)repc * NJ : R.q + (size_t)wp_clamp(R, max(start_index - 1, 0)) * NJ;
#pragma unroll
        for (int i = 0; i < NJ; ++i) { cq[i] = q0[i]; cdq[i] = 0.0f; cpa[i] = 0.0f; }
    }
    const int final_end = min(end_index, R.n - 1);
    int prefix = 0;
    bool broken = false;
    unsigned long long steps = 0;
    unsigned word = 0u;
    for (int idx = start_index; idx <= final_end; ++idx) {
        EnvRegs s;
        s.flags = 0u;
        float gq_out[NJ], gq[NJ];
        const float* g = R.q + (size_t)wp_clamp(R, idx) * NJ;
#pragma unroll
        for (int i = 0; i < NJ; ++i) gq[i] = __ldg(g + i);
        reset_core(P, s, KIN_MODE_APPROACH, cq, cdq, cpa, gq, nullptr, gq_out);
        RouteRegs rr{idx, 0, idx, 0};
        RouteOut ro;
        ro.done = 0u;
        bool running = active;
        while (__any_sync(0xffffffffu, running)) {
            if (running) {
                float o56[OBS], o[ROBS], act[ACT];
                StepOut so;
                build_obs(P, s, KIN_MODE_APPROACH, o56);
                build_route_obs(P, R, s, rr.index, o56, o);
                mlp_forward<ROBS, ACT, RT_THREADS>(sw, o, act, scratch);
#pragma unroll
                for (int i = 0; i < ACT; ++i) act[i] = clampf(act[i], -1.0f, 1.0f);
                route_step_core<false, false>(P, R, nullptr, s, rr, act, true, so, ro, nullptr);   // eval needs no off-route term
                steps += 1;
                running = !(ro.done & (KIN_DONE_TERMINATED | KIN_DONE_TRUNCATED));
            }
        }
        const bool ok = (ro.done & KIN_DONE_SUCCESS) != 0;
        if (ok && !broken) prefix += 1; else broken = true;
        const int k = idx - start_index;
        if (ok) word |= 1u << (k & 31);
        if (success_bits && active && ((k & 31) == 31 || idx == final_end)) {
            success_bits[(size_t)rep * words + (k >> 5)] = word;
            word = 0u;
        }
#pragma unroll
        for (int i = 0; i < NJ; ++i) { cq[i] = s.q[i]; cdq[i] = s.dq[i]; cpa[i] = s.pa[i]; }
    }
    if (active) prefix_out[rep] = prefix;
    if (env_steps) {
        unsigned long long mine = active ? steps : 0ull;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, off);
        if ((tid & 31) == 0 && mine) atomicAdd(env_steps, mine);
    }
}

static bool route_ok(const KinRouteTable* r) { return r && r->n_waypoints >= 2 && r->n_waypoints <= 65535 && r->q_goal && r->pose6 && r->next_q_delta && r->progress_m; }
static RouteView view_of(const KinRouteTable* r) { return RouteView{r->n_waypoints, r->q_goal, r->pose6, r->next_q_delta, r->progress_m}; }

}  // namespace kin

using namespace kin;

extern "C" int kin_route_reset(void* handle, const KinRouteTable* host_route, float* state, int stride, int n_envs, const int* env_ids,
                               int n_reset, const int* route_index, const int* start_route_index, const int* last_route_index,
                               const float* initial_q, const float* initial_dq, const float* initial_prev_action, float* obs, void* stream) {
    KinHandle* h = kin_handle(handle);
    if (!h || !route_ok(host_route)) return kin_fail(KIN_ERR_INVALID_ARG, "kin_route_reset: bad handle or route table");
    if (!state || !route_index || n_envs <= 0 || stride < n_envs || (stride % 32) != 0 || n_reset < 0) return kin_fail(KIN_ERR_INVALID_ARG, "kin_route_reset: bad sizes");
    if (obs && ((uintptr_t)obs & 15u)) return kin_fail(KIN_ERR_INVALID_ARG, "kin_route_reset: obs must be 16-byte aligned");
    if (n_reset == 0) return KIN_OK;
    kin_route_reset_kernel<<<(n_reset + 127) / 128, 128, 0, (cudaStream_t)stream>>>(h->params, view_of(host_route), state, stride, n_envs, env_ids, n_reset,
                                                                                    route_index, start_route_index, last_route_index, initial_q,
                                                                                    initial_dq, initial_prev_action, obs);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? KIN_OK : kin_fail_cuda(e, "kin_route_reset");
}

extern "C" int kin_route_step(void* handle, const KinRouteTable* host_route, float* state, int stride, int n_envs, const float* action,
                              float* obs, float* reward, uint8_t* done, float* raux, float* rcomp, int sequence_mode,
                              int reset_ready_streak_on_advance, void* stream) {
    KinHandle* h = kin_handle(handle);
    if (!h || !route_ok(host_route)) return kin_fail(KIN_ERR_INVALID_ARG, "kin_route_step: bad handle or route table");
    if (!state || !action || !obs || !reward || !done || n_envs <= 0 || stride < n_envs || (stride % 32) != 0) return kin_fail(KIN_ERR_INVALID_ARG, "kin_route_step: bad buffers / sizes");
    if (((uintptr_t)obs & 15u)) return kin_fail(KIN_ERR_INVALID_ARG, "kin_route_step: obs must be 16-byte aligned");
    const RouteView R = view_of(host_route);
    const size_t smem = (size_t)(RT_WARPS * ROBS_TILE_FLOATS + min(R.n, ROUTE_MAX_SMEM_WP) * NJ) * sizeof(float);
    const int blocks = (n_envs + RT_THREADS - 1) / RT_THREADS;
    cudaStream_t st = (cudaStream_t)stream;
#define KIN_RLAUNCH(SEQ, COMP)                                                                                                          \
    do {                                                                                                                                \
        cudaError_t ea = cudaFuncSetAttribute(kin_route_step_kernel<SEQ, COMP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (ea != cudaSuccess) return kin_fail_cuda(ea, "kin_route_step: smem attribute");                                              \
        kin_route_step_kernel<SEQ, COMP><<<blocks, RT_THREADS, smem, st>>>(h->params, R, state, stride, n_envs, action, obs, reward, done, \
                                                                          raux, rcomp, reset_ready_streak_on_advance);                   \
    } while (0)
    if (sequence_mode) { if (rcomp) KIN_RLAUNCH(true, true); else KIN_RLAUNCH(true, false); }
    else { if (rcomp) KIN_RLAUNCH(false, true); else KIN_RLAUNCH(false, false); }
#undef KIN_RLAUNCH
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? KIN_OK : kin_fail_cuda(e, "kin_route_step");
}

extern "C" int kin_route_probe(void* handle, const KinRouteTable* host_route, const KinPolicyWeights* w, const float* start_q, int start_index,
                               int end_index, int n, int* prefix, uint32_t* success_bits, unsigned long long* env_steps, void* stream) {
    KinHandle* h = kin_handle(handle);
    if (!h || !route_ok(host_route)) return kin_fail(KIN_ERR_INVALID_ARG, "kin_route_probe: bad handle or route table");
    if (!w || w->in_dim != ROBS || !w->pi_w0 || !w->pi_b0 || !w->pi_w1 || !w->pi_b1 || !w->act_w || !w->act_b) return kin_fail(KIN_ERR_INVALID_ARG, "kin_route_probe: need an 80-input actor");
    if (!prefix || n <= 0 || start_index < 1 || end_index < start_index) return kin_fail(KIN_ERR_INVALID_ARG, "kin_route_probe: bad sizes / indices");
    const size_t smem = (size_t)(MlpSmem<ROBS>::FLOATS + HID * RT_THREADS) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(kin_route_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return kin_fail_cuda(e, "kin_route_probe: smem attribute");
    const int final_end = end_index < host_route->n_waypoints - 1 ? end_index : host_route->n_waypoints - 1;
    const int words = (final_end - start_index + 1 + 31) / 32;
    DevPolicyR p{w->pi_w0, w->pi_b0, w->pi_w1, w->pi_b1, w->act_w, w->act_b};
    kin_route_probe_kernel<<<(n + RT_THREADS - 1) / RT_THREADS, RT_THREADS, smem, (cudaStream_t)stream>>>(h->params, view_of(host_route), p, start_q,
                                                                                                        start_index, end_index, n, prefix, success_bits,
                                                                                                        words, env_steps);
    e = cudaGetLastError();
    return e == cudaSuccess ? KIN_OK : kin_fail_cuda(e, "kin_route_probe");
}
