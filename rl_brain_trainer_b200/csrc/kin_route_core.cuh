// kin_route_core.cuh -- device-side core of the dense holder-route wrappers (route observation, route-ready predicate, the
// route step with in-episode waypoint advance and the 13-term route reward), shared by kin_route.cu and kin_route_tc.cu.
// Reference: route/route_env.py:49-212, route/route_sequence_env.py:96-278, route/reward_route.py:36-143,
// route/route_observation.py:31-61 (paths under kinematic_phase1/).
#pragma once

#include "kin_internal.h"
#include "kin_state.cuh"

namespace kin {

constexpr int ROBS = KIN_ROUTE_OBS_DIM;
constexpr int ROBS_TILE_FLOATS = WARP * ROBS;   // 2560 floats = 10 240 B per warp
constexpr int RT_THREADS = 128;
constexpr int RT_WARPS = RT_THREADS / WARP;
constexpr int ROUTE_PRUNE_RINGS = 24;            // pruned nearest-waypoint scan: rings (2 candidates each) checked around the pivot
constexpr int ROUTE_DESCENT_MAX = 32;            // ... and moves of the descent along the route index that finds the pivot
constexpr int ROUTE_MAX_SMEM_WP = 1024;         // waypoint joint vectors staged in smem for the nearest-waypoint scan

struct RouteView {
    int n;
    const float* q;      // [n][7]
    const float* pose;   // [n][6]
    const float* tan;    // [n][7] next_q_delta
    const float* prog;   // [n]
    const float* lb;     // [n][lb_k] pruning bounds of the nearest-waypoint scan (KinRouteTable::nearest_lb), or nullptr
    int lb_k;
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
        // route_env.py:135: min over ALL waypoints of |q_w - q|.  Every candidate's squared distance is computed by the same
        // expression in either branch and min is order-independent, so the pruned scan returns the full scan's value bit for bit.
        auto d2 = [&](int w) {
            const float* qw = q_table + w * NJ;
            float acc = 0.0f;
#pragma unroll
            for (int i = 0; i < NJ; ++i) acc = fmaf(qw[i] - s.q[i], qw[i] - s.q[i], acc);
            return acc;
        };
        // Pruned form (needs KinRouteTable::nearest_lb): (1) from the current target walk along the route index while the distance
        // falls -- a replica tracking the route sits a few waypoints behind its target, so this finds a pivot j next to it in a
        // handful of evaluations; (2) ring check around the pivot: every waypoint w with |w - j| >= k satisfies
        // |q_w - q| >= |q_w - q_j| - |q_j - q| >= lb[j][k] - r, so once lb[j][k] - r >= sqrt(best) nothing further out can win.
        // Any pivot makes (2) a proof; (1) only makes it short.  Replicas far off the route (the rings run out) take the plain scan.
        bool done_scan = false;
        if (R.lb != nullptr && R.lb_k > 1) {
            int j = wp_clamp(R, target);
            float best = d2(j);
            {
                const float dl = j > 0 ? d2(j - 1) : CUDART_INF_F, dr = j + 1 < R.n ? d2(j + 1) : CUDART_INF_F;
                const int dir = dl < dr ? -1 : 1;
                float dn = fminf(dl, dr);
                for (int moves = 0; dn < best && moves < ROUTE_DESCENT_MAX; ++moves) {
                    best = dn;
                    j += dir;
                    const int nx = j + dir;
                    dn = (nx >= 0 && nx < R.n) ? d2(nx) : CUDART_INF_F;
                }
            }
            const float r = sqrtf(best) * 1.00001f;
            const float* lb = R.lb + (size_t)j * R.lb_k;
            int rings = min(R.lb_k - 1, ROUTE_PRUNE_RINGS);
            if (!(__ldg(lb + rings) * 0.99999f - r >= r)) rings = 0;      // the table cannot settle this replica: straight to the plain scan
            for (int k = 1; k <= rings; ++k) {
                const float slack = __ldg(lb + k) * 0.99999f - r;
                if (slack > 0.0f && slack * slack >= best) { done_scan = true; break; }     // also on +inf: no waypoint that far from j
                if (j - k >= 0) best = fminf(best, d2(j - k));
                if (j + k < R.n) best = fminf(best, d2(j + k));
            }
            nearest = sqrtf(best);
        }
        // Replicas the ring check could not settle get the plain scan over ALL waypoints -- done by the whole warp for one such
        // lane at a time (its q broadcast, every lane takes the waypoints w = lane, lane + 32, ..., shuffle-min at the end), so a
        // single far-off replica costs its warp ~16 strided iterations instead of a 483-iteration serial loop.
        // NOTE: warp-collective -- every lane of the warp calls route_step_core (the kernels keep inactive lanes in step).
        unsigned todo = __ballot_sync(0xffffffffu, !done_scan);
        const int lane = (int)(threadIdx.x & 31u);
        if (__popc(todo) > 16) {          // most of the warp is off the route: the classic loop, one broadcast waypoint per iteration
            if (!done_scan) {
                float m = CUDART_INF_F;
                for (int w = 0; w < R.n; ++w) m = fminf(m, d2(w));
                nearest = sqrtf(m);
            }
            todo = 0u;
        }
        while (todo) {
            const int src = __ffs((int)todo) - 1;
            todo &= todo - 1u;
            float qs[NJ];
#pragma unroll
            for (int i = 0; i < NJ; ++i) qs[i] = __shfl_sync(0xffffffffu, s.q[i], src);
            float m = CUDART_INF_F;
            for (int w = lane; w < R.n; w += 32) {
                const float* qw = q_table + w * NJ;
                float acc = 0.0f;
#pragma unroll
                for (int i = 0; i < NJ; ++i) acc = fmaf(qw[i] - qs[i], qw[i] - qs[i], acc);
                m = fminf(m, acc);
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, off));
            if (lane == src) nearest = sqrtf(m);
        }
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

// sample_route_reset + RouteKinematicEnv.reset for one slot (route/route_reset_samplers.py:43-117, route/route_env.py:60-97,
// route/route_sequence_env.py:120-124): draws from `rng`, fills the env registers and the route registers, returns goal_q.
__device__ __forceinline__ void sample_route_reset_dev(const KinEnvParams& P, const RouteView& R, const KinRouteResetParams& C, Philox& rng, EnvRegs& s,
                                                       RouteRegs& rr, float* gq_out) {
    int mode = 4;
    {
        const float u = rng.uniform();
#pragma unroll
        for (int m = 3; m >= 0; --m) if (u < C.mode_cdf[m]) mode = m;
    }
    if (C.forced_mode >= 0) mode = C.forced_mode;
    int ri = rng.integers(C.index_lo[mode], C.index_hi[mode]);
    const int start = mode == 0 ? 0 : max(ri - 1, 0);
    const int src = mode == 4 ? ri : start;
    int last = ri;
    if (C.sequence_length > 0) {
        ri = min(max(ri, 1), C.max_route_index);
        last = min(ri + C.sequence_length - 1, C.max_route_index);
    }
    float nz[24];
#pragma unroll
    for (int k = 0; k < 24; k += 2) nz[k] = gauss_pair(rng, &nz[k + 1]);
    float r_iq[NJ], r_idq[NJ], r_ipa[NJ], r_gq[NJ];
    const float* sq = R.q + (size_t)wp_clamp(R, src) * NJ;
    const float* gq = R.q + (size_t)wp_clamp(R, ri) * NJ;
#pragma unroll
    for (int k = 0; k < NJ; ++k) {
        r_iq[k] = clampf(fmaf(nz[k], C.q_noise_std, __ldg(sq + k)), P.joint_lower[k], P.joint_upper[k]);
        r_idq[k] = nz[7 + k] * C.dq_noise_std;
        r_ipa[k] = clampf(nz[14 + k] * C.prev_action_noise_std, -1.0f, 1.0f);
        r_gq[k] = __ldg(gq + k);
    }
    s.flags = 0u;
    reset_core(P, s, KIN_MODE_APPROACH, r_iq, r_idq, r_ipa, r_gq, nullptr, gq_out);
    rr.index = ri;
    rr.streak = 0;
    rr.last = last;
    rr.completed = 0;
}

static inline bool route_ok(const KinRouteTable* r) { return r && r->n_waypoints >= 2 && r->n_waypoints <= 65535 && r->q_goal && r->pose6 && r->next_q_delta && r->progress_m; }
static inline RouteView view_of(const KinRouteTable* r) {
    return RouteView{r->n_waypoints, r->q_goal, r->pose6, r->next_q_delta, r->progress_m, r->nearest_lb, r->nearest_lb ? r->nearest_lb_k : 0};
}

}  // namespace kin
