/*
 * kin_oracle_route.c -- CPU fp64 oracle for the dense holder-route wrappers
 * (TEST INFRASTRUCTURE, see kin_oracle.h).  Follows, relative to
 * hrl_ws/src/hrl_trainer/hrl_trainer/kinematic_phase1/:
 *   route/route_dataset.py:73-99, route/route_env.py:124-212,
 *   route/route_sequence_env.py:139-257, route/reward_route.py:36-143,
 *   route/route_observation.py:31-61, eval/eval_route_curriculum.py:55-125,188-218.
 */
#include "kin_oracle.h"

#include <math.h>
#include <string.h>

#define NJ 7

static double clipd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }
static double maxd(double a, double b) { return a > b ? a : b; }
static double normn(const double *v, int n) {
    double acc = 0.0;
    for (int i = 0; i < n; ++i) acc += v[i] * v[i];
    return sqrt(acc);
}
static double dist7(const double *a, const double *b) {
    double acc = 0.0;
    for (int i = 0; i < NJ; ++i) acc += (a[i] - b[i]) * (a[i] - b[i]);
    return sqrt(acc);
}
static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* route/route_dataset.py:73-99 */
void kor_route_build(const double *q_goal, int n, double *pose6, double *next_q_delta, double *progress_m) {
    for (int i = 0; i < n; ++i) kor_fk_pose6(q_goal + (size_t)i * NJ, pose6 + (size_t)i * 6);
    progress_m[0] = 0.0;
    for (int i = 1; i < n; ++i) {
        double d[3];
        for (int k = 0; k < 3; ++k) d[k] = pose6[(size_t)i * 6 + k] - pose6[(size_t)(i - 1) * 6 + k];
        progress_m[i] = progress_m[i - 1] + sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    }
    for (int i = 0; i < n; ++i) {
        int j = i + 1 < n ? i + 1 : n - 1;
        for (int k = 0; k < NJ; ++k) next_q_delta[(size_t)i * NJ + k] = q_goal[(size_t)j * NJ + k] - q_goal[(size_t)i * NJ + k];
    }
}

static const double *wp_q(const kor_route *r, int idx) { return r->q_goal + (size_t)clampi(idx, 0, r->n_waypoints - 1) * NJ; }
static const double *wp_pose(const kor_route *r, int idx) { return r->pose6 + (size_t)clampi(idx, 0, r->n_waypoints - 1) * 6; }
static const double *wp_tangent(const kor_route *r, int idx) { return r->next_q_delta + (size_t)clampi(idx, 0, r->n_waypoints - 1) * NJ; }
static double wp_progress(const kor_route *r, int idx) { return r->progress_m[clampi(idx, 0, r->n_waypoints - 1)]; }

/* explicit reset: route/route_env.py:49-97 + eval/eval_route_curriculum.py:67-87 */
void kor_route_reset(const kor_params *p, const kor_route *r, kor_route_state *s, int route_index,
                     int start_route_index, const double *initial_q, const double *initial_dq,
                     const double *initial_prev_action) {
    memset(s, 0, sizeof(*s));
    const double *iq = initial_q ? initial_q : wp_q(r, start_route_index);
    kor_reset(p, &s->base, KOR_MODE_APPROACH, iq, initial_dq, initial_prev_action, wp_q(r, route_index), NULL);
    s->route_index = route_index;
    s->start_route_index = start_route_index;
    s->ready_streak = 0;
    s->last_route_index = route_index;
    s->completed_waypoints = 0;
    memcpy(s->prev_q, s->base.q, sizeof(s->prev_q));
    memcpy(s->prev_dq, s->base.dq, sizeof(s->prev_dq));
}

/* route/route_observation.py:31-61 appended to the base 56 in alphabetical key order:
 * ... q 40:47, route_q_error 47:54, route_q_goal 54:61, route_scalar 61:64, route_tangent 64:71,
 * task_type 71:74, wp_ori_err 74:77, wp_pos_err 77:80 */
void kor_route_observation(const kor_params *p, const kor_route *r, const kor_route_state *s, float obs[80]) {
    float base[56];
    kor_observation(p, &s->base, base);
    for (int i = 0; i < 80; ++i) obs[i] = 0.0f;
    for (int i = 0; i < 47; ++i) obs[i] = base[i];
    const double *goal = wp_q(r, s->route_index);
    const double *tan = wp_tangent(r, s->route_index - 1 > 0 ? s->route_index - 1 : 0);
    for (int i = 0; i < NJ; ++i) {
        double span = maxd(p->joint_upper[i] - p->joint_lower[i], 1e-9);
        double dl = maxd(p->joint_delta_limit[i], 1e-9);
        obs[47 + i] = (float)clipd((goal[i] - s->base.q[i]) / dl, -1.0, 1.0);
        obs[54 + i] = (float)clipd(2.0 * ((goal[i] - p->joint_lower[i]) / span) - 1.0, -1.0, 1.0);
        obs[64 + i] = (float)clipd(tan[i] / dl, -1.0, 1.0);
    }
    int max_idx = r->n_waypoints - 1;
    obs[61] = (float)clipd((double)s->route_index / (double)(max_idx > 1 ? max_idx : 1), 0.0, 1.0);
    obs[62] = (float)clipd(wp_progress(r, s->route_index) / maxd(wp_progress(r, max_idx), 1e-9), 0.0, 1.0);
    obs[63] = 0.0f;
    obs[71] = base[47]; obs[72] = base[48]; obs[73] = base[49];
    for (int k = 0; k < 6; ++k) obs[74 + k] = base[50 + k];
}

/* route/reward_route.py:36-51 */
static int route_ready(const kor_params *p, double qe, double pos, double ori, double an, double dqn) {
    return qe <= p->rr_route_ready_q_threshold && pos <= p->rr_route_ready_pos_threshold_m &&
           ori <= p->rr_route_ready_ori_threshold_rad && an <= p->rr_route_ready_action_threshold &&
           dqn <= p->rr_route_ready_dq_threshold;
}

/* route/reward_route.py:54-143 */
static double route_reward(const kor_params *p, const double *prev_q, const double *curr_q, const double *goal_q,
                           const double *prev_pose6, const double *curr_pose6, const double *goal_pose6,
                           const double *tangent, const double *action, const double *prev_action,
                           const double *curr_dq, int ready_streak, double nearest, double c[17]) {
    double prev_q_err = dist7(goal_q, prev_q), curr_q_err = dist7(goal_q, curr_q);
    double pe[3], oe[3];
    kor_pose_error(prev_pose6, goal_pose6, pe, oe);
    double prev_pos = normn(pe, 3), prev_ori = normn(oe, 3);
    kor_pose_error(curr_pose6, goal_pose6, pe, oe);
    double curr_pos = normn(pe, 3), curr_ori = normn(oe, 3);
    double an = normn(action, NJ), dqn = normn(curr_dq, NJ), tn = normn(tangent, NJ);
    double dot = 0.0;
    for (int i = 0; i < NJ; ++i) dot += (curr_q[i] - prev_q[i]) * tangent[i];
    double tangent_progress = tn > 0.0 ? dot / maxd(tn, 1e-9) : 0.0;
    int ready = route_ready(p, curr_q_err, curr_pos, curr_ori, an, dqn);
    double low_motion = 0.0;
    if (curr_pos <= 2.0 * p->rr_route_ready_pos_threshold_m && curr_ori <= 2.0 * p->rr_route_ready_ori_threshold_rad) {
        double a_clean = maxd(1.0 - an / maxd(p->rr_route_ready_action_threshold, 1e-9), 0.0);
        double d_clean = maxd(1.0 - dqn / maxd(p->rr_route_ready_dq_threshold, 1e-9), 0.0);
        low_motion = p->rr_low_motion_near_waypoint_bonus * 0.5 * (a_clean + d_clean);
    }
    double msq = 0.0, dmsq = 0.0;
    for (int i = 0; i < NJ; ++i) { msq += action[i] * action[i]; dmsq += (action[i] - prev_action[i]) * (action[i] - prev_action[i]); }
    msq /= NJ; dmsq /= NJ;
    double smooth = -p->rr_action_magnitude_weight * msq;
    smooth += -p->rr_action_delta_weight * dmsq;
    c[0] = p->rr_q_goal_progress_weight * (prev_q_err - curr_q_err);
    c[1] = p->rr_ee_position_progress_weight * (prev_pos - curr_pos);
    c[2] = p->rr_ee_orientation_progress_weight * (prev_ori - curr_ori);
    c[3] = p->rr_route_tangent_progress_weight * maxd(tangent_progress, 0.0);
    c[4] = ready ? p->rr_same_step_route_ready_bonus : 0.0;
    c[5] = (ready && ready_streak >= 1) ? p->rr_route_ready_dwell_bonus : 0.0;
    c[6] = low_motion;
    c[7] = -p->rr_orientation_regression_penalty_weight * maxd(curr_ori - prev_ori, 0.0);
    c[8] = -p->rr_q_route_regression_penalty_weight * maxd(curr_q_err - prev_q_err, 0.0);
    c[9] = -p->rr_off_route_penalty_weight * maxd(nearest, 0.0);
    c[10] = smooth;
    c[11] = -p->rr_dq_penalty_weight * dqn;
    c[12] = (curr_q_err >= prev_q_err && curr_pos >= prev_pos && curr_ori >= prev_ori) ? -p->rr_no_progress_penalty : 0.0;
    c[13] = curr_q_err; c[14] = curr_pos; c[15] = curr_ori; c[16] = (double)ready;
    double reward = 0.0;
    for (int i = 0; i < 13; ++i) reward += c[i];
    return reward;
}

/* route/route_env.py:124-192 (sequence_mode == 0) and route/route_sequence_env.py:139-232 (!= 0) */
void kor_route_step(const kor_params *p, const kor_route *r, kor_route_state *s, const double action[7],
                    int sequence_mode, int reset_ready_streak_on_advance, kor_route_step_out *out, float obs[80]) {
    double prev_q[NJ], prev_dq[NJ], prev_action[NJ], prev_pose6[6], curr_pose6[6];
    memcpy(prev_q, s->prev_q, sizeof(prev_q));
    memcpy(prev_dq, s->prev_dq, sizeof(prev_dq));
    memcpy(prev_action, s->base.prev_action, sizeof(prev_action));
    kor_fk_pose6(prev_q, prev_pose6);
    int target_index = s->route_index;
    const double *goal_q = wp_q(r, target_index);
    const double *goal_pose6 = wp_pose(r, target_index);
    const double *tangent = wp_tangent(r, target_index - 1 > 0 ? target_index - 1 : 0);

    kor_step(p, &s->base, action, &out->base, NULL);
    const double *curr_q = s->base.q, *curr_dq = s->base.dq;
    kor_fk_pose6(curr_q, curr_pose6);
    double q_err = dist7(goal_q, curr_q), prev_q_err = dist7(goal_q, prev_q);
    double an = normn(action, NJ), dqn = normn(curr_dq, NJ);
    double nearest = INFINITY;
    for (int i = 0; i < r->n_waypoints; ++i) {
        double d = dist7(r->q_goal + (size_t)i * NJ, curr_q);
        if (d < nearest) nearest = d;
    }
    int ready = route_ready(p, q_err, out->base.position_error_norm, out->base.orientation_error_norm, an, dqn);
    s->ready_streak = ready ? s->ready_streak + 1 : 0;
    out->route_reward = route_reward(p, prev_q, curr_q, goal_q, prev_pose6, curr_pose6, goal_pose6, tangent, action,
                                     prev_action, curr_dq, s->ready_streak, nearest, out->route_components);
    int wp_success = ready && s->ready_streak >= p->term_success_dwell_steps;
    int terminated, success;
    out->waypoint_success = wp_success;
    out->route_ready = ready;
    if (!sequence_mode) {
        success = wp_success;
        terminated = out->base.terminated;
        if (out->base.terminated && out->base.reason == KOR_REASON_SUCCESS && !success) terminated = 0;
        if (success && p->term_terminate_on_success) terminated = 1;
    } else {
        success = 0; terminated = 0;
        if (wp_success) {
            s->completed_waypoints += 1;
            if (target_index >= s->last_route_index) {
                success = 1; terminated = 1;
            } else { /* _advance_target, route_sequence_env.py:253-257 */
                s->route_index = target_index + 1;
                memcpy(s->base.goal_q, wp_q(r, s->route_index), sizeof(double) * NJ);
                memcpy(s->base.goal_pose6, wp_pose(r, s->route_index), sizeof(double) * 6);
                double pe[3], oe[3];
                kor_pose_error(s->base.ee_pose6, s->base.goal_pose6, pe, oe);
                s->base.entry_position_error_norm = normn(pe, 3);
                s->base.entry_orientation_error_norm = normn(oe, 3);
                s->base.entry_action_l2 = normn(s->base.prev_action, NJ);
                s->base.entry_dq_norm = normn(s->base.dq, NJ);
                if (reset_ready_streak_on_advance) s->ready_streak = 0;
            }
        }
        if (out->base.terminated && !terminated && out->base.reason != KOR_REASON_SUCCESS) terminated = 1;
    }
    out->route_ready_streak = s->ready_streak;
    out->route_q_error_norm = q_err;
    out->nearest_route_q_distance = nearest;
    out->success = success;
    out->terminated = terminated;
    out->route_regression = q_err > prev_q_err;
    out->route_orientation_hit = out->base.orientation_error_norm <= p->rr_route_ready_ori_threshold_rad;
    out->route_index = s->route_index;
    memcpy(s->prev_q, s->base.q, sizeof(s->prev_q));
    memcpy(s->prev_dq, s->base.dq, sizeof(s->prev_dq));
    if (obs) kor_route_observation(p, r, s, obs);
}

/* eval/eval_route_curriculum.py:55-125 (_roll_one) chained as in :188-218.
 * success_flags[idx - start_index], final_errors[(idx-start)*3 + {pos, ori, q}]; returns longest prefix. */
int kor_route_sequential_probe(const kor_params *p, const kor_route *r, const kor_mlp *policy, int start_index,
                               int end_index, const double *start_q, int *success_flags, double *final_errors,
                               long long *env_steps_out) {
    double cq[NJ], cdq[NJ], cpa[NJ];
    const double *q0 = start_q ? start_q : wp_q(r, start_index - 1 > 0 ? start_index - 1 : 0);
    for (int i = 0; i < NJ; ++i) { cq[i] = q0[i]; cdq[i] = 0.0; cpa[i] = 0.0; }
    int final_end = end_index < r->n_waypoints - 1 ? end_index : r->n_waypoints - 1;
    long long steps = 0;
    int prefix = 0, broken = 0;
    for (int idx = start_index; idx <= final_end; ++idx) {
        kor_route_state s;
        kor_route_step_out so;
        float obs[80], af[7];
        double a[NJ];
        kor_route_reset(p, r, &s, idx, 0, cq, cdq, cpa);
        kor_route_observation(p, r, &s, obs);
        int terminated = 0, truncated = 0;
        memset(&so, 0, sizeof(so));
        while (!(terminated || truncated)) {
            kor_mlp_forward(policy, obs, af, NULL);
            for (int i = 0; i < NJ; ++i) a[i] = (double)af[i];
            kor_route_step(p, r, &s, a, 0, 1, &so, obs);
            terminated = so.terminated; truncated = so.base.truncated;
            steps += 1;
        }
        int k = idx - start_index;
        if (success_flags) success_flags[k] = so.success;
        if (final_errors) {
            final_errors[3 * k + 0] = so.base.position_error_norm;
            final_errors[3 * k + 1] = so.base.orientation_error_norm;
            final_errors[3 * k + 2] = so.route_q_error_norm;
        }
        if (so.success && !broken) prefix += 1; else broken = 1;
        for (int i = 0; i < NJ; ++i) { cq[i] = s.base.q[i]; cdq[i] = s.base.dq[i]; cpa[i] = s.base.prev_action[i]; }
    }
    if (env_steps_out) *env_steps_out = steps;
    return prefix;
}
