/*
 * kin_oracle.c -- CPU fp64 oracle (TEST INFRASTRUCTURE, see kin_oracle.h).
 *
 * Restates, in plain C and double precision, the arithmetic of the reference's pure-Python
 * kinematic env.  Reference paths are relative to
 *   hrl_ws/src/hrl_trainer/hrl_trainer/   (jerry102102102/RL_brain_trainer).
 * Operation order follows the reference so that fp64 results agree to ~1e-13.
 */
#include "kin_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define NJ 7
static const double KOR_PI = 3.14159265358979323846;

/* ------------------------------------------------------------------------------------
 * URDF-style chain constants -- v5_1/ee_fk.py:14-61 (joint types, origin xyz / rpy, local axes)
 * ---------------------------------------------------------------------------------- */
static const int JOINT_PRISMATIC[NJ] = {1, 0, 0, 0, 0, 0, 0};
static const double ORIGIN_XYZ[NJ][3] = {
    {0.00715921043213119, 0.0000809621375843506, -0.0635},
    {-0.021178, 0.0, 0.1868},
    {-0.0633967414837172, 0.000642782425827271, 0.0602000000000009},
    {-0.000134989688424625, 0.425, 0.0133123982251372},
    {-0.0000850456535865796, -0.39225, -0.0083864861805065},
    {0.0475482889721905, -0.000817137634885778, -0.0805958577476871},
    {0.0436977540622506, 0.000443046177049933, -0.0521517110277254},
};
static const double ORIGIN_RPY[NJ][3] = {
    {0.0, 0.0, 0.0},
    {0.0, 0.0, 0.0},
    {1.5707963267949, 0.0, 1.5707963267949},
    {3.14159265358979, 0.0, 0.0},
    {3.14159265358979, 0.0, -1.5707963267949},
    {3.14159265358979, 1.5707963267949, 0.0},
    {-1.5707963267949, 0.0, -1.5707963267949},
};
static const double AXES_LOCAL[NJ][3] = {
    {1.0, 0.0, 0.0},
    {0.0, 0.0, 1.0},
    {0.0101382310641698, 0.0, -0.999948606814815},
    {0.010138231064165, 0.0, 0.999948606814815},
    {0.0, -0.0101382310641647, -0.999948606814815},
    {0.0, 0.0, -1.0},
    {-0.0101384515502096, 0.0, 0.999948604579338},
};

static void mat3_mul(const double a[9], const double b[9], double out[9]) {
    double t[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            double acc = 0.0;
            for (int k = 0; k < 3; ++k) acc += a[3 * r + k] * b[3 * k + c];
            t[3 * r + c] = acc;
        }
    memcpy(out, t, sizeof(t));
}

/* v5_1/ee_fk.py:64-71  R = Rz(yaw) @ Ry(pitch) @ Rx(roll) */
static void rpy_to_rot(double roll, double pitch, double yaw, double R[9]) {
    double cr = cos(roll), sr = sin(roll);
    double cp = cos(pitch), sp = sin(pitch);
    double cy = cos(yaw), sy = sin(yaw);
    double rx[9] = {1, 0, 0, 0, cr, -sr, 0, sr, cr};
    double ry[9] = {cp, 0, sp, 0, 1, 0, -sp, 0, cp};
    double rz[9] = {cy, -sy, 0, sy, cy, 0, 0, 0, 1};
    double t[9];
    mat3_mul(rz, ry, t);
    mat3_mul(t, rx, R);
}

/* v5_1/ee_fk.py:74-88  Rodrigues about a re-normalised local axis */
static void rot_axis_local(const double axis[3], double angle, double R[9]) {
    double n = sqrt(axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2]) + 1e-12;
    double x = axis[0] / n, y = axis[1] / n, z = axis[2] / n;
    double c = cos(angle), s = sin(angle), C = 1.0 - c;
    R[0] = c + x * x * C;     R[1] = x * y * C - z * s; R[2] = x * z * C + y * s;
    R[3] = y * x * C + z * s; R[4] = c + y * y * C;     R[5] = y * z * C - x * s;
    R[6] = z * x * C - y * s; R[7] = z * y * C + x * s; R[8] = c + z * z * C;
}

/* T <- T @ [R p; 0 1]  (4x4 homogeneous product, v5_1/ee_fk.py:91-95,107-117) */
static void chain_apply(double R_w[9], double p_w[3], const double R[9], const double p[3]) {
    double np_[3];
    for (int r = 0; r < 3; ++r)
        np_[r] = R_w[3 * r] * p[0] + R_w[3 * r + 1] * p[1] + R_w[3 * r + 2] * p[2] + p_w[r];
    mat3_mul(R_w, R, R_w);
    p_w[0] = np_[0]; p_w[1] = np_[1]; p_w[2] = np_[2];
}

/* v5_1/ee_fk.py:98-117 */
void kor_fk_matrix(const double q[7], double T[16]) {
    double R_w[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    double p_w[3] = {0, 0, 0};
    const double zero3[3] = {0, 0, 0};
    const double eye[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int i = 0; i < NJ; ++i) {
        double Ro[9];
        rpy_to_rot(ORIGIN_RPY[i][0], ORIGIN_RPY[i][1], ORIGIN_RPY[i][2], Ro);
        chain_apply(R_w, p_w, Ro, ORIGIN_XYZ[i]);
        if (JOINT_PRISMATIC[i]) {
            double d[3] = {AXES_LOCAL[i][0] * q[i], AXES_LOCAL[i][1] * q[i], AXES_LOCAL[i][2] * q[i]};
            chain_apply(R_w, p_w, eye, d);
        } else {
            double Rj[9];
            rot_axis_local(AXES_LOCAL[i], q[i], Rj);
            chain_apply(R_w, p_w, Rj, zero3);
        }
    }
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) T[4 * r + c] = R_w[3 * r + c];
        T[4 * r + 3] = p_w[r];
    }
    T[12] = 0; T[13] = 0; T[14] = 0; T[15] = 1;
}

/* v5_1/ee_fk.py:120-134  ZYX Euler extraction */
void kor_fk_pose6(const double q[7], double pose6[6]) {
    double T[16];
    kor_fk_matrix(q, T);
    double roll = atan2(T[9], T[10]);
    double pitch = atan2(-T[8], sqrt(T[0] * T[0] + T[4] * T[4]));
    double yaw = atan2(T[4], T[0]);
    pose6[0] = T[3]; pose6[1] = T[7]; pose6[2] = T[11];
    pose6[3] = roll; pose6[4] = pitch; pose6[5] = yaw;
}

/* kinematics/pose_utils.py:11-12  (numpy floored modulo: result has the divisor's sign) */
double kor_wrap_to_pi(double v) {
    double two_pi = 2.0 * KOR_PI;
    double m = fmod(v + KOR_PI, two_pi);
    if (m != 0.0 && m < 0.0) m += two_pi;
    return m - KOR_PI;
}

/* kinematics/pose_utils.py:15-26 */
void kor_pose_error(const double curr6[6], const double goal6[6], double pos_err[3], double ori_err[3]) {
    for (int k = 0; k < 3; ++k) {
        pos_err[k] = goal6[k] - curr6[k];
        ori_err[k] = kor_wrap_to_pi(goal6[3 + k] - curr6[3 + k]);
    }
}

static double norm3(const double v[3]) { return sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
static double normn(const double *v, int n) {
    double acc = 0.0;
    for (int i = 0; i < n; ++i) acc += v[i] * v[i];
    return sqrt(acc);
}
static double clipd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }
static double maxd(double a, double b) { return a > b ? a : b; }
static double mind(double a, double b) { return a < b ? a : b; }

/* kinematics/joint_limits.py:37-47 (URDF absent from the snapshot -> defaults) */
void kor_default_joint_specs(double lower[7], double upper[7], double delta_limit[7]) {
    const double dl[NJ] = {0.08, 0.30, 0.24, 0.24, 0.30, 0.40, 0.30};
    lower[0] = -0.385; upper[0] = 0.385;
    for (int i = 1; i < NJ; ++i) { lower[i] = -KOR_PI; upper[i] = KOR_PI; }
    for (int i = 0; i < NJ; ++i) delta_limit[i] = dl[i];
}

/* kinematics/joint_limits.py:165-174 */
static double joint_limit_margin_i(const kor_params *p, double q, int i) {
    double span = maxd(p->joint_upper[i] - p->joint_lower[i], 1e-9);
    double left = (q - p->joint_lower[i]) / span;
    double right = (p->joint_upper[i] - q) / span;
    return clipd(2.0 * mind(left, right), 0.0, 1.0);
}
static double joint_limit_margin_min(const kor_params *p, const double q[7]) {
    double m = joint_limit_margin_i(p, q[0], 0);
    for (int i = 1; i < NJ; ++i) m = mind(m, joint_limit_margin_i(p, q[i], i));
    return m;
}

/* envs/arm_kinematic_env.py:432-444 -- thresholds ALWAYS from reward_config (ar_*) */
static int is_near_goal(const kor_params *p, double pos, double ori) {
    if (pos > p->ar_near_goal_pos_threshold_m) return 0;
    if (p->ar_use_orientation_gate && ori > p->ar_near_goal_ori_threshold_rad) return 0;
    return 1;
}
static int is_pre_near_goal(const kor_params *p, double pos, double ori) {
    if (pos > p->ar_pre_near_goal_pos_threshold_m) return 0;
    if (p->ar_use_orientation_gate && ori > p->ar_near_goal_ori_threshold_rad) return 0;
    return 1;
}

/* envs/arm_kinematic_env.py:489-507 */
static double interp_control(double pos, double near_thr, double far_thr, double near_v, double far_v,
                             double fallback) {
    if (near_thr <= 0.0 || far_thr <= near_thr) return fallback;
    if (pos <= near_thr) return near_v;
    if (pos >= far_thr) return far_v;
    double alpha = (pos - near_thr) / maxd(far_thr - near_thr, 1e-9);
    return near_v + alpha * (far_v - near_v);
}

/* envs/arm_kinematic_env.py:425-430 */
static void capture_entry_metrics(kor_state *s) {
    double pe[3], oe[3];
    kor_pose_error(s->ee_pose6, s->goal_pose6, pe, oe);
    s->entry_position_error_norm = norm3(pe);
    s->entry_orientation_error_norm = norm3(oe);
    s->entry_action_l2 = normn(s->prev_action, NJ);
    s->entry_dq_norm = normn(s->dq, NJ);
}

/* envs/arm_kinematic_env.py:102-211, explicit-options branch */
void kor_reset(const kor_params *p, kor_state *s, int mode, const double *initial_q,
               const double *initial_dq, const double *initial_prev_action, const double *goal_q,
               const double *goal_pose6) {
    memset(s, 0, sizeof(*s));
    s->min_pos_error = INFINITY;
    s->mode = mode;
    for (int i = 0; i < NJ; ++i) {
        s->q[i] = clipd(initial_q[i], p->joint_lower[i], p->joint_upper[i]);
        s->dq[i] = initial_dq ? initial_dq[i] : 0.0;
        s->prev_action[i] = initial_prev_action ? initial_prev_action[i] : 0.0;
    }
    kor_fk_pose6(s->q, s->ee_pose6);
    if (goal_pose6) { /* explicit goal_pose6 wins; goal_q stored unclipped (:190-192) */
        memcpy(s->goal_pose6, goal_pose6, sizeof(double) * 6);
        for (int i = 0; i < NJ; ++i) s->goal_q[i] = goal_q ? goal_q[i] : 0.0;
    } else {
        for (int i = 0; i < NJ; ++i) s->goal_q[i] = clipd(goal_q[i], p->joint_lower[i], p->joint_upper[i]);
        kor_fk_pose6(s->goal_q, s->goal_pose6);
    }
    capture_entry_metrics(s);
}

/* envs/observation_builder.py:29-94 flattened in SB3's alphabetical key order:
 * dq 0:7, goal_ori_err 7:10, goal_pos_err 10:13, joint_limit_margin 13:20, mode_flag 20:24,
 * next_wp_ori_err 24:27, next_wp_pos_err 27:30, prev_action 30:37, progress 37:40, q 40:47,
 * task_type 47:50, wp_ori_err 50:53, wp_pos_err 53:56 */
void kor_observation(const kor_params *p, const kor_state *s, float obs[56]) {
    double pe[3], oe[3];
    for (int i = 0; i < 56; ++i) obs[i] = 0.0f;
    kor_pose_error(s->ee_pose6, s->goal_pose6, pe, oe);
    for (int i = 0; i < NJ; ++i) {
        double span = maxd(p->joint_upper[i] - p->joint_lower[i], 1e-9);
        obs[0 + i] = (float)clipd(s->dq[i] / maxd(p->joint_delta_limit[i], 1e-9), -1.0, 1.0);
        obs[13 + i] = (float)joint_limit_margin_i(p, s->q[i], i);
        obs[30 + i] = (float)clipd(s->prev_action[i], -1.0, 1.0);
        obs[40 + i] = (float)clipd(2.0 * ((s->q[i] - p->joint_lower[i]) / span) - 1.0, -1.0, 1.0);
    }
    for (int k = 0; k < 3; ++k) {
        obs[7 + k] = (float)clipd(oe[k] / p->obs_ori_err_scale_rad, -1.0, 1.0);
        obs[10 + k] = (float)clipd(pe[k] / p->obs_pos_err_scale_m, -1.0, 1.0);
    }
    int mode_index = s->mode < 0 ? 0 : (s->mode > 3 ? 3 : s->mode);
    obs[20 + mode_index] = 1.0f;
    int ep_len = p->episode_length > 1 ? p->episode_length : 1;
    int dw_tgt = p->dwell_steps_target > 1 ? p->dwell_steps_target : 1;
    obs[37] = (float)clipd((double)s->episode_step / (double)ep_len, 0.0, 1.0);
    obs[38] = (float)clipd((double)s->dwell_count / (double)dw_tgt, 0.0, 1.0);
    obs[39] = 0.0f;
    obs[47] = 1.0f; /* task_type = [1,0,0] */
}

/* envs/termination.py:20-57 */
static void evaluate_termination(const kor_params *p, int step_count, double pos, double ori, int dwell,
                                 int *terminated, int *truncated, int *success, int *reason) {
    *terminated = 0; *truncated = 0; *success = 0; *reason = KOR_REASON_RUNNING;
    int met = pos <= p->term_success_pos_threshold_m &&
              (!p->term_require_orientation || ori <= p->term_success_ori_threshold_rad) &&
              dwell >= p->term_success_dwell_steps;
    if (!isfinite(pos) || !isfinite(ori)) {
        *terminated = 1; *reason = KOR_REASON_INVALID_STATE;
    } else if (met) {
        *success = 1;
        if (p->term_terminate_on_success) { *terminated = 1; *reason = KOR_REASON_SUCCESS; }
    }
    if (!*terminated && step_count >= p->term_max_episode_steps) {
        *truncated = 1; *reason = KOR_REASON_MAX_STEPS;
    }
}

typedef struct reward_in {
    double prev_pos, prev_ori, curr_pos, curr_ori;
    const double *action, *prev_action;
    int curr_in_pre_near, prev_in_near, curr_in_near;
    int dwell, entry_count, drift_count, success;
    double margin_min, dq_norm, prev_dq_norm, delta_q_change_l2;
    double entry_pos, entry_ori, entry_action, entry_dq;
} reward_in;

static double mean_sq(const double *a, int n) {
    double acc = 0.0;
    for (int i = 0; i < n; ++i) acc += a[i] * a[i];
    return acc / (double)n;
}
static double mean_sq_diff(const double *a, const double *b, int n) {
    double acc = 0.0;
    for (int i = 0; i < n; ++i) acc += (a[i] - b[i]) * (a[i] - b[i]);
    return acc / (double)n;
}

/* envs/reward_approach.py:75-373.  c[] is filled in the reference's dict order. */
static double approach_reward(const kor_params *p, const reward_in *in, double c[64]) {
    double prev_pos = in->prev_pos, curr_pos = in->curr_pos, prev_ori = in->prev_ori, curr_ori = in->curr_ori;
    int pre = in->curr_in_pre_near, cn = in->curr_in_near, pn = in->prev_in_near, dwell = in->dwell;

    double position_progress = p->ar_position_progress_weight * (prev_pos - curr_pos);
    double global_ori = p->ar_orientation_progress_weight * (prev_ori - curr_ori);
    double near_field_ori = pre ? p->ar_near_field_orientation_progress_weight * (prev_ori - curr_ori) : 0.0;
    double orientation_progress = global_ori + near_field_ori;
    double milestone = 0.0;
    if (pre)
        for (int i = 0; i < p->ar_n_milestones; ++i)
            if (curr_ori <= p->ar_orientation_milestone_thresholds_rad[i]) milestone += p->ar_orientation_milestone_bonuses[i];
    double nf_center = pre ? -p->ar_near_field_orientation_center_weight * curr_ori : 0.0;

    double pre_near_goal = (pre && !cn) ? p->ar_pre_near_goal_bonus : 0.0;
    int ec = in->entry_count - 1; if (ec < 0) ec = 0;
    double bonus_scale = pow(p->ar_near_goal_bonus_decay, (double)ec);
    double near_goal = (cn && !pn) ? p->ar_near_goal_bonus * bonus_scale : 0.0;
    double inner_progress = (pre && !cn) ? p->ar_pre_near_to_near_progress_weight * maxd(prev_pos - curr_pos, 0.0) : 0.0;
    double coarse_bonus = (pre && curr_ori <= p->ar_coarse_orientation_bonus_threshold_rad) ? p->ar_coarse_orientation_bonus : 0.0;

    int curr_ho = p->ar_handover_pos_threshold_m > 0.0 && curr_pos <= p->ar_handover_pos_threshold_m &&
                  (p->ar_handover_ori_threshold_rad <= 0.0 || curr_ori <= p->ar_handover_ori_threshold_rad);
    int prev_ho = p->ar_handover_pos_threshold_m > 0.0 && prev_pos <= p->ar_handover_pos_threshold_m &&
                  (p->ar_handover_ori_threshold_rad <= 0.0 || prev_ori <= p->ar_handover_ori_threshold_rad);
    double ho_bonus = (curr_ho && !prev_ho) ? p->ar_handover_bonus : 0.0;
    double ho_ret = (curr_ho && prev_ho) ? p->ar_handover_retention_bonus : 0.0;
    double ho_dwell = (curr_ho && dwell >= 2) ? p->ar_handover_dwell_bonus : 0.0;
    double ho_leave = (prev_ho && !curr_ho) ? -p->ar_handover_leave_penalty : 0.0;
    double regress = maxd(curr_pos - prev_pos, 0.0) + maxd(curr_ori - prev_ori, 0.0);
    double ho_regr = (prev_ho || curr_ho) ? -p->ar_handover_regression_weight * regress : 0.0;
    double dwell_b = (cn && dwell >= 2) ? p->ar_dwell_bonus : 0.0;
    int esc = in->drift_count - p->ar_drift_penalty_escalation_start; if (esc < 0) esc = 0;
    double drift_scale = 1.0 + p->ar_drift_penalty_escalation_per_count * (double)esc;
    double drift_w = p->ar_drift_penalty_weight * drift_scale;
    double drift_pen = pn ? -drift_w * maxd(curr_pos - prev_pos, 0.0) : 0.0;
    double leave_pen = (pn && !cn) ? -p->ar_near_goal_leave_penalty : 0.0;
    double action_norm = normn(in->action, NJ);
    double prev_action_norm = normn(in->prev_action, NJ);
    double dqn = in->dq_norm, pdqn = in->prev_dq_norm;

    int dc_en = p->ar_dock_coarse_ready_pos_threshold_m > 0.0 && p->ar_dock_coarse_ready_ori_threshold_rad > 0.0;
    int curr_dc_pose = dc_en && curr_pos <= p->ar_dock_coarse_ready_pos_threshold_m && curr_ori <= p->ar_dock_coarse_ready_ori_threshold_rad;
    int prev_dc_pose = dc_en && prev_pos <= p->ar_dock_coarse_ready_pos_threshold_m && prev_ori <= p->ar_dock_coarse_ready_ori_threshold_rad;
    int curr_dc_motion = (p->ar_dock_coarse_ready_action_threshold <= 0.0 || action_norm <= p->ar_dock_coarse_ready_action_threshold) &&
                         (p->ar_dock_coarse_ready_dq_threshold <= 0.0 || dqn <= p->ar_dock_coarse_ready_dq_threshold);
    int prev_dc_motion = (p->ar_dock_coarse_ready_action_threshold <= 0.0 || prev_action_norm <= p->ar_dock_coarse_ready_action_threshold) &&
                         (p->ar_dock_coarse_ready_dq_threshold <= 0.0 || pdqn <= p->ar_dock_coarse_ready_dq_threshold);
    int curr_dc = curr_dc_pose && curr_dc_motion, prev_dc = prev_dc_pose && prev_dc_motion;

    int fr_en = p->ar_finisher_ready_pos_threshold_m > 0.0 && p->ar_finisher_ready_ori_threshold_rad > 0.0;
    int curr_fr_pose = fr_en && curr_pos <= p->ar_finisher_ready_pos_threshold_m && curr_ori <= p->ar_finisher_ready_ori_threshold_rad;
    int prev_fr_pose = fr_en && prev_pos <= p->ar_finisher_ready_pos_threshold_m && prev_ori <= p->ar_finisher_ready_ori_threshold_rad;
    int curr_fr_motion = (p->ar_finisher_ready_action_threshold <= 0.0 || action_norm <= p->ar_finisher_ready_action_threshold) &&
                         (p->ar_finisher_ready_dq_threshold <= 0.0 || dqn <= p->ar_finisher_ready_dq_threshold);
    int prev_fr_motion = (p->ar_finisher_ready_action_threshold <= 0.0 || prev_action_norm <= p->ar_finisher_ready_action_threshold) &&
                         (p->ar_finisher_ready_dq_threshold <= 0.0 || pdqn <= p->ar_finisher_ready_dq_threshold);
    int curr_fr = curr_fr_pose && curr_fr_motion, prev_fr = prev_fr_pose && prev_fr_motion;

    int nh_en = p->ar_near_handoff_pos_threshold_m > 0.0 && p->ar_near_handoff_ori_threshold_rad > 0.0;
    int nh = nh_en && curr_pos <= p->ar_near_handoff_pos_threshold_m && curr_ori <= p->ar_near_handoff_ori_threshold_rad;
    int prev_nh = nh_en && prev_pos <= p->ar_near_handoff_pos_threshold_m && prev_ori <= p->ar_near_handoff_ori_threshold_rad;

    double dc_bonus = (curr_dc && !prev_dc) ? p->ar_dock_coarse_ready_bonus : 0.0;
    double dc_ret = (curr_dc && prev_dc) ? p->ar_dock_coarse_ready_retention_bonus : 0.0;
    double dc_dwell = (curr_dc && dwell >= 2) ? p->ar_dock_coarse_ready_dwell_bonus : 0.0;
    double dc_leave = (prev_dc && !curr_dc) ? -p->ar_dock_coarse_ready_leave_penalty : 0.0;
    double dc_regr = (nh || prev_nh || curr_dc_pose || prev_dc_pose) ? -p->ar_dock_coarse_ready_regression_weight * regress : 0.0;
    double fr_bonus = (curr_fr && !prev_fr) ? p->ar_finisher_ready_bonus : 0.0;
    double fr_ret = (curr_fr && prev_fr) ? p->ar_finisher_ready_retention_bonus : 0.0;
    double fr_dwell = (curr_fr && dwell >= 2) ? p->ar_finisher_ready_dwell_bonus : 0.0;
    double fr_leave = (prev_fr && !curr_fr) ? -p->ar_finisher_ready_leave_penalty : 0.0;
    double fr_regr = (nh || prev_nh || curr_fr_pose || prev_fr_pose) ? -p->ar_finisher_ready_regression_weight * regress : 0.0;

    int nh_any = nh || curr_dc_pose || curr_fr_pose;
    double act_msq = mean_sq(in->action, NJ);
    double nh_action_pen = nh_any ? -p->ar_near_handoff_action_weight * act_msq : 0.0;
    double nh_dq_pen = nh_any ? -p->ar_near_handoff_dq_weight * dqn : 0.0;
    double nh_motion = 0.0, nh_settle = 0.0;
    if (nh_any) {
        /* python `a or b` fallbacks, reward_approach.py:253-254 */
        double a_thr = p->ar_finisher_ready_action_threshold != 0.0 ? p->ar_finisher_ready_action_threshold : p->ar_dock_coarse_ready_action_threshold;
        double d_thr = p->ar_finisher_ready_dq_threshold != 0.0 ? p->ar_finisher_ready_dq_threshold : p->ar_dock_coarse_ready_dq_threshold;
        double a_scale = maxd(a_thr, 1e-9), d_scale = maxd(d_thr, 1e-9);
        double a_clean = a_thr > 0 ? maxd(1.0 - action_norm / a_scale, 0.0) : 0.0;
        double d_clean = d_thr > 0 ? maxd(1.0 - dqn / d_scale, 0.0) : 0.0;
        nh_motion = p->ar_near_handoff_motion_bonus_weight * (0.5 * a_clean + 0.5 * d_clean);
        nh_settle = p->ar_near_handoff_settle_bonus_weight *
                    (0.5 * maxd(prev_action_norm - action_norm, 0.0) + 0.5 * maxd(pdqn - dqn, 0.0));
    }
    double same_step = (curr_pos < prev_pos && curr_ori < prev_ori && (pre || nh)) ? p->ar_same_step_alignment_bonus : 0.0;
    double smooth_mult = (curr_ho || prev_ho) ? p->ar_handover_smoothness_multiplier : 1.0;
    double smooth = smooth_mult * (-p->ar_action_magnitude_weight * act_msq -
                                   p->ar_action_delta_weight * mean_sq_diff(in->action, in->prev_action, NJ));
    double jl_pen = -p->ar_joint_limit_penalty_weight * (maxd(0.25 - in->margin_min, 0.0) / 0.25);
    double succ = in->success ? p->ar_success_bonus : 0.0;

    int k = 0;
    c[k++] = position_progress;            /* 0 */
    c[k++] = global_ori;                   /* 1 (not summed) */
    c[k++] = near_field_ori;               /* 2 (not summed) */
    c[k++] = orientation_progress;         /* 3 */
    c[k++] = milestone;                    /* 4 */
    c[k++] = nf_center;                    /* 5 */
    c[k++] = pre_near_goal;                /* 6 */
    c[k++] = near_goal;                    /* 7 */
    c[k++] = inner_progress;               /* 8 */
    c[k++] = (cn && !pn) ? bonus_scale : 0.0; /* 9 (not summed) */
    c[k++] = coarse_bonus;                 /* 10 */
    c[k++] = ho_bonus; c[k++] = ho_ret; c[k++] = ho_dwell; c[k++] = ho_leave; c[k++] = ho_regr; /* 11-15 */
    c[k++] = dc_bonus; c[k++] = dc_ret; c[k++] = dc_dwell; c[k++] = dc_leave; c[k++] = dc_regr; /* 16-20 */
    c[k++] = fr_bonus; c[k++] = fr_ret; c[k++] = fr_dwell; c[k++] = fr_leave; c[k++] = fr_regr; /* 21-25 */
    c[k++] = nh_action_pen; c[k++] = nh_dq_pen; c[k++] = nh_motion; c[k++] = nh_settle;         /* 26-29 */
    c[k++] = same_step;                    /* 30 */
    c[k++] = dwell_b;                      /* 31 */
    c[k++] = drift_pen;                    /* 32 */
    c[k++] = leave_pen;                    /* 33 */
    c[k++] = drift_scale;                  /* 34 (not summed) */
    c[k++] = (double)in->entry_count;      /* 35 */
    c[k++] = (double)in->drift_count;      /* 36 */
    c[k++] = smooth;                       /* 37 */
    c[k++] = smooth_mult;                  /* 38 (not summed) */
    c[k++] = jl_pen;                       /* 39 */
    c[k++] = succ;                         /* 40 */
    c[k++] = curr_pos; c[k++] = curr_ori; c[k++] = action_norm; c[k++] = dqn; c[k++] = (double)dwell; /* 41-45 */
    c[k++] = (double)pre; c[k++] = (double)cn; c[k++] = (double)curr_ho; c[k++] = (double)curr_dc;     /* 46-49 */
    c[k++] = (double)curr_dc_pose; c[k++] = (double)curr_fr; c[k++] = (double)curr_fr_pose; c[k++] = (double)nh; /* 50-53 */

    /* sum in the reference's order (reward_approach.py:333-371) */
    static const int summed[34] = {0, 3, 4, 5, 6, 7, 8, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22,
                                   23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 37, 39, 40};
    double reward = 0.0;
    for (int i = 0; i < 34; ++i) reward += c[summed[i]];
    return reward;
}

/* envs/reward_dock.py:105-120 */
static double entry_penalty_scale(double pos, double near_thr, double far_thr, double near_m, double far_m) {
    if (near_thr <= 0.0 || far_thr <= near_thr) return 1.0;
    if (pos <= near_thr) return near_m;
    if (pos >= far_thr) return far_m;
    double alpha = (pos - near_thr) / maxd(far_thr - near_thr, 1e-9);
    return near_m + alpha * (far_m - near_m);
}

/* envs/reward_dock.py:123-484.  c[] in the reference's dict order. */
static double dock_reward(const kor_params *p, const reward_in *in, double c[64]) {
    double prev_pos = in->prev_pos, curr_pos = in->curr_pos, prev_ori = in->prev_ori, curr_ori = in->curr_ori;
    int cn = in->curr_in_near, pn = in->prev_in_near, dwell = in->dwell;
    double dqn = in->dq_norm;

    double position_progress = p->dr_position_progress_weight * (prev_pos - curr_pos);
    double orientation_progress = p->dr_orientation_progress_weight * (prev_ori - curr_ori);
    double stay = cn ? p->dr_stay_in_zone_bonus : 0.0;
    int dm1 = dwell - 1 > 0 ? dwell - 1 : 0;
    double dwell_bonus = cn ? p->dr_dwell_bonus * (double)dm1 : 0.0;
    double wr_bonus = cn ? p->dr_working_range_bonus : 0.0;
    int wrd = dwell - p->dr_working_range_dwell_start + 1; if (wrd < 0) wrd = 0;
    double wr_dwell = (cn && dwell >= p->dr_working_range_dwell_start) ? p->dr_working_range_dwell_bonus * (double)wrd : 0.0;
    int curr_tight = curr_pos <= p->dr_tight_pose_pos_threshold_m && curr_ori <= p->dr_tight_pose_ori_threshold_rad;
    int prev_tight = prev_pos <= p->dr_tight_pose_pos_threshold_m && prev_ori <= p->dr_tight_pose_ori_threshold_rad;
    double ns_pos = p->dr_near_strict_pos_threshold_m != 0.0 ? p->dr_near_strict_pos_threshold_m : p->dr_tight_pose_pos_threshold_m * 2.0;
    double ns_ori = p->dr_near_strict_ori_threshold_rad != 0.0 ? p->dr_near_strict_ori_threshold_rad : p->dr_tight_pose_ori_threshold_rad * 3.0;
    int curr_ns = curr_pos <= ns_pos && curr_ori <= ns_ori;
    int prev_ns = prev_pos <= ns_pos && prev_ori <= ns_ori;
    double s_pc = maxd(1.0 - curr_pos / maxd(p->dr_tight_pose_pos_threshold_m, 1e-9), 0.0);
    double s_oc = maxd(1.0 - curr_ori / maxd(p->dr_tight_pose_ori_threshold_rad, 1e-9), 0.0);
    double s_close = (0.8 * s_pc + 0.2 * s_oc); s_close = s_close * s_close;
    double tight_bonus = curr_tight ? p->dr_tight_pose_bonus : 0.0;
    double tight_dwell = curr_tight ? p->dr_tight_pose_dwell_bonus * (double)dm1 : 0.0;
    double strict_leave = (prev_tight && !curr_tight) ? -p->dr_strict_pose_leave_penalty : 0.0;
    double sc_reward = curr_tight ? p->dr_strict_center_reward_weight * s_close : 0.0;
    double r_p = curr_pos / maxd(p->dr_tight_pose_pos_threshold_m, 1e-9);
    double r_o = curr_ori / maxd(p->dr_tight_pose_ori_threshold_rad, 1e-9);
    double sc_pos_pen = p->dr_strict_center_position_weight > 0.0 ? -p->dr_strict_center_position_weight * (r_p * r_p) : 0.0;
    double sc_ori_pen = p->dr_strict_center_orientation_weight > 0.0 ? -p->dr_strict_center_orientation_weight * (r_o * r_o) : 0.0;
    double act_msq = mean_sq(in->action, NJ);
    double action_rms = sqrt(act_msq);
    double sc_small = 0.0;
    if (p->dr_strict_center_small_action_bonus_weight > 0.0 && p->dr_strict_center_small_action_pos_radius_m > 0.0 &&
        p->dr_strict_center_small_action_ori_radius_rad > 0.0 && p->dr_strict_center_small_action_scale > 0.0) {
        double cpc = maxd(1.0 - curr_pos / p->dr_strict_center_small_action_pos_radius_m, 0.0);
        double coc = maxd(1.0 - curr_ori / p->dr_strict_center_small_action_ori_radius_rad, 0.0);
        double cc = pow(0.8 * cpc + 0.2 * coc, p->dr_strict_center_small_action_power);
        double smallness = maxd(1.0 - action_rms / p->dr_strict_center_small_action_scale, 0.0);
        sc_small = curr_tight ? p->dr_strict_center_small_action_bonus_weight * cc * smallness : 0.0;
    }
    double sc_dwell = 0.0;
    if (curr_tight && p->dr_strict_center_dwell_bonus_weight > 0.0 && dwell >= p->dr_strict_center_dwell_start) {
        int es = dwell - p->dr_strict_center_dwell_escalation_start; if (es < 0) es = 0;
        double dscale = 1.0 + p->dr_strict_center_dwell_escalation_per_step * (double)es;
        sc_dwell = p->dr_strict_center_dwell_bonus_weight * s_close * dscale;
    }
    double tp_shape = p->dr_tight_position_shaping_radius_m > 0.0
                          ? p->dr_tight_position_shaping_weight * maxd(1.0 - curr_pos / maxd(p->dr_tight_position_shaping_radius_m, 1e-9), 0.0) : 0.0;
    double to_shape = p->dr_tight_orientation_shaping_radius_rad > 0.0
                          ? p->dr_tight_orientation_shaping_weight * maxd(1.0 - curr_ori / maxd(p->dr_tight_orientation_shaping_radius_rad, 1e-9), 0.0) : 0.0;
    double conv_pos = (p->dr_convergence_position_radius_m > 0.0 && mind(prev_pos, curr_pos) <= p->dr_convergence_position_radius_m)
                          ? p->dr_convergence_position_progress_weight * (prev_pos - curr_pos) : 0.0;
    double gate_scale = (p->dr_position_first_orientation_pos_threshold_m > 0.0 && curr_pos > p->dr_position_first_orientation_pos_threshold_m)
                            ? p->dr_position_first_orientation_pre_scale : 1.0;
    double conv_ori = (p->dr_convergence_orientation_radius_rad > 0.0 && mind(prev_ori, curr_ori) <= p->dr_convergence_orientation_radius_rad)
                          ? gate_scale * p->dr_convergence_orientation_progress_weight * (prev_ori - curr_ori) : 0.0;
    double leave_zone = (pn && !cn) ? -p->dr_leave_zone_penalty : 0.0;
    double wr_exit = (pn && !cn) ? -p->dr_working_range_exit_penalty : 0.0;
    double drift = -p->dr_drift_penalty_position_weight * maxd(curr_pos - prev_pos, 0.0);
    drift += -p->dr_drift_penalty_orientation_weight * maxd(curr_ori - prev_ori, 0.0);
    if (curr_tight || prev_tight) drift *= p->dr_strict_zone_drift_penalty_multiplier;

    double action_l2 = normn(in->action, NJ);
    double e_scale = entry_penalty_scale(maxd(prev_pos, curr_pos), p->dr_entry_action_penalty_near_pos_threshold_m,
                                         p->dr_entry_action_penalty_far_pos_threshold_m,
                                         p->dr_entry_action_penalty_near_multiplier, p->dr_entry_action_penalty_far_multiplier);
    double ad_msq = mean_sq_diff(in->action, in->prev_action, NJ);
    double smooth = -p->dr_action_magnitude_weight * act_msq;
    smooth += -p->dr_action_delta_weight * ad_msq;
    if (curr_tight) smooth *= p->dr_strict_zone_action_penalty_multiplier;
    smooth *= e_scale;
    double ad_rms = sqrt(ad_msq);
    double adv_pen = (p->dr_action_delta_violation_weight > 0.0 && p->dr_action_delta_violation_threshold > 0.0)
                         ? -p->dr_action_delta_violation_weight * e_scale * maxd(ad_rms - p->dr_action_delta_violation_threshold, 0.0) : 0.0;
    double dqc_pen = (p->dr_delta_q_change_penalty_weight > 0.0 && p->dr_delta_q_change_penalty_threshold > 0.0)
                         ? -p->dr_delta_q_change_penalty_weight * e_scale * maxd(in->delta_q_change_l2 - p->dr_delta_q_change_penalty_threshold, 0.0) : 0.0;
    double entry_pos = in->entry_pos, entry_ori = in->entry_ori, entry_action = in->entry_action, entry_dq = in->entry_dq;
    double preserve = 0.0;
    if (p->dr_preserve_state_bonus > 0.0 && (curr_ns || curr_tight)) {
        int pos_ok = curr_pos <= entry_pos + p->dr_preserve_position_tolerance_m;
        int ori_ok = curr_ori <= entry_ori + p->dr_preserve_orientation_tolerance_rad;
        if (pos_ok && ori_ok) preserve = p->dr_preserve_state_bonus;
    }
    double strict_hold = curr_tight ? p->dr_strict_hold_bonus * (double)dm1 : 0.0;
    double low_motion = 0.0;
    if (p->dr_low_motion_bonus > 0.0 && curr_ns &&
        (p->dr_low_motion_action_threshold <= 0.0 || action_l2 <= p->dr_low_motion_action_threshold) &&
        (p->dr_low_motion_dq_threshold <= 0.0 || dqn <= p->dr_low_motion_dq_threshold))
        low_motion = p->dr_low_motion_bonus;
    double tiny = 0.0;
    if (p->dr_tiny_correction_bonus > 0.0 && curr_ns && !curr_tight) {
        int improved = curr_pos <= prev_pos && curr_ori <= prev_ori;
        int small_a = p->dr_tiny_correction_action_threshold <= 0.0 || action_l2 <= p->dr_tiny_correction_action_threshold;
        if (improved && small_a) tiny = p->dr_tiny_correction_bonus;
    }
    double worse = 0.0;
    worse += -p->dr_worse_than_entry_position_weight * maxd(curr_pos - entry_pos - p->dr_worse_than_entry_position_tolerance_m, 0.0);
    worse += -p->dr_worse_than_entry_orientation_weight * maxd(curr_ori - entry_ori - p->dr_worse_than_entry_orientation_tolerance_rad, 0.0);
    double ns_regr = 0.0;
    if (curr_ns || prev_ns)
        ns_regr = -p->dr_near_strict_regression_multiplier *
                  (p->dr_drift_penalty_position_weight * maxd(curr_pos - prev_pos, 0.0) +
                   p->dr_drift_penalty_orientation_weight * maxd(curr_ori - prev_ori, 0.0));
    double aggr_scale = curr_ns ? p->dr_near_strict_action_penalty_multiplier : 1.0;
    double aggr = (p->dr_aggressive_action_weight > 0.0 && p->dr_aggressive_action_threshold > 0.0)
                      ? -p->dr_aggressive_action_weight * aggr_scale * maxd(action_l2 - p->dr_aggressive_action_threshold, 0.0) : 0.0;
    double dqp_scale = curr_ns ? p->dr_near_strict_dq_penalty_multiplier : 1.0;
    double dq_pen = (p->dr_dq_penalty_weight > 0.0 && p->dr_dq_penalty_threshold > 0.0)
                        ? -p->dr_dq_penalty_weight * dqp_scale * maxd(dqn - p->dr_dq_penalty_threshold, 0.0) : 0.0;
    double jl_pen = -p->dr_joint_limit_penalty_weight * (maxd(0.25 - in->margin_min, 0.0) / 0.25);
    double succ = in->success ? p->dr_success_bonus : 0.0;

    double b_outer = 0, b_inner = 0, b_dwell = 0, b_outer_exit = 0, b_inner_exit = 0, b_break = 0, b_drift = 0;
    int zone = 0;
    if (p->dr_basin_outer_radius_m > 0.0 && p->dr_basin_inner_radius_m > 0.0 && p->dr_basin_dwell_radius_m > 0.0) {
        double outer_r = maxd(p->dr_basin_outer_radius_m, 1e-9), inner_r = maxd(p->dr_basin_inner_radius_m, 1e-9),
               dwell_r = maxd(p->dr_basin_dwell_radius_m, 1e-9);
        int po = prev_pos <= outer_r, pi_ = prev_pos <= inner_r, pd = prev_pos <= dwell_r;
        int co = curr_pos <= outer_r, ci = curr_pos <= inner_r, cd = curr_pos <= dwell_r;
        zone = cd ? 3 : (ci ? 2 : (co ? 1 : 0));
        if (co) b_outer = p->dr_basin_outer_bonus * (1.0 + maxd(1.0 - curr_pos / outer_r, 0.0));
        if (ci) b_inner = p->dr_basin_inner_bonus * (1.0 + maxd(1.0 - curr_pos / inner_r, 0.0));
        if (cd) b_dwell = p->dr_basin_dwell_bonus * (1.0 + maxd(1.0 - curr_pos / dwell_r, 0.0));
        b_outer_exit = (po && !co) ? -p->dr_basin_outer_exit_penalty : 0.0;
        b_inner_exit = (pi_ && !ci) ? -p->dr_basin_inner_exit_penalty : 0.0;
        b_break = (pd && !cd) ? -p->dr_basin_dwell_break_penalty : 0.0;
        b_drift = (po || co) ? -p->dr_basin_drift_penalty_weight * maxd(curr_pos - prev_pos, 0.0) : 0.0;
    }

    int k = 0;
    c[k++] = position_progress; c[k++] = orientation_progress; c[k++] = stay; c[k++] = dwell_bonus;      /* 0-3 */
    c[k++] = wr_bonus; c[k++] = wr_dwell; c[k++] = tight_bonus; c[k++] = tight_dwell;                    /* 4-7 */
    c[k++] = strict_leave; c[k++] = sc_reward; c[k++] = sc_pos_pen; c[k++] = sc_ori_pen;                 /* 8-11 */
    c[k++] = sc_small; c[k++] = sc_dwell; c[k++] = tp_shape; c[k++] = to_shape;                          /* 12-15 */
    c[k++] = conv_pos; c[k++] = conv_ori;                                                               /* 16-17 */
    c[k++] = gate_scale; c[k++] = e_scale;                                                              /* 18-19 (not summed) */
    c[k++] = leave_zone; c[k++] = wr_exit; c[k++] = drift; c[k++] = smooth; c[k++] = adv_pen; c[k++] = dqc_pen; /* 20-25 */
    c[k++] = preserve; c[k++] = strict_hold; c[k++] = low_motion; c[k++] = tiny; c[k++] = worse;          /* 26-30 */
    c[k++] = ns_regr; c[k++] = aggr; c[k++] = dq_pen; c[k++] = jl_pen; c[k++] = succ;                     /* 31-35 */
    c[k++] = b_outer; c[k++] = b_inner; c[k++] = b_dwell; c[k++] = b_outer_exit; c[k++] = b_inner_exit;   /* 36-40 */
    c[k++] = b_break; c[k++] = b_drift;                                                                  /* 41-42 */
    c[k++] = (double)zone;                                                                               /* 43 */
    c[k++] = curr_pos; c[k++] = curr_ori; c[k++] = (double)dwell; c[k++] = (double)curr_tight; c[k++] = (double)curr_ns; /* 44-48 */
    c[k++] = entry_pos; c[k++] = entry_ori; c[k++] = entry_action; c[k++] = entry_dq;                     /* 49-52 */
    c[k++] = curr_pos - entry_pos; c[k++] = curr_ori - entry_ori; c[k++] = action_l2 - entry_action; c[k++] = dqn - entry_dq; /* 53-56 */
    c[k++] = (double)in->entry_count; c[k++] = (double)in->drift_count; c[k++] = (double)cn;              /* 57-59 */

    /* reference sum order (reward_dock.py:437-482): 0..17, 20..42 */
    double reward = 0.0;
    for (int i = 0; i <= 17; ++i) reward += c[i];
    for (int i = 20; i <= 42; ++i) reward += c[i];
    return reward;
}

/* envs/arm_kinematic_env.py:213-365 */
void kor_step(const kor_params *p, kor_state *s, const double action_in[7], kor_step_out *out, float obs[56]) {
    double action[NJ], prev_action[NJ], prev_pose6[6];
    for (int i = 0; i < NJ; ++i) { action[i] = clipd(action_in[i], -1.0, 1.0); prev_action[i] = s->prev_action[i]; }
    memcpy(prev_pose6, s->ee_pose6, sizeof(prev_pose6));
    double pe[3], oe[3];
    kor_pose_error(prev_pose6, s->goal_pose6, pe, oe);
    double prev_pos = norm3(pe), prev_ori = norm3(oe);
    int dock = (s->mode == KOR_MODE_DOCK);
    double dock_limit = clipd(p->dock_residual_action_limit, 0.0, 1.0);
    double dqc_scale = maxd(p->dock_delta_q_change_limit_scale, 0.0);
    if (dock) {
        dock_limit = clipd(interp_control(prev_pos, p->dock_dynamic_action_limit_near_pos_threshold_m,
                                          p->dock_dynamic_action_limit_far_pos_threshold_m,
                                          p->dock_dynamic_residual_action_limit_near,
                                          p->dock_dynamic_residual_action_limit_far, p->dock_residual_action_limit), 0.0, 1.0);
        dqc_scale = maxd(interp_control(prev_pos, p->dock_dynamic_action_limit_near_pos_threshold_m,
                                        p->dock_dynamic_action_limit_far_pos_threshold_m,
                                        p->dock_dynamic_delta_q_change_limit_scale_near,
                                        p->dock_dynamic_delta_q_change_limit_scale_far, p->dock_delta_q_change_limit_scale), 0.0);
        for (int i = 0; i < NJ; ++i) action[i] = clipd(action[i], -dock_limit, dock_limit);
    }
    int prev_in_near = is_near_goal(p, prev_pos, prev_ori);
    double scale = p->action_delta_scale;
    if (dock && p->dock_action_delta_scale > 0.0) {
        scale = p->dock_action_delta_scale;
    } else if (!dock) {
        if (p->dynamic_action_delta_scale_enabled) {
            double mult = interp_control(prev_pos, p->dynamic_action_delta_scale_near_pos_threshold_m,
                                         p->dynamic_action_delta_scale_far_pos_threshold_m,
                                         p->dynamic_action_delta_scale_near_multiplier,
                                         p->dynamic_action_delta_scale_far_multiplier, 1.0);
            scale = p->action_delta_scale * maxd(mult, 0.0);
        }
    }
    double q_next[NJ], dq_next[NJ], dchg[NJ];
    for (int i = 0; i < NJ; ++i) {
        double max_dq = p->joint_delta_limit[i] * scale;
        double cmd = action[i] * max_dq;
        if (dock && dqc_scale > 0.0) {
            double lim = max_dq * dqc_scale;
            cmd = s->dq[i] + clipd(cmd - s->dq[i], -lim, lim);
            cmd = clipd(cmd, -max_dq, max_dq);
        }
        q_next[i] = clipd(s->q[i] + cmd, p->joint_lower[i], p->joint_upper[i]);
        dq_next[i] = q_next[i] - s->q[i];
        dchg[i] = dq_next[i] - s->dq[i];
    }
    double delta_q_change_l2 = normn(dchg, NJ);
    double ee_next[6];
    kor_fk_pose6(q_next, ee_next);
    kor_pose_error(ee_next, s->goal_pose6, pe, oe);
    double curr_pos = norm3(pe), curr_ori = norm3(oe);
    int curr_pre = is_pre_near_goal(p, curr_pos, curr_ori);
    int curr_near = is_near_goal(p, curr_pos, curr_ori);
    s->min_pos_error = mind(s->min_pos_error, curr_pos);
    if (curr_pre) s->pre_near_goal_hit = 1;
    if (curr_near && !prev_in_near) s->near_goal_entry_count += 1;
    if (curr_near) s->dwell_count += 1; else s->dwell_count = 0;
    if (prev_in_near && curr_pos > prev_pos) s->near_goal_drift_count += 1;

    int terminated, truncated, success, reason;
    evaluate_termination(p, s->episode_step + 1, curr_pos, curr_ori, s->dwell_count, &terminated, &truncated, &success, &reason);

    reward_in in;
    in.prev_pos = prev_pos; in.prev_ori = prev_ori; in.curr_pos = curr_pos; in.curr_ori = curr_ori;
    in.action = action; in.prev_action = prev_action;
    in.curr_in_pre_near = curr_pre; in.prev_in_near = prev_in_near; in.curr_in_near = curr_near;
    in.dwell = s->dwell_count; in.entry_count = s->near_goal_entry_count; in.drift_count = s->near_goal_drift_count;
    in.success = success; in.margin_min = joint_limit_margin_min(p, q_next);
    in.dq_norm = normn(dq_next, NJ); in.prev_dq_norm = normn(s->dq, NJ); in.delta_q_change_l2 = delta_q_change_l2;
    in.entry_pos = s->entry_position_error_norm; in.entry_ori = s->entry_orientation_error_norm;
    in.entry_action = s->entry_action_l2; in.entry_dq = s->entry_dq_norm;

    memset(out->components, 0, sizeof(out->components));
    if (dock) { out->reward = dock_reward(p, &in, out->components); out->n_components = KOR_N_DOCK_COMPONENTS; }
    else      { out->reward = approach_reward(p, &in, out->components); out->n_components = KOR_N_APPROACH_COMPONENTS; }

    s->episode_step += 1;
    for (int i = 0; i < NJ; ++i) { s->q[i] = q_next[i]; s->dq[i] = dq_next[i]; s->prev_action[i] = action[i]; }
    memcpy(s->ee_pose6, ee_next, sizeof(ee_next));
    if (curr_near) s->near_goal_hit = 1;

    out->position_error_norm = curr_pos;
    out->orientation_error_norm = curr_ori;
    out->action_l2 = normn(action, NJ);
    out->executed_delta_q_l2 = in.dq_norm;
    out->delta_q_change_l2 = delta_q_change_l2;
    out->dock_action_limit = dock_limit;
    out->dock_delta_q_change_limit_scale = dqc_scale;
    out->joint_limit_margin_min = in.margin_min;
    out->terminated = terminated; out->truncated = truncated; out->success = success; out->reason = reason;
    out->curr_in_pre_near_goal = curr_pre; out->curr_in_near_goal = curr_near;
    out->pad0 = 0;
    if (obs) kor_observation(p, s, obs);
}

void kor_step_batch(const kor_params *p, kor_state *s, const double *actions, int n, kor_step_out *outs, float *obs) {
    for (int e = 0; e < n; ++e) kor_step(p, &s[e], actions + (size_t)e * NJ, &outs[e], obs ? obs + (size_t)e * 56 : NULL);
}

/* ------------------------------------------------------------------------------------
 * Policy: SB3 MultiInputPolicy (flatten-concat -> tanh MLP) restated, fp32 like torch
 * (call site eval/eval_three_stage.py:25-27).  predict(deterministic=True)
 * returns the mean action clipped to the Box bounds [-1, 1].
 * ---------------------------------------------------------------------------------- */
static void dense_tanh(const float *w, const float *b, const float *x, int in_dim, int out_dim, float *y, int act) {
    for (int o = 0; o < out_dim; ++o) {
        float acc = b[o];
        const float *wr = w + (size_t)o * in_dim;
        for (int i = 0; i < in_dim; ++i) acc += wr[i] * x[i];
        y[o] = act ? tanhf(acc) : acc;
    }
}

void kor_mlp_forward(const kor_mlp *m, const float *obs, float action[7], float *value) {
    float h0[64], h1[64];
    dense_tanh(m->pi_w0, m->pi_b0, obs, m->in_dim, 64, h0, 1);
    dense_tanh(m->pi_w1, m->pi_b1, h0, 64, 64, h1, 1);
    dense_tanh(m->act_w, m->act_b, h1, 64, 7, action, 0);
    for (int i = 0; i < 7; ++i) action[i] = action[i] < -1.0f ? -1.0f : (action[i] > 1.0f ? 1.0f : action[i]);
    if (value && m->has_value) {
        dense_tanh(m->vf_w0, m->vf_b0, obs, m->in_dim, 64, h0, 1);
        dense_tanh(m->vf_w1, m->vf_b1, h0, 64, 64, h1, 1);
        dense_tanh(m->val_w, m->val_b, h1, 64, 1, value, 0);
    }
}

/* ------------------------------------------------------------------------------------
 * Approach -> Finisher evaluation of ONE episode.
 *   eval/eval_pipeline_ablation.py:60-147  (_run_approach_with_handoff)
 *   eval/eval_three_stage.py:41-56,59-125  (_dock_coarse_ready, _run_policy)
 *   eval/eval_approach_finisher.py:24-32   (_finisher_ready)
 *   eval/eval_workspace_expansion.py:126-147 (final-settled overrides first-confirmed)
 * ---------------------------------------------------------------------------------- */
static int dock_coarse_ready(const kor_params *p, double pos, double ori, double an, double dqn) {
    return p->ar_dock_coarse_ready_pos_threshold_m > 0.0 && p->ar_dock_coarse_ready_ori_threshold_rad > 0.0 &&
           pos <= p->ar_dock_coarse_ready_pos_threshold_m && ori <= p->ar_dock_coarse_ready_ori_threshold_rad &&
           (p->ar_dock_coarse_ready_action_threshold <= 0.0 || an <= p->ar_dock_coarse_ready_action_threshold) &&
           (p->ar_dock_coarse_ready_dq_threshold <= 0.0 || dqn <= p->ar_dock_coarse_ready_dq_threshold);
}
static int finisher_ready(const kor_params *p, double pos, double ori, double an, double dqn) {
    return p->ar_finisher_ready_pos_threshold_m > 0.0 && p->ar_finisher_ready_ori_threshold_rad > 0.0 &&
           pos <= p->ar_finisher_ready_pos_threshold_m && ori <= p->ar_finisher_ready_ori_threshold_rad &&
           (p->ar_finisher_ready_action_threshold <= 0.0 || an <= p->ar_finisher_ready_action_threshold) &&
           (p->ar_finisher_ready_dq_threshold <= 0.0 || dqn <= p->ar_finisher_ready_dq_threshold);
}

static long long eval_one(const kor_params *pa, const kor_params *pf, const kor_mlp *approach, const kor_mlp *finisher,
                          const double *iq, const double *idq, const double *ipa, const double *gq, const double *gp6,
                          int confirm, kor_episode_result *r) {
    kor_state s, snap;
    kor_step_out so;
    float obs[56], af[7];
    double a[7];
    long long steps_total = 0;
    memset(r, 0, sizeof(*r));
    kor_reset(pa, &s, KOR_MODE_APPROACH, iq, idq, ipa, gq, gp6);
    kor_observation(pa, &s, obs);
    double min_pos = s.entry_position_error_norm, min_ori = s.entry_orientation_error_norm;
    int terminated = 0, truncated = 0, steps = 0, streak = 0, max_streak = 0, first_ready = -1, ready_hit = 0;
    int have_snap = 0, snap_step = -1;
    double snap_pos = 0, snap_ori = 0, last_an = 0, last_dqn = 0;
    memset(&so, 0, sizeof(so));
    while (!(terminated || truncated)) {
        kor_mlp_forward(approach, obs, af, NULL);
        for (int i = 0; i < 7; ++i) a[i] = (double)af[i];
        double an = normn(a, 7);
        kor_step(pa, &s, a, &so, obs);
        steps += 1;
        terminated = so.terminated; truncated = so.truncated;
        double dqn = so.executed_delta_q_l2, pos = so.position_error_norm, ori = so.orientation_error_norm;
        min_pos = mind(min_pos, pos); min_ori = mind(min_ori, ori);
        if (dock_coarse_ready(pa, pos, ori, an, dqn)) {
            ready_hit = 1;
            if (first_ready < 0) first_ready = steps;
            streak += 1;
        } else {
            streak = 0;
        }
        if (streak > max_streak) max_streak = streak;
        if (!have_snap && streak >= confirm) { have_snap = 1; snap = s; snap_step = steps; snap_pos = pos; snap_ori = ori; }
        last_an = an; last_dqn = dqn;
    }
    steps_total += steps;
    r->approach_steps = steps;
    r->approach_success = so.success;
    r->approach_final_position_error = so.position_error_norm;
    r->approach_final_orientation_error = so.orientation_error_norm;
    r->min_position_error = min_pos;
    r->min_orientation_error = min_ori;
    r->first_ready_step = first_ready;
    r->max_ready_streak = max_streak;
    int final_ready = finisher_ready(pa, so.position_error_norm, so.orientation_error_norm, last_an, last_dqn);
    r->final_ready = final_ready;
    r->ready_hit = ready_hit || final_ready;
    r->ready_dwell = (max_streak >= confirm) || final_ready;
    r->success = so.success;
    r->final_position_error = so.position_error_norm;
    r->final_orientation_error = so.orientation_error_norm;
    r->final_action_magnitude = last_an;
    r->final_dq_norm = last_dqn;
    r->handoff_kind = 0;
    r->handoff_step = -1;
    const kor_state *h = NULL;
    if (final_ready) { h = &s; r->handoff_kind = 2; r->handoff_step = steps; r->handoff_position_error = so.position_error_norm; r->handoff_orientation_error = so.orientation_error_norm; }
    else if (have_snap) { h = &snap; r->handoff_kind = 1; r->handoff_step = snap_step; r->handoff_position_error = snap_pos; r->handoff_orientation_error = snap_ori; }
    for (int i = 0; i < 7; ++i) r->final_q[i] = s.q[i];
    if (h && finisher && pf) {
        kor_state f;
        /* _state_reset_options: final_q/final_dq/final_prev_action/goal_q/goal_pose6 (eval_three_stage.py:30-38) */
        kor_reset(pf, &f, KOR_MODE_DOCK, h->q, h->dq, h->prev_action, h->goal_q, h->goal_pose6);
        kor_observation(pf, &f, obs);
        terminated = truncated = 0; steps = 0;
        double an = 0.0;
        while (!(terminated || truncated)) {
            kor_mlp_forward(finisher, obs, af, NULL);
            for (int i = 0; i < 7; ++i) a[i] = (double)af[i];
            an = normn(a, 7);
            kor_step(pf, &f, a, &so, obs);
            steps += 1;
            terminated = so.terminated; truncated = so.truncated;
        }
        steps_total += steps;
        r->finisher_steps = steps;
        r->success = so.success;
        r->final_position_error = so.position_error_norm;
        r->final_orientation_error = so.orientation_error_norm;
        r->final_action_magnitude = an;
        r->final_dq_norm = so.executed_delta_q_l2;
        for (int i = 0; i < 7; ++i) r->final_q[i] = f.q[i];
    }
    return steps_total;
}

typedef struct eval_job {
    const kor_params *pa, *pf;
    const kor_mlp *approach, *finisher;
    const double *initial_q, *initial_dq, *initial_prev_action, *goal_q, *goal_pose6;
    int n, confirm;
    kor_episode_result *results;
    int *next;            /* shared work counter (chunks of 16 episodes) */
    pthread_mutex_t *mu;
    long long steps;
} eval_job;

static void *eval_worker(void *arg) {
    eval_job *j = (eval_job *)arg;
    for (;;) {
        pthread_mutex_lock(j->mu);
        int lo = *j->next;
        *j->next = lo + 16;
        pthread_mutex_unlock(j->mu);
        if (lo >= j->n) break;
        int hi = lo + 16 < j->n ? lo + 16 : j->n;
        for (int e = lo; e < hi; ++e)
            j->steps += eval_one(j->pa, j->pf, j->approach, j->finisher, j->initial_q + (size_t)e * 7,
                                 j->initial_dq ? j->initial_dq + (size_t)e * 7 : NULL,
                                 j->initial_prev_action ? j->initial_prev_action + (size_t)e * 7 : NULL,
                                 j->goal_q + (size_t)e * 7, j->goal_pose6 ? j->goal_pose6 + (size_t)e * 6 : NULL,
                                 j->confirm, &j->results[e]);
    }
    return NULL;
}

/* Runs n independent episodes on n_threads host threads (pthreads; 0/1 -> calling thread). */
void kor_eval_approach_finisher(const kor_params *pa, const kor_params *pf, const kor_mlp *approach,
                                const kor_mlp *finisher, const double *initial_q, const double *initial_dq,
                                const double *initial_prev_action, const double *goal_q, const double *goal_pose6,
                                int n, int handoff_confirm_steps, int n_threads, kor_episode_result *results,
                                long long *env_steps_out) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    int next = 0;
    pthread_mutex_t mu;
    pthread_mutex_init(&mu, NULL);
    eval_job jobs[256];
    pthread_t tids[256];
    for (int t = 0; t < n_threads; ++t) {
        eval_job j = {pa, pf, approach, finisher, initial_q, initial_dq, initial_prev_action, goal_q, goal_pose6,
                      n, handoff_confirm_steps, results, &next, &mu, 0};
        jobs[t] = j;
    }
    for (int t = 1; t < n_threads; ++t) pthread_create(&tids[t], NULL, eval_worker, &jobs[t]);
    eval_worker(&jobs[0]);
    long long total = jobs[0].steps;
    for (int t = 1; t < n_threads; ++t) { pthread_join(tids[t], NULL); total += jobs[t].steps; }
    pthread_mutex_destroy(&mu);
    if (env_steps_out) *env_steps_out = total;
}
