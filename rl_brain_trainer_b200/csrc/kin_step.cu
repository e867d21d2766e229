// kin_step.cu -- K1: the fused env-step kernel, plus reset / observe / FK kernels.
//
// One thread per env, struct-of-arrays state in HBM, fully coalesced 4-byte row accesses
// (148 B read + 124 B written per env), the row-major [n,7] action and [n,56] observation tensors
// staged per warp through shared memory, the observation tile written by the TMA engine
// (cp.async.bulk).  Algorithmic traffic per env-step: 532 B approach / 548 B dock.
//
// Replaces ArmKinematicEnv.step / reset / current_observation
// (kinematic_phase1/envs/arm_kinematic_env.py:102-381) for a batch of envs.
#include <cstdio>
#include <cstring>

#include "kin_internal.h"
#include "kin_state.cuh"

namespace kin {

constexpr int STEP_THREADS = 128;
constexpr int STEP_WARPS = STEP_THREADS / WARP;
// resident CTAs per SM the plain step is compiled for (registers <= 65536 / (128 * n)): the kernel is HBM-bound, more warps in
// flight hide more of the load latency; the info-heavy variants (components / aux rows) keep the compiler's own choice
#ifndef KIN_STEP_MIN_BLOCKS
#define KIN_STEP_MIN_BLOCKS 5
#endif

template <int MODE, bool COMP, bool AUX, bool AUTORESET, bool BULK>
__global__ void __launch_bounds__(STEP_THREADS, (COMP || AUX) ? 1 : KIN_STEP_MIN_BLOCKS)
kin_step_kernel(const __grid_constant__ KinEnvParams P, const KinSamplerParams* __restrict__ S, float* __restrict__ state,
                int stride, int n, const float* __restrict__ action, float* __restrict__ obs, float* __restrict__ reward,
                uint8_t* __restrict__ done, float* __restrict__ aux, float* __restrict__ comps, uint64_t seed,
                float* __restrict__ terminal_obs) {
    __shared__ __align__(128) float tiles[STEP_WARPS][OBS_TILE_FLOATS];
    const int lane = threadIdx.x & (WARP - 1);
    const int warp = threadIdx.x >> 5;
    const int env0 = (blockIdx.x * STEP_WARPS + warp) * WARP;  // first env of this warp
    if (env0 >= n) return;                                      // warp-uniform
    const int env = env0 + lane;
    const bool active = env < n;
    const int envc = active ? env : n - 1;  // inactive lanes shadow the last env (loads stay in bounds, stores are masked)
    float* tile = tiles[warp];

    float a[NJ];
    load_action_tile(action, env0, n, tile, lane, a);

    EnvRegs s;
    load_env<MODE != KIN_MODE_APPROACH>(state, stride, envc, s);

    StepOut out;
    float c[COMP ? KIN_MAX_COMPONENTS : 1];
    if (COMP) {
#pragma unroll
        for (int k = 0; k < KIN_MAX_COMPONENTS; ++k) c[k] = 0.0f;
    }
    step_core<MODE, COMP>(P, s, a, out, c);
    const int mode = (MODE == KIN_MODE_PER_ENV) ? (int)((s.flags >> KIN_FLAG_MODE_SHIFT) & 3u) : MODE;

    if (active) {
        reward[env] = out.reward;
        if (AUX) {
            aux[(size_t)KIN_AUX_POS_ERR * stride + env] = out.pos;
            aux[(size_t)KIN_AUX_ORI_ERR * stride + env] = out.ori;
            aux[(size_t)KIN_AUX_ACTION_L2 * stride + env] = out.action_l2;
            aux[(size_t)KIN_AUX_DQ_L2 * stride + env] = out.dq_l2;
            aux[(size_t)KIN_AUX_DQ_CHANGE_L2 * stride + env] = out.dq_change_l2;
            aux[(size_t)KIN_AUX_DOCK_LIMIT * stride + env] = out.dock_limit;
            aux[(size_t)KIN_AUX_DQC_SCALE * stride + env] = out.dqc_scale;
            aux[(size_t)KIN_AUX_MARGIN_MIN * stride + env] = out.margin_min;
        }
        if (COMP) {
#pragma unroll
            for (int k = 0; k < KIN_MAX_COMPONENTS; ++k) comps[(size_t)k * stride + env] = c[k];
        }
    }

    const int tile_bytes = min(WARP, n - env0) * OBS * 4;
    float o[OBS];
    unsigned done_bits = out.done;
    if (AUTORESET) {
        const bool finished = active && (out.done & (KIN_DONE_TERMINATED | KIN_DONE_TRUNCATED));
        const unsigned any = __ballot_sync(0xffffffffu, finished);
        if (any) {
            if (terminal_obs) {  // VecEnv "terminal_observation": the last observation of the finished episode
                build_obs_from(P, s, mode, out.pe, out.oe, out.margin, o);
                stage_obs_row(tile, lane, o);
                if (BULK) {
                    bulk_store_tile(terminal_obs + (size_t)env0 * OBS, tile, tile_bytes, lane);
                    bulk_store_wait_read(lane);
                } else {
                    lsu_store_tile(terminal_obs + (size_t)env0 * OBS, tile, tile_bytes, lane);
                }
            }
            if (finished) {
                const unsigned episode = ld_row_u(state, stride, KIN_ROW_EPISODE, env) + 1u;
                Philox rng(seed, (unsigned)env, episode);
                ResetDraw d;
                sample_reset(P, *S, rng, mode, d);
                float gq[NJ];
                reset_core(P, s, mode, d.iq, d.idq, d.ipa, d.gq, d.has_gpose ? d.gpose : nullptr, gq);
                s.flags = (s.flags & ~(0xfu << KIN_FLAG_STAGE_SHIFT)) | ((unsigned)d.stage << KIN_FLAG_STAGE_SHIFT);
                store_env_reset(state, stride, env, s, gq);
                st_row_u(state, stride, KIN_ROW_EPISODE, env, episode);
                done_bits |= KIN_DONE_AUTORESET;
            }
        }
        if (active && !(done_bits & KIN_DONE_AUTORESET)) store_env_step(state, stride, env, s);
    } else {
        if (active) store_env_step(state, stride, env, s);
    }
    if (active) done[env] = (uint8_t)done_bits;

    if (AUTORESET && (done_bits & KIN_DONE_AUTORESET)) build_obs(P, s, mode, o);   // fresh episode: recompute errors / margins
    else build_obs_from(P, s, mode, out.pe, out.oe, out.margin, o);
    stage_obs_row(tile, lane, o);
    if (BULK) {
        bulk_store_tile(obs + (size_t)env0 * OBS, tile, tile_bytes, lane);
        bulk_store_wait_read(lane);
    } else {
        lsu_store_tile(obs + (size_t)env0 * OBS, tile, tile_bytes, lane);
    }
}

// explicit-options reset (AKE:102-211): option rows are [n_reset,7|6] row-major, slots optionally scattered
__global__ void __launch_bounds__(128)
kin_reset_kernel(const __grid_constant__ KinEnvParams P, float* __restrict__ state, int stride, int n_envs,
                 const int* __restrict__ env_ids, int n_reset, int mode, const float* __restrict__ iq,
                 const float* __restrict__ idq, const float* __restrict__ ipa, const float* __restrict__ gq,
                 const float* __restrict__ gpose, float* __restrict__ obs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_reset) return;
    const int env = env_ids ? env_ids[i] : i;
    if (env < 0 || env >= n_envs) return;
    float r_iq[NJ], r_idq[NJ], r_ipa[NJ], r_gq[NJ], r_gp[6], gq_out[NJ];
#pragma unroll
    for (int k = 0; k < NJ; ++k) {
        r_iq[k] = iq[(size_t)i * NJ + k];
        r_idq[k] = idq ? idq[(size_t)i * NJ + k] : 0.0f;
        r_ipa[k] = ipa ? ipa[(size_t)i * NJ + k] : 0.0f;
        r_gq[k] = gq ? gq[(size_t)i * NJ + k] : 0.0f;
    }
    if (gpose) {
#pragma unroll
        for (int k = 0; k < 6; ++k) r_gp[k] = gpose[(size_t)i * 6 + k];
    }
    EnvRegs s;
    s.flags = ld_row_u(state, stride, KIN_ROW_FLAGS, env) & (0xfu << KIN_FLAG_STAGE_SHIFT);  // keep the stage tag
    reset_core(P, s, mode, r_iq, r_idq, r_ipa, r_gq, gpose ? r_gp : nullptr, gq_out);
    store_env_reset(state, stride, env, s, gq_out);
    if (obs) {
        float o[OBS];
        build_obs(P, s, mode, o);
        float4* dst = reinterpret_cast<float4*>(obs + (size_t)i * OBS);
#pragma unroll
        for (int k = 0; k < OBS / 4; ++k) dst[k] = make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
    }
}

// device-sampled reset of the masked slots (sample_approach_reset / sample_dock_reset + reset)
__global__ void __launch_bounds__(128)
kin_reset_sampled_kernel(const __grid_constant__ KinEnvParams P, const KinSamplerParams* __restrict__ S, float* __restrict__ state,
                         int stride, int n_envs, const uint8_t* __restrict__ mask, int mode, uint64_t seed,
                         float* __restrict__ obs) {
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= n_envs) return;
    if (mask && !mask[env]) return;
    const unsigned episode = ld_row_u(state, stride, KIN_ROW_EPISODE, env) + 1u;
    Philox rng(seed, (unsigned)env, episode);
    ResetDraw d;
    sample_reset(P, *S, rng, mode, d);
    EnvRegs s;
    s.flags = 0u;
    float gq[NJ];
    reset_core(P, s, mode, d.iq, d.idq, d.ipa, d.gq, d.has_gpose ? d.gpose : nullptr, gq);
    s.flags |= (unsigned)d.stage << KIN_FLAG_STAGE_SHIFT;
    store_env_reset(state, stride, env, s, gq);
    st_row_u(state, stride, KIN_ROW_EPISODE, env, episode);
    if (obs) {
        float o[OBS];
        build_obs(P, s, mode, o);
        float4* dst = reinterpret_cast<float4*>(obs + (size_t)env * OBS);
#pragma unroll
        for (int k = 0; k < OBS / 4; ++k) dst[k] = make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
    }
}

// current_observation() for the whole batch (AKE:381)
__global__ void __launch_bounds__(STEP_THREADS)
kin_observe_kernel(const __grid_constant__ KinEnvParams P, const float* __restrict__ state, int stride, int n, float* __restrict__ obs) {
    __shared__ __align__(128) float tiles[STEP_WARPS][OBS_TILE_FLOATS];
    const int lane = threadIdx.x & (WARP - 1);
    const int warp = threadIdx.x >> 5;
    const int env0 = (blockIdx.x * STEP_WARPS + warp) * WARP;
    if (env0 >= n) return;
    const int envc = min(env0 + lane, n - 1);
    EnvRegs s;
    load_env<false>(state, stride, envc, s);
    float o[OBS];
    build_obs(P, s, (int)((s.flags >> KIN_FLAG_MODE_SHIFT) & 3u), o);
    stage_obs_row(tiles[warp], lane, o);
    bulk_store_tile(obs + (size_t)env0 * OBS, tiles[warp], min(WARP, n - env0) * OBS * 4, lane);
    bulk_store_wait_read(lane);
}

__global__ void __launch_bounds__(128)
kin_fk_kernel(const __grid_constant__ KinEnvParams P, const float* __restrict__ q, float* __restrict__ pose6, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float r[NJ], p[6];
#pragma unroll
    for (int k = 0; k < NJ; ++k) r[k] = q[(size_t)i * NJ + k];
    fk_pose6(P, r, p);
#pragma unroll
    for (int k = 0; k < 6; ++k) pose6[(size_t)i * 6 + k] = p[k];
}

template <int MODE, bool AUTORESET>
static cudaError_t launch_step_mode(const KinHandle* h, float* state, int stride, int n, const float* action, float* obs, float* reward,
                                    uint8_t* done, float* aux, float* comps, uint64_t seed, float* terminal_obs, cudaStream_t st) {
    const int blocks = (n + STEP_THREADS - 1) / STEP_THREADS;
    const bool bulk = !kin_env_flag("KIN_NO_BULK_STORE");
#define KIN_LAUNCH(COMP, AUX, BULK)                                                                                                    \
    kin_step_kernel<MODE, COMP, AUX, AUTORESET, BULK><<<blocks, STEP_THREADS, 0, st>>>(h->params, h->d_sampler, state, stride, n, action, obs, \
                                                                                       reward, done, aux, comps, seed, terminal_obs)
    if (comps) {
        if (bulk) KIN_LAUNCH(true, true, true); else KIN_LAUNCH(true, true, false);
    } else if (aux) {
        if (bulk) KIN_LAUNCH(false, true, true); else KIN_LAUNCH(false, true, false);
    } else {
        if (bulk) KIN_LAUNCH(false, false, true); else KIN_LAUNCH(false, false, false);
    }
#undef KIN_LAUNCH
    return cudaGetLastError();
}

}  // namespace kin

using namespace kin;

extern "C" int kin_env_step(void* handle, float* state, int stride, int n_envs, int mode_hint, const float* action, float* obs,
                            float* reward, uint8_t* done, float* aux, float* components, int auto_reset, uint64_t seed,
                            float* terminal_obs, void* stream) {
    KinHandle* h = kin_handle(handle);
    if (!h) return kin_fail(KIN_ERR_INVALID_ARG, "kin_env_step: bad handle");
    if (!state || !action || !obs || !reward || !done) return kin_fail(KIN_ERR_INVALID_ARG, "kin_env_step: null buffer");
    if (n_envs <= 0 || stride < n_envs || (stride % 32) != 0) return kin_fail(KIN_ERR_INVALID_ARG, "kin_env_step: need 0 < n_envs <= stride, stride % 32 == 0");
    if ((components && !aux)) return kin_fail(KIN_ERR_INVALID_ARG, "kin_env_step: components requires aux");
    if (auto_reset && !h->d_sampler) return kin_fail(KIN_ERR_INVALID_ARG, "kin_env_step: auto_reset needs kin_params_set_sampler");
    if (((uintptr_t)obs & 15u) || (terminal_obs && ((uintptr_t)terminal_obs & 15u))) return kin_fail(KIN_ERR_INVALID_ARG, "kin_env_step: obs must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
    if (auto_reset) {
        if (mode_hint == KIN_MODE_APPROACH) e = launch_step_mode<KIN_MODE_APPROACH, true>(h, state, stride, n_envs, action, obs, reward, done, aux, components, seed, terminal_obs, st);
        else if (mode_hint == KIN_MODE_DOCK) e = launch_step_mode<KIN_MODE_DOCK, true>(h, state, stride, n_envs, action, obs, reward, done, aux, components, seed, terminal_obs, st);
        else return kin_fail(KIN_ERR_UNSUPPORTED, "kin_env_step: auto_reset needs a uniform mode (approach or dock)");
    } else {
        if (mode_hint == KIN_MODE_APPROACH) e = launch_step_mode<KIN_MODE_APPROACH, false>(h, state, stride, n_envs, action, obs, reward, done, aux, components, seed, nullptr, st);
        else if (mode_hint == KIN_MODE_DOCK) e = launch_step_mode<KIN_MODE_DOCK, false>(h, state, stride, n_envs, action, obs, reward, done, aux, components, seed, nullptr, st);
        else if (mode_hint == KIN_MODE_PER_ENV) e = launch_step_mode<KIN_MODE_PER_ENV, false>(h, state, stride, n_envs, action, obs, reward, done, aux, components, seed, nullptr, st);
        else return kin_fail(KIN_ERR_INVALID_ARG, "kin_env_step: bad mode_hint");
    }
    return e == cudaSuccess ? KIN_OK : kin_fail_cuda(e, "kin_env_step");
}

extern "C" int kin_env_reset(void* handle, float* state, int stride, int n_envs, const int* env_ids, int n_reset, int mode,
                             const float* initial_q, const float* initial_dq, const float* initial_prev_action, const float* goal_q,
                             const float* goal_pose6, float* obs, void* stream) {
    KinHandle* h = kin_handle(handle);
    if (!h) return kin_fail(KIN_ERR_INVALID_ARG, "kin_env_reset: bad handle");
    if (!state || !initial_q || (!goal_q && !goal_pose6)) return kin_fail(KIN_ERR_INVALID_ARG, "kin_env_reset: need state, initial_q and goal_q or goal_pose6");
    if (n_envs <= 0 || stride < n_envs || (stride % 32) != 0 || n_reset < 0 || (!env_ids && n_reset > n_envs)) return kin_fail(KIN_ERR_INVALID_ARG, "kin_env_reset: bad sizes");
    if (mode != KIN_MODE_APPROACH && mode != KIN_MODE_DOCK) return kin_fail(KIN_ERR_UNSUPPORTED, "kin_env_reset: mode must be approach (0) or dock (1)");
    if (obs && ((uintptr_t)obs & 15u)) return kin_fail(KIN_ERR_INVALID_ARG, "kin_env_reset: obs must be 16-byte aligned");
    if (n_reset == 0) return KIN_OK;
    kin_reset_kernel<<<(n_reset + 127) / 128, 128, 0, (cudaStream_t)stream>>>(h->params, state, stride, n_envs, env_ids, n_reset, mode, initial_q,
                                                                              initial_dq, initial_prev_action, goal_q, goal_pose6, obs);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? KIN_OK : kin_fail_cuda(e, "kin_env_reset");
}

extern "C" int kin_env_reset_sampled(void* handle, float* state, int stride, int n_envs, const uint8_t* mask, int mode, uint64_t seed,
                                     float* obs, void* stream) {
    KinHandle* h = kin_handle(handle);
    if (!h) return kin_fail(KIN_ERR_INVALID_ARG, "kin_env_reset_sampled: bad handle");
    if (!h->d_sampler) return kin_fail(KIN_ERR_INVALID_ARG, "kin_env_reset_sampled: call kin_params_set_sampler first");
    if (!state || n_envs <= 0 || stride < n_envs || (stride % 32) != 0) return kin_fail(KIN_ERR_INVALID_ARG, "kin_env_reset_sampled: bad sizes");
    if (mode != KIN_MODE_APPROACH && mode != KIN_MODE_DOCK) return kin_fail(KIN_ERR_UNSUPPORTED, "kin_env_reset_sampled: mode must be approach (0) or dock (1)");
    kin_reset_sampled_kernel<<<(n_envs + 127) / 128, 128, 0, (cudaStream_t)stream>>>(h->params, h->d_sampler, state, stride, n_envs, mask, mode, seed, obs);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? KIN_OK : kin_fail_cuda(e, "kin_env_reset_sampled");
}

extern "C" int kin_env_observe(void* handle, const float* state, int stride, int n_envs, float* obs, void* stream) {
    KinHandle* h = kin_handle(handle);
    if (!h) return kin_fail(KIN_ERR_INVALID_ARG, "kin_env_observe: bad handle");
    if (!state || !obs || n_envs <= 0 || stride < n_envs || (stride % 32) != 0 || ((uintptr_t)obs & 15u)) return kin_fail(KIN_ERR_INVALID_ARG, "kin_env_observe: bad args");
    kin_observe_kernel<<<(n_envs + STEP_THREADS - 1) / STEP_THREADS, STEP_THREADS, 0, (cudaStream_t)stream>>>(h->params, state, stride, n_envs, obs);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? KIN_OK : kin_fail_cuda(e, "kin_env_observe");
}

extern "C" int kin_fk_pose6(void* handle, const float* q, float* pose6, int n, void* stream) {
    KinHandle* h = kin_handle(handle);
    if (!h) return kin_fail(KIN_ERR_INVALID_ARG, "kin_fk_pose6: bad handle");
    if (!q || !pose6 || n < 0) return kin_fail(KIN_ERR_INVALID_ARG, "kin_fk_pose6: bad args");
    if (n == 0) return KIN_OK;
    kin_fk_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(h->params, q, pose6, n);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? KIN_OK : kin_fail_cuda(e, "kin_fk_pose6");
}
