// kin_route_collect.cu -- K4-route: the fused PPO rollout collection of the 80-input route policy.  ONE launch runs n_steps of
//   actor + critic forward (tcgen05 kind::f16, K = 96) -> Gaussian action sample + log-prob -> route wrapper step (base env step,
//   route-ready streak, in-episode waypoint advance, 13-term route reward) -> in-register sampled route reset of finished slots
//   -> rollout-buffer writes
// for every replica, the env + route state in registers for the whole rollout and the waypoint table in shared memory.
//
// Replaces SB3's collect_rollouts over a VecEnv of RouteSequenceKinematicEnv / RouteKinematicEnv (call site
// kinematic_phase1/train_route_curriculum.py:113-145; env: route/route_sequence_env.py:96-278, route/route_env.py:49-212, reset:
// route/route_reset_samplers.py:43-117).  The per-step-launch path of ppo.py (_collect_route: kin_policy_act -> kin_route_step ->
// kin_ppo_bootstrap -> kin_route_reset_sampled, five launches per time step) stays as the restatement this kernel is replayed
// against (tests/test_gpu_ppo.py::test_fused_route_collection_replays_through_the_step_kernels).
//
// Mapping as in kin_collect.cu: one replica <-> one thread <-> one row of the A operand <-> one TMEM lane; a CTA holds TILES
// independent 128-replica tiles on named barriers sharing one bf16 copy of the weights.  Per step and tile:
//   X = [obs80 | 1 | 0]  bf16, two SWIZZLE_128B K-major tiles (columns 0..63 | 64..127, 96 used)
//   Z = X [W0a;W0c]^T            128 x 128 x 96   -> tanh -> H1 (actor half over X's first tile, critic half over its second)
//   Z = H1a W1a^T | H1c W1c^T    2 x (128 x 64 x 64) -> tanh(+b1) -> H2 in place
//   O = H2a WOa^T + H2c WOc^T    128 x 16 x 64    -> 7 action means + value
// The observation goes to the rollout buffer as fp32 rows ([T + 1][n][80]: what kin_ppo_grad_tc<80> consumes).  Episodes that hit
// the time limit append their terminal observation to a list; kin_ppo_bootstrap_list adds gamma * V(terminal_obs) afterwards.
#include "kin_ppo_layout.cuh"
#include "kin_route_core.cuh"
#include "kin_umma.cuh"

namespace kin {

using namespace umma;

constexpr int RC_ROWS = 128;
constexpr int RC_TILE_BYTES = RC_ROWS * 128;
constexpr int RC_IN = KIN_ROUTE_OBS_DIM;      // 80
constexpr float kHalfLog2PiR = 0.91893853320467274178f;

template <int TILES>
struct __align__(1024) RouteCollectSmem {
    unsigned char XH[TILES][2][RC_TILE_BYTES];   // [.][0]: X columns 0..63, then actor H1 / H2; [.][1]: X columns 64..127, then critic H1 / H2
    unsigned char W0[2][RC_TILE_BYTES];          // [k tile][128 rows: actor 0..63 | critic 64..127][64 cols]; column 80 (tile 1, col 16) = b0
    unsigned char W1[2][64 * 128];
    unsigned char WO[2][16 * 128];
    float b1[128];
    float bo[8];
    float ls[8];
    float sig[8];
    unsigned long long mbar[TILES];
    unsigned tmem_base;
    alignas(16) float q_table[ROUTE_MAX_SMEM_WP * KIN_NJ];   // waypoint joint vectors for the nearest-waypoint scan (n <= 1024)
};

struct RouteCollectOut {
    float* obs;               // [T + 1][n][80]
    float* action;            // [T][n][7]
    float* logp;              // [T][n]
    float* value;             // [T][n]
    float* reward;            // [T][n]
    uint8_t* done;            // [T][n]
    uint8_t* episode_start;   // [T][n]
    int* flags;               // [T][n]  route flags of the step (bit0 ready, bit1 regression, bit2 orientation hit, bit3 waypoint success)
    uint8_t* start_io;        // [n]
    float* last_value;        // [n]
    int* boot_count;
    int* boot_index;          // [cap]  t * n + env
    float* boot_obs;          // [cap][80]
    int boot_cap;
};

struct RTile {
    unsigned char *X0, *X1;
    unsigned aX0, aX1, aW0a, aW0b, aW1a, aW1c, aWOa, aWOc, mb, tz_mma, tz_row;
    unsigned par;
    int tile, row;
};

__device__ __forceinline__ bool rc_elect() {
    unsigned pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0u;
}
__device__ __forceinline__ void rc_tile_bar(int tile) { asm volatile("bar.sync %0, %1;" ::"r"(tile + 1), "r"(RC_ROWS) : "memory"); }

// actor + critic forward of the tile's 128 observations (80 floats each); collective over the tile's 128 threads.
// out[0..6] = action means, out[7] = value.
template <int TILES>
__device__ __forceinline__ void route_policy_forward_tc(RouteCollectSmem<TILES>& S, RTile& c, const float* o, float* out) {
    // ---- X image: 80 obs | 1 | zeros, two K tiles
#pragma unroll
    for (int ch = 0; ch < 8; ++ch)
        *reinterpret_cast<uint4*>(c.X0 + sw_chunk(c.row, ch)) = make_uint4(pack_bf16(o[8 * ch], o[8 * ch + 1]), pack_bf16(o[8 * ch + 2], o[8 * ch + 3]),
                                                                            pack_bf16(o[8 * ch + 4], o[8 * ch + 5]), pack_bf16(o[8 * ch + 6], o[8 * ch + 7]));
#pragma unroll
    for (int ch = 0; ch < 2; ++ch)
        *reinterpret_cast<uint4*>(c.X1 + sw_chunk(c.row, ch)) = make_uint4(pack_bf16(o[64 + 8 * ch], o[65 + 8 * ch]), pack_bf16(o[66 + 8 * ch], o[67 + 8 * ch]),
                                                                            pack_bf16(o[68 + 8 * ch], o[69 + 8 * ch]), pack_bf16(o[70 + 8 * ch], o[71 + 8 * ch]));
    *reinterpret_cast<uint4*>(c.X1 + sw_chunk(c.row, 2)) = make_uint4(0x00003F80u, 0u, 0u, 0u);      // column 80 = 1 (bias carrier)
    *reinterpret_cast<uint4*>(c.X1 + sw_chunk(c.row, 3)) = make_uint4(0u, 0u, 0u, 0u);               // columns 88..95 (the critic's H wrote here)
    fence_async_smem();
    fence_before();
    rc_tile_bar(c.tile);
    if (c.row < 32 && rc_elect()) {
        fence_after();
        constexpr unsigned id = idesc_bf16(128, 128, false, false);
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_bf16(c.tz_mma, desc_k(c.aX0 + k * 32), desc_k(c.aW0a + k * 32), id, k > 0);
#pragma unroll
        for (int k = 0; k < 2; ++k) mma_bf16(c.tz_mma, desc_k(c.aX1 + k * 32), desc_k(c.aW0b + k * 32), id, 1u);
        commit(c.mb);
    }
    mbar_wait(c.mb, c.par);
    c.par ^= 1u;
    fence_after();
    // ---- H1 = tanh(Z): columns 0..63 actor -> X0's place, 64..127 critic -> X1's place
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float v[32];
        tmem_ld32(c.tz_row + q * 32, v);
        unsigned char* dst = (q < 2) ? c.X0 : c.X1;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            unsigned p[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) p[e] = pack_bf16(tanh_fast(v[8 * j + 2 * e]), tanh_fast(v[8 * j + 2 * e + 1]));
            *reinterpret_cast<uint4*>(dst + sw_chunk(c.row, (q & 1) * 4 + j)) = make_uint4(p[0], p[1], p[2], p[3]);
        }
    }
    fence_async_smem();
    fence_before();
    rc_tile_bar(c.tile);
    if (c.row < 32 && rc_elect()) {
        fence_after();
        constexpr unsigned id = idesc_bf16(128, 64, false, false);
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_bf16(c.tz_mma, desc_k(c.aX0 + k * 32), desc_k(c.aW1a + k * 32), id, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_bf16(c.tz_mma + 64, desc_k(c.aX1 + k * 32), desc_k(c.aW1c + k * 32), id, k > 0);
        commit(c.mb);
    }
    mbar_wait(c.mb, c.par);
    c.par ^= 1u;
    fence_after();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float v[32];
        tmem_ld32(c.tz_row + q * 32, v);
        unsigned char* dst = (q < 2) ? c.X0 : c.X1;
        const float* b = S.b1 + q * 32;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            unsigned p[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
                p[e] = pack_bf16(tanh_fast(v[8 * j + 2 * e] + b[8 * j + 2 * e]), tanh_fast(v[8 * j + 2 * e + 1] + b[8 * j + 2 * e + 1]));
            *reinterpret_cast<uint4*>(dst + sw_chunk(c.row, (q & 1) * 4 + j)) = make_uint4(p[0], p[1], p[2], p[3]);
        }
    }
    fence_async_smem();
    fence_before();
    rc_tile_bar(c.tile);
    if (c.row < 32 && rc_elect()) {
        fence_after();
        constexpr unsigned id = idesc_bf16(128, 16, false, false);
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_bf16(c.tz_mma, desc_k(c.aX0 + k * 32), desc_k(c.aWOa + k * 32), id, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_bf16(c.tz_mma, desc_k(c.aX1 + k * 32), desc_k(c.aWOc + k * 32), id, 1u);
        commit(c.mb);
    }
    mbar_wait(c.mb, c.par);
    c.par ^= 1u;
    fence_after();
    float r[16];
    tmem_ld16(c.tz_row, r);
    fence_before();
#pragma unroll
    for (int d = 0; d < 8; ++d) out[d] = r[d] + S.bo[d];
}

__device__ __forceinline__ void store_obs_row80(float* dst, const float* o) {
    float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
    for (int k = 0; k < RC_IN / 4; ++k) d4[k] = make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
}

template <int TILES, bool SEQ>
__global__ void __launch_bounds__(TILES * RC_ROWS, 1)
kin_route_collect_kernel(const __grid_constant__ KinEnvParams P, RouteView R, const __grid_constant__ KinRouteResetParams C, float* __restrict__ state,
                         int stride, int n, const float* __restrict__ params, int T, uint64_t noise_seed, uint32_t step0, uint64_t reset_seed,
                         int reset_streak, RouteCollectOut out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    RouteCollectSmem<TILES>& S = *reinterpret_cast<RouteCollectSmem<TILES>*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    const PpoOffsets O = ppo_offsets(RC_IN);
    const int tid = threadIdx.x, warp = tid >> 5;
    constexpr int NT = TILES * RC_ROWS;

    // ---- weights -> bf16 operand tiles (both nets), converted here from the flat fp32 parameters -----------------------------
    for (int i = tid; i < 2 * 128 * 64; i += NT) {       // W0: [k tile][net row][col]
        const int kt = i >> 13, nn = (i >> 6) & 127, kc = i & 63, k = kt * 64 + kc, u = nn & 63;
        const int wbase = (nn < 64) ? O.pi_w0 : O.vf_w0, bbase = (nn < 64) ? O.pi_b0 : O.vf_b0;
        const float v = k < RC_IN ? __ldg(params + wbase + u * RC_IN + k) : (k == RC_IN ? __ldg(params + bbase + u) : 0.0f);
        *reinterpret_cast<unsigned short*>(S.W0[kt] + sw_elem(nn, kc)) = (unsigned short)(pack_bf16(v, 0.0f) & 0xffffu);
    }
    for (int i = tid; i < 2 * 4096; i += NT) {
        const int nt = i >> 12, u = (i >> 6) & 63, k = i & 63;
        *reinterpret_cast<unsigned short*>(S.W1[nt] + sw_elem(u, k)) =
            (unsigned short)(pack_bf16(__ldg(params + (nt ? O.vf_w1 : O.pi_w1) + u * 64 + k), 0.0f) & 0xffffu);
    }
    {
        uint4* zw = reinterpret_cast<uint4*>(S.WO);
        for (int i = tid; i < 2 * 16 * 128 / 16; i += NT) zw[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    for (int i = tid; i < 7 * 64; i += NT)
        *reinterpret_cast<unsigned short*>(S.WO[0] + sw_elem(i >> 6, i & 63)) = (unsigned short)(pack_bf16(__ldg(params + O.act_w + i), 0.0f) & 0xffffu);
    if (tid < 64) *reinterpret_cast<unsigned short*>(S.WO[1] + sw_elem(7, tid)) = (unsigned short)(pack_bf16(__ldg(params + O.val_w + tid), 0.0f) & 0xffffu);
    if (tid < 128) S.b1[tid] = __ldg(params + (tid < 64 ? O.pi_b1 + tid : O.vf_b1 + tid - 64));
    if (tid < 8) {
        const float ls = tid < 7 ? __ldg(params + O.log_std + tid) : 0.0f;
        S.bo[tid] = tid < 7 ? __ldg(params + O.act_b + tid) : __ldg(params + O.val_b);
        S.ls[tid] = ls;
        S.sig[tid] = expf(ls);
    }
    load_q_table(S.q_table, R, tid, NT);
    if (warp == 0) tmem_alloc(smem_u32(&S.tmem_base), TILES * 128);
    if (tid < TILES) {
        mbar_init(smem_u32(&S.mbar[tid]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();

    RTile c;
    c.tile = tid >> 7;
    c.row = tid & 127;
    c.X0 = S.XH[c.tile][0];
    c.X1 = S.XH[c.tile][1];
    c.aX0 = smem_u32(c.X0);
    c.aX1 = smem_u32(c.X1);
    c.aW0a = smem_u32(S.W0[0]);
    c.aW0b = smem_u32(S.W0[1]);
    c.aW1a = smem_u32(S.W1[0]);
    c.aW1c = smem_u32(S.W1[1]);
    c.aWOa = smem_u32(S.WO[0]);
    c.aWOc = smem_u32(S.WO[1]);
    c.mb = smem_u32(&S.mbar[c.tile]);
    c.tz_mma = S.tmem_base + c.tile * 128;
    c.tz_row = c.tz_mma + ((unsigned)((warp & 3) * 32) << 16);
    c.par = 0u;

    const int gtile = blockIdx.x * TILES + c.tile;
    const int n_tiles = n / RC_ROWS;
    const bool live = gtile < n_tiles;                        // tile-uniform
    const int env = (live ? gtile : 0) * RC_ROWS + c.row;
    const float* table = R.n <= ROUTE_MAX_SMEM_WP ? S.q_table : R.q;

    EnvRegs s;
    load_env<true>(state, stride, env, s);
    RouteRegs rr;
    {
        const unsigned r0 = ld_row_u(state, stride, KIN_ROW_ROUTE, env), r1 = ld_row_u(state, stride, KIN_ROW_ROUTE2, env);
        rr.index = (int)(r0 & 0xffffu); rr.streak = (int)(r0 >> 16); rr.last = (int)(r1 & 0xffffu); rr.completed = (int)(r1 >> 16);
    }
    unsigned start = live ? out.start_io[env] : 0u;

    if (live) {
        // observation of the current (env, route) registers -- what kin_route_step / kin_route_reset_sampled leave in their obs rows
        auto observe = [&](float* o) {
            float o56[OBS];
            build_obs(P, s, KIN_MODE_APPROACH, o56);
            build_route_obs(P, R, s, rr.index, o56, o);
        };
        bool goal_dirty = false;       // an in-episode waypoint advance rewrote goal pose / entry metrics since the last reset
        for (int t = 0; t < T; ++t) {
            const size_t idx = (size_t)t * n + env;
            float mv[8];
            {
                float o[RC_IN];
                observe(o);
                store_obs_row80(out.obs + idx * RC_IN, o);
                route_policy_forward_tc<TILES>(S, c, o, mv);
            }
            // ---- a = mean + sigma * eps, log N(a): the draws of kin_policy_act
            float a[NJ], lp = 0.0f;
            {
                Philox rng(noise_seed, (unsigned)env, step0 + (uint32_t)t);
                float eps[8];
#pragma unroll
                for (int k = 0; k < 8; k += 2) eps[k] = gauss_pair(rng, &eps[k + 1]);
#pragma unroll
                for (int k = 0; k < NJ; ++k) {
                    a[k] = fmaf(S.sig[k], eps[k], mv[k]);
                    lp += -0.5f * eps[k] * eps[k] - S.ls[k] - kHalfLog2PiR;
                    out.action[idx * NJ + k] = a[k];
                }
            }
            out.logp[idx] = lp;
            out.value[idx] = mv[7];
            out.episode_start[idx] = (uint8_t)start;
            // ---- route wrapper step (warp-collective: every lane of a live tile is here)
            StepOut so;
            RouteOut ro;
            const int index_before = rr.index;
            route_step_core<SEQ, false>(P, R, table, s, rr, a, reset_streak != 0, so, ro, nullptr);
            if (SEQ && rr.index != index_before) goal_dirty = true;
            const bool finished = (ro.done & (KIN_DONE_TERMINATED | KIN_DONE_TRUNCATED)) != 0u;
            if (finished) {
                if ((ro.done & KIN_DONE_TRUNCATED) && !(ro.done & KIN_DONE_TERMINATED)) {   // TimeLimit bootstrap list: the terminal observation
                    const int slot = atomicAdd(out.boot_count, 1);
                    if (slot < out.boot_cap) {
                        float o[RC_IN];
                        observe(o);
                        out.boot_index[slot] = (int)idx;
                        store_obs_row80(out.boot_obs + (size_t)slot * RC_IN, o);
                    }
                }
                Philox rng(reset_seed, (uint32_t)env, step0 + (uint32_t)t);
                float gq[NJ];
                sample_route_reset_dev(P, R, C, rng, s, rr, gq);
                store_env_reset(state, stride, env, s, gq);       // goal pose / goal_q / entry rows are only written at resets and advances
                goal_dirty = false;
            }
            out.reward[idx] = ro.reward;
            out.done[idx] = (uint8_t)ro.done;
            out.flags[idx] = (int)ro.flags;
            start = finished ? 1u : 0u;
        }
        // ---- value of the observation after the last step (GAE bootstrap), final observation row, state write-back
        {
            float o[RC_IN], mv[8];
            observe(o);
            store_obs_row80(out.obs + ((size_t)T * n + env) * RC_IN, o);
            route_policy_forward_tc<TILES>(S, c, o, mv);
            out.last_value[env] = mv[7];
        }
        out.start_io[env] = (uint8_t)start;
        store_env_step(state, stride, env, s);
        if (goal_dirty) {      // the last in-episode advance since the last reset: goal pose, entry metrics and goal_q rows
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
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(S.tmem_base, TILES * 128);
}

template <int TILES>
static cudaError_t launch_route_collect(const KinHandle* h, const RouteView& R, const KinRouteResetParams& C, float* state, int stride, int n, const float* params,
                                        int T, uint64_t noise_seed, uint32_t step0, uint64_t reset_seed, int seq, int reset_streak, const RouteCollectOut& out,
                                        cudaStream_t st) {
    const size_t smem = sizeof(RouteCollectSmem<TILES>) + 1024;
    const int n_tiles = n / RC_ROWS;
    const int grid = (n_tiles + TILES - 1) / TILES;
    cudaError_t e;
    if (seq) {
        e = cudaFuncSetAttribute(kin_route_collect_kernel<TILES, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kin_route_collect_kernel<TILES, true><<<grid, TILES * RC_ROWS, smem, st>>>(h->params, R, C, state, stride, n, params, T, noise_seed, step0, reset_seed,
                                                                                 reset_streak, out);
    } else {
        e = cudaFuncSetAttribute(kin_route_collect_kernel<TILES, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kin_route_collect_kernel<TILES, false><<<grid, TILES * RC_ROWS, smem, st>>>(h->params, R, C, state, stride, n, params, T, noise_seed, step0, reset_seed,
                                                                                  reset_streak, out);
    }
    return cudaGetLastError();
}

}  // namespace kin

using namespace kin;

extern "C" int kin_route_collect(void* handle, const KinRouteTable* host_route, const KinRouteResetParams* host_reset, float* state, int stride, int n_envs,
                                 const float* params, int n_steps, uint64_t noise_seed, uint32_t first_step, uint64_t reset_seed, int sequence_mode,
                                 int reset_ready_streak_on_advance, float* obs, float* action, float* logp, float* value, float* reward, uint8_t* done,
                                 uint8_t* episode_start, int* route_flags, uint8_t* start_io, float* last_value, int* boot_count, int* boot_index,
                                 float* boot_obs, int boot_cap, int tiles_per_cta, void* stream) {
    KinHandle* h = kin_handle(handle);
    if (!h || !route_ok(host_route) || !host_reset) return kin_fail(KIN_ERR_INVALID_ARG, "kin_route_collect: bad handle, route table or reset parameters");
    if (!state || !params || !obs || !action || !logp || !value || !reward || !done || !episode_start || !route_flags || !start_io || !last_value ||
        !boot_count || !boot_index || !boot_obs || boot_cap <= 0 || n_steps <= 0)
        return kin_fail(KIN_ERR_INVALID_ARG, "kin_route_collect: null buffer or bad sizes");
    if (n_envs <= 0 || (n_envs % RC_ROWS) != 0 || stride < n_envs || (stride % 32) != 0)
        return kin_fail(KIN_ERR_INVALID_ARG, "kin_route_collect: n_envs must be a positive multiple of 128 (one GEMM tile), stride % 32 == 0");
    if (((uintptr_t)obs & 15u) || ((uintptr_t)boot_obs & 15u)) return kin_fail(KIN_ERR_INVALID_ARG, "kin_route_collect: obs / boot_obs must be 16-byte aligned");
    for (int m = 0; m < 5; ++m)
        if (host_reset->index_lo[m] < 0 || host_reset->index_hi[m] >= host_route->n_waypoints || host_reset->index_lo[m] > host_reset->index_hi[m])
            return kin_fail(KIN_ERR_INVALID_ARG, "kin_route_collect: waypoint range outside the route");
    cudaStream_t st = (cudaStream_t)stream;
    int tiles = tiles_per_cta;
    if (tiles <= 0) {   // fewest tiles per CTA that still fits the batch in one wave of CTAs (one CTA per SM)
        int sms = 148, dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int n_tiles = n_envs / RC_ROWS;
        tiles = n_tiles <= sms ? 1 : (n_tiles <= 2 * sms ? 2 : 4);
    }
    RouteCollectOut out{obs, action, logp, value, reward, done, episode_start, route_flags, start_io, last_value, boot_count, boot_index, boot_obs, boot_cap};
    cudaError_t e = cudaMemsetAsync(boot_count, 0, sizeof(int), st);
    if (e != cudaSuccess) return kin_fail_cuda(e, "kin_route_collect: memset");
    const RouteView R = view_of(host_route);
    if (tiles == 1) e = launch_route_collect<1>(h, R, *host_reset, state, stride, n_envs, params, n_steps, noise_seed, first_step, reset_seed, sequence_mode, reset_ready_streak_on_advance, out, st);
    else if (tiles == 2) e = launch_route_collect<2>(h, R, *host_reset, state, stride, n_envs, params, n_steps, noise_seed, first_step, reset_seed, sequence_mode, reset_ready_streak_on_advance, out, st);
    else if (tiles == 4) e = launch_route_collect<4>(h, R, *host_reset, state, stride, n_envs, params, n_steps, noise_seed, first_step, reset_seed, sequence_mode, reset_ready_streak_on_advance, out, st);
    else return kin_fail(KIN_ERR_INVALID_ARG, "kin_route_collect: tiles_per_cta must be 0 (auto), 1, 2 or 4");
    return e == cudaSuccess ? KIN_OK : kin_fail_cuda(e, "kin_route_collect");
}
