// kin_ppo_layout.cuh -- flat parameter layout of the PPO trainer (PARAM_ORDER in ppo.py) shared by the update kernels.
#pragma once

#include "kin_internal.h"

namespace kin {

struct PpoOffsets {
    int pi_w0, pi_b0, pi_w1, pi_b1, act_w, act_b, vf_w0, vf_b0, vf_w1, vf_b1, val_w, val_b, log_std, total;
};
__host__ __device__ inline PpoOffsets ppo_offsets(int in_dim) {
    PpoOffsets o;
    int p = 0;
    o.pi_w0 = p; p += 64 * in_dim;
    o.pi_b0 = p; p += 64;
    o.pi_w1 = p; p += 4096;
    o.pi_b1 = p; p += 64;
    o.act_w = p; p += 7 * 64;
    o.act_b = p; p += 7;
    o.vf_w0 = p; p += 64 * in_dim;
    o.vf_b0 = p; p += 64;
    o.vf_w1 = p; p += 4096;
    o.vf_b1 = p; p += 64;
    o.val_w = p; p += 64;
    o.val_b = p; p += 1;
    o.log_std = p; p += 7;
    o.total = p;
    return o;
}

// grad[p] = sum over CTAs of partials[c][p] (rows of P + KIN_PPO_STATS + 8 floats); stats (nullable) likewise, scaled
int kin_ppo_reduce_launch(const float* partials, int n_cta, int P, float* grad, float* stats, float inv_global_batch, cudaStream_t st);

}  // namespace kin
