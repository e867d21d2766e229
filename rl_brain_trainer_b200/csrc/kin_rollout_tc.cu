// kin_rollout_tc.cu -- K2 (tensor-core variant) placeholder; see DESIGN.md "K2-TC".
#include "kin_internal.h"

using namespace kin;

int kin_rollout_tc_launch(const KinHandle*, const KinHandle*, const KinPolicyWeights*, const KinPolicyWeights*, const float*, const float*,
                          const float*, const float*, const float*, int, int, int, int, uint32_t*, unsigned long long*, cudaStream_t) {
    return kin_fail(KIN_ERR_UNSUPPORTED, "kin_rollout_approach_finisher: tensor-core variant not built yet");
}
