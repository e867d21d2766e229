// kin_internal.h -- host-side handle and error plumbing shared by the .cu translation units.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

#include "kin_b200.h"

namespace kin {

constexpr uint32_t KIN_HANDLE_MAGIC = 0x4b494e42u;  // "KINB"

struct KinHandle {
    uint32_t magic;
    KinEnvParams params;               // host copy; passed to kernels by value (constant bank)
    KinSamplerParams* d_sampler;       // device copy (indexed per lane by stage), nullptr until set
    KinSamplerParams h_sampler;
};

inline KinHandle* kin_handle(void* p) {
    KinHandle* h = static_cast<KinHandle*>(p);
    return (h && h->magic == KIN_HANDLE_MAGIC) ? h : nullptr;
}

// cudaFuncSetAttribute is per device: one-time guards are indexed by the current device (a process may drive several GPUs)
constexpr int KIN_MAX_DEVICES = 64;
inline int kin_device_slot() {
    int d = 0;
    cudaGetDevice(&d);
    return (d >= 0 && d < KIN_MAX_DEVICES) ? d : 0;
}

int kin_fail(int code, const char* msg);
int kin_fail_cuda(cudaError_t e, const char* where);
bool kin_env_flag(const char* name);

}  // namespace kin
