// kin_peer.cu -- the PPO gradient exchange over NVLink peer memory (the one real exchange step of the training path:
// SB3 single-process PPO sums the minibatch gradient over all samples; with one rank per GPU that is an all-reduce per minibatch,
// 128 of them per update, each only 64 KB -- latency-bound, so the cost is launch + synchronisation, not bandwidth).
//
// Instead of  [reduce CTA partials] -> NCCL all-reduce -> [Adam]  the reduction kernel itself PUSHES the rank's gradient into
// every peer's receive buffer (posted NVLink stores, fire and forget) and bumps a per-sender counter there; the gather kernel
// of each rank waits for the counters of all senders and adds the slots in rank order, so every rank computes the bitwise
// identical sum and the parameters never drift apart.  One exchange = two small kernels, no host involvement, no ring.
//
// The receive-buffer layout and the in-kernel form of the exchange (the gradient kernel's tail) live in kin_peer.cuh; this file
// keeps the buffer management and the two-kernel form (push + gather), which the strict-fp32 update still uses.
//   A slot is rewritten two exchanges later; by then every reader has passed the wait of the exchange in between, which its
//   sender only reaches after its own gather of this exchange (stream order), so two slots suffice.
#include <cstdlib>
#include <cstring>

#include "kin_peer.cuh"

namespace kin {

// reduce the per-CTA partial gradients (as kin_ppo_reduce_kernel) and store the result into slot[parity][rank] of EVERY rank
__global__ void __launch_bounds__(256)
kin_peer_push_kernel(const float* __restrict__ partials, int n_cta, int P, float inv_global_batch, PeerTable peers, int rank, int world,
                     unsigned epoch) {
    __shared__ float part[8][32];
    __shared__ int last;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int p = blockIdx.x * 32 + lane;
    const int prow = P + KIN_PPO_STATS + 8;
    float a0 = 0.0f, a1 = 0.0f;
    if (p < P + 5) {
        int c = w;
        for (; c + 8 < n_cta; c += 16) {
            a0 += __ldg(partials + (size_t)c * prow + p);
            a1 += __ldg(partials + (size_t)(c + 8) * prow + p);
        }
        if (c < n_cta) a0 += __ldg(partials + (size_t)c * prow + p);
    }
    part[w][lane] = a0 + a1;
    __syncthreads();
    // warp q of the block stores the 32 values into peer q's slot: one 128-byte posted store per peer
    if (w < world && p < P + 5) {
        float a = part[0][lane];
#pragma unroll
        for (int k = 1; k < 8; ++k) a += part[k][lane];
        if (p >= P) a *= inv_global_batch;
        float* slot = reinterpret_cast<float*>(peers.base[w] + PEER_HEADER) + ((size_t)(epoch & 1u) * world + rank) * peer_row(P);
        slot[p] = a;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();      // cumulative: after the block barrier it also orders the other threads' peer stores
        unsigned* count = reinterpret_cast<unsigned*>(peers.base[rank] + 64);
        last = atomicAdd(count, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last) {      // every CTA's stores are ordered before its increment: tell the peers this push has landed
        __threadfence_system();
        if (threadIdx.x < world) {
            unsigned* arrived = reinterpret_cast<unsigned*>(peers.base[threadIdx.x]) + rank;
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(arrived), "r"(epoch) : "memory");
        }
        if (threadIdx.x == 0) *reinterpret_cast<unsigned*>(peers.base[rank] + 64) = 0u;
    }
}

// wait until the pushes of all ranks for `epoch` have landed in this rank's buffer, then grad = sum over ranks (rank order)
__global__ void __launch_bounds__(256)
kin_peer_gather_kernel(const unsigned char* __restrict__ local, int world, unsigned epoch, int P, float* __restrict__ grad, float* __restrict__ stats,
                       unsigned long long timeout_cycles, int* __restrict__ timed_out) {
    if (threadIdx.x < world && *reinterpret_cast<volatile int*>(timed_out) == 0) {      // sticky: after one timeout nobody waits again
        const unsigned* arrived = reinterpret_cast<const unsigned*>(local) + threadIdx.x;
        const long long t0 = clock64();
        // counters only grow; (int) difference tolerates wrap-around
        while ((int)(ld_acquire_sys(arrived) - epoch) < 0) {
            if ((unsigned long long)(clock64() - t0) > timeout_cycles) {     // a peer died: report instead of hanging the GPU
                atomicExch(timed_out, 1);
                break;
            }
            __nanosleep(64);
        }
    }
    __syncthreads();
    // a peer that never delivered leaves stale slots behind: mark the minibatch so kin_ppo_adam skips it on this rank (the sticky
    // flag makes every later exchange skip too -- parameters stay where they were until the host raises, they never diverge)
    if (blockIdx.x == 0 && threadIdx.x == 0 && stats) stats[KIN_PPO_STAT_SKIP] = *reinterpret_cast<volatile int*>(timed_out) ? 1.0f : 0.0f;
    const int p = blockIdx.x * 256 + threadIdx.x;
    if (p >= P + 5) return;
    const float* slot = reinterpret_cast<const float*>(local + PEER_HEADER) + (size_t)(epoch & 1u) * world * peer_row(P);
    float a = 0.0f;
    for (int r = 0; r < world; ++r) a += __ldcg(slot + (size_t)r * peer_row(P) + p);      // L2 is where the peers' stores land
    if (p < P) grad[p] = a;
    else if (stats) stats[p - P] = a;
}

}  // namespace kin

// device-side wait limit in SM clocks: KIN_PEER_TIMEOUT_S seconds (default 30) at ~2 GHz; first-call module loads, checkpoint writes
// or a throttled peer can skew ranks by seconds
unsigned long long kin::kin_peer_timeout_cycles() {
    static unsigned long long timeout = 0ull;
    if (timeout == 0ull) {
        double sec = 30.0;
        if (const char* v = getenv("KIN_PEER_TIMEOUT_S")) { const double f = atof(v); if (f > 0.0) sec = f; }
        timeout = (unsigned long long)(sec * 2.0e9);
    }
    return timeout;
}

using namespace kin;

extern "C" int kin_peer_buffer_bytes(int n_params, int world) {
    return (int)peer_buffer_size(n_params, world);
}

extern "C" int kin_peer_buffer_create(int n_params, int world, void** buffer, unsigned char* ipc_handle) {
    if (!buffer || !ipc_handle || world < 1 || world > PEER_MAX || n_params <= 0) return kin_fail(KIN_ERR_INVALID_ARG, "kin_peer_buffer_create: bad arguments");
    const size_t bytes = (size_t)kin_peer_buffer_bytes(n_params, world);
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) e = cudaMemset(p, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        if (p) cudaFree(p);
        return kin_fail_cuda(e, "kin_peer_buffer_create");
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == KIN_PEER_HANDLE_BYTES, "IPC handle size");
    memcpy(ipc_handle, &h, sizeof(h));
    *buffer = p;
    return KIN_OK;
}

extern "C" int kin_peer_buffer_open(const unsigned char* ipc_handle, void** buffer) {
    if (!ipc_handle || !buffer) return kin_fail(KIN_ERR_INVALID_ARG, "kin_peer_buffer_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, sizeof(h));
    cudaError_t e = cudaIpcOpenMemHandle(buffer, h, cudaIpcMemLazyEnablePeerAccess);
    return e == cudaSuccess ? KIN_OK : kin_fail_cuda(e, "kin_peer_buffer_open (peers must be GPUs of one node with P2P access)");
}

extern "C" int kin_peer_buffer_close(void* buffer) {
    cudaError_t e = cudaIpcCloseMemHandle(buffer);
    return e == cudaSuccess ? KIN_OK : kin_fail_cuda(e, "kin_peer_buffer_close");
}

extern "C" int kin_peer_buffer_destroy(void* buffer) {
    cudaError_t e = cudaFree(buffer);
    return e == cudaSuccess ? KIN_OK : kin_fail_cuda(e, "kin_peer_buffer_destroy");
}

extern "C" int kin_peer_grad_push(const float* partials, int n_cta, int n_params, long long global_batch, void* const* peer_buffers, int rank, int world,
                                  unsigned epoch, void* stream) {
    if (!partials || !peer_buffers || n_cta <= 0 || n_params <= 0 || global_batch <= 0 || world < 1 || world > PEER_MAX || rank < 0 || rank >= world || epoch == 0u)
        return kin_fail(KIN_ERR_INVALID_ARG, "kin_peer_grad_push: bad arguments (epoch counts exchanges from 1)");
    PeerTable t{};
    for (int i = 0; i < world; ++i) {
        if (!peer_buffers[i]) return kin_fail(KIN_ERR_INVALID_ARG, "kin_peer_grad_push: null peer buffer");
        t.base[i] = static_cast<unsigned char*>(peer_buffers[i]);
    }
    const int blocks = (n_params + 5 + 31) / 32;
    kin_peer_push_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(partials, n_cta, n_params, 1.0f / (float)global_batch, t, rank, world, epoch);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? KIN_OK : kin_fail_cuda(e, "kin_peer_grad_push");
}

extern "C" int kin_peer_grad_gather(const void* local_buffer, int n_params, int world, unsigned epoch, float* grad, float* stats, int* timed_out,
                                    void* stream) {
    if (!local_buffer || !grad || !timed_out || n_params <= 0 || world < 1 || world > PEER_MAX || epoch == 0u)
        return kin_fail(KIN_ERR_INVALID_ARG, "kin_peer_grad_gather: bad arguments");
    const int blocks = (n_params + 5 + 255) / 256;
    const unsigned long long timeout = kin_peer_timeout_cycles();
    kin_peer_gather_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(static_cast<const unsigned char*>(local_buffer), world, epoch, n_params, grad, stats,
                                                                     timeout, timed_out);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? KIN_OK : kin_fail_cuda(e, "kin_peer_grad_gather");
}
