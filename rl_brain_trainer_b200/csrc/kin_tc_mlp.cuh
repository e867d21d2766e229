// kin_tc_mlp.cuh -- the tcgen05 (kind::tf32) building blocks of the policy-in-the-loop kernels: operand layouts, descriptors,
// MMA issue, TMEM loads, the per-tile named barrier.  Shared by kin_rollout_tc.cu (56-input Approach / Finisher policies) and
// kin_route_tc.cu (80-input route policy).
#pragma once

#include "kin_internal.h"
#include "kin_state.cuh"

namespace kin {

constexpr int TC_TILE = 128;               // episodes per tile == UMMA M == TMEM lanes
constexpr int TC_TILES = 4;                // tiles per CTA (the last one may be partial: 1..4 warps)
constexpr int TC_MAX_THREADS = TC_TILE * TC_TILES;
constexpr int TC_K = 64;                   // padded reduction width of every layer
constexpr int TC_HID = 64;
constexpr int CHUNK_FLOATS_A = TC_TILE * 32;   // one 128-byte-wide K chunk of an A tile: 128 rows x 32 floats
constexpr int A_TILE_FLOATS = 2 * CHUNK_FLOATS_A;
constexpr int CHUNK_FLOATS_W = TC_HID * 32;    // 64 rows x 32 floats
constexpr int W_FLOATS = 2 * CHUNK_FLOATS_W;
constexpr int CHUNK_FLOATS_WO = 8 * 32;        // output layer: 8 rows (7 actions + zero row)
constexpr int WO_FLOATS = 2 * CHUNK_FLOATS_WO;

struct TcSmem {
    float W0[W_FLOATS];                 // 16 KB  [64][64]: 56 inputs | bias column | zero pad
    float W1[W_FLOATS];                 // 16 KB
    float WO[WO_FLOATS];                // 2 KB   [8][64]
    float b1[TC_HID];
    float bo[8];
    unsigned long long mbar[TC_TILES];
    unsigned tmem_base;
    int run_flags[2][TC_TILES][4];   // double-buffered by step parity: written before, read after the layer-1 barrier
    alignas(1024) float A[1][A_TILE_FLOATS];   // one 32 KB A tile per tile of the CTA (1..4, sized at launch), 1024-byte aligned
};

struct DevPolicyTc {
    const float *w0, *b0, *w1, *b1, *wo, *bo;
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// Round-to-nearest (ties away) TF32: cvt.rna.tf32 is "add half an ulp of the 10-bit mantissa, clear the low 13 bits"; the
// tensor core ignores those 13 bits of a kind::tf32 operand, so the add alone feeds it the same operand (values here are finite
// and far from overflow: observations in [-1, 1], tanh outputs, trained weights).
__device__ __forceinline__ float to_tf32(float x) { return __uint_as_float(__float_as_uint(x) + 0x1000u); }
__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// float offset of element (row, k) inside a K-major SWIZZLE_128B operand whose K chunks hold `rows` rows each
__device__ __forceinline__ int sw128_offset(int row, int k, int chunk_floats) {
    const int chunk = k >> 5, kk = k & 31;
    return chunk * chunk_floats + row * 32 + ((((kk >> 2) ^ (row & 7)) << 2) | (kk & 3));
}

// UMMA shared-memory descriptor, K-major, SWIZZLE_128B: start>>4 | LBO(=1)<<16 | SBO(=1024 B>>4)<<32 | version 1<<46 | layout 2<<61
__device__ __forceinline__ unsigned long long umma_desc(unsigned saddr) {
    return (unsigned long long)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D fp32, A/B tf32, both K-major, N>>3 at bit 17, M>>4 at bit 24
__host__ __device__ constexpr unsigned umma_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned mbar_saddr) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar_saddr) : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned saddr, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(saddr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned saddr, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}"
        ::"r"(saddr), "r"(parity) : "memory");
}
// one lane of a converged warp (the MMA issue is then compiled with uniform-register descriptors instead of per-MMA R2UR chains)
__device__ __forceinline__ bool elect_one_tc() {
    unsigned pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0u;
}
__device__ __forceinline__ void tile_barrier(int tile, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(tile + 1), "r"(threads) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// this thread's 32 consecutive accumulator columns [col0, col0+32) of its own TMEM lane
__device__ __forceinline__ void tmem_ld32(unsigned taddr, float* v) {
    unsigned r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(unsigned taddr, float* v) {
    unsigned r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// stage one policy's actor into the swizzled B-operand images (all threads of the CTA)
__device__ __forceinline__ void load_weights_tc(TcSmem& S, const DevPolicyTc& p, int tid, int nthreads) {
    for (int i = tid; i < TC_HID * TC_K; i += nthreads) {
        const int n = i >> 6, k = i & 63;
        float v0 = k < KIN_OBS_DIM ? __ldg(p.w0 + n * KIN_OBS_DIM + k) : (k == KIN_OBS_DIM ? __ldg(p.b0 + n) : 0.0f);
        S.W0[sw128_offset(n, k, CHUNK_FLOATS_W)] = to_tf32(v0);
        S.W1[sw128_offset(n, k, CHUNK_FLOATS_W)] = to_tf32(__ldg(p.w1 + n * TC_HID + k));
    }
    for (int i = tid; i < 8 * TC_K; i += nthreads) {
        const int n = i >> 6, k = i & 63;
        S.WO[sw128_offset(n, k, CHUNK_FLOATS_WO)] = n < KIN_NJ ? to_tf32(__ldg(p.wo + n * TC_HID + k)) : 0.0f;
    }
    for (int i = tid; i < TC_HID; i += nthreads) S.b1[i] = __ldg(p.b1 + i);     // a CTA may be a single warp
    if (tid < 8) S.bo[tid] = tid < KIN_NJ ? __ldg(p.bo + tid) : 0.0f;
}

// write 4 consecutive K elements [k4*4, k4*4+4) of this thread's A row (already TF32-rounded)
__device__ __forceinline__ void a_store4(float* A, int row, int k4, float x0, float x1, float x2, float x3) {
    const int chunk = k4 >> 3, j = k4 & 7;
    float4* dst = reinterpret_cast<float4*>(A + chunk * CHUNK_FLOATS_A + row * 32 + ((j ^ (row & 7)) << 2));
    *dst = make_float4(x0, x1, x2, x3);
}

// one GEMM of the tile: D[128 x N] (TMEM) = A[128 x 64] (smem) * W[N x 64]^T (smem); issued by one thread
__device__ __forceinline__ void issue_layer(unsigned a_saddr, unsigned w_saddr, int w_chunk_bytes, unsigned tmem_d, unsigned idesc, unsigned mbar_saddr) {
    tc_fence_after();
#pragma unroll
    for (int k = 0; k < TC_K / 8; ++k) {
        const unsigned a_off = (k >> 2) * (CHUNK_FLOATS_A * 4) + (k & 3) * 32;
        const unsigned w_off = (k >> 2) * w_chunk_bytes + (k & 3) * 32;
        umma_tf32(tmem_d, umma_desc(a_saddr + a_off), umma_desc(w_saddr + w_off), idesc, k > 0 ? 1u : 0u);
    }
    umma_commit(mbar_saddr);
}

}  // namespace kin
