// Micro-test: tcgen05.mma kind::f16 (bf16) with MN-major A and B taken from K-major-style tiles ([128 rows][64 bf16 = 128 B], SWIZZLE_128B):
// D[64 x N] = A^T B, A[128 k][64 m], B[128 k][N<=64 n].  Checks the M = 64 TMEM lane map (row m -> lane m%16 + 32*(m/16)).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long mk_desc(unsigned saddr, unsigned lbo_bytes, unsigned sbo_bytes) {
    return (unsigned long long)((saddr >> 4) & 0x3FFFu) | ((unsigned long long)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((unsigned long long)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (2ull << 61);
}
__global__ void k(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, int M, int N) {
    extern __shared__ unsigned char raw[];
    __nv_bfloat16* sm = (__nv_bfloat16*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __nv_bfloat16* sA = sm;            // [128 rows][64] bf16, 16 KB
    __nv_bfloat16* sB = sm + 8192;
    __shared__ unsigned long long mbar;
    __shared__ unsigned tbase;
    int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 128 * 64; i += 128) {
        int r = i / 64, c = i % 64;
        int off = r * 64 + ((((c >> 3) ^ (r & 7)) << 3) | (c & 7));   // 16-byte units of 8 bf16
        sA[off] = A[i];
        sB[off] = B[i];
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    unsigned tb = tbase;
    if (tid == 0) {
        // kind::f16: c fp32 (1<<4), a/b bf16 (1<<7, 1<<10), both MN-major (bits 15, 16)
        unsigned idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
        for (int kk = 0; kk < 8; ++kk) {   // K = 128 rows, 16 per MMA = two 1024-byte atoms
            unsigned long long da = mk_desc(smem_u32(sA) + kk * 2048, 16384, 1024), db = mk_desc(smem_u32(sB) + kk * 2048, 16384, 1024);
            unsigned acc = kk > 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tb), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    asm volatile("{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n\t@P1 bra DN;\n\tbra W;\n\tDN:\n\t}" ::"r"(smem_u32(&mbar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 8) {
        unsigned r[8];
        unsigned taddr = tb + ((unsigned)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; ++j) D[tid * N + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(64) : "memory");
}
int main() {
    for (int cfg = 0; cfg < 3; ++cfg) {
        int M = cfg == 2 ? 128 : 64, N = cfg == 1 ? 8 : 64;
        if (M == 128) continue;
        std::vector<__nv_bfloat16> A(128 * 64), B(128 * 64);
        std::vector<float> Af(128 * 64), Bf(128 * 64), D(128 * 64, 0.f);
        srand(1 + cfg);
        for (size_t i = 0; i < A.size(); ++i) {
            A[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.0f); Af[i] = __bfloat162float(A[i]);
            B[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.0f); Bf[i] = __bfloat162float(B[i]);
        }
        __nv_bfloat16 *dA, *dB; float* dD;
        cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, D.size() * 4);
        cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
        cudaMemset(dD, 0, D.size() * 4);
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
        k<<<1, 128, 40000>>>(dA, dB, dD, M, N);
        cudaError_t e = cudaDeviceSynchronize();
        printf("bf16 MN-major M=%d N=%d: %s\n", M, N, cudaGetErrorString(e));
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0; int bad = 0;
        for (int m = 0; m < M; ++m) {
            int lane = (m % 16) + 32 * (m / 16);
            for (int n = 0; n < N; ++n) {
                double ref = 0;
                for (int kk = 0; kk < 128; ++kk) ref += (double)Af[kk * 64 + m] * Bf[kk * 64 + n];
                double err = fabs(ref - D[lane * N + n]);
                if (err > maxerr) maxerr = err;
                if (err > 1e-3) ++bad;
            }
        }
        printf("  max err %.3e, bad %d of %d (D[0][0..2] = %.4f %.4f %.4f)\n", maxerr, bad, M * N, D[0], D[1], D[2]);
        cudaFree(dA); cudaFree(dB); cudaFree(dD);
    }
    return 0;
}
