// Micro-test: tcgen05.mma kind::tf32 with MN-major A and B (both stored [k rows][mn contiguous], SWIZZLE_128B) and M = 64.
// D[64 x 64] = A^T B with A[128 k][64 m], B[128 k][64 n]; prints max error vs a host reference and the TMEM lane map.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long desc_mn(unsigned saddr, unsigned lbo_bytes) {
    return (unsigned long long)((saddr >> 4) & 0x3FFFu) | ((unsigned long long)((lbo_bytes >> 4) & 0x3FFFu) << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__global__ void k(const float* A, const float* B, float* D, int M, int N, int mode) {
    extern __shared__ unsigned char raw[];
    float* sm = (float*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    float* sA = sm;                 // [2 chunks][128 rows][32]
    float* sB = sm + 2 * 4096;
    __shared__ unsigned long long mbar;
    __shared__ unsigned tbase;
    int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 128 * 64; i += 128) {
        int r = i / 64, c = i % 64;
        int off = (c >> 5) * 4096 + r * 32 + (((((c & 31) >> 2) ^ (r & 7)) << 2) | (c & 3));
        sA[off] = A[i];
        sB[off] = B[i];
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)) : "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    unsigned tb = tbase;
    if (tid == 0) {
        unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((mode & 1) ? (1u << 15) : 0u) | ((mode & 2) ? (1u << 16) : 0u) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
        int nk = (mode == 0) ? 8 : 16;
        for (int kk = 0; kk < nk; ++kk) {   // MN-major: K = 128 rows, 8 per MMA = one 1024-byte atom; K-major: K = 64 cols, 8 per MMA = 32 bytes
            unsigned koff_k = (kk >> 2) * 16384 + (kk & 3) * 32;
            unsigned long long da = (mode & 1) ? desc_mn(smem_u32(sA) + kk * 1024, 16384) : desc_mn(smem_u32(sA) + koff_k, 16);
            unsigned long long db = (mode & 2) ? desc_mn(smem_u32(sB) + kk * 1024, 16384) : desc_mn(smem_u32(sB) + koff_k, 16);
            unsigned acc = kk > 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tb), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
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
        for (int j = 0; j < 8; ++j) D[tid * N + c0 + j] = __uint_as_float(r[j]);   // D indexed by TMEM lane
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(64) : "memory");
}
int main() {
    for (int cfg = 0; cfg < 5; ++cfg) {
        // cfg 0: K-major both, M=128 N=64 (known-good shape of the rollout kernel); 1: A MN-major only (M=64, B K-major N=64 K=128?? skipped)
        int mode = cfg == 0 ? 0 : 3;
        int M = cfg == 0 ? 128 : 64, N = (cfg == 2) ? 8 : 64;
        if (cfg == 3) { M = 128; }       // MN-major with M = 128 would need 128 m-columns: skip
        if (cfg == 3 || cfg == 4) continue;
        int Mcols = 64;
        std::vector<float> A(128 * 64), B(128 * 64), D(128 * 64, 0.f);
        srand(1 + cfg);
        auto tf = [](float x) { unsigned u; memcpy(&u, &x, 4); u &= 0xffffe000u; float y; memcpy(&y, &u, 4); return y; };
        for (auto& v : A) v = tf((rand() % 2001 - 1000) / 1000.0f);
        for (auto& v : B) v = tf((rand() % 2001 - 1000) / 1000.0f);
        float *dA, *dB, *dD;
        cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
        cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
        cudaMemset(dD, 0, D.size() * 4);
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
        k<<<1, 128, 70000>>>(dA, dB, dD, M, N, mode);
        cudaError_t e = cudaDeviceSynchronize();
        printf("cfg M=%d N=%d: %s\n", M, N, cudaGetErrorString(e));
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        // reference and lane map: row m expected at lane (m % 16) + 32 * (m / 16)
        double maxerr = 0; int bad = 0;
        for (int m = 0; m < M; ++m) {
            int lane = (M == 64) ? (m % 16) + 32 * (m / 16) : m;
            for (int n = 0; n < N; ++n) {
                double ref = 0;
                if (mode == 0) { for (int kk = 0; kk < 64; ++kk) ref += (double)A[m * 64 + kk] * B[n * 64 + kk]; }
                else { for (int kk = 0; kk < 128; ++kk) ref += (double)A[kk * 64 + m] * B[kk * 64 + n]; }
                double err = fabs(ref - D[lane * N + n]);
                if (err > maxerr) maxerr = err;
                if (err > 1e-3) ++bad;
            }
        }
        printf("  max err %.3e, bad %d of %d  (D[lane0][0..3] = %.4f %.4f %.4f %.4f)\n", maxerr, bad, M * N, D[0], D[1], D[2], D[3]);
        // where did the data go?  nonzeros per lane, and search for ref(0,0), ref(1,0), ref(0,1), ref(17,3)
        int nzl = 0; for (int l = 0; l < 128; ++l) { int nz = 0; for (int n = 0; n < N; ++n) nz += D[l * N + n] != 0.f; if (nz) { if (nzl < 12) printf("  lane %d: %d nonzero, first %.4f\n", l, nz, D[l * N]); ++nzl; } }
        printf("  lanes with data: %d\n", nzl);
        int probes[4][2] = {{0, 0}, {1, 0}, {0, 1}, {17, 3}};
        for (auto& pr : probes) {
            double ref = 0; for (int kk = 0; kk < 128; ++kk) ref += (double)A[kk * 64 + pr[0]] * B[kk * 64 + pr[1]];
            for (int i = 0; i < 128 * N; ++i) if (fabs(D[i] - ref) < 2e-4) printf("  ref(%d,%d)=%.4f found at lane %d col %d\n", pr[0], pr[1], ref, i / N, i % N);
        }
        (void)Mcols;
        cudaFree(dA); cudaFree(dB); cudaFree(dD);
    }
    return 0;
}
