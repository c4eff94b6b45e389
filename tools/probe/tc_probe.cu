// tc_probe.cu -- de-risks the tcgen05 plumbing the tensor-core render kernel needs (round 2 experiment):
//   1. D1[128][N] = A[128][K] * B[N][K]^T, kind::tf32, A and B from shared memory (K-major, no swizzle: [K/4][rows][4 floats]),
//      accumulators in TMEM, read back with tcgen05.ld 32x32b;
//   2. D2[128][N2] = T[128][N] * W[N2][N]^T with T = max(D1, 0) written back to TMEM (tcgen05.st) and used as the A operand.
// Values are small integers (exact in tf32), so the results must match the host bit for bit.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_probe tc_probe.cu ; run on a B200.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128, K1 = 16, N1 = 32, N2 = 16;

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(const void *base, unsigned lbo_bytes, unsigned sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(base) >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46; // descriptor version (Blackwell)
    return d;               // base offset 0, LBO mode 0, layout type 0 = no swizzle
}
__host__ __device__ constexpr uint32_t make_idesc(int m, int n)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); // F32 acc, TF32 x TF32, K-major both
}
__device__ __forceinline__ void mma_ss(unsigned d_tmem, uint64_t a, uint64_t b, uint32_t idesc, unsigned acc)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                 :: "r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(unsigned d_tmem, unsigned a_tmem, uint64_t b, uint32_t idesc, unsigned acc)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n"
                 :: "r"(d_tmem), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(unsigned long long *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool wait_bounded(unsigned long long *bar, unsigned parity)
{
    for (unsigned spins = 0; spins < (1u << 22); ++spins) {
        unsigned done;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) return true;
    }
    return false;
}
__device__ __forceinline__ void tmem_ld16(unsigned taddr, float (&v)[16])
{
    unsigned r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st16(unsigned taddr, const float (&v)[16])
{
    unsigned r[16];
    for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(v[i]);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                    "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// layout of an operand with R rows and K columns: [K/4][R][4 floats]
__host__ __device__ inline int opnd(int R, int r, int k) { return (k >> 2) * (R * 4) + r * 4 + (k & 3); }

__global__ void __launch_bounds__(128, 1) probe(const float *A, const float *B, const float *W, float *D1, float *D2, int *status)
{
    __shared__ __align__(128) float sA[M * K1], sB[N1 * K1], sW[N2 * N1];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ unsigned tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < M * K1; i += 128) sA[opnd(M, i / K1, i % K1)] = A[i];
    for (int i = tid; i < N1 * K1; i += 128) sB[opnd(N1, i / K1, i % K1)] = B[i];
    for (int i = tid; i < N2 * N1; i += 128) sW[opnd(N2, i / N1, i % N1)] = W[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic-proxy smem writes -> visible to the tensor core (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tb = tmem_base_s;
    const unsigned d1 = tb, ta = tb + 32, d2 = tb + 64; // columns: D1 [0,32), T [32,64), D2 [64,80)

    if (tid == 0) {
        for (int ks = 0; ks < K1 / 8; ++ks) {
            const uint64_t da = make_desc(sA + ks * 2 * (M * 4), M * 16, 128);
            const uint64_t db = make_desc(sB + ks * 2 * (N1 * 4), N1 * 16, 128);
            mma_ss(d1, da, db, make_idesc(M, N1), ks > 0);
        }
        commit(&bar);
    }
    bool ok = wait_bounded(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (!ok) { if (tid == 0) status[0] = 1; }
    const unsigned lane_base = (unsigned)(warp * 32) << 16;
    if (ok) {
        for (int c = 0; c < N1; c += 16) {
            float v[16];
            tmem_ld16(d1 + lane_base + c, v);
            for (int i = 0; i < 16; ++i) D1[tid * N1 + c + i] = v[i];
            for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.0f);
            tmem_st16(ta + lane_base + c, v);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (ok && tid == 0) {
        for (int ks = 0; ks < N1 / 8; ++ks) {
            const uint64_t dw = make_desc(sW + ks * 2 * (N2 * 4), N2 * 16, 128);
            mma_ts(d2, ta + ks * 8, dw, make_idesc(M, N2), ks > 0);
        }
        commit(&bar);
    }
    if (ok) {
        ok = wait_bounded(&bar, 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (!ok) { if (tid == 0) status[0] = 2; }
    }
    if (ok) {
        float v[16];
        tmem_ld16(d2 + lane_base, v);
        for (int i = 0; i < 16; ++i) D2[tid * N2 + i] = v[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tb), "r"(128u) : "memory");
}

int main()
{
    std::vector<float> A(M * K1), B(N1 * K1), W(N2 * N1), D1(M * N1), D2(M * N2), R1(M * N1), R2(M * N2);
    unsigned s = 7;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (float)((int)((s >> 20) % 9) - 4); };
    for (auto &x : A) x = rnd();
    for (auto &x : B) x = rnd();
    for (auto &x : W) x = rnd();
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N1; ++n) { float a = 0; for (int k = 0; k < K1; ++k) a += A[m * K1 + k] * B[n * K1 + k]; R1[m * N1 + n] = a; }
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N2; ++n) { float a = 0; for (int k = 0; k < N1; ++k) a += (R1[m * N1 + k] > 0 ? R1[m * N1 + k] : 0) * W[n * N1 + k]; R2[m * N2 + n] = a; }
    float *dA, *dB, *dW, *dD1, *dD2; int *dS;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dW, W.size() * 4);
    cudaMalloc(&dD1, D1.size() * 4); cudaMalloc(&dD2, D2.size() * 4); cudaMalloc(&dS, 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD1, 0xff, D1.size() * 4); cudaMemset(dD2, 0xff, D2.size() * 4); cudaMemset(dS, 0, 4);
    probe<<<1, 128>>>(dA, dB, dW, dD1, dD2, dS);
    cudaError_t e = cudaDeviceSynchronize();
    int st = -1;
    cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(D1.data(), dD1, D1.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(D2.data(), dD2, D2.size() * 4, cudaMemcpyDeviceToHost);
    double e1 = 0, e2 = 0;
    for (size_t i = 0; i < D1.size(); ++i) e1 = std::max(e1, (double)fabsf(D1[i] - R1[i]));
    for (size_t i = 0; i < D2.size(); ++i) e2 = std::max(e2, (double)fabsf(D2[i] - R2[i]));
    printf("tc_probe: cuda %s, status %d, max |D1 - ref| = %g, max |D2 - ref| = %g  (D1[0..3] = %g %g %g %g, ref %g %g %g %g)\n", cudaGetErrorString(e), st, e1,
           e2, D1[0], D1[1], D1[2], D1[3], R1[0], R1[1], R1[2], R1[3]);
    printf("          D1 row 1: %g %g %g %g | ref %g %g %g %g ; D1 row 9: %g %g | ref %g %g\n", D1[N1], D1[N1 + 1], D1[N1 + 2], D1[N1 + 3], R1[N1], R1[N1 + 1],
           R1[N1 + 2], R1[N1 + 3], D1[9 * N1], D1[9 * N1 + 1], R1[9 * N1], R1[9 * N1 + 1]);
    return (e == cudaSuccess && st == 0 && e1 == 0 && e2 == 0) ? 0 : 1;
}
