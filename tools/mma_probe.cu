// Stand-alone probe of tcgen05.mma kind::i8 (u8 x u8 -> s32) with hand-built shared-memory
// descriptors in the no-swizzle K-major ("interleaved core matrix") layout, accumulator in TMEM,
// read back with tcgen05.ld.32x32b.  Used to pin the descriptor fields before k_score_mma was
// written.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o mma_probe mma_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

// A: [128][K] bytes row-major, B: [N][K] bytes row-major, D: [128][N] s32
__global__ void __launch_bounds__(128) probe(const uint8_t* A, const uint8_t* B, int N, int K, int swap, int32_t* D) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_base;
    __shared__ __align__(8) uint64_t bar;
    const int nk = K / 32;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* sA = smem;
    uint8_t* sB = smem + (size_t)nk * 2 * 128 * 16;
    // layout per K=32 step: [kc(2)][rows][16 bytes]
    for (int i = tid; i < 128 * K; i += 128) {
        const int m = i / K, k = i % K;
        sA[(size_t)(k / 32) * (2 * 128 * 16) + ((k % 32) / 16) * (128 * 16) + m * 16 + (k % 16)] = A[i];
    }
    for (int i = tid; i < N * K; i += 128) {
        const int m = i / K, k = i % K;
        sB[(size_t)(k / 32) * (2 * N * 16) + ((k % 32) / 16) * (N * 16) + m * 16 + (k % 16)] = B[i];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        mbar_init(smem_u32(&bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = (2u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        for (int k = 0; k < nk; ++k) {
            const uint32_t a0 = smem_u32(sA + (size_t)k * 2 * 128 * 16), b0 = smem_u32(sB + (size_t)k * 2 * N * 16);
            const uint32_t lboA = 128 * 16, lboB = N * 16, sbo = 128;
            const uint64_t da = swap ? make_desc(a0, sbo, lboA) : make_desc(a0, lboA, sbo);
            const uint64_t db = swap ? make_desc(b0, sbo, lboB) : make_desc(b0, lboB, sbo);
            const uint32_t acc = k > 0 ? 1u : 0u;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait(smem_u32(&bar), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t r[16];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                       "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) D[(size_t)(warp * 32 + lane) * N + c0 + j] = (int32_t)r[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

int main() {
    const int Ns[] = {16, 96, 192, 240};
    const int K = 128;
    int bad_total = 0;
    for (int swap = 0; swap < 2; ++swap)
        for (int N : Ns) {
            std::vector<uint8_t> A(128 * K), B(N * K);
            uint32_t s = 12345u + N;
            auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (s >> 16) & 1u; };
            for (auto& v : A) v = (uint8_t)rnd();
            for (auto& v : B) v = (uint8_t)rnd();
            std::vector<int32_t> ref(128 * N, 0), out(128 * N, -1);
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < N; ++n) {
                    int acc = 0;
                    for (int k = 0; k < K; ++k) acc += A[m * K + k] * B[n * K + k];
                    ref[m * N + n] = acc;
                }
            uint8_t *dA, *dB; int32_t* dD;
            CK(cudaMalloc(&dA, A.size())); CK(cudaMalloc(&dB, B.size())); CK(cudaMalloc(&dD, out.size() * 4));
            CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice));
            CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
            CK(cudaMemset(dD, 0xff, out.size() * 4));
            const size_t smem = (size_t)(K / 32) * 2 * 16 * (128 + N);
            CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            probe<<<1, 128, smem>>>(dA, dB, N, K, swap, dD);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("swap=%d N=%d: kernel error %s\n", swap, N, cudaGetErrorString(e)); return 3; }
            CK(cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost));
            int bad = 0;
            for (size_t i = 0; i < out.size(); ++i) bad += out[i] != ref[i];
            printf("swap=%d N=%d: %d / %zu mismatches; D[0][0..3] = %d %d %d %d (ref %d %d %d %d); D[1][0]=%d (ref %d) D[127][N-1]=%d (ref %d)\n",
                   swap, N, bad, out.size(), out[0], out[1], out[2], out[3], ref[0], ref[1], ref[2], ref[3],
                   out[N], ref[N], out[127 * N + N - 1], ref[127 * N + N - 1]);
            if (swap == 0) bad_total += bad;
            cudaFree(dA); cudaFree(dB); cudaFree(dD);
        }
    printf(bad_total == 0 ? "PROBE OK (lbo = K-chunk stride, sbo = 8-row group stride)\n" : "PROBE MISMATCH\n");
    return 0;
}
