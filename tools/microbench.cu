// Pipe-rate micro-benchmarks that fix the integer roofline of the scoring pass
// (POPC is not listed in the B300 notes, "measure it on the box").
// Each kernel runs register-resident dependent chains; reported as lane-ops/clk/SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define UNROLL 8

template <int OP>
__global__ void k(uint32_t* out, uint32_t seed, long long* cycles) {
    uint32_t a[UNROLL];
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) a[i] = seed + threadIdx.x * 7 + i * 13 + blockIdx.x;
    uint32_t b = seed ^ threadIdx.x, c = seed * 3 + 1;
    float fa[UNROLL]; double da[UNROLL];
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) { fa[i] = 1.0f + a[i] * 1e-9f; da[i] = 1.0 + a[i] * 1e-12; }
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < UNROLL; ++i) {
            if (OP == 0) a[i] = __popc(a[i]) + b;                       // POPC + IADD
            if (OP == 1) a[i] = (a[i] & b) ^ c;                         // LOP3
            if (OP == 2) a[i] = a[i] + b;                               // IADD
            if (OP == 3) a[i] = __popc(a[i] & b) + a[i];                // AND+POPC+IADD (naive inner loop)
            if (OP == 4) a[i] = __match_any_sync(0xffffffffu, a[i] & 7u) + a[i];
            if (OP == 5) a[i] = __reduce_or_sync(0xffffffffu, a[i]) + 1;
            if (OP == 6) fa[i] = __fdiv_rn(fa[i], 1.0000001f);
            if (OP == 7) da[i] = __dadd_rn(__dmul_rn(da[i], 1.0000000001), 1e-9);
            if (OP == 8) da[i] = __ddiv_rn(da[i], 1.0000000001);
            if (OP == 9) fa[i] = __fadd_rn(__fmul_rn(fa[i], 1.0000001f), 1e-9f);
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) s += a[i] + (uint32_t)fa[i] + (uint32_t)da[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

__global__ void k_atoms(uint32_t* out, int spread, long long* cycles) {
    __shared__ uint32_t sm[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = 0;
    __syncthreads();
    uint32_t x = threadIdx.x * 2654435761u;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
        x = x * 1664525u + 1013904223u;
        const uint32_t idx = spread ? (x >> 8) & 8191u : ((threadIdx.x >> 5) * 32 + (x >> 27));
        atomicOr(&sm[idx], 1u << (x & 31));
    }
    long long t1 = clock64();
    __syncthreads();
    out[blockIdx.x * blockDim.x + threadIdx.x] = sm[threadIdx.x];
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int OP>
void run(const char* name, double ops_per_iter) {
    uint32_t* out; long long* cyc; long long h;
    const int threads = 1024, blocks = 148 * 2;
    cudaMalloc(&out, blocks * threads * 4); cudaMalloc(&cyc, 8);
    k<OP><<<blocks, threads>>>(out, 12345u, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP><<<blocks, threads>>>(out, 12345u, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    // 2 blocks of 1024 threads per SM run concurrently: lane-ops per SM over the block's cycles
    double lane_ops_per_sm = 2.0 * threads * (double)ITERS * UNROLL * ops_per_iter;
    printf("%-34s %8.3f ms  block cycles %10lld  -> %7.2f lane-ops/clk/SM (chip %.2f Tops/s)\n", name, ms, h,
           lane_ops_per_sm / (double)h, 148.0 * lane_ops_per_sm / (ms * 1e-3) / 1e12);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("device %s, %d SMs, clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    run<0>("POPC+IADD", 1);
    run<1>("LOP3", 1);
    run<2>("IADD", 1);
    run<3>("AND+POPC+IADD (per pair-word)", 1);
    run<4>("MATCH.ANY (+IADD)", 1);
    run<5>("REDUX.OR (+IADD)", 1);
    run<6>("FDIV.RN fp32", 1);
    run<7>("DMUL+DADD fp64 (2 ops)", 2);
    run<8>("DDIV.RN fp64", 1);
    run<9>("FMUL+FADD fp32 (2 ops)", 2);
    for (int spread = 0; spread < 2; ++spread) {
        uint32_t* out; long long* cyc; long long h;
        cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
        k_atoms<<<148, 1024>>>(out, spread, cyc); cudaDeviceSynchronize();
        k_atoms<<<148, 1024>>>(out, spread, cyc); cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("smem atomicOr %-20s block cycles %10lld -> %6.2f cyc/warp-instr/SM\n",
               spread ? "(random 32 KB)" : "(warp-local 128 B)", h, (double)h / (ITERS * 32.0));
        cudaFree(out); cudaFree(cyc);
    }
    return 0;
}
