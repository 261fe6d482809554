// Issue-rate micro-benchmarks of the instructions k_project's inner loop is made of (B200, sm_100a).
// Every op is an `asm volatile` statement on runtime operands, so nothing is folded, merged or
// eliminated; each thread keeps 8 independent chains (ILP 8), one CTA of 1024 threads per SM
// (8 warps per scheduler).  Reported: cycles per warp-instruction per SM sub-partition (1.0 = the
// issue limit), from clock64 of one CTA.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_build/pipes_bench tools/pipes_bench.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define ITERS 2048
#define NCH 8

enum Op { FFMA, FFMA2, FADD, FMUL, IMAD, LOP3, IADD3, SHF, FSETP_SEL, FMNMX, MUFU, RED_SMEM, RED_SMEM_PRED_OFF,
          MIX_FFMA_LOP3, MIX_FFMA_FADD, MIX_FFMA_IMAD, MIX_3FFMA_1LOP3, FFMA_SAT, FFMA_IMM, FSETP_ONLY, MIX_FFMA2_LOP3,
          MIX_FFMA_MUFU8, POPC, F2I, I2F, IMNMX, MUFU_DEP, FSETP_CHAIN, N_OPS };

template <int OP>
__global__ void __launch_bounds__(1024, 1) k(uint32_t* out, float fb, float fc, uint32_t ub, uint32_t uc,
                                              unsigned long long* t_first, unsigned long long* t_last) {
    __shared__ uint32_t sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 0;
    __syncthreads();
    float a[NCH];
    uint32_t u[NCH];
    unsigned long long w[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        a[i] = 1.0f + (threadIdx.x + i) * 1e-6f;
        u[i] = threadIdx.x * 97u + i;
        w[i] = ((unsigned long long)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i]);
    }
    const unsigned long long wb = ((unsigned long long)__float_as_uint(fb) << 32) | __float_as_uint(fb);
    const unsigned long long wc = ((unsigned long long)__float_as_uint(fc) << 32) | __float_as_uint(fc);
    const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(sm) + 4u * ((threadIdx.x * 33u) & 2047u);   // + 128 * i stays inside
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(fb), "f"(fc));
            if (OP == FFMA_SAT) asm volatile("fma.rn.sat.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(fb), "f"(fc));
            if (OP == FFMA_IMM) asm volatile("fma.rn.f32 %0, %0, %1, 0f4B400000;" : "+f"(a[i]) : "f"(fb));
            if (OP == FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(w[i]) : "l"(wb), "l"(wc));
            if (OP == FADD) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(fc));
            if (OP == FMUL) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(fb));
            if (OP == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(ub), "r"(uc));
            if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(ub), "r"(uc));
            if (OP == IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(ub));
            if (OP == SHF) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(ub), "r"(uc));
            if (OP == FSETP_SEL)
                asm volatile("{.reg .pred p; setp.gt.f32 p, %0, %1; selp.f32 %0, %2, %0, p;}" : "+f"(a[i]) : "f"(fb), "f"(fc));
            if (OP == FSETP_ONLY)
                asm volatile("{.reg .pred p; setp.gt.f32 p, %1, %2; @p add.u32 %0, %0, 1;}" : "+r"(u[i]) : "f"(a[i]), "f"(fb));
            if (OP == FMNMX) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(fb));
            if (OP == MUFU) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (OP == MUFU_DEP) {                 // rcp of a moving value (a plain rcp chain has a two-cycle fixed point)
                asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
                asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(fb));
            }
            if (OP == POPC) asm volatile("popc.b32 %0, %0;" : "+r"(u[i]));
            if (OP == F2I) asm volatile("{.reg .s32 t; cvt.rzi.s32.f32 t, %1; add.s32 %0, %0, t;}" : "+r"(u[i]) : "f"(a[i]));
            if (OP == I2F) asm volatile("{.reg .f32 t; cvt.rn.f32.s32 t, %1; add.rn.f32 %0, %0, t;}" : "+f"(a[i]) : "r"(u[i]));
            if (OP == IMNMX) asm volatile("min.s32 %0, %0, %1;" : "+r"(u[i]) : "r"(ub));
            if (OP == FSETP_CHAIN)                // the compare depends on the moving value: it cannot be hoisted
                asm volatile("{.reg .pred p; setp.gt.f32 p, %0, %1; @p add.rn.f32 %0, %0, %2;}" : "+f"(a[i]) : "f"(fb), "f"(fc));
            if (OP == RED_SMEM) asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(saddr + 128u * i), "r"(u[i]) : "memory");
            if (OP == RED_SMEM_PRED_OFF)
                asm volatile("{.reg .pred p; setp.eq.u32 p, %2, 12345; @p red.shared.or.b32 [%0], %1;}" ::"r"(saddr + 128u * i), "r"(u[i]), "r"(ub) : "memory");
            if (OP == MIX_FFMA_LOP3) {
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(fb), "f"(fc));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(ub), "r"(uc));
            }
            if (OP == MIX_FFMA2_LOP3) {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(w[i]) : "l"(wb), "l"(wc));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(ub), "r"(uc));
            }
            if (OP == MIX_FFMA_FADD) {
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(fb), "f"(fc));
                asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(fc));
            }
            if (OP == MIX_FFMA_IMAD) {
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(fb), "f"(fc));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(ub), "r"(uc));
            }
            if (OP == MIX_3FFMA_1LOP3) {
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(fb), "f"(fc));
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(fc), "f"(fb));
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(fb), "f"(fc));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(ub), "r"(uc));
            }
            if (OP == MIX_FFMA_MUFU8) {          // 8 FFMA per MUFU (k_project's ratio is ~13:1)
#pragma unroll
                for (int q = 0; q < 8; ++q) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(fb), "f"(fc));
                asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            }
        }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) s += u[i] + __float_as_uint(a[i]) + (uint32_t)(w[i] >> 32) + (uint32_t)w[i];
    __syncthreads();
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + sm[threadIdx.x];
    // the CTA's span: first warp to start .. last warp to finish (a single warp's clock favours whichever
    // warp the scheduler prefers)
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) {
        atomicMin(t_first, (unsigned long long)t0);
        atomicMax(t_last, (unsigned long long)t1);
    }
}

template <int OP>
void run(const char* name, int instr_per_chain_step) {
    uint32_t* out;
    unsigned long long *cyc, hv[2];
    long long h = 0;
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaMalloc(&out, (size_t)sms * 1024 * 4);
    cudaMalloc(&cyc, 16);
    for (int rep = 0; rep < 2; ++rep) {
        hv[0] = ~0ull; hv[1] = 0ull;
        cudaMemcpy(cyc, hv, 16, cudaMemcpyHostToDevice);
        k<OP><<<sms, 1024>>>(out, 1.0000001f, 1e-9f, 0x9e3779b9u, 0x7f4a7c15u, cyc, cyc + 1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%-28s FAILED: %s\n", name, cudaGetErrorString(e)); exit(1); }
        cudaMemcpy(hv, cyc, 16, cudaMemcpyDeviceToHost);
        h = (long long)(hv[1] - hv[0]);
    }
    const double winstr_per_smsp = 8.0 * ITERS * NCH * instr_per_chain_step;     // 8 warps per scheduler
    printf("%-28s %2d instr/step  %10lld cycles  -> %6.3f cycles per warp-instr per SMSP  (%5.1f lane-instr/clk/SM)\n", name,
           instr_per_chain_step, h, (double)h / winstr_per_smsp, 128.0 * winstr_per_smsp / (double)h);
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("device %s, %d SMs\n", p.name, p.multiProcessorCount);
    run<FFMA>("FFMA (3 reg)", 1);
    run<FFMA_SAT>("FFMA.SAT", 1);
    run<FFMA_IMM>("FFMA (imm addend)", 1);
    run<FFMA2>("FFMA2 (f32x2)", 1);
    run<FADD>("FADD", 1);
    run<FMUL>("FMUL", 1);
    run<IMAD>("IMAD", 1);
    run<LOP3>("LOP3", 1);
    run<IADD3>("IADD3", 1);
    run<SHF>("SHF.L.W", 1);
    run<FSETP_SEL>("FSETP + SEL", 2);
    run<FSETP_ONLY>("FSETP + @p IADD", 2);
    run<FMNMX>("FMNMX", 1);
    run<MUFU_DEP>("MUFU.RCP + FADD", 2);
    run<POPC>("POPC", 1);
    run<F2I>("F2I.TRUNC + IADD3", 2);
    run<I2F>("I2F + FADD", 2);
    run<IMNMX>("IMNMX (min.s32)", 1);
    run<FSETP_CHAIN>("FSETP + @p FADD", 2);
    run<RED_SMEM>("RED.shared.or (no conflict)", 1);
    run<RED_SMEM_PRED_OFF>("ISETP + @!p RED.shared", 2);
    run<MIX_FFMA_LOP3>("FFMA + LOP3", 2);
    run<MIX_FFMA2_LOP3>("FFMA2 + LOP3", 2);
    run<MIX_FFMA_FADD>("FFMA + FADD", 2);
    run<MIX_FFMA_IMAD>("FFMA + IMAD", 2);
    run<MIX_3FFMA_1LOP3>("3 FFMA + LOP3", 4);
    run<MIX_FFMA_MUFU8>("8 FFMA + MUFU.RCP", 9);
    return 0;
}
