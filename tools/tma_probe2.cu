// usage: tma_probe2 <mode>   0: 1-D bulk copy   1: TMA 2-D on a [1024][1024] f32 tensor, box 32x32
//                            2: TMA 2-D on [4800][20] u32, box 8x32    3: same as 2 but tensor rows padded to 32 words
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k_bulk(const uint32_t* src, int bytes, uint32_t* out) {
    extern __shared__ __align__(1024) uint32_t tile[];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t b = smem_u32(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(tile)), "l"(src), "r"(bytes), "r"(b) : "memory");
    }
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(b), "r"(0) : "memory");
    } while (!ok);
    for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) out[i] = tile[i];
}

__global__ void k_tma(const __grid_constant__ CUtensorMap map, int bytes, int c0, int r0, uint32_t* out) {
    extern __shared__ __align__(1024) uint32_t tile[];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t b = smem_u32(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
            ::"r"(smem_u32(tile)), "l"(reinterpret_cast<uint64_t>(&map)), "r"(c0), "r"(r0), "r"(b) : "memory");
    }
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(b), "r"(0) : "memory");
    } while (!ok);
    for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) out[i] = tile[i];
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int mode = argc > 1 ? atoi(argv[1]) : 0;
    int rows = 4800, pitch = 20, bw = 8, R = 32;
    CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_UINT32;
    if (mode == 1) { rows = 1024; pitch = 1024; bw = 32; R = 32; dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT32; }
    if (mode == 3) { pitch = 32; }
    std::vector<uint32_t> h((size_t)rows * pitch);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint32_t)i;
    uint32_t *d, *out;
    cudaMalloc(&d, h.size() * 4); cudaMalloc(&out, 8192);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    const size_t smem = 49152 + 64;
    cudaError_t e;
    if (mode == 0) {
        cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_bulk<<<1, 128, smem>>>(d + 400, 4096, out);
        e = cudaDeviceSynchronize();
        printf("mode 0 (1-D bulk): %s", cudaGetErrorString(e));
        if (e == cudaSuccess) { uint32_t g[4]; cudaMemcpy(g, out, 16, cudaMemcpyDeviceToHost); printf("  first=%u (want 400)", g[0]); }
        printf("\n");
        return e != cudaSuccess;
    }
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeFn fn = (EncodeFn)p;
    CUtensorMap map;
    cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)pitch * 4};
    cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)R};
    cuuint32_t es[2] = {1, 1};
    CUresult r = fn(&map, dt, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode -> %d; ", (int)r);
    cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_tma<<<1, 128, smem>>>(map, bw * R * 4, argc > 2 ? atoi(argv[2]) : 4, 100, out);
    e = cudaDeviceSynchronize();
    printf("mode %d (tma 2-D %dx%d box %dx%d): %s", mode, rows, pitch, bw, R, cudaGetErrorString(e));
    if (e == cudaSuccess) { uint32_t g[4]; cudaMemcpy(g, out, 16, cudaMemcpyDeviceToHost); printf("  first=%u (want %d)", g[0], 100 * pitch + 4); }
    printf("\n");
    return e != cudaSuccess;
}
