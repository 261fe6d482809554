// Isolates the TMA 2-D load used by k_score_tma: uint32 [rows][pitch] tensor, box {bw, R}.
// usage: tma_probe <variant 0..3> <bw>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>

struct Maps { CUtensorMap m[4]; int bw[4]; int rows[4]; };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void body(const CUtensorMap* map, int bytes, int c0, int r0, uint32_t* out) {
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
            "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
            ::"r"(smem_u32(tile)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(r0), "r"(b) : "memory");
    }
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(b), "r"(0) : "memory");
    } while (!ok);
    for (int i = threadIdx.x; i < 256; i += blockDim.x) out[i] = tile[i];
}

__global__ void probe_single(const __grid_constant__ CUtensorMap map, int bytes, int c0, int r0, uint32_t* out) {
    body(&map, bytes, c0, r0, out);
}
__global__ void probe_struct0(const __grid_constant__ Maps maps, int bytes, int c0, int r0, uint32_t* out) {
    body(&maps.m[0], bytes, c0, r0, out);
}
__global__ void probe_struct_rt(const __grid_constant__ Maps maps, int sel, int bytes, int c0, int r0, uint32_t* out) {
    body(&maps.m[sel], bytes, c0, r0, out);
}
__global__ void probe_global(const CUtensorMap* map, int bytes, int c0, int r0, uint32_t* out) {
    body(map, bytes, c0, r0, out);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    const int bw = argc > 2 ? atoi(argv[2]) : 8;
    const int l2 = argc > 3 ? atoi(argv[3]) : 0;
    const int pitch = 20, rows = 4800, R = 256 / bw;
    std::vector<uint32_t> h((size_t)rows * pitch);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint32_t)i;
    uint32_t *d, *out;
    cudaMalloc(&d, h.size() * 4); cudaMalloc(&out, 1024);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeFn fn = (EncodeFn)p;
    Maps maps;
    cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)pitch * 4};
    cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)R};
    cuuint32_t es[2] = {1, 1};
    for (int i = 0; i < 4; ++i) {
        CUresult r = fn(&maps.m[i], CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d, dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        l2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r) { printf("encode failed %d\n", (int)r); return 2; }
    }
    const int bytes = bw * R * 4, c0 = 3, r0 = 100;
    const size_t smem = 49152 + 64;
    cudaError_t e;
    if (variant == 0) {
        cudaFuncSetAttribute(probe_single, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        probe_single<<<1, 128, smem>>>(maps.m[0], bytes, c0, r0, out);
    } else if (variant == 1) {
        cudaFuncSetAttribute(probe_struct0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        probe_struct0<<<1, 128, smem>>>(maps, bytes, c0, r0, out);
    } else if (variant == 2) {
        cudaFuncSetAttribute(probe_struct_rt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        probe_struct_rt<<<1, 128, smem>>>(maps, 2, bytes, c0, r0, out);
    } else {
        CUtensorMap* dm; cudaMalloc(&dm, sizeof(CUtensorMap));
        cudaMemcpy(dm, &maps.m[0], sizeof(CUtensorMap), cudaMemcpyHostToDevice);
        cudaFuncSetAttribute(probe_global, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        probe_global<<<1, 128, smem>>>(dm, bytes, c0, r0, out);
    }
    e = cudaDeviceSynchronize();
    printf("variant=%d bw=%d l2=%d: %s", variant, bw, l2, cudaGetErrorString(e));
    if (e != cudaSuccess) { printf("\n"); return 1; }
    uint32_t got[256]; cudaMemcpy(got, out, 1024, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r = 0; r < R; ++r)
        for (int c = 0; c < bw; ++c) {
            uint32_t want = (c0 + c < pitch) ? (uint32_t)((r0 + r) * pitch + c0 + c) : 0u;
            if (got[r * bw + c] != want) ++bad;
        }
    printf("  mismatches=%d first=%u\n", bad, got[0]);
    return 0;
}
