// standalone probe: which 3-D TMA box configurations load correctly on this GPU/driver
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tmap, int c0, int c1, int c2, uint32_t bytes, uint32_t* out, int nwords) {
    extern __shared__ __align__(128) uint8_t raw[];
    uint32_t* buf = reinterpret_cast<uint32_t*>(raw);
    __shared__ alignas(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(smem_u32(buf)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(&bar)) : "memory");
    }
    __syncthreads();
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(smem_u32(&bar)) : "memory");
    for (int i = threadIdx.x; i < nwords; i += blockDim.x) out[i] = buf[i];
}
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int run(EncodeFn enc, int w, int h, int n, int boxw, int boxh, int c0, int c1, CUtensorMapDataType dt, int esz, const char* label) {
    size_t bytes = (size_t)w * 3 * h * n;
    uint8_t* d; cudaMalloc(&d, bytes);
    std::vector<uint8_t> hsrc(bytes); for (size_t i = 0; i < bytes; ++i) hsrc[i] = (uint8_t)(i * 7 + 3);
    cudaMemcpy(d, hsrc.data(), bytes, cudaMemcpyHostToDevice);
    CUtensorMap tm; memset(&tm, 0, sizeof(tm));
    cuuint64_t dims[3] = {(cuuint64_t)w * 3 / esz, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t strides[2] = {(cuuint64_t)w * 3, (cuuint64_t)w * 3 * h};
    cuuint32_t box[3] = {(cuuint32_t)boxw, (cuuint32_t)boxh, 1}; cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&tm, dt, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    uint32_t tile = boxw * esz * boxh; int nwords = tile / 4;
    uint32_t* dout; cudaMalloc(&dout, tile); cudaMemset(dout, 0xEE, tile);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000);
    k<<<1, 128, tile + 128>>>(tm, c0, c1, 0, tile, dout, nwords);
    cudaError_t e = cudaDeviceSynchronize();
    int bad = -1;
    if (e == cudaSuccess) {
        std::vector<uint8_t> ho(tile); cudaMemcpy(ho.data(), dout, tile, cudaMemcpyDeviceToHost);
        bad = 0;
        for (int y = 0; y < boxh; ++y) for (int xb = 0; xb < boxw * esz; ++xb) {
            long gy = c1 + y, gxb = (long)c0 * esz + xb;
            uint8_t want = (gy >= 0 && gy < h && gxb >= 0 && gxb < (long)w * 3) ? hsrc[(size_t)gy * w * 3 + gxb] : 0;
            if (ho[(size_t)y * boxw * esz + xb] != want) ++bad;
        }
    }
    printf("%-40s encode=%d run=%s mismatches=%d\n", label, (int)r, cudaGetErrorString(e), bad);
    if (e != cudaSuccess) { cudaDeviceReset(); return 1; }
    cudaFree(d); cudaFree(dout);
    return 0;
}
int main(int argc, char** argv) {
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    EncodeFn enc = (EncodeFn)fn;
    int which = argc > 1 ? atoi(argv[1]) : 0;
    if (which == 0) return run(enc, 640, 426, 2, 64, 16, 96, 8, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, "u32 64x16 interior");
    if (which == 1) return run(enc, 640, 426, 2, 104, 67, 93, 62, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, "u32 104x67 interior");
    if (which == 2) return run(enc, 640, 426, 2, 104, 67, -3, -2, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, "u32 104x67 negative origin");
    if (which == 3) return run(enc, 32, 32, 1, 104, 67, -3, -2, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, "u32 104x67 box > tensor (32x32)");
    if (which == 4) return run(enc, 640, 426, 2, 208, 35, -12, -2, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, "u8 208x35 negative origin");
    if (which == 5) return run(enc, 640, 426, 2, 104, 67, 400, 400, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, "u32 104x67 overhang bottom/right");
    if (which == 6) return run(enc, 32, 32, 1, 24, 32, 0, 0, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, "u32 24x32 box == tensor (32x32)");
    return 0;
}
