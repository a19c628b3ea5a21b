// tools/tma_probe.cu -- which TMA box parameters does sm_100a accept for u8 2-D tiled loads?
// usage: tma_probe <box_w> <x> <y>    (box_h = 32, tensor = 1536 x 2048 bytes)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap tmap, int x, int y, int box_bytes, uint8_t* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(box_bytes) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
            ::"r"(smem_u32(smem)), "l"(&tmap), "r"(x), "r"(y), "r"(smem_u32(&bar)) : "memory");
    }
    int ok = 0;
    for (int spin = 0; spin < (1 << 22) && !ok; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    }
    if (!ok && threadIdx.x == 0) printf("TIMEOUT\n");
    for (int i = threadIdx.x; i < box_bytes; i += blockDim.x) out[i] = smem[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int box_w = argc > 1 ? atoi(argv[1]) : 112, x = argc > 2 ? atoi(argv[2]) : 6, y = argc > 3 ? atoi(argv[3]) : 2;
    const int W = 1536, H = 2048, box_h = 32;
    std::vector<uint8_t> h((size_t)W * H);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)((i * 2654435761u) >> 13);
    uint8_t *d, *dout;
    cudaMalloc(&d, h.size());
    cudaMalloc(&dout, box_w * box_h);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {(cuuint64_t)W, (cuuint64_t)H};
    const cuuint64_t gstride[1] = {(cuuint64_t)W};
    const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = ((EncodeTiledFn)p)(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, gdim, gstride, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("box %dx%d at (%d,%d): encode=%d ", box_w, box_h, x, y, (int)r);
    if (r != CUDA_SUCCESS) { printf("\n"); return 0; }
    probe<<<1, 128, box_w * box_h>>>(tmap, x, y, box_w * box_h, dout);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel=%s ", cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<uint8_t> o(box_w * box_h);
        cudaMemcpy(o.data(), dout, o.size(), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int rr = 0; rr < box_h; ++rr)
            for (int b = 0; b < box_w; ++b) {
                const int row = y + rr, col = x + b;
                const uint8_t want = (row >= 0 && row < H && col >= 0 && col < W) ? h[(size_t)row * W + col] : 0;
                bad += o[rr * box_w + b] != want;
            }
        printf("mismatches=%d", bad);
    }
    printf("\n");
    return 0;
}
