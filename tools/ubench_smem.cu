// tools/ubench_smem.cu -- shared-memory pipe probes (sm_100a): bytes/clk/SM of LDS/STS at 64 and 128 bits, and of
// the mixes the fused PDQ kernel issues.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_smem ubench_smem.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum { LD64, LD128, ST64, ST128, LD64_ST64, LD128_ST128, LD64_ST64_SAME };

template <int OP>
__global__ void probe(float* out, int iters, long long* cycles) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // each warp owns 8 KB: rows of 272 B (lane = row) or contiguous (lane = column)
    unsigned base = (unsigned)__cvta_generic_to_shared(smem) + w * 8704;
    const unsigned row = base + lane * 272, col = base + lane * 8;
    float2 a = make_float2(1.0f, 2.0f);
    float4 b = make_float4(1.f, 2.f, 3.f, 4.f);
    float acc = 0.f;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (OP == LD64 || OP == LD64_ST64 || OP == LD64_ST64_SAME) {
                float2 v;
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(col + k * 272));
                acc += v.x;
            }
            if (OP == LD128 || OP == LD128_ST128) {
                float4 v;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(row + k * 16));
                acc += v.x;
            }
            if (OP == ST64 || OP == LD64_ST64)
                asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(col + (k + 16) * 272), "f"(a.x), "f"(a.y));
            if (OP == LD64_ST64_SAME)
                asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(col + k * 272), "f"(a.x), "f"(a.y));
            if (OP == ST128 || OP == LD128_ST128)
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(row + 256 - k * 16), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w));
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int OP>
void run(const char* name, int warps, double bytes_per_inner) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 2048;
    float* out;
    long long* cyc;
    cudaMalloc(&out, sizeof(float) * warps * 32 * sms);
    cudaMalloc(&cyc, sizeof(long long));
    cudaFuncSetAttribute(probe<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8704 * 16);
    probe<OP><<<sms, warps * 32, 8704 * warps>>>(out, 8, cyc);
    probe<OP><<<sms, warps * 32, 8704 * warps>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, cyc, sizeof c, cudaMemcpyDeviceToHost);
    const double bytes = (double)warps * iters * 16 * bytes_per_inner;
    printf("%-16s warps %2d : %9lld cyc  %7.1f B/clk/SM  (%.2f cyc per warp-instruction pair/op)\n", name, warps, c,
           bytes / (double)c, (double)c / ((double)warps * iters * 16));
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    for (int warps : {8, 16}) {
        if (warps == 8) {
            run<LD64>("LDS.64", 8, 256); run<LD128>("LDS.128", 8, 512); run<ST64>("STS.64", 8, 256); run<ST128>("STS.128", 8, 512);
            run<LD64_ST64>("LDS.64+STS.64", 8, 512); run<LD128_ST128>("LDS.128+STS.128", 8, 1024);
            run<LD64_ST64_SAME>("LDS.64+STS.64 same", 8, 512);
        } else {
            run<LD64>("LDS.64", 16, 256); run<LD128>("LDS.128", 16, 512); run<ST64>("STS.64", 16, 256); run<ST128>("STS.128", 16, 512);
            run<LD64_ST64>("LDS.64+STS.64", 16, 512); run<LD128_ST128>("LDS.128+STS.128", 16, 1024);
        }
    }
    return 0;
}
