// tools/ubench.cu -- instruction-throughput probes for the Hamming kernels' roofline (POPC / LOP3 / IADD3
// issue rate per SM per clock on sm_100a).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP>
__global__ void probe(uint32_t* out, uint32_t seed, int iters, long long* cycles) {
    uint32_t a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = seed * (threadIdx.x + 1) + k * 0x9E3779B9u;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (OP == 0) a[k] = __popc(a[k]) + seed;                      // POPC + IADD
            if (OP == 1) a[k] = (a[k] ^ seed) + 0x1234567u;               // LOP3 + IADD
            if (OP == 2) a[k] = a[k] + seed;                               // IADD only
            if (OP == 3) asm volatile("popc.b32 %0, %0;" : "+r"(a[k]));     // POPC only (dependent chain x8 ILP)
            if (OP == 4) a[k] = __popc(a[k] ^ seed);                       // LOP3 + POPC
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int OP>
void run(const char* name, int ops_per_inner) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int threads = 1024, blocks = sms * 2, iters = 4096;
    uint32_t* out;
    long long* cyc;
    cudaMalloc(&out, sizeof(uint32_t) * threads * blocks);
    cudaMalloc(&cyc, sizeof(long long));
    probe<OP><<<blocks, threads>>>(out, 12345u, 16, cyc);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<OP><<<blocks, threads>>>(out, 12345u, iters, cyc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long c = 0;
    cudaMemcpy(&c, cyc, sizeof c, cudaMemcpyDeviceToHost);
    // one block's view: 2 blocks of 1024 threads share an SM -> 2048 threads * iters * 8 inner per SM
    const double inner_per_sm = 2048.0 * iters * 8;
    printf("%-14s %8.3f ms  block0 cycles %lld  -> %.1f inner-ops/clk/SM (x%d instr each), %.2f T inner-ops/s chip\n",
           name, ms, c, inner_per_sm / (double)c, ops_per_inner, inner_per_sm * sms / (ms * 1e-3) / 1e12);
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs, L2 %d MB, clock %d MHz\n", p.name, p.multiProcessorCount, p.l2CacheSize >> 20, p.clockRate / 1000);
    run<0>("popc+iadd", 2);
    run<1>("lop3+iadd", 2);
    run<2>("iadd", 1);
    run<3>("popc", 1);
    run<4>("lop3+popc", 2);
    return 0;
}
