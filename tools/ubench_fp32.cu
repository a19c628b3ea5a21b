// tools/ubench_fp32.cu -- issue-rate / latency probes behind the fused PDQ kernel's design (sm_100a):
// scalar FADD/FFMA vs the packed FADD2/FFMA2/FMUL2 forms, PRMT on the ALU pipe, and mixes of them, at the
// occupancies the kernel runs at (2 and 4 warps per scheduler).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_fp32 ubench_fp32.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum { FADD1, FADD2P, FFMA1, FFMA2P, FMUL2P, PRMT1, MIX_FADD2_PRMT, MIX_FADD2_FADD, MIX_LUMA2, CHAIN2x4 };

template <int OP, int ILP>
__global__ void probe(float* out, float seed, int iters, long long* cycles) {
    float2 a[ILP];
    uint32_t u[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) {
        a[k] = make_float2(seed * (threadIdx.x + 1) + k, seed * 0.5f + k);
        u[k] = __float_as_uint(a[k].x) * 2654435761u;
    }
    const float2 b = make_float2(seed, seed * 1.25f);
    const float2 c = make_float2(1.0f - seed * 1e-6f, 1.0f + seed * 1e-6f);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
            if (OP == FADD1) a[k].x = __fadd_rn(a[k].x, b.x);
            if (OP == FADD2P) a[k] = __fadd2_rn(a[k], b);
            if (OP == FFMA1) a[k].x = __fmaf_rn(a[k].x, c.x, b.x);
            if (OP == FFMA2P) a[k] = __ffma2_rn(a[k], c, b);
            if (OP == FMUL2P) a[k] = __fmul2_rn(a[k], c);
            if (OP == PRMT1) u[k] = __byte_perm(u[k], 0x4B000000u, 0x7541u);
            if (OP == MIX_FADD2_PRMT) {
                a[k] = __fadd2_rn(a[k], b);
                u[k] = __byte_perm(u[k], 0x4B000000u, 0x7541u);
            }
            if (OP == MIX_FADD2_FADD) {
                a[k] = __fadd2_rn(a[k], b);
                c.x == 0.0f ? (void)0 : (void)0;
                u[k] = __float_as_uint(__fadd_rn(__uint_as_float(u[k]), b.y));
            }
            if (OP == MIX_LUMA2) {
                // the per-pixel-pair instruction mix of the packed kernel's P1 role:
                // 6 PRMT + 3 FFMA2 + 2 FADD2 (luma) + 2 FADD2 (chain) + 1 FMUL2 (scale)
                const uint32_t w = u[k];
                const float2 mr = make_float2(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540u)),
                                              __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7541u)));
                const float2 mg = make_float2(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7542u)),
                                              __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7543u)));
                const float2 mb = make_float2(__uint_as_float(__byte_perm(w ^ 0x55u, 0x4B000000u, 0x7540u)),
                                              __uint_as_float(__byte_perm(w ^ 0x55u, 0x4B000000u, 0x7541u)));
                const float2 r = __ffma2_rn(make_float2(0.299f, 0.299f), mr, make_float2(-2508193.75f, -2508193.75f));
                const float2 g = __ffma2_rn(make_float2(0.587f, 0.587f), mg, make_float2(-4924113.0f, -4924113.0f));
                const float2 bl = __ffma2_rn(make_float2(0.114f, 0.114f), mb, make_float2(-956301.3125f, -956301.3125f));
                const float2 l = __fadd2_rn(__fadd2_rn(r, g), bl);
                a[k] = __fadd2_rn(a[k], l);
                a[k] = __fadd2_rn(a[k], b);
                const float2 y = __fmul2_rn(a[k], make_float2(0.25f, 0.25f));
                u[k] = w + __float_as_uint(y.x) + __float_as_uint(y.y);
            }
            if (OP == CHAIN2x4) {  // dependent add, sub per chain (the running-sum box filter), ILP chains
                a[k] = __fadd2_rn(a[k], b);
                a[k] = __fadd2_rn(a[k], c);
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += a[k].x + a[k].y + __uint_as_float(u[k]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int OP, int ILP>
void run(const char* name, int threads_per_sm, int instr_per_inner) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int iters = 8192;
    float* out;
    long long* cyc;
    cudaMalloc(&out, sizeof(float) * threads_per_sm * sms);
    cudaMalloc(&cyc, sizeof(long long));
    probe<OP, ILP><<<sms, threads_per_sm>>>(out, 1.5f, 16, cyc);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<OP, ILP><<<sms, threads_per_sm>>>(out, 1.5f, iters, cyc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long c = 0;
    cudaMemcpy(&c, cyc, sizeof c, cudaMemcpyDeviceToHost);
    const double warps = threads_per_sm / 32.0;
    const double winstr = warps * iters * ILP * instr_per_inner;  // warp-instructions per SM
    printf("%-22s thr/SM %4d ILP %d : %9lld cyc, %6.3f warp-instr/clk/SM (%.3f per scheduler), %.2f cyc per inner op per warp\n",
           name, threads_per_sm, ILP, c, winstr / (double)c, winstr / (double)c / 4.0, (double)c / (iters * (double)ILP));
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs, clock %d MHz\n", p.name, p.multiProcessorCount, p.clockRate / 1000);
    // throughput, full occupancy
    run<FADD1, 8>("FADD", 1024, 1);
    run<FADD2P, 8>("FADD2", 1024, 1);
    run<FFMA1, 8>("FFMA", 1024, 1);
    run<FFMA2P, 8>("FFMA2", 1024, 1);
    run<FMUL2P, 8>("FMUL2", 1024, 1);
    run<PRMT1, 8>("PRMT", 1024, 1);
    run<MIX_FADD2_PRMT, 8>("FADD2+PRMT", 1024, 2);
    run<MIX_FADD2_FADD, 8>("FADD2+FADD", 1024, 2);
    // dependent latency: one warp per SM, one chain
    run<FADD1, 1>("FADD lat", 32, 1);
    run<FADD2P, 1>("FADD2 lat", 32, 1);
    run<FFMA2P, 1>("FFMA2 lat", 32, 1);
    run<PRMT1, 1>("PRMT lat", 32, 1);
    // the kernel's occupancies: 8 warps (2 per scheduler) and 16 warps (4 per scheduler)
    run<FADD2P, 4>("FADD2", 256, 1);
    run<FADD2P, 4>("FADD2", 512, 1);
    run<FADD1, 4>("FADD", 256, 1);
    run<FADD1, 4>("FADD", 512, 1);
    run<CHAIN2x4, 4>("chain(add,sub) x4", 256, 2);
    run<CHAIN2x4, 4>("chain(add,sub) x4", 512, 2);
    run<CHAIN2x4, 2>("chain(add,sub) x2", 256, 2);
    run<MIX_LUMA2, 2>("P1 mix (14 instr)", 256, 14);
    run<MIX_LUMA2, 2>("P1 mix (14 instr)", 512, 14);
    run<MIX_LUMA2, 4>("P1 mix (14 instr)", 256, 14);
    return 0;
}
