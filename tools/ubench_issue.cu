// tools/ubench_issue.cu -- what does an instruction COST next to a stream of scalar FADDs at the systolic kernel's
// occupancy (8 warps per SM = 2 per scheduler)?  Every test runs 32 FADDs (4 dependent chains x 8) per body plus N
// instructions of another kind, and reports cycles per body per scheduler-warp-pair.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_issue ubench_issue.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum { NONE, PRMT16, LOP16, SEL16, FADD2_16, FMUL16, IMMA4, SHFL8, PRMT48, IMMA12, IADD16, FFMA2_16, PRMT_ONLY16, IMMA_ONLY4, FADD2_ONLY16, I2F16 };

__device__ __forceinline__ void imma(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int OP>
__global__ void __launch_bounds__(256, 1) probe(float* out, float seed, int iters, long long* cycles, int with_fadd) {
    float c[4];
    uint32_t u[16];
    float2 p[8];
    int d[4][4];
    uint32_t av[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) c[k] = seed * (threadIdx.x + 1) + k;
#pragma unroll
    for (int k = 0; k < 16; ++k) u[k] = (threadIdx.x + k) * 2654435761u;
#pragma unroll
    for (int k = 0; k < 8; ++k) p[k] = make_float2(seed + k, seed * 2 + k);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        av[k] = u[k];
#pragma unroll
        for (int j = 0; j < 4; ++j) d[k][j] = k + j;
    }
    const float b = seed * 1e-3f;
    const bool pr = (threadIdx.x & 1) != 0;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        if (OP < PRMT_ONLY16) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int k = 0; k < 4; ++k) c[k] = __fadd_rn(c[k], b);
        }
        if (OP == PRMT16 || OP == PRMT_ONLY16) {
#pragma unroll
            for (int k = 0; k < 16; ++k) u[k] = __byte_perm(u[k], 0x4B000000u, 0x7541u);
        }
        if (OP == PRMT48) {
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int k = 0; k < 16; ++k) u[k] = __byte_perm(u[k], 0x4B000000u, 0x7541u);
        }
        if (OP == LOP16) {
#pragma unroll
            for (int k = 0; k < 16; ++k) u[k] = (u[k] & 0xFF00FFu) | (u[(k + 1) & 15] & 0x4B000000u);
        }
        if (OP == IADD16) {
#pragma unroll
            for (int k = 0; k < 16; ++k) u[k] = u[k] + u[(k + 1) & 15] + 12345u;
        }
        if (OP == SEL16) {
#pragma unroll
            for (int k = 0; k < 16; ++k) u[k] = pr ? u[(k + 1) & 15] : u[(k + 5) & 15];
        }
        if (OP == FADD2_16 || OP == FADD2_ONLY16) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int k = 0; k < 8; ++k) p[k] = __fadd2_rn(p[k], make_float2(b, b));
        }
        if (OP == FFMA2_16) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int k = 0; k < 8; ++k) p[k] = __ffma2_rn(p[k], make_float2(1.0000001f, 1.0000001f), make_float2(b, b));
        }
        if (OP == FMUL16) {
#pragma unroll
            for (int k = 0; k < 16; ++k) u[k] = __float_as_uint(__fmul_rn(__uint_as_float(u[k]), 1.0000001f));
        }
        if (OP == I2F16) {
#pragma unroll
            for (int k = 0; k < 16; ++k) u[k] = __float_as_uint((float)(u[k] & 0xFFu)) + u[k];
        }
        if (OP == IMMA4 || OP == IMMA_ONLY4) {
#pragma unroll
            for (int k = 0; k < 4; ++k) imma(d[k], av, u[4], u[5]);
        }
        if (OP == IMMA12) {
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int k = 0; k < 4; ++k) imma(d[k], av, u[4], u[5]);
        }
        if (OP == SHFL8) {
#pragma unroll
            for (int k = 0; k < 8; ++k) u[k] = __shfl_sync(0xffffffffu, u[k], (threadIdx.x + 31) & 31);
        }
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) s += c[k] + d[k][0] + d[k][1] + d[k][2] + d[k][3];
#pragma unroll
    for (int k = 0; k < 16; ++k) s += __uint_as_float(u[k]);
#pragma unroll
    for (int k = 0; k < 8; ++k) s += p[k].x + p[k].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int OP>
void run(const char* name, int threads) {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(float));
    cudaMalloc(&cyc, sizeof(long long));
    const int iters = 20000;
    probe<OP><<<148, threads>>>(out, 1.5f, 100, cyc, 1);
    probe<OP><<<148, threads>>>(out, 1.5f, iters, cyc, 1);
    cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    const int warps_per_sched = threads / 128;
    printf("%-34s %4d thr/SM: %7.1f cycles per body per warp = %6.1f per body-of-one-warp at the scheduler (%s)\n", name, threads,
           (double)h / iters, (double)h / iters / warps_per_sched, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int threads : {128, 256, 512}) {
        run<NONE>("32 FADD", threads);
        run<PRMT16>("32 FADD + 16 PRMT", threads);
        run<PRMT48>("32 FADD + 48 PRMT", threads);
        run<LOP16>("32 FADD + 16 LOP3", threads);
        run<IADD16>("32 FADD + 16 IADD3", threads);
        run<SEL16>("32 FADD + 16 SEL", threads);
        run<FMUL16>("32 FADD + 16 FMUL", threads);
        run<FADD2_16>("32 FADD + 16 FADD2", threads);
        run<FFMA2_16>("32 FADD + 16 FFMA2", threads);
        run<I2F16>("32 FADD + 16 (LOP3 + I2F + IADD)", threads);
        run<IMMA4>("32 FADD + 4 IMMA.16832.U8", threads);
        run<IMMA12>("32 FADD + 12 IMMA.16832.U8", threads);
        run<SHFL8>("32 FADD + 8 SHFL", threads);
        run<PRMT_ONLY16>("16 PRMT", threads);
        run<FADD2_ONLY16>("16 FADD2", threads);
        run<IMMA_ONLY4>("4 IMMA.16832.U8", threads);
    }
    return 0;
}
