// tests/legacy/pdq_lines.cu -- TEST-ONLY: the round-1 PDQ pipelines (v1 "line" kernels, k4_colpass_finalize, and the
// launchers of the two tiled fused kernels next to this file), kept as an independent CUDA implementation that the
// parity tests cross-check the product against (tests/test_parity_hash.py::test_legacy_pipelines_agree).  Built into
// tests/legacy/libvpdq_b200_legacy.so; nothing here is part of libvpdq_b200.so.
//
// Why "lines": the Jarosz box filter is a RUNNING SUM (s += x[r]; s -= x[l]; y = s / n), so every
// output carries the rounding history of its whole row/column prefix.  A line cannot be tiled or
// tree-reduced without changing bits; what CAN run in parallel are the 512 independent lines of a
// pass.  So each pass is: one thread per line, 512-step dependent chain in registers, results written
// TRANSPOSED so that the next (orthogonal) pass again reads contiguous lines and every store is a
// coalesced 128-byte warp transaction.
//
//   K1  luma + row pass 1      thread = (frame, row)     RGB row (1536 B)   -> P1^T [512 j][512 i] f32
//   K2  column pass 1          thread = (frame, column)  P1^T line          -> P2   [512 i][512 j] f32
//   K3  row pass 2, decimated  thread = (frame, row)     P2 line            -> P3^T [64 jj][512 i] f32
//   K4  column pass 2 + 64x64 decimate + quality + 64->16 DCT + median + bits   CTA = frame
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "legacy.cuh"

namespace vpdq {

__constant__ float c_dct[16 * 64];

// ---------------------------------------------------------------------------------------------------
// Running-sum box filter, window 4 (512-wide lines): Appendix A step 3 with x[k] = 0 outside [0, 512).
//   feed(x[r]) returns the window sum belonging to output index o = r - 2:
//       s += x[r];  s -= x[r-4];        (adding / subtracting +0.0f is exact, so the four upstream
//                                        phases collapse into this one step plus two drains)
//   divisors: o = 0 -> 3, o = 1..509 -> 4, o = 510 -> 3, o = 511 -> 2.
// ---------------------------------------------------------------------------------------------------
struct BoxChain {
    float s, r0, r1, r2, r3;  // r0 = x[r-4] ... r3 = x[r-1]
    __device__ __forceinline__ void init() { s = r0 = r1 = r2 = r3 = 0.0f; }
    __device__ __forceinline__ float feed(float x) {
        s = fadd(s, x);
        s = fsub(s, r0);
        r0 = r1; r1 = r2; r2 = r3; r3 = x;
        return s;
    }
    __device__ __forceinline__ float drain() {
        s = fsub(s, r0);
        r0 = r1; r1 = r2; r2 = r3; r3 = 0.0f;
        return s;
    }
};

// u8 -> fp32 product without an I2F: byte b is spliced into the mantissa of 2^23 (PRMT), giving the
// float M = 2^23 + b exactly; then fma(c, M, -c*2^23) = RN(c*b) -- the exact real product rounded
// once, i.e. bit-identical to __fmul_rn(c, (float)b).  (c*2^23 is exact: a power-of-two scaling.)
__device__ __forceinline__ float byte_magic(uint32_t word, int k) {
    return __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7540u + k));
}
__device__ __forceinline__ float luma_rgb(float mr, float mg, float mb) {
    const float cr = 0.299f, cg = 0.587f, cb = 0.114f, two23 = 8388608.0f;
    const float r = __fmaf_rn(cr, mr, -(cr * two23));
    const float g = __fmaf_rn(cg, mg, -(cg * two23));
    const float b = __fmaf_rn(cb, mb, -(cb * two23));
    return fadd(fadd(r, g), b);  // (0.299 R + 0.587 G) + 0.114 B
}

// 16 pixels of luma from 48 (RGB) or 16 (gray) consecutive bytes of one row
template <int CH>
__device__ __forceinline__ void load_luma16(const uint8_t* row, int it, float (&x)[16]) {
    if (CH == 3) {
        const uint4* p = reinterpret_cast<const uint4*>(row) + 3 * it;
        const uint4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
        const uint32_t w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
#pragma unroll
        for (int px = 0; px < 16; ++px) {
            const int b0 = 3 * px, b1 = 3 * px + 1, b2 = 3 * px + 2;
            x[px] = luma_rgb(byte_magic(w[b0 >> 2], b0 & 3), byte_magic(w[b1 >> 2], b1 & 3),
                             byte_magic(w[b2 >> 2], b2 & 3));
        }
    } else {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(row) + it);
        const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int px = 0; px < 16; ++px) {
            const float m = byte_magic(w[px >> 2], px & 3);
            x[px] = luma_rgb(m, m, m);
        }
    }
}

__device__ __forceinline__ void load_f16(const float* line, int it, float (&x)[16]) {
    const float4* p = reinterpret_cast<const float4*>(line) + 4 * it;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 v = __ldg(p + q);
        x[4 * q + 0] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
    }
}

// Full-resolution pass: all 512 outputs of the line, written with stride 512 (transposed).
template <int SRC /*3 = RGB, 1 = gray, 0 = float line*/>
__device__ __forceinline__ void box_line_full(const void* src, float* dst /* + o*512 */) {
    BoxChain c;
    c.init();
#pragma unroll 1
    for (int it = 0; it < 32; ++it) {
        float x[16];
        if (SRC == 0)
            load_f16(static_cast<const float*>(src), it, x);
        else
            load_luma16<SRC>(static_cast<const uint8_t*>(src), it, x);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float v = c.feed(x[k]);
            const int o = it * 16 + k - 2;
            if (k >= 2 || it > 0) {
                float y;
                if (k == 2 && it == 0)
                    y = fdiv(v, 3.0f);  // o == 0
                else
                    y = fmul(v, 0.25f);  // == v / 4.0f exactly
                dst[(size_t)o * kDim] = y;
            }
        }
    }
    dst[(size_t)510 * kDim] = fdiv(c.drain(), 3.0f);
    dst[(size_t)511 * kDim] = fmul(c.drain(), 0.5f);
}

template <int CH>
__global__ void __launch_bounds__(128) k1_luma_rowpass(const uint8_t* __restrict__ frames, float* __restrict__ p1t,
                                                       int64_t n_lines) {
    const int64_t line = (int64_t)blockIdx.x * 128 + threadIdx.x;  // frame*512 + row
    if (line >= n_lines) return;
    const int64_t f = line >> 9;
    const int i = (int)(line & 511);
    box_line_full<CH>(frames + (size_t)line * (kDim * CH), p1t + (size_t)f * kPlane + i);
}

__global__ void __launch_bounds__(128) k2_colpass(const float* __restrict__ p1t, float* __restrict__ p2,
                                                  int64_t n_lines) {
    const int64_t line = (int64_t)blockIdx.x * 128 + threadIdx.x;  // frame*512 + column
    if (line >= n_lines) return;
    const int64_t f = line >> 9;
    const int j = (int)(line & 511);
    box_line_full<0>(p1t + (size_t)line * kDim, p2 + (size_t)f * kPlane + j);
}

// Decimated pass: only outputs o = 8*m + 4 (m = 0..63) are produced; they sit at k = 6 and k = 14 of
// each 16-step group (o = 16*it + k - 2).  emit(m, value).
template <typename Emit>
__device__ __forceinline__ void box_line_dec(const float* line, Emit emit) {
    BoxChain c;
    c.init();
#pragma unroll 1
    for (int it = 0; it < 32; ++it) {
        float x[16];
        load_f16(line, it, x);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float v = c.feed(x[k]);
            if (k == 6) emit(2 * it, fmul(v, 0.25f));
            if (k == 14) emit(2 * it + 1, fmul(v, 0.25f));
        }
    }
}

__global__ void __launch_bounds__(128) k3_rowpass_dec(const float* __restrict__ p2, float* __restrict__ p3t,
                                                      int64_t n_lines) {
    const int64_t line = (int64_t)blockIdx.x * 128 + threadIdx.x;  // frame*512 + row
    if (line >= n_lines) return;
    const int64_t f = line >> 9;
    const int i = (int)(line & 511);
    float* dst = p3t + (size_t)f * (kDec * kDim) + i;
    box_line_dec(p2 + (size_t)line * kDim, [&](int m, float y) { dst[(size_t)m * kDim] = y; });
}

// x / 255.0f, correctly rounded, branch free: q = RN(x*c), r = x - 255q (exact, FMA), q' = RN(q + r*c) with
// c = RN(1/255).  Equal to IEEE x / 255.0f for EVERY finite float (tests/emu/div3_check.c 255).
__device__ __forceinline__ float div255(float x) {
    const float c = 0.00392156886f;  // 0x3B808081
    const float q = fmul(x, c);
    const float r = __fmaf_rn(-255.0f, q, x);
    return __fmaf_rn(r, c, q);
}

// ---------------------------------------------------------------------------------------------------
// K4: one CTA (256 threads) per frame.
// ---------------------------------------------------------------------------------------------------
constexpr int kDP = 65;  // padded pitch of the 16x64 tables in shared memory

// FROM_A = false: `in` is p3t [n][64][512] (v1 line kernels): run column pass 2 here.
// FROM_A = true : `in` is a64 [n][64][64] (fused kernel already did column pass 2 + decimation).
template <bool FROM_A>
__global__ void __launch_bounds__(256) k4_colpass_finalize(const float* __restrict__ in,
                                                           uint8_t* __restrict__ hashes,
                                                           int32_t* __restrict__ quality,
                                                           float* __restrict__ a64_dbg,
                                                           float* __restrict__ b16_dbg) {
    __shared__ __align__(16) float A[kDec][kDec];   // decimated 64x64 plane
    __shared__ float D[16][kDP];      // DCT rows
    __shared__ float T[16][kDP];      // D * A
    __shared__ __align__(16) float B[256];
    __shared__ int g_sum;
    __shared__ float med;

    const int t = threadIdx.x;
    const int64_t f = blockIdx.x;

    for (int e = t; e < 16 * 64; e += 256) D[e >> 6][e & 63] = c_dct[e];
    if (t == 0) g_sum = 0;

    if (FROM_A) {
        const float4* src = reinterpret_cast<const float4*>(in + (size_t)f * (kDec * kDec));
        float4* dst = reinterpret_cast<float4*>(&A[0][0]);
#pragma unroll
        for (int e = 0; e < 4; ++e) dst[t + 256 * e] = __ldg(src + t + 256 * e);
    } else if (t < kDec) {
        // column pass 2 on the 64 surviving columns; outputs at rows 8m+4 are the decimated plane
        const float* line = in + (size_t)f * (kDec * kDim) + (size_t)t * kDim;
        box_line_dec(line, [&](int m, float y) { A[m][t] = y; });
    }
    __syncthreads();

    if (a64_dbg)
        for (int e = t; e < kDec * kDec; e += 256) a64_dbg[(size_t)f * 4096 + e] = A[e >> 6][e & 63];

    // quality: sum of |trunc((u - v) * 100 / 255)| over vertical and horizontal neighbours
    {
        int g = 0;
        for (int e = t; e < kDec * kDec; e += 256) {
            const int i = e >> 6, j = e & 63;
            const float u = A[i][j];
            if (i < 63) g += abs(__float2int_rz(div255(fmul(fsub(u, A[i + 1][j]), 100.0f))));
            if (j < 63) g += abs(__float2int_rz(div255(fmul(fsub(u, A[i][j + 1]), 100.0f))));
        }
        g = __reduce_add_sync(0xffffffffu, g);
        if ((t & 31) == 0) atomicAdd(&g_sum, g);
    }

    // T = D * A : thread -> column j, four rows i; sequential in k, separate multiply and add
    {
        const int j = t & 63, i0 = (t >> 6) * 4;
        float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 8
        for (int k = 0; k < 64; ++k) {
            const float a = A[k][j];
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[u] = fadd(acc[u], fmul(D[i0 + u][k], a));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) T[i0 + u][j] = acc[u];
    }
    __syncthreads();

    // B = T * D^T : thread -> (i, j)
    float bv;
    {
        const int i = t >> 4, j = t & 15;
        float acc = 0.0f;
#pragma unroll 8
        for (int k = 0; k < 64; ++k) acc = fadd(acc, fmul(T[i][k], D[j][k]));
        bv = acc;
        B[t] = acc;
    }
    __syncthreads();
    if (b16_dbg) b16_dbg[(size_t)f * 256 + t] = bv;

    // median = 128-th smallest of the 256 values (what Torben's method returns for n = 256):
    // the value v with  #{B < v} < 128 <= #{B <= v}
    {
        int lt = 0, le = 0;
        const float4* b4 = reinterpret_cast<const float4*>(B);
#pragma unroll 8
        for (int u = 0; u < 64; ++u) {
            const float4 b = b4[u];
            lt += (b.x < bv) + (b.y < bv) + (b.z < bv) + (b.w < bv);
            le += (b.x <= bv) + (b.y <= bv) + (b.z <= bv) + (b.w <= bv);
        }
        if (lt < 128 && le >= 128) med = bv;  // all writers hold the same value
    }
    __syncthreads();

    // bit k = 16 i + j = t  ->  byte t >> 3, bit t & 7: eight little-endian 32-bit ballots
    const unsigned word = __ballot_sync(0xffffffffu, bv > med);
    if ((t & 31) == 0) reinterpret_cast<uint32_t*>(hashes)[f * 8 + (t >> 5)] = word;
    if (t == 0) {
        const int q = g_sum / 90;
        quality[f] = q > 100 ? 100 : q;
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
static int legacy_upload_tables() {
    static std::mutex mu;
    static bool done[64] = {};
    int dev = 0;
    VPDQ_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    if (dev >= 0 && dev < 64 && done[dev]) return VPDQ_B200_OK;
    VPDQ_CUDA(cudaMemcpyToSymbol(c_dct, pdq_host_dct(), sizeof(float) * 16 * 64));
    if (dev >= 0 && dev < 64) done[dev] = true;
    return VPDQ_B200_OK;
}

constexpr size_t kLinesScratchPerFrame = ((size_t)2 * kPlane + (size_t)kDec * kDim) * sizeof(float);
constexpr int64_t kMaxChunk = 256;

}  // namespace vpdq

using namespace vpdq;

// impl: 0 = v1 line kernels, 1 = one-frame fused kernel (RGB24 only) + k4, 2 = frame-pair fused kernel + k4.
// d_scratch: at least legacy_scratch_bytes(n_frames).
extern "C" __attribute__((visibility("default"))) size_t legacy_scratch_bytes(long long n_frames) {
    const long long c = n_frames < 1 ? 1 : (n_frames > kMaxChunk ? kMaxChunk : n_frames);
    const size_t lines = (size_t)c * kLinesScratchPerFrame, fused = (size_t)(n_frames < 1 ? 1 : n_frames) * 16384;
    return lines > fused ? lines : fused;
}

extern "C" __attribute__((visibility("default"))) int legacy_debug_flags(int* flags) {
    int f1 = 0, f2 = 0;
    int rc = fused_debug_flags(&f1);
    if (rc == 0) rc = fused2_debug_flags(&f2);
    *flags = f1 | f2;
    return rc;
}

extern "C" __attribute__((visibility("default"))) const char* legacy_last_error() { return legacy_error(); }

extern "C" __attribute__((visibility("default"))) int legacy_pdq_stages_dev(
    int impl, const uint8_t* d_frames, int channels, long long n_frames, uint8_t* d_hashes, int32_t* d_quality,
    float* d_a64, float* d_b16, void* d_scratch, size_t scratch_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_frames == 0) return VPDQ_B200_OK;
    if (impl == 1 && channels != 3) return VPDQ_B200_ERR_UNSUPPORTED;
    const bool fused = impl != 0;
    const size_t per_frame = fused ? 16384 : kLinesScratchPerFrame;
    long long chunk = (long long)(scratch_bytes / per_frame);
    if (chunk < 1) return VPDQ_B200_ERR_INVALID;
    if (!fused && chunk > kMaxChunk) chunk = kMaxChunk;
    int rc = legacy_upload_tables();
    if (rc) return rc;
    const size_t frame_bytes = (size_t)kPlane * channels;
    for (long long f0 = 0; f0 < n_frames; f0 += chunk) {
        const long long nf = (n_frames - f0 < chunk) ? (n_frames - f0) : chunk;
        const uint8_t* src = d_frames + (size_t)f0 * frame_bytes;
        uint8_t* hp = d_hashes + (size_t)f0 * 32;
        int32_t* qp = d_quality + f0;
        float* adbg = d_a64 ? d_a64 + (size_t)f0 * 4096 : nullptr;
        float* bdbg = d_b16 ? d_b16 + (size_t)f0 * 256 : nullptr;
        if (fused) {
            float* a64 = static_cast<float*>(d_scratch);
            rc = impl == 2 ? fused2_jarosz_launch(src, channels, nf, a64, stream) : fused_jarosz_launch(src, nf, a64, stream);
            if (rc) return rc;
            k4_colpass_finalize<true><<<(unsigned)nf, 256, 0, stream>>>(a64, hp, qp, adbg, bdbg);
        } else {
            float* p1t = static_cast<float*>(d_scratch);
            float* p2 = p1t + (size_t)nf * kPlane;
            float* p3t = p2 + (size_t)nf * kPlane;
            const long long n_lines = nf * kDim;
            const unsigned grid = (unsigned)((n_lines + 127) / 128);
            if (channels == 3)
                k1_luma_rowpass<3><<<grid, 128, 0, stream>>>(src, p1t, n_lines);
            else
                k1_luma_rowpass<1><<<grid, 128, 0, stream>>>(src, p1t, n_lines);
            k2_colpass<<<grid, 128, 0, stream>>>(p1t, p2, n_lines);
            k3_rowpass_dec<<<grid, 128, 0, stream>>>(p2, p3t, n_lines);
            k4_colpass_finalize<false><<<(unsigned)nf, 256, 0, stream>>>(p3t, hp, qp, adbg, bdbg);
        }
    }
    VPDQ_CUDA(cudaGetLastError());
    return VPDQ_B200_OK;
}
