// pdq_kernels.cu -- PDQ frame hash on sm_100a, "line" formulation (v1).
//
// Replaces the native work behind hvdaccelerators.vpdq.VideoHasher.hash_frame
// (reference call site: src/hydrusvideodeduplicator/vpdqpy/vpdqpy.py:118); algorithm and op order are
// SURVEY.md Appendix A (Meta ThreatExchange PDQ).  Every fp32 operation that the CPU path performs is
// performed here in the same order with the same single rounding; nothing is re-associated.
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
//
// Passes 3/4 only produce the outputs the 64x64 decimation reads (rows/cols 8k+4) but still run the
// full running sums, as exactness demands.
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "common.cuh"

namespace vpdq {

__constant__ float c_dct[16 * 64];
__device__ float g_dct[16 * 64];  // the same table in global memory (coalesced fill of k5's shared copies)

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
// K5: quality + 64->16 DCT + median + bits from the decimated plane a64 [n][64][64] the fused kernels emit.
// One CTA (256 threads) per frame; same arithmetic and order as k4_colpass_finalize<true> (which stays as
// the A/B reference, VPDQ_B200_FINALIZE=k4), about half the instructions:
//   * T = D*A with two accumulator pairs per thread in packed FMUL2 / FADD2 (4 D values per 128-bit load
//     from a transposed table), B = T*D^T with 128-bit loads along k;
//   * quality: (u - down, u - right) as one packed pair, exact /255, and |trunc(x)| taken as the mantissa of
//     RZ(|x| + 2^23) -- the integer bit patterns are summed as they are (256 threads x 32 terms: the 2^23
//     exponent offsets cancel mod 2^32);
//   * median: #{B < v} only: the 128-th smallest value is the largest v with #{B < v} <= 127.
// ---------------------------------------------------------------------------------------------------
constexpr int kDJ = 68;  // pitch of the D / T rows: 128-bit k-chunks of 8 consecutive rows fall in different banks

__device__ __forceinline__ float2 fadd2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) { return __fmul2_rn(a, b); }

__global__ void __launch_bounds__(256) k5_finalize(const float* __restrict__ a64, uint8_t* __restrict__ hashes,
                                                   int32_t* __restrict__ quality, float* __restrict__ a64_dbg,
                                                   float* __restrict__ b16_dbg) {
    __shared__ __align__(16) float A[kDec][kDec];  // decimated 64x64 plane
    __shared__ __align__(16) float Dt[kDec][16];   // Dt[k][i] = D[i][k]
    __shared__ __align__(16) float Dj[16][kDJ];    // D[j][k]
    __shared__ __align__(16) float T[16][kDJ];     // D * A
    __shared__ __align__(16) float B[256];
    __shared__ unsigned g_sum;
    __shared__ int med_key;

    const int t = threadIdx.x;
    const int64_t f = blockIdx.x;

#pragma unroll
    for (int e = t; e < 16 * 64; e += 256) {
        Dj[e >> 6][e & 63] = __ldg(g_dct + e);
        Dt[e >> 4][e & 15] = __ldg(g_dct + (e & 15) * 64 + (e >> 4));  // consecutive lanes -> consecutive words of Dt
    }
    if (t == 0) {
        g_sum = 0u;
        med_key = INT_MIN;
    }
    {
        const float4* src = reinterpret_cast<const float4*>(a64 + (size_t)f * (kDec * kDec));
        float4* dst = reinterpret_cast<float4*>(&A[0][0]);
#pragma unroll
        for (int e = 0; e < 4; ++e) dst[t + 256 * e] = __ldg(src + t + 256 * e);
    }
    __syncthreads();

    if (a64_dbg)
        for (int e = t; e < kDec * kDec; e += 256) a64_dbg[(size_t)f * 4096 + e] = A[e >> 6][e & 63];

    // quality: sum of |trunc((u - v) * 100 / 255)| over vertical and horizontal neighbours; thread = column j,
    // 16 consecutive rows (an absent neighbour is replaced by u itself: difference 0, term 0)
    {
        const int j = t & 63, i0 = (t >> 6) * 16;
        const int jr = j < 63 ? j + 1 : j;
        const float c255 = 0.00392156886f;  // RN(1/255), see div255()
        unsigned g = 0u;
        float u = A[i0][j];
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const int i = i0 + r;
            const float down = A[i < 63 ? i + 1 : i][j];
            const float right = A[i][jr];
            const float2 x = fmul2(fadd2(make_float2(u, u), make_float2(-down, -right)), make_float2(100.0f, 100.0f));
            const float2 q0 = fmul2(x, make_float2(c255, c255));
            const float2 rr = __ffma2_rn(make_float2(-255.0f, -255.0f), q0, x);
            const float2 q = __ffma2_rn(rr, make_float2(c255, c255), q0);  // = x / 255.0f, correctly rounded
            g += __float_as_uint(__fadd_rz(fabsf(q.x), 8388608.0f)) + __float_as_uint(__fadd_rz(fabsf(q.y), 8388608.0f));
            u = down;
        }
        g = __reduce_add_sync(0xffffffffu, g);
        if ((t & 31) == 0) atomicAdd(&g_sum, g);
    }

    // T = D * A : thread -> column j, four rows i0..i0+3; sequential in k, separate multiply and add
    {
        const int j = t & 63, i0 = (t >> 6) * 4;
        // (products packed, sums scalar: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under
        //  --fmad=false, which would change the bits; it leaves this form alone -- checked in the SASS and by the
        //  parity tests)
        float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 16
        for (int k = 0; k < 64; ++k) {
            const float a = A[k][j];
            const float4 d = *reinterpret_cast<const float4*>(&Dt[k][i0]);
            const float2 p01 = fmul2(make_float2(d.x, d.y), make_float2(a, a));
            const float2 p23 = fmul2(make_float2(d.z, d.w), make_float2(a, a));
            acc[0] = fadd(acc[0], p01.x);
            acc[1] = fadd(acc[1], p01.y);
            acc[2] = fadd(acc[2], p23.x);
            acc[3] = fadd(acc[3], p23.y);
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
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const float4 tv = *reinterpret_cast<const float4*>(&T[i][4 * q]);
            const float4 dv = *reinterpret_cast<const float4*>(&Dj[j][4 * q]);
            acc = fadd(acc, fmul(tv.x, dv.x));
            acc = fadd(acc, fmul(tv.y, dv.y));
            acc = fadd(acc, fmul(tv.z, dv.z));
            acc = fadd(acc, fmul(tv.w, dv.w));
        }
        bv = acc;
        B[t] = acc;
    }
    __syncthreads();
    if (b16_dbg) b16_dbg[(size_t)f * 256 + t] = bv;

    // median = 128-th smallest of the 256 values (what Torben's method returns for n = 256) = the largest value v
    // with #{B < v} <= 127.  Keys: an order-preserving map float -> int (of v + 0.0f, so that -0 == +0).
    int key = __float_as_int(fadd(bv, 0.0f));
    key ^= (key >> 31) & 0x7fffffff;
    {
        int lt = 0;
        const float4* b4 = reinterpret_cast<const float4*>(B);
#pragma unroll 16
        for (int u = 0; u < 64; ++u) {
            const float4 b = b4[u];
            lt += (b.x < bv) + (b.y < bv) + (b.z < bv) + (b.w < bv);
        }
        const int cand = __reduce_max_sync(0xffffffffu, lt <= 127 ? key : INT_MIN);
        if ((t & 31) == 0) atomicMax(&med_key, cand);
    }
    __syncthreads();

    // bit k = 16 i + j = t  ->  byte t >> 3, bit t & 7: eight little-endian 32-bit ballots
    const unsigned word = __ballot_sync(0xffffffffu, key > med_key);
    if ((t & 31) == 0) reinterpret_cast<uint32_t*>(hashes)[f * 8 + (t >> 5)] = word;
    if (t == 0) {
        const int q = (int)(g_sum / 90u);
        quality[f] = q > 100 ? 100 : q;
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
static float h_dct[16 * 64];
static std::once_flag h_dct_once;

const float* pdq_host_dct() {
    std::call_once(h_dct_once, [] {
        // Appendix A step 6: scale rounded to fp32 first, product in double, stored as fp32
        const float scale = (float)sqrt(2.0 / 64.0);
        for (int i = 0; i < 16; i++)
            for (int j = 0; j < 64; j++)
                h_dct[i * 64 + j] = (float)(scale * cos((M_PI / 2 / 64.0) * (i + 1) * (2 * j + 1)));
    });
    return h_dct;
}

int pdq_upload_tables() {
    static std::mutex mu;
    static bool done[64] = {};
    int dev = 0;
    VPDQ_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    if (dev >= 0 && dev < 64 && done[dev]) return VPDQ_B200_OK;
    VPDQ_CUDA(cudaMemcpyToSymbol(c_dct, pdq_host_dct(), sizeof(h_dct)));
    VPDQ_CUDA(cudaMemcpyToSymbol(g_dct, pdq_host_dct(), sizeof(h_dct)));
    if (dev >= 0 && dev < 64) done[dev] = true;
    return VPDQ_B200_OK;
}

constexpr size_t kScratchPerFrame = ((size_t)2 * kPlane + (size_t)kDec * kDim) * sizeof(float);
constexpr int64_t kMaxChunk = 256;

size_t pdq_scratch_bytes(int64_t n_frames) {
    const int64_t c = n_frames < 1 ? 1 : (n_frames > kMaxChunk ? kMaxChunk : n_frames);
    return (size_t)c * kScratchPerFrame;
}

// which pipeline hashes RGB frames: the frame-pair fused kernel (default), the one-frame fused kernel
// (VPDQ_B200_PDQ_IMPL=fused) or the v1 line kernels (=lines); all are CUDA and bit-identical -- the switch
// exists for A/B measurements
static std::atomic<int> g_pdq_impl{-1};

int pdq_impl() {
    int v = g_pdq_impl.load(std::memory_order_relaxed);
    if (v < 0) {
        const char* e = getenv("VPDQ_B200_PDQ_IMPL");
        v = (e && strcmp(e, "lines") == 0) ? 0 : (e && strcmp(e, "fused") == 0) ? 1 : (e && strcmp(e, "fused2") == 0) ? 2 : 3;
        g_pdq_impl.store(v, std::memory_order_relaxed);
    }
    return v;
}

int pdq_set_impl(int impl) {
    if (impl < 0 || impl > 3) {
        set_error("set_pdq_impl: %d is not one of 0 (lines), 1 (fused), 2 (fused2), 3 (systolic)", impl);
        return VPDQ_B200_ERR_INVALID;
    }
    g_pdq_impl.store(impl, std::memory_order_relaxed);
    return VPDQ_B200_OK;
}

int pdq_timeout_flags_async(int* h_flags, cudaStream_t stream) {
    int rc = systolic_timeout_flag_async(h_flags + 0, stream);
    if (rc == 0) rc = fused2_timeout_flag_async(h_flags + 1, stream);
    if (rc == 0) rc = fused_timeout_flag_async(h_flags + 2, stream);
    return rc;
}

int pdq_force_timeout_flags(int value) {
    int rc = systolic_debug_force_timeout(value);
    if (rc == 0) rc = fused2_force_timeout(value);
    if (rc == 0) rc = fused_force_timeout(value);
    return rc;
}

int pdq_launch(const uint8_t* d_frames, int channels, int64_t n_frames, uint8_t* d_hashes, int32_t* d_quality,
               float* d_a64, float* d_b16, void* d_scratch, size_t scratch_bytes, cudaStream_t stream) {
    if (n_frames == 0) return VPDQ_B200_OK;
    // gray frames: the frame-pair fused kernel too (the one-frame fused kernel is RGB24 only -> line kernels)
    const int impl = (channels == 3 || pdq_impl() >= 2) ? pdq_impl() : 0;
    const bool fused = impl != 0;
    const size_t per_frame = fused ? fused_scratch_per_frame() : kScratchPerFrame;
    int64_t chunk = (int64_t)(scratch_bytes / per_frame);
    if (chunk < 1) {
        set_error("pdq: scratch too small (%zu bytes, need >= %zu)", scratch_bytes, per_frame);
        return VPDQ_B200_ERR_INVALID;
    }
    if (!fused && chunk > kMaxChunk) chunk = kMaxChunk;
    int rc = pdq_upload_tables();
    if (rc) return rc;

    const size_t frame_bytes = (size_t)kPlane * channels;
    for (int64_t f0 = 0; f0 < n_frames; f0 += chunk) {
        const int64_t nf = (n_frames - f0 < chunk) ? (n_frames - f0) : chunk;
        const uint8_t* src = d_frames + (size_t)f0 * frame_bytes;
        uint8_t* hp = d_hashes + (size_t)f0 * 32;
        int32_t* qp = d_quality + f0;
        float* adbg = d_a64 ? d_a64 + (size_t)f0 * 4096 : nullptr;
        float* bdbg = d_b16 ? d_b16 + (size_t)f0 * 256 : nullptr;
        if (fused) {
            float* a64 = static_cast<float*>(d_scratch);
            rc = impl == 3   ? systolic_jarosz_launch(src, channels, nf, a64, stream)
                 : impl == 2 ? fused2_jarosz_launch(src, channels, nf, a64, stream)
                             : fused_jarosz_launch(src, nf, a64, stream);
            if (rc) return rc;
            static const bool use_k4 = [] {  // A/B switch: the previous finalize kernel
                const char* e = getenv("VPDQ_B200_FINALIZE");
                return e && strcmp(e, "k4") == 0;
            }();
            if (use_k4)
                k4_colpass_finalize<true><<<(unsigned)nf, 256, 0, stream>>>(a64, hp, qp, adbg, bdbg);
            else
                k5_finalize<<<(unsigned)nf, 256, 0, stream>>>(a64, hp, qp, adbg, bdbg);
        } else {
            float* p1t = static_cast<float*>(d_scratch);
            float* p2 = p1t + (size_t)nf * kPlane;
            float* p3t = p2 + (size_t)nf * kPlane;
            const int64_t n_lines = nf * kDim;
            const unsigned grid = (unsigned)((n_lines + 127) / 128);
            if (channels == 3)
                k1_luma_rowpass<3><<<grid, 128, 0, stream>>>(src, p1t, n_lines);
            else
                k1_luma_rowpass<1><<<grid, 128, 0, stream>>>(src, p1t, n_lines);
            k2_colpass<<<grid, 128, 0, stream>>>(p1t, p2, n_lines);
            k3_rowpass_dec<<<grid, 128, 0, stream>>>(p2, p3t, n_lines);
            g_launches += 3;
            k4_colpass_finalize<false><<<(unsigned)nf, 256, 0, stream>>>(p3t, hp, qp, adbg, bdbg);
        }
        g_launches += 1;
    }
    VPDQ_CUDA(cudaGetLastError());
    return VPDQ_B200_OK;
}

}  // namespace vpdq
