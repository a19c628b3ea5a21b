// pdq_finalize.cuh -- the second half of the PDQ frame hash as a device function for 256 cooperating threads:
// the decimated 64x64 plane -> quality, 64->16 DCT (both directions), median, 256 bits.  Used by k5_finalize
// (pdq_kernels.cu).  The synchronisation of the 256 threads is a parameter: an experiment ran this function in eight
// extra "finalize warps" inside the persistent Jarosz kernel (register file re-split with setmaxnreg, planes handed over
// through flags in shared memory) -- bit-exact, but 15 % slower than the two kernels back to back: compiled against
// the 512-thread launch bound the Jarosz warps' code needs 12 % more instructions ("tried").
//
// Same arithmetic and order as the straightforward k4_colpass_finalize<true> kept in tests/legacy/pdq_lines.cu
// (cross-checked by the parity tests), about half the instructions:
//   * T = D*A with packed products (FMUL2; 4 D values per 128-bit load from a transposed table) and scalar sequential
//     sums, B = T*D^T with 128-bit loads along k.  (Products packed, sums scalar: ptxas contracts mul.rn.f32x2 +
//     add.rn.f32x2 into FFMA2 even under --fmad=false, which would change the bits; it leaves this form alone --
//     checked in the SASS and by the parity tests);
//   * quality: (u - down, u - right) as one packed pair, exact /255 (Markstein), and |trunc(x)| taken as the mantissa
//     of RZ(|x| + 2^23) -- the integer bit patterns are summed as they are (256 threads x 32 terms: the 2^23
//     exponent offsets cancel mod 2^32);
//   * median: #{B < v} only: the 128-th smallest value is the largest v with #{B < v} <= 127.
#pragma once
#include <limits.h>

#include "common.cuh"

namespace vpdq {

constexpr int kDJ = 68;  // pitch of the D / T rows: 128-bit k-chunks of 8 consecutive rows fall in different banks

struct FinalizeSmem {
    alignas(16) float A[kDec][kDec];  // decimated 64x64 plane
    alignas(16) float Dt[kDec][16];   // Dt[k][i] = D[i][k]
    alignas(16) float Dj[16][kDJ];    // D[j][k]
    alignas(16) float T[16][kDJ];     // D * A
    alignas(16) float B[256];
    unsigned g_sum;
    int med_key;
};

// fill the two shared copies of the 16 x 64 DCT table (once per CTA); t = 0..255
__device__ __forceinline__ void finalize_load_tables(FinalizeSmem& sm, const float* __restrict__ dct, int t) {
#pragma unroll
    for (int e = t; e < 16 * 64; e += 256) {
        sm.Dj[e >> 6][e & 63] = __ldg(dct + e);
        sm.Dt[e >> 4][e & 15] = __ldg(dct + (e & 15) * 64 + (e >> 4));  // consecutive lanes -> consecutive words of Dt
    }
}

// One frame.  plane: 4096 floats in global memory.  sync(): a barrier over the 256 threads.  The caller must sync() once more
// before the shared buffers are reused.
template <typename Sync>
__device__ __forceinline__ void finalize_frame(FinalizeSmem& sm, const float* __restrict__ plane, uint8_t* __restrict__ hash_out,
                                               int32_t* __restrict__ quality_out, float* __restrict__ a64_dbg,
                                               float* __restrict__ b16_dbg, int t, Sync sync) {
    if (t == 0) {
        sm.g_sum = 0u;
        sm.med_key = INT_MIN;
    }
    {
        const float4* src = reinterpret_cast<const float4*>(plane);
        float4* dst = reinterpret_cast<float4*>(&sm.A[0][0]);
#pragma unroll
        for (int e = 0; e < 4; ++e) dst[t + 256 * e] = __ldcg(src + t + 256 * e);
    }
    sync();

    if (a64_dbg)
        for (int e = t; e < kDec * kDec; e += 256) a64_dbg[e] = sm.A[e >> 6][e & 63];

    // quality: sum of |trunc((u - v) * 100 / 255)| over vertical and horizontal neighbours; thread = column j,
    // 16 consecutive rows (an absent neighbour is replaced by u itself: difference 0, term 0)
    {
        const int j = t & 63, i0 = (t >> 6) * 16;
        const int jr = j < 63 ? j + 1 : j;
        const float c255 = 0.00392156886f;  // RN(1/255)
        unsigned g = 0u;
        float u = sm.A[i0][j];
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const int i = i0 + r;
            const float down = sm.A[i < 63 ? i + 1 : i][j];
            const float right = sm.A[i][jr];
            const float2 x = __fmul2_rn(__fadd2_rn(make_float2(u, u), make_float2(-down, -right)), make_float2(100.0f, 100.0f));
            const float2 q0 = __fmul2_rn(x, make_float2(c255, c255));
            const float2 rr = __ffma2_rn(make_float2(-255.0f, -255.0f), q0, x);
            const float2 q = __ffma2_rn(rr, make_float2(c255, c255), q0);  // = x / 255.0f, correctly rounded
            g += __float_as_uint(__fadd_rz(fabsf(q.x), 8388608.0f)) + __float_as_uint(__fadd_rz(fabsf(q.y), 8388608.0f));
            u = down;
        }
        g = __reduce_add_sync(0xffffffffu, g);
        if ((t & 31) == 0) atomicAdd(&sm.g_sum, g);
    }

    // T = D * A : thread -> column j, four rows i0..i0+3; sequential in k, separate multiply and add
    {
        const int j = t & 63, i0 = (t >> 6) * 4;
        float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 16
        for (int k = 0; k < 64; ++k) {
            const float a = sm.A[k][j];
            const float4 d = *reinterpret_cast<const float4*>(&sm.Dt[k][i0]);
            const float2 p01 = __fmul2_rn(make_float2(d.x, d.y), make_float2(a, a));
            const float2 p23 = __fmul2_rn(make_float2(d.z, d.w), make_float2(a, a));
            acc[0] = fadd(acc[0], p01.x);
            acc[1] = fadd(acc[1], p01.y);
            acc[2] = fadd(acc[2], p23.x);
            acc[3] = fadd(acc[3], p23.y);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) sm.T[i0 + u][j] = acc[u];
    }
    sync();

    // B = T * D^T : thread -> (i, j)
    float bv;
    {
        const int i = t >> 4, j = t & 15;
        float acc = 0.0f;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const float4 tv = *reinterpret_cast<const float4*>(&sm.T[i][4 * q]);
            const float4 dv = *reinterpret_cast<const float4*>(&sm.Dj[j][4 * q]);
            acc = fadd(acc, fmul(tv.x, dv.x));
            acc = fadd(acc, fmul(tv.y, dv.y));
            acc = fadd(acc, fmul(tv.z, dv.z));
            acc = fadd(acc, fmul(tv.w, dv.w));
        }
        bv = acc;
    }
    if (b16_dbg) b16_dbg[t] = bv;

    // median = 128-th smallest of the 256 values (what Torben's method returns for n = 256) = the largest value v
    // with #{B < v} <= 127.  #{B < v} per thread without the 256 x 256 comparisons: every warp sorts its 32 values
    // (bitonic network over shuffles) and publishes the list; a thread then counts, per list, the values below its
    // own with a branch-free lower bound -- the list in the warp's registers, probed with index shuffles.
    // Keys: an order-preserving map float -> int (of v + 0.0f, so that -0 == +0).
    int key = __float_as_int(fadd(bv, 0.0f));
    key ^= (key >> 31) & 0x7fffffff;
    {
        const int lane = t & 31;
        float sv = bv;
#pragma unroll
        for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
            for (int j = k >> 1; j >= 1; j >>= 1) {
                const float o = __shfl_xor_sync(0xffffffffu, sv, j);
                const bool take_min = ((lane & k) == 0) == ((lane & j) == 0);
                sv = take_min ? fminf(sv, o) : fmaxf(sv, o);
            }
        }
        sm.B[t] = sv;  // [warp][rank in the warp]
        sync();
        int lt = 0;
#pragma unroll
        for (int wl = 0; wl < 8; ++wl) {
            const float lv = sm.B[32 * wl + lane];
            int pos = 0;
#pragma unroll
            for (int step = 16; step >= 1; step >>= 1) {
                const float e = __shfl_sync(0xffffffffu, lv, pos + step - 1);
                pos += e < bv ? step : 0;
            }
            const float e = __shfl_sync(0xffffffffu, lv, pos);
            lt += pos + (e < bv ? 1 : 0);
        }
        const int cand = __reduce_max_sync(0xffffffffu, lt <= 127 ? key : INT_MIN);
        if ((t & 31) == 0) atomicMax(&sm.med_key, cand);
    }
    sync();

    // bit k = 16 i + j = t  ->  byte t >> 3, bit t & 7: eight little-endian 32-bit ballots
    const unsigned word = __ballot_sync(0xffffffffu, key > sm.med_key);
    if ((t & 31) == 0) reinterpret_cast<uint32_t*>(hash_out)[t >> 5] = word;
    if (t == 0) {
        const int q = (int)(sm.g_sum / 90u);
        *quality_out = q > 100 ? 100 : q;
    }
}

}  // namespace vpdq
