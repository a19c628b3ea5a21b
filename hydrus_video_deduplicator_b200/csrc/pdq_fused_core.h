// pdq_fused_core.h -- per-lane arithmetic and the tile schedule of the fused PDQ kernel (kx_fused_p123).
//
// Compiled twice: by nvcc into the kernel (pdq_fused.cu) and by g++ into a CPU emulator
// (tests/emu/pdq_fused_emu.cpp) that executes the very same schedule, warp by warp and lane by lane,
// so that every index in here is checked against the oracle without a GPU.  The emulator is test
// infrastructure; the product only ever runs the CUDA build.
//
// Geometry.  A 512x512 frame is cut into 16 row BANDS x 16 column STRIPS of 32x32 tiles.  One CTA of
// 16 warps streams its frames through three chained roles; warp w plays the row role for band w and the
// column role for strip w:
//
//   P1 (row role, lane = row)     luma + row pass 1 over one tile  -> writes the tile into slot[w]
//   P2 (column role, lane = col)  column pass 1 over one tile, IN PLACE in slot[band]
//   P3 (row role, lane = row)     row pass 2 over one tile of slot[w]; emits only the 4 decimated
//                                 columns 32*strip + {4, 12, 20, 28}
//
// Wavefront: tile (band b, strip s) of frame n is produced by P1 at step 16n + b + s + 1, consumed by P2
// in the same step (after a CTA barrier) and by P3 in the next one.  Each warp therefore walks its band
// left to right (row roles) and its strip top to bottom (column role), one tile per step, which is the
// order the running sums need; chain state lives in registers across steps.
//
// The 2-sample lag of the box filter (feeding x[r] yields the output for index r-2) is absorbed by
// shifting what is FED rather than what is produced, so every tile holds 32 aligned outputs:
//   columns: P1 feeds pixels 32s+2 .. 32s+33 = bytes 96s+6 .. 96s+101 of the row.  TMA needs a 16-byte
//            aligned box start (measured: tools/tma_probe.cu), so the box is bytes 96s .. 96s+111 and the
//            pixels start at byte 6 of every staged row; pixels 0,1 are a per-row prologue, pixels 512,513
//            are TMA out-of-bounds zeros (= the two drain steps);
//   rows:    band b of P1 holds image rows 32b+2 .. 32b+33; rows 512,513 of frame n ARE rows 0,1 of
//            frame n+1 in memory (frames are contiguous, a CTA owns a contiguous frame range), which
//            P2 stashes as the prologue of the next frame while feeding zeros (drains) to frame n.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define VPDQ_HD __host__ __device__ __forceinline__
#else
#define VPDQ_HD inline
#endif

namespace vpdq_core {

#if defined(__CUDA_ARCH__)
VPDQ_HD float f_add(float a, float b) { return __fadd_rn(a, b); }
VPDQ_HD float f_sub(float a, float b) { return __fsub_rn(a, b); }
VPDQ_HD float f_mul(float a, float b) { return __fmul_rn(a, b); }
VPDQ_HD float f_div(float a, float b) { return __fdiv_rn(a, b); }
VPDQ_HD float f_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
VPDQ_HD float bits_to_float(uint32_t u) { return __uint_as_float(u); }
VPDQ_HD uint32_t byte_splice(uint32_t word, int k) { return __byte_perm(word, 0x4B000000u, 0x7540u + k); }
#else
// host build (emulator): compile with -ffp-contract=off; fmaf() is a correctly rounded fused op
}  // namespace vpdq_core
#include <math.h>
namespace vpdq_core {
VPDQ_HD float f_add(float a, float b) { volatile float r = a + b; return r; }
VPDQ_HD float f_sub(float a, float b) { volatile float r = a - b; return r; }
VPDQ_HD float f_mul(float a, float b) { volatile float r = a * b; return r; }
VPDQ_HD float f_div(float a, float b) { volatile float r = a / b; return r; }
VPDQ_HD float f_fma(float a, float b, float c) { return fmaf(a, b, c); }
VPDQ_HD float bits_to_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
VPDQ_HD uint32_t byte_splice(uint32_t word, int k) { return 0x4B000000u | ((word >> (8 * k)) & 0xFFu); }
#endif

constexpr int kTile = 32;
constexpr int kBands = 16;
constexpr int kRawPitch = 112;            // bytes per staged row (7 x 16): bytes 6..101 used; the pitch also makes LDS.128 conflict free
constexpr int kRawWords = kRawPitch / 4;  // 28
constexpr int kRawSkip = 6;               // first used byte of a staged row
constexpr int kRawBoxBytes = kRawPitch * kTile;  // 3584 B per warp per stage

// running-sum box filter, window 4 (see pdq_kernels.cu / SURVEY.md Appendix A step 3)
struct Chain {
    float s, r0, r1, r2, r3;
    VPDQ_HD void init() { s = r0 = r1 = r2 = r3 = 0.0f; }
    VPDQ_HD float feed(float x) {
        s = f_add(s, x);
        s = f_sub(s, r0);
        r0 = r1; r1 = r2; r2 = r3; r3 = x;
        return s;
    }
};

// divisor by output index: 0 -> 3, 1..509 -> 4, 510 -> 3, 511 -> 2
VPDQ_HD float scale_out(float v, int o) {
    if (o == 0 || o == 510) return f_div(v, 3.0f);
    if (o == 511) return f_mul(v, 0.5f);
    return f_mul(v, 0.25f);
}

VPDQ_HD float luma3(uint32_t wr, int kr, uint32_t wg, int kg, uint32_t wb, int kb) {
    const float cr = 0.299f, cg = 0.587f, cb = 0.114f, two23 = 8388608.0f;
    const float r = f_fma(cr, bits_to_float(byte_splice(wr, kr)), -(cr * two23));
    const float g = f_fma(cg, bits_to_float(byte_splice(wg, kg)), -(cg * two23));
    const float b = f_fma(cb, bits_to_float(byte_splice(wb, kb)), -(cb * two23));
    return f_add(f_add(r, g), b);
}

// luma of the pixel whose R byte sits at byte offset b0 of a little-endian word array
template <int N>
VPDQ_HD float luma_at(const uint32_t (&w)[N], int b0) {
    const int b1 = b0 + 1, b2 = b0 + 2;
    return luma3(w[b0 >> 2], b0 & 3, w[b1 >> 2], b1 & 3, w[b2 >> 2], b2 & 3);
}

// word index of element (row l, column c) inside a 32x32 fp32 tile: 16-byte chunks XOR-swizzled by the
// row so that lane=row float4 accesses and lane=column scalar accesses are both bank-conflict free
VPDQ_HD int tile_idx(int l, int c) { return l * kTile + ((((c >> 2) ^ (l & 7)) << 2) | (c & 3)); }

// ---- schedule: which (frame n, tile index) a role of warp w works on at step T -------------------
// P1 and P2 share u = T - 1 - w ; P3 uses u = T - 2 - w ; n = floor(u / 16), index = u mod 16.
// n = -1 is the virtual frame in front of the CTA's range: only warp 15's P1 (it yields rows 0,1 of the
// first real frame) and every warp's P2 stash at band 15 are live there.
VPDQ_HD int sched_u12(int T, int w) { return T - 1 - w; }
VPDQ_HD int sched_u3(int T, int w) { return T - 2 - w; }
VPDQ_HD int floor_div16(int u) { return u >> 4; }  // arithmetic shift: floor for negatives too
VPDQ_HD int num_steps(int n_frames_cta) { return 16 * n_frames_cta + 17; }
VPDQ_HD bool p1_live(int u, int w, int F) { return u < 16 * F && (u >= 0 || (w == 15 && u >= -16)); }
VPDQ_HD bool p2_live(int u, int F) { return u < 16 * F && u >= -1; }
VPDQ_HD bool p3_live(int u, int F) { return u < 16 * F && u >= 0; }
// first image row (global, over the whole batch) fed by lane 0 of band w for CTA-local frame n
VPDQ_HD long long p1_row0(long long f_begin, int n, int w) { return (f_begin + n) * 512 + 32 * w + 2; }
VPDQ_HD int p1_box_x(int strip) { return 96 * strip; }  // 16-byte aligned; pixel 32*strip + 2 is at byte 6 of the box

// ---- P1: one lane (row) of one tile -------------------------------------------------------------
// raw: the lane's staged row, 28 words (pixels 32*strip+2 .. 32*strip+33 start at byte kRawSkip); tile: slot[w]
// first2: the row's first 8 bytes (pixels 0,1) -- only read when strip == 0
VPDQ_HD void p1_lane(Chain& ch, const uint32_t (&raw)[kRawWords], const uint32_t (&first2)[2], float* tile, int l,
                     int strip) {
    if (strip == 0) {
        ch.init();
        ch.feed(luma_at(first2, 0));
        ch.feed(luma_at(first2, 3));
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int g = 0; g < 2; ++g) {
        float y[16];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int k = 0; k < 16; ++k) {
            const int kk = 16 * g + k;
            const float v = ch.feed(luma_at(raw, kRawSkip + 3 * kk));
            // only output columns 0, 510, 511 deviate from the x0.25 scale: kk = 0 @ strip 0; 30, 31 @ strip 15
            if ((kk == 0 && strip == 0) || (kk >= 30 && strip == 15))
                y[k] = scale_out(v, 32 * strip + kk);
            else
                y[k] = f_mul(v, 0.25f);
        }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int q = 0; q < 4; ++q) {
            float* dst = tile + tile_idx(l, 16 * g + 4 * q);
            dst[0] = y[4 * q + 0]; dst[1] = y[4 * q + 1]; dst[2] = y[4 * q + 2]; dst[3] = y[4 * q + 3];
        }
    }
}

// ---- P2: one lane (column) of one tile, in place -------------------------------------------------
// p0/p1: P1 rows 0,1 of the NEXT frame, carried from band 15 of the previous one
VPDQ_HD void p2_lane(Chain& ch, float& p0, float& p1, float* tile, int c, int band, bool virtual_frame) {
    if (virtual_frame) {  // n = -1: nothing to filter, just pick up rows 0,1 of the first real frame
        p0 = tile[tile_idx(30, c)];
        p1 = tile[tile_idx(31, c)];
        return;
    }
    if (band == 0) {
        ch.init();
        ch.feed(p0);
        ch.feed(p1);
    }
    float n0 = 0.0f, n1 = 0.0f;
    if (band == 15) {
        n0 = tile[tile_idx(30, c)];
        n1 = tile[tile_idx(31, c)];
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int l = 0; l < kTile; ++l) {
        float x = tile[tile_idx(l, c)];
        if (band == 15 && l >= 30) x = 0.0f;  // image rows 512, 513 do not exist: the two drain steps
        const float v = ch.feed(x);
        float y;
        if ((l == 0 && band == 0) || (l >= 30 && band == 15))
            y = scale_out(v, 32 * band + l);
        else
            y = f_mul(v, 0.25f);
        tile[tile_idx(l, c)] = y;
    }
    if (band == 15) {
        p0 = n0;
        p1 = n1;
    }
}

// ---- P3: one lane (row) of one tile; emits 4 decimated outputs ------------------------------------
// out points at p3t[frame][0][row]; column jj lives at out[jj * 512]
VPDQ_HD void p3_lane(Chain& ch, const float* tile, int l, int strip, float* out) {
    if (strip == 0) ch.init();
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int q = 0; q < 8; ++q) {
        const float* src = tile + tile_idx(l, 4 * q);
        const float x[4] = {src[0], src[1], src[2], src[3]};
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < 4; ++j) {
            const float v = ch.feed(x[j]);
            const int kk = 4 * q + j;              // fed column 32*strip + kk -> output column 32*strip + kk - 2
            if ((kk & 7) == 6) out[(size_t)(4 * strip + (kk >> 3)) * 512] = f_mul(v, 0.25f);
        }
    }
}

}  // namespace vpdq_core
