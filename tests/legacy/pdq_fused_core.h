// pdq_fused_core.h -- per-lane arithmetic and the tile schedule of the fused PDQ kernel (kx_fused_jarosz).
//
// Compiled twice: by nvcc into the kernel (pdq_fused.cu) and by g++ into a CPU emulator
// (tests/emu/pdq_fused_emu.cpp) that executes the very same schedule, warp by warp and lane by lane,
// so that every index in here is checked against the oracle without a GPU.  The emulator is test
// infrastructure; the product only ever runs the CUDA build.
//
// Geometry.  A 512x512 frame is cut into 16 row BANDS x 16 column STRIPS of 32x32 tiles.  One CTA of
// 16 warps streams its frames through four chained roles; warp w plays the row roles for band w and the
// column roles for strip w:
//
//   P1 (row role, lane = row)     luma + row pass 1 over one tile   -> writes the tile into slot[T&1][w]
//   P2 (column role, lane = col)  column pass 1 over one tile, IN PLACE in slot[(T-1)&1][band]
//   P3 (row role, lane = row)     row pass 2 over one tile of slot[T&1][w] (read before P1 overwrites it);
//                                 emits only the 4 decimated columns 32*strip + {4,12,20,28} into t3[T&1]
//   P4 (column role, 4 columns)   column pass 2 over t3[(T-1)&1][w]; emits the 4 decimated rows
//                                 32*band + {4,12,20,28}: 16 values of the 64x64 plane per step
//
// Wavefront: tile (band b, strip s) of frame n is written by P1 at step 16n+b+s+1, filtered by P2 at +2,
// read by P3 at +3, and its 4x32 decimated slice is consumed by P4 at +4.  All four roles of a warp run in
// the SAME step on different tiles -- four independent dependent-add chains interleaved instruction by
// instruction (a GPU warp issues in order, so the interleaving is done here, in the source) -- and ONE CTA
// barrier separates steps.  Each warp walks its band left to right (row roles) and its strip top to bottom
// (column roles), one tile per step, which is the order the running sums need; chain state lives in
// registers across steps.
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
//   P3/P4 only need outputs 8m+4 <= 508, i.e. inputs <= 510: no prologue shift, no drains.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define VPDQ_HD __host__ __device__ __forceinline__
#else
#define VPDQ_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define VPDQ_UNROLL _Pragma("unroll")
#else
#define VPDQ_UNROLL
#endif

namespace vpdq_core {

#if defined(__CUDA_ARCH__)
VPDQ_HD float f_add(float a, float b) { return __fadd_rn(a, b); }
VPDQ_HD float f_sub(float a, float b) { return __fsub_rn(a, b); }
VPDQ_HD float f_mul(float a, float b) { return __fmul_rn(a, b); }
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
VPDQ_HD float f_fma(float a, float b, float c) { return fmaf(a, b, c); }
VPDQ_HD float bits_to_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
VPDQ_HD uint32_t byte_splice(uint32_t word, int k) { return 0x4B000000u | ((word >> (8 * k)) & 0xFFu); }
#endif

// v / 3.0f, correctly rounded, branch free (Markstein: q = RN(v*c), r = v - 3q exactly by FMA, q' = RN(q + r*c)
// with c = RN(1/3)).  Verified equal to IEEE v / 3.0f for EVERY finite positive float (tests/emu/div3_check.c).
VPDQ_HD float div3(float v) {
    const float c = 0.333333343267440796f;  // 0x3EAAAAAB
    const float q = f_mul(v, c);
    const float r = f_fma(-3.0f, q, v);
    return f_fma(r, c, q);
}

// two independent fp32 lanes in one register pair: on sm_100a add/sub map to the packed FADD2 instruction
// (same IEEE round-to-nearest result per component as two scalar FADDs, half the issue slots)
struct F2 {
    float x, y;
};
#if defined(__CUDA_ARCH__)
VPDQ_HD F2 f2_add(F2 a, F2 b) {
    const float2 r = __fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
    return F2{r.x, r.y};
}
VPDQ_HD F2 f2_sub(F2 a, F2 b) {
    const float2 r = __fadd2_rn(make_float2(a.x, a.y), make_float2(-b.x, -b.y));
    return F2{r.x, r.y};
}
VPDQ_HD F2 f2_fma(F2 a, F2 b, F2 c) {
    const float2 r = __ffma2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y), make_float2(c.x, c.y));
    return F2{r.x, r.y};
}
#else
VPDQ_HD F2 f2_fma(F2 a, F2 b, F2 c) { return F2{f_fma(a.x, b.x, c.x), f_fma(a.y, b.y, c.y)}; }
VPDQ_HD F2 f2_add(F2 a, F2 b) { return F2{f_add(a.x, b.x), f_add(a.y, b.y)}; }
VPDQ_HD F2 f2_sub(F2 a, F2 b) { return F2{f_sub(a.x, b.x), f_sub(a.y, b.y)}; }
#endif

constexpr int kTile = 32;
constexpr int kBands = 16;
constexpr int kRawPitch = 112;                   // bytes per staged row (7 x 16): bytes 6..101 used; LDS.128 conflict free
constexpr int kRawWords = kRawPitch / 4;         // 28
constexpr int kRawSkip = 6;                      // first used byte of a staged row
constexpr int kRawBoxBytes = kRawPitch * kTile;  // 3584 B per warp
constexpr int kT3Pitch = kTile + 4;               // floats per decimated column in t3 (+4: the 4 columns a warp reads land in different banks)
constexpr int kT3Strip = 4 * kT3Pitch;           // floats per strip in a t3 buffer: [q = 0..3][l = 0..31]

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

// two chains advanced together (x = P1's row chain, y = P2's column chain): 2 FADD2 per step instead of 4 FADD
struct Chain2 {
    F2 s, r0, r1, r2, r3;
    VPDQ_HD void init_x() { s.x = r0.x = r1.x = r2.x = r3.x = 0.0f; }
    VPDQ_HD void init_y() { s.y = r0.y = r1.y = r2.y = r3.y = 0.0f; }
    VPDQ_HD F2 feed(F2 v) {
        s = f2_add(s, v);
        s = f2_sub(s, r0);
        r0 = r1; r1 = r2; r2 = r3; r3 = v;
        return s;
    }
    VPDQ_HD void feed_x(float v) {  // prologue steps of the x chain alone
        s.x = f_sub(f_add(s.x, v), r0.x);
        r0.x = r1.x; r1.x = r2.x; r2.x = r3.x; r3.x = v;
    }
    VPDQ_HD void feed_y(float v) {
        s.y = f_sub(f_add(s.y, v), r0.y);
        r0.y = r1.y; r1.y = r2.y; r2.y = r3.y; r3.y = v;
    }
};

// full-resolution output scale at tile-local index k of tile number t (= strip for rows, band for columns):
// divisor 3 at global index 0 and 510, 2 at 511, else 4
VPDQ_HD float scale_edge(float v, int k, int t) {
    if (k == 0) return t == 0 ? div3(v) : f_mul(v, 0.25f);
    if (k == 30) return t == 15 ? div3(v) : f_mul(v, 0.25f);
    if (k == 31) return f_mul(v, t == 15 ? 0.5f : 0.25f);
    return f_mul(v, 0.25f);
}

VPDQ_HD float luma3(uint32_t wr, int kr, uint32_t wg, int kg, uint32_t wb, int kb) {
    // u8 -> fp32 product without an I2F: the byte is spliced into the mantissa of 2^23 (PRMT): M = 2^23 + b
    // exactly; fma(c, M, -c*2^23) = RN(c*b), bit-identical to __fmul_rn(c, (float)b)  (c*2^23 is exact)
    const float cr = 0.299f, cg = 0.587f, cb = 0.114f, two23 = 8388608.0f;
    const float r = f_fma(cr, bits_to_float(byte_splice(wr, kr)), -(cr * two23));
    const float g = f_fma(cg, bits_to_float(byte_splice(wg, kg)), -(cg * two23));
    const float b = f_fma(cb, bits_to_float(byte_splice(wb, kb)), -(cb * two23));
    return f_add(f_add(r, g), b);  // (0.299 R + 0.587 G) + 0.114 B
}

// luma of the pixel whose R byte sits at byte offset b0 of a little-endian word array
template <int N>
VPDQ_HD float luma_at(const uint32_t (&w)[N], int b0) {
    const int b1 = b0 + 1, b2 = b0 + 2;
    return luma3(w[b0 >> 2], b0 & 3, w[b1 >> 2], b1 & 3, w[b2 >> 2], b2 & 3);
}

// luma of the two consecutive pixels whose R bytes sit at byte offsets b0 and b0 + 3 (packed FFMA2 / FADD2:
// 5 instructions per pixel pair instead of 10, each component rounded exactly as in luma3)
template <int N>
VPDQ_HD F2 luma_pair_at(const uint32_t (&w)[N], int b0) {
    const float cr = 0.299f, cg = 0.587f, cb = 0.114f, two23 = 8388608.0f;
    const int r0 = b0, g0 = b0 + 1, bl0 = b0 + 2, r1 = b0 + 3, g1 = b0 + 4, bl1 = b0 + 5;
    const F2 mr{bits_to_float(byte_splice(w[r0 >> 2], r0 & 3)), bits_to_float(byte_splice(w[r1 >> 2], r1 & 3))};
    const F2 mg{bits_to_float(byte_splice(w[g0 >> 2], g0 & 3)), bits_to_float(byte_splice(w[g1 >> 2], g1 & 3))};
    const F2 mb{bits_to_float(byte_splice(w[bl0 >> 2], bl0 & 3)), bits_to_float(byte_splice(w[bl1 >> 2], bl1 & 3))};
    const F2 r = f2_fma(F2{cr, cr}, mr, F2{-(cr * two23), -(cr * two23)});
    const F2 g = f2_fma(F2{cg, cg}, mg, F2{-(cg * two23), -(cg * two23)});
    const F2 b = f2_fma(F2{cb, cb}, mb, F2{-(cb * two23), -(cb * two23)});
    return f2_add(f2_add(r, g), b);
}

// word index of element (row l, column c) inside a 32x32 fp32 tile: 16-byte chunks XOR-swizzled by the
// row so that lane=row float4 accesses and lane=column scalar accesses are both bank-conflict free
VPDQ_HD int tile_idx(int l, int c) { return l * kTile + ((((c >> 2) ^ (l & 7)) << 2) | (c & 3)); }

// ---- schedule: which tile a role of warp w works on at step T: u = T - role - w, role = 1..4 ------
// frame n = floor(u / 16) (CTA-local), tile index = u mod 16.  n = -1 is the virtual frame in front of the
// CTA's range: there only warp 15's P1 (it yields rows 0,1 of the first real frame) and every warp's P2 at
// band 15 (which stashes them) are live.
VPDQ_HD int sched_u(int T, int role, int w) { return T - role - w; }
VPDQ_HD int floor_div16(int u) { return u >> 4; }  // arithmetic shift: floor for negatives too
VPDQ_HD int num_steps(int n_frames_cta) { return 16 * n_frames_cta + 20; }
VPDQ_HD bool p1_live(int u, int w, int F) { return u < 16 * F && (u >= 0 || (w == 15 && u >= -16)); }
VPDQ_HD bool p2_live(int u, int F) { return u < 16 * F && u >= -1; }
VPDQ_HD bool p34_live(int u, int F) { return u < 16 * F && u >= 0; }
VPDQ_HD int p1_first(int w) { return w == 15 ? -16 : 0; }
// first image row (global, over the whole batch) fed by lane 0 of band w for CTA-local frame n
VPDQ_HD long long p1_row0(long long f_begin, int n, int w) { return (f_begin + n) * 512 + 32 * w + 2; }
VPDQ_HD int p1_box_x(int strip) { return 96 * strip; }  // 16-byte aligned; pixel 32*strip + 2 is at byte 6 of the box

struct LaneState {
    Chain2 c12;    // x: P1 row chain, y: P2 column chain
    Chain c3, c4;
    float p0, p1;  // P1 rows 0,1 of the next frame (this lane's column), carried from band 15
    VPDQ_HD void init() {
        c12.init_x(); c12.init_y(); c3.init(); c4.init();
        p0 = p1 = 0.0f;
    }
};

struct StepArgs {
    bool live1, live2, live3, live4;
    int s1, b2, s3, b4;     // P1 strip, P2 band, P3 strip, P4 band
    float* tile_a;          // slot[T&1][w]        P3 reads it, then P1 overwrites it
    float* tile_b;          // slot[(T-1)&1][b2]   P2, in place
    float* t3_w;            // t3[T&1] + s3*kT3Strip      P3 writes [q][lane]
    const float* t3_r;      // t3[(T-1)&1] + w*kT3Strip   P4 reads  [lane&3][0..31]
    float* a_out;           // a64 + frame4*4096 + 4*w + (lane&3); row m at a_out[m*64]; written by lanes < 4
};

// One lane's work for one step: the four roles, interleaved element by element.
// raw: the lane's staged RGB row (28 words, pixels from byte kRawSkip); first2: bytes 0..7 of that image row
// after_loads(): called once the step's up-front shared-memory loads have been issued (the kernel uses it to
// hand the staging buffer back to TMA a little later than the raw loads, hiding their latency)
template <typename Hook>
VPDQ_HD void fused_step(LaneState& st, const StepArgs& a, const uint32_t (&raw)[kRawWords],
                        const uint32_t (&first2)[2], int lane, Hook after_loads) {
    // ---- per-role prologues (warp-uniform conditions) ----
    if (a.s1 == 0) {  // new row: pixels 0,1 are fed without output
        st.c12.init_x();
        st.c12.feed_x(luma_at(first2, 0));
        st.c12.feed_x(luma_at(first2, 3));
    }
    if (a.b2 == 0) {  // new column: P1 rows 0,1 were stashed from the previous frame's band 15
        st.c12.init_y();
        st.c12.feed_y(st.p0);
        st.c12.feed_y(st.p1);
    }
    if (a.s3 == 0) st.c3.init();
    if (a.b4 == 0) st.c4.init();

    // P3's tile row is read ahead of P1's stores to the same row (P1 writes columns k-3..k at step k):
    // columns 0..15 up front, columns 16..31 at k = 12 (before P1 reaches column 16)
    float x3[32];
    VPDQ_UNROLL
    for (int q = 0; q < 4; ++q) {
        const float* src = a.tile_a + tile_idx(lane, 4 * q);
        x3[4 * q + 0] = src[0]; x3[4 * q + 1] = src[1]; x3[4 * q + 2] = src[2]; x3[4 * q + 3] = src[3];
    }
    after_loads();
    float n0 = 0.0f, n1 = 0.0f;
    float y1[4];
    const float* t3_lane = a.t3_r + (lane & 3) * kT3Pitch;
    float x4[4];

    VPDQ_UNROLL
    for (int k = 0; k < kTile; ++k) {
        if (k == 12) {
            VPDQ_UNROLL
            for (int q = 4; q < 8; ++q) {
                const float* src = a.tile_a + tile_idx(lane, 4 * q);
                x3[4 * q + 0] = src[0]; x3[4 * q + 1] = src[1]; x3[4 * q + 2] = src[2]; x3[4 * q + 3] = src[3];
            }
        }
        // P3: row pass 2, fed column 32*s3 + k -> output column 32*s3 + k - 2; keep k = 6, 14, 22, 30
        {
            const float v = st.c3.feed(x3[k]);
            if ((k & 7) == 6 && a.live3) a.t3_w[(k >> 3) * kT3Pitch + lane] = f_mul(v, 0.25f);
        }
        // P1 (x): luma + row pass 1, fed pixel 32*s1 + 2 + k -> output column 32*s1 + k
        // P2 (y): column pass 1 in place, fed P1 row 32*b2 + 2 + k -> output row 32*b2 + k
        {
            float* p = a.tile_b + tile_idx(k, lane);
            float x2 = *p;
            if (k == 30) n0 = x2;
            if (k == 31) n1 = x2;
            if (k >= 30 && a.b2 == 15) x2 = 0.0f;  // image rows 512, 513 do not exist: the two drain steps
            const F2 v = st.c12.feed(F2{luma_at(raw, kRawSkip + 3 * k), x2});
            y1[k & 3] = scale_edge(v.x, k, a.s1);
            if ((k & 3) == 3 && a.live1) {
                float* dst = a.tile_a + tile_idx(lane, k - 3);
                dst[0] = y1[0]; dst[1] = y1[1]; dst[2] = y1[2]; dst[3] = y1[3];
            }
            const float y2 = scale_edge(v.y, k, a.b2);
            if (a.live2) *p = y2;
        }
        // P4: column pass 2 on this strip's decimated column lane&3, fed row 32*b4 + k -> output row 32*b4 + k - 2
        {
            if ((k & 3) == 0) {
                const float* src = t3_lane + k;
                x4[0] = src[0]; x4[1] = src[1]; x4[2] = src[2]; x4[3] = src[3];
            }
            const float v = st.c4.feed(x4[k & 3]);
            if ((k & 7) == 6 && a.live4 && lane < 4) a.a_out[(size_t)(4 * a.b4 + (k >> 3)) * 64] = f_mul(v, 0.25f);
        }
    }
    if (a.b2 == 15 && a.live2) {  // rows 512, 513 of this frame = rows 0, 1 of the next one
        st.p0 = n0;
        st.p1 = n1;
    }
}

}  // namespace vpdq_core
