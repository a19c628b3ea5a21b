// pdq_fused2_core.h -- per-lane arithmetic and the tile schedule of the frame-PAIR fused PDQ kernel
// (kx_fused_jarosz2, pdq_fused2.cu): every lane carries the same line of TWO frames in the two halves of a
// 64-bit register pair, so each running-sum step, each luma product and each shared-memory access is one
// packed instruction (FADD2 / FFMA2 / FMUL2, LDS.64 / STS.64 / .128) for both frames.  Rounding is per
// component, IEEE round-to-nearest -- bit-identical to the scalar op order of the CPU oracle.
//
// Compiled twice: by nvcc into the kernel and by g++ into the CPU emulator (tests/emu/pdq_fused2_emu.cpp)
// that executes the very same schedule warp by warp and lane by lane, so every index in here is checked
// against the oracle without a GPU.  The emulator is test infrastructure; the product runs the CUDA build.
//
// Geometry.  A 512x512 frame = 16 row BANDS x 16 column STRIPS of 32x32 tiles.  A CTA owns a contiguous frame
// range, split in two halves A = [f0, f0+FA) and B = [f0+FA, f0+FA+FB), FA = ceil(F/2); pair n = (A_n, B_n).
// Eight MAIN warps run three chained roles, one tile per role per step; warp w owns bands w and w+8 for the
// row roles and strips w and w+8 for the column roles:
//
//   P1 (row role, lane = row)     luma + row pass 1 over one tile            -> slot[T&1][w]
//   P2 (column role, lane = col)  column pass 1 over one tile, IN PLACE in slot[(T-1)&1][band & 7]
//   P3 (row role, lane = row)     row pass 2 over slot[T&1][w] (each chunk read before P1 overwrites it);
//                                 emits only the 4 decimated columns 32*strip + {4,12,20,28} -> t3[T&1]
// and a ninth warp runs
//   P4 (lane = one of the 8 strips x 4 decimated columns live in the step) column pass 2 over t3[(T-1)&1],
//                                 emitting the decimated rows 32*band + {4,12,20,28}: the 64x64 plane.
//
// Schedule: role r of warp w works at step T on u = T - r - w.  Row roles decode u as pair n = u >> 5,
// half h = (u >> 4) & 1, strip = u & 15, band = 8h + w: a warp walks band w left to right (16 steps), then
// band w + 8, then the next pair.  Column roles decode u as n = u >> 5, bh = (u >> 4) & 1, sh = (u >> 3) & 1,
// bl = u & 7: band = 8 bh + bl, strip = 8 sh + w: 8 bands down strip w, 8 bands down strip w + 8, then the
// lower halves -- exactly one step behind the row role that produced each tile.  A warp therefore carries
// TWO column-chain states (strips w, w+8) and swaps them every 8 steps.  One CTA barrier per step.
//
// Deferred scaling.  Interior outputs of the box filter are s/4; powers of two commute exactly with fp32
// rounding (no overflow / subnormals here: |values| < 2^17, all multiples of 2^-40), so the planes are kept
// UNSCALED (x4 after P1, x16 after P2, x64 after P3, x256 after P4) and the single multiply by 2^-8 happens
// on the 4096 emitted values.  Edge outputs (divisors 3, 3, 2 at indices 0, 510, 511) become 4*div3(s),
// 4*div3(s), 2*s.  This removes one multiply per element per pass from the inner loop.
//
// The 2-sample lag of the box filter (feeding x[r] yields the output for index r-2) is absorbed by shifting
// what is FED, as in pdq_fused_core.h: P1 is fed pixels 32s+2.. (RGB24: TMA box at byte 96s, pixels from byte 6;
// gray: box at byte 32s, pixels from byte 2; pixels 0,1 = the bytes in front of them in strip 0's box = the
// per-row prologue; pixels 512,513 = TMA out-of-bounds zeros = the drain steps) and its bands hold image rows
// 32b+2..32b+33, rows 512,513 of a frame being rows 0,1 of the next frame of the same half, which P2 stashes as
// that frame's prologue while feeding zeros to the current one.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define VPDQ2_HD __host__ __device__ __forceinline__
#else
#define VPDQ2_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define VPDQ2_UNROLL _Pragma("unroll")
#else
#define VPDQ2_UNROLL
#endif

namespace vpdq_core2 {

struct alignas(8) F2 {
    float x, y;  // x: frame of half A, y: frame of half B
};
struct alignas(16) F4 {
    F2 lo, hi;
};

#if defined(__CUDA_ARCH__)
VPDQ2_HD F2 f2_add(F2 a, F2 b) {
    const float2 r = __fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
    return F2{r.x, r.y};
}
VPDQ2_HD F2 f2_sub(F2 a, F2 b) {
    const float2 r = __fadd2_rn(make_float2(a.x, a.y), make_float2(-b.x, -b.y));
    return F2{r.x, r.y};
}
VPDQ2_HD F2 f2_mul(F2 a, F2 b) {
    const float2 r = __fmul2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
    return F2{r.x, r.y};
}
VPDQ2_HD F2 f2_fma(F2 a, F2 b, F2 c) {
    const float2 r = __ffma2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y), make_float2(c.x, c.y));
    return F2{r.x, r.y};
}
VPDQ2_HD float bits_to_float(uint32_t u) { return __uint_as_float(u); }
VPDQ2_HD uint32_t byte_splice(uint32_t word, int k) { return __byte_perm(word, 0x4B000000u, 0x7540u + k); }
#else
}  // namespace vpdq_core2
#include <math.h>
namespace vpdq_core2 {
// host build (emulator): compile with -ffp-contract=off; fmaf() is a correctly rounded fused op
inline float h_add(float a, float b) { volatile float r = a + b; return r; }
inline float h_sub(float a, float b) { volatile float r = a - b; return r; }
inline float h_mul(float a, float b) { volatile float r = a * b; return r; }
VPDQ2_HD F2 f2_add(F2 a, F2 b) { return F2{h_add(a.x, b.x), h_add(a.y, b.y)}; }
VPDQ2_HD F2 f2_sub(F2 a, F2 b) { return F2{h_sub(a.x, b.x), h_sub(a.y, b.y)}; }
VPDQ2_HD F2 f2_mul(F2 a, F2 b) { return F2{h_mul(a.x, b.x), h_mul(a.y, b.y)}; }
VPDQ2_HD F2 f2_fma(F2 a, F2 b, F2 c) { return F2{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)}; }
VPDQ2_HD float bits_to_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
VPDQ2_HD uint32_t byte_splice(uint32_t word, int k) { return 0x4B000000u | ((word >> (8 * k)) & 0xFFu); }
#endif

VPDQ2_HD F2 f2_splat(float v) { return F2{v, v}; }

// v / 3.0f per component, correctly rounded, branch free (Markstein: q = RN(v*c), r = v - 3q exactly by FMA,
// q' = RN(q + r*c), c = RN(1/3)); equal to IEEE v / 3.0f for EVERY finite positive float
// (tests/emu/div3_check.c), and invariant under power-of-two scaling of v.
VPDQ2_HD F2 div3(F2 v) {
    const F2 c = f2_splat(0.333333343267440796f);  // 0x3EAAAAAB
    const F2 q = f2_mul(v, c);
    const F2 r = f2_fma(f2_splat(-3.0f), q, v);
    return f2_fma(r, c, q);
}

constexpr int kMainWarps = 8;
constexpr int kWarps = kMainWarps + 1;            // + the P4 warp
constexpr int kTile = 32;
// staged raw rows, by channel count CH (3: RGB24, 1: 8-bit gray == R = G = B): a step's 32 pixels plus the two
// in front of them, rounded up to the 16-byte granularity of a TMA box
template <int CH>
struct Raw {
    static constexpr int kPitch = CH == 3 ? 112 : 48;    // bytes per staged row: bytes 2*CH .. 34*CH - 1 used; at both
                                                          // pitches the lane = row 128-bit reads are bank-conflict free
    static constexpr int kWords = kPitch / 4;            // 28 / 12
    static constexpr int kSkip = 2 * CH;                 // first byte of the step's 32 pixels in a staged row
    static constexpr int kBoxBytes = kPitch * 32;        // 3584 / 1536 B per warp per frame
    static constexpr int kStripBytes = 32 * CH;          // 16-byte aligned box start of strip s: kStripBytes * s
};
constexpr int kRawBoxBytesMax = Raw<3>::kBoxBytes;
constexpr int kPitch = kTile + 2;                 // F2 per tile row (272 B = 17 x 16 B): lane=row 128-bit and
                                                  // lane=column 64-bit accesses are both bank-conflict free
constexpr int kSlotF2 = kTile * kPitch;           // 1088 F2 = 8704 B per tile slot
constexpr int kT3Pitch = kTile + 4;               // F2 per decimated column in t3 (+4: the 4 columns of a strip
                                                  // read by the P4 warp land in different banks)
constexpr int kT3Strip = 4 * kT3Pitch + 2;        // F2 per strip in a t3 buffer: [q = 0..3][row 0..31]; +2 (16 B): the
                                                  // P4 warp's 8 strips x 4 columns 128-bit reads are conflict free
constexpr int kTStart = -8;                       // first step (warp 7's P1 on the virtual pair -1)

// running-sum box filter, window 4, two frames at once: feeding x[r] returns
// the (unscaled) window sum whose output index is r - 2
struct Chain2 {
    F2 s, r0, r1, r2, r3;
    VPDQ2_HD void init() { s = r0 = r1 = r2 = r3 = F2{0.0f, 0.0f}; }
    VPDQ2_HD F2 feed(F2 v) {
        s = f2_add(s, v);
        s = f2_sub(s, r0);
        r0 = r1; r1 = r2; r2 = r3; r3 = v;
        return s;
    }
};

// full-resolution output at tile-local index k of tile number t (= strip for rows, band for columns), in the
// deferred-scale representation (4x the true output): divisor 3 at global index 0 and 510, 2 at 511, else 4
VPDQ2_HD F2 edge_scale(F2 v, int k, int t) {
    if (k == 0 && t == 0) return f2_mul(div3(v), f2_splat(4.0f));
    if (k == 30 && t == 15) return f2_mul(div3(v), f2_splat(4.0f));
    if (k == 31 && t == 15) return f2_mul(v, f2_splat(2.0f));
    return v;
}

// luma of the pixel whose first byte sits at byte offset b0 of the two frames' little-endian word arrays.
// u8 -> fp32 product without an I2F: the byte is spliced into the mantissa of 2^23 (PRMT): M = 2^23 + b exactly;
// fma(c, M, -c*2^23) = RN(c*b), bit-identical to __fmul_rn(c, (float)b)  (c*2^23 is exact).
// CH == 1: gray == R = G = B, the same three-term expression on one byte (SURVEY.md 8 note a-1).
template <int CH, int N>
VPDQ2_HD F2 luma_pair_at(const uint32_t (&wa)[N], const uint32_t (&wb)[N], int b0) {
    const float cr = 0.299f, cg = 0.587f, cb = 0.114f, two23 = 8388608.0f;
    const int b1 = CH == 3 ? b0 + 1 : b0, b2 = CH == 3 ? b0 + 2 : b0;
    const F2 mr{bits_to_float(byte_splice(wa[b0 >> 2], b0 & 3)), bits_to_float(byte_splice(wb[b0 >> 2], b0 & 3))};
    const F2 mg = CH == 3 ? F2{bits_to_float(byte_splice(wa[b1 >> 2], b1 & 3)), bits_to_float(byte_splice(wb[b1 >> 2], b1 & 3))} : mr;
    const F2 mb = CH == 3 ? F2{bits_to_float(byte_splice(wa[b2 >> 2], b2 & 3)), bits_to_float(byte_splice(wb[b2 >> 2], b2 & 3))} : mr;
    const F2 r = f2_fma(f2_splat(cr), mr, f2_splat(-(cr * two23)));
    const F2 g = f2_fma(f2_splat(cg), mg, f2_splat(-(cg * two23)));
    const F2 b = f2_fma(f2_splat(cb), mb, f2_splat(-(cb * two23)));
    return f2_add(f2_add(r, g), b);  // (0.299 R + 0.587 G) + 0.114 B
}

// ---- schedule ---------------------------------------------------------------------------------------
VPDQ2_HD int sched_u(int T, int role, int w) { return T - role - w; }
VPDQ2_HD int u_pair(int u) { return u >> 5; }  // arithmetic shift: floor for negatives too
// row roles (P1, P3) of main warp w
VPDQ2_HD int row_band(int u, int w) { return 8 * ((u >> 4) & 1) + w; }
VPDQ2_HD int row_strip(int u) { return u & 15; }
// column roles (P2 of main warp w, P4 of lane group w)
VPDQ2_HD int col_band(int u) { return 8 * ((u >> 4) & 1) + (u & 7); }
VPDQ2_HD int col_strip(int u, int w) { return 8 * ((u >> 3) & 1) + w; }
VPDQ2_HD bool col_swap(int u) { return (u & 7) == 0; }  // the column role changes strip: swap the chain states
// last step (inclusive) for FA pairs
VPDQ2_HD int last_step(int FA) { return 32 * FA + 10; }
// pair -1 is virtual: there only warp 7's P1 on band 15 (it yields rows 0,1 of the first real frames) and every
// warp's P2 at band 15 (which stashes them) are live
VPDQ2_HD bool p1_live(int u, int w, int FA) { return u < 32 * FA && (u >= 0 || (w == 7 && u >= -16)); }
VPDQ2_HD bool p2_live(int u, int FA) { return u < 32 * FA && (u >= 0 || u == -9 || u == -1); }
VPDQ2_HD bool p34_live(int u, int FA) { return u < 32 * FA && u >= 0; }
VPDQ2_HD int p1_first(int w) { return w == 7 ? -16 : 0; }
// first image row (global, over the whole batch) fed by lane 0 of main warp w for tile u of the half whose
// first frame is `half_begin`
VPDQ2_HD long long p1_row0(long long half_begin, int u, int w) {
    return (half_begin + u_pair(u)) * 512 + 32 * row_band(u, w) + 2;
}
template <int CH>
VPDQ2_HD int p1_box_x(int u) { return Raw<CH>::kStripBytes * row_strip(u); }  // 16-byte aligned; pixel 32*strip + 2 is at byte kSkip

struct LaneState {
    Chain2 c1, c2, c2_parked, c3;
    F2 p0, p1, p0_parked, p1_parked;  // P1 rows 0,1 of the next frame (this lane's column), carried from band 15
    VPDQ2_HD void init() {
        c1.init(); c2.init(); c2_parked.init(); c3.init();
        p0 = p1 = p0_parked = p1_parked = F2{0.0f, 0.0f};
    }
};

struct StepArgs {
    bool live1, live2, live3, swap2;
    int s1, b2, s3;  // P1 strip, P2 band, P3 strip
    F2* tile_a;      // slot[T&1][w]               P3 reads it, P1 overwrites it
    F2* tile_b;      // slot[(T-1)&1][b2 & 7]      P2, in place
    F2* t3_w;        // t3[T&1] + (s3 & 7)*kT3Strip   P3 writes [q][lane]
};

VPDQ2_HD void swap_chain(Chain2& a, Chain2& b) { const Chain2 t = a; a = b; b = t; }
VPDQ2_HD void swap_f2(F2& a, F2& b) { const F2 t = a; a = b; b = t; }

// One lane's work for one step of a main warp: the three roles, interleaved element by element.
// raw_a / raw_b: the lane's staged rows of the two frames (Raw<CH>::kWords words; the step's 32 pixels start at
// byte kSkip; the bytes in front are the two previous pixels -- for strip 0 that is pixels 0, 1 of the image row,
// the prologue of the running sum, so no separate load is needed for them).
// (The kernel pulls them into registers at the END of the previous step and hands the staging buffers back to
// TMA at once, so that the copy for the step after next has a whole step to land.)
// probe(): called at k == kProbeAt; the kernel polls (without blocking) whether the NEXT step's boxes have landed,
// so that the poll's latency is hidden behind the rest of the step.
constexpr int kProbeAt = 26;
template <int CH, typename Probe>
VPDQ2_HD void main_step(LaneState& st, const StepArgs& a, const uint32_t (&raw_a)[Raw<CH>::kWords],
                        const uint32_t (&raw_b)[Raw<CH>::kWords], int lane, Probe probe) {
    // ---- per-role prologues (warp-uniform conditions) ----
    if (a.swap2) {
        swap_chain(st.c2, st.c2_parked);
        swap_f2(st.p0, st.p0_parked);
        swap_f2(st.p1, st.p1_parked);
    }
    if (a.s1 == 0) {  // new row: pixels 0,1 are fed without output
        st.c1.init();
        st.c1.feed(luma_pair_at<CH>(raw_a, raw_b, 0));
        st.c1.feed(luma_pair_at<CH>(raw_a, raw_b, CH));
    }
    if (a.b2 == 0) {  // new column: P1 rows 0,1 were stashed from the previous frame's band 15
        st.c2.init();
        st.c2.feed(st.p0);
        st.c2.feed(st.p1);
    }
    if (a.s3 == 0) st.c3.init();

    F2* row_a = a.tile_a + lane * kPitch;  // this lane's tile row: P3 reads each 4-column chunk before P1 overwrites it
    F2* col_b = a.tile_b + lane;           // this lane's tile column; row k at col_b[k * kPitch]
    F2 x3[2][4];
    {
        const F4 v0 = *reinterpret_cast<const F4*>(row_a), v1 = *reinterpret_cast<const F4*>(row_a + 2);
        x3[0][0] = v0.lo; x3[0][1] = v0.hi; x3[0][2] = v1.lo; x3[0][3] = v1.hi;
    }
    F2 x2_q[2] = {col_b[0], col_b[kPitch]};  // P2's input, fetched two rows ahead
    F2 n0{0.0f, 0.0f}, n1{0.0f, 0.0f};
    F2 y1[4];

    VPDQ2_UNROLL
    for (int k = 0; k < kTile; ++k) {
        if (k == kProbeAt) probe();
        if ((k & 3) == 0 && k + 4 < kTile) {  // P3's next chunk (columns k+4..k+7): P1 stores there at k+7
            const F4 v0 = *reinterpret_cast<const F4*>(row_a + k + 4), v1 = *reinterpret_cast<const F4*>(row_a + k + 6);
            F2(&dst)[4] = x3[((k >> 2) + 1) & 1];
            dst[0] = v0.lo; dst[1] = v0.hi; dst[2] = v1.lo; dst[3] = v1.hi;
        }
        // P3: row pass 2, fed column 32*s3 + k -> output column 32*s3 + k - 2; keep k = 6, 14, 22, 30
        {
            const F2 v = st.c3.feed(x3[(k >> 2) & 1][k & 3]);
            if ((k & 7) == 6 && a.live3) a.t3_w[(k >> 3) * kT3Pitch + lane] = v;
        }
        // P1: luma + row pass 1, fed pixel 32*s1 + 2 + k -> output column 32*s1 + k
        {
            const F2 v = st.c1.feed(luma_pair_at<CH>(raw_a, raw_b, Raw<CH>::kSkip + CH * k));
            y1[k & 3] = edge_scale(v, k, a.s1);
            if ((k & 3) == 3 && a.live1) {
                *reinterpret_cast<F4*>(row_a + k - 3) = F4{y1[0], y1[1]};
                *reinterpret_cast<F4*>(row_a + k - 1) = F4{y1[2], y1[3]};
            }
        }
        // P2: column pass 1 in place, fed P1 row 32*b2 + 2 + k -> output row 32*b2 + k
        {
            F2 x2 = x2_q[k & 1];
            if (k + 2 < kTile) x2_q[k & 1] = col_b[(k + 2) * kPitch];
            if (k == 30) n0 = x2;
            if (k == 31) n1 = x2;
            if (k >= 30 && a.b2 == 15) x2 = F2{0.0f, 0.0f};  // image rows 512, 513 do not exist: the two drain steps
            const F2 v = st.c2.feed(x2);
            const F2 y2 = edge_scale(v, k, a.b2);
            if (a.live2) col_b[k * kPitch] = y2;
        }
    }
    if (a.b2 == 15 && a.live2) {  // rows 512, 513 of this frame = rows 0, 1 of the next one of the same half
        st.p0 = n0;
        st.p1 = n1;
    }
}

// ---- the P4 warp: lane = (group g = lane >> 2 playing column-role warp g, decimated column q = lane & 3) ----
struct P4State {
    Chain2 c4, c4_parked;
    VPDQ2_HD void init() { c4.init(); c4_parked.init(); }
};

struct P4Args {
    bool live_a, live_b, swap4;
    int b4;            // band
    const F2* t3_r;    // t3[(T-1)&1] + g*kT3Strip + q*kT3Pitch: rows 0..31 of this lane's decimated column
    float* out_a;      // a64 + frame_a*4096 + 4*strip + q ; decimated row m at out_a[m * 64]
    float* out_b;
};

VPDQ2_HD void p4_step(P4State& st, const P4Args& a) {
    if (a.swap4) swap_chain(st.c4, st.c4_parked);
    if (a.b4 == 0) st.c4.init();
    VPDQ2_UNROLL
    for (int k = 0; k < kTile; k += 2) {
        const F4 x = *reinterpret_cast<const F4*>(a.t3_r + k);
        // column pass 2, fed row 32*b4 + k -> output row 32*b4 + k - 2; keep k = 6, 14, 22, 30 (all even)
        const F2 v = st.c4.feed(x.lo);
        if ((k & 7) == 6) {
            const F2 y = f2_mul(v, f2_splat(0.00390625f));  // the deferred 4^-4
            const int m = 4 * a.b4 + (k >> 3);
            if (a.live_a) a.out_a[m * 64] = y.x;
            if (a.live_b) a.out_b[m * 64] = y.y;
        }
        st.c4.feed(x.hi);
    }
}

}  // namespace vpdq_core2
