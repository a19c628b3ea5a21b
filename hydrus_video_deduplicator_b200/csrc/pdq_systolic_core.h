// pdq_systolic_core.h -- per-lane arithmetic and the schedule of the WARP-PER-FRAME systolic PDQ kernel
// (kx_systolic_jarosz, pdq_systolic.cu): luma + the four Jarosz box-filter passes + the 64x64 decimation with
// NO shared-memory transposition between the passes.
//
// Why.  The Jarosz filter is a running sum (s += x[r]; s -= x[l]; y = s / n), so every output carries the
// rounding history of its whole line prefix and each line is an inherently serial fp32 chain.
// The tiled kernels (pdq_fused*.cu) gave every lane whole lines and therefore had to transpose the plane
// through shared memory between the row and the column passes: 4 tile crossings + 2 raw crossings, 5.6 MB per
// frame through a 128 B/clk pipe -- the measured bound of those kernels ("What's weak" 2).
// Here the data never moves; the CHAIN STATE does:
//
//   * one warp owns a frame; lane l owns image columns 16 l .. 16 l + 15 for all four passes;
//   * column passes (P2, P4) are private to a lane: 16 (resp. 2) independent running sums in registers,
//     fed one image row per step;
//   * row passes (P1, P3) run ALONG the lanes: lane l extends the running sum of a row over its 16 columns
//     and hands the chain state (sum + the last four inputs = 5 floats) to lane l + 1 with one rotate-shuffle
//     per step.  Lane l therefore works on stream row t - l at step t (a systolic skew of one row per lane),
//     and everything a lane produces is consumed by the same lane: P1 -> P2 -> P3 -> P4 stay in registers.
//
// Per 512 pixels the shared-memory/shuffle pipe sees 10 shuffles + the raw bytes once in (TMA) and once out
// (4 LDS.128 per lane) instead of 176 wavefronts.
//
// Index algebra (all verified against the oracle by tests/emu/pdq_systolic_emu.cpp, which compiles this very
// header with g++ and executes the schedule step by step, incl. the TMA ring with poisoned slots):
//   * a frame is a STREAM of kStepsPerFrame = 516 rows: the 512 image rows, then 4 zero rows.  The zero rows
//     drain the column chains (P2 needs one: output row 510; the rest flush the 4-deep histories so that the
//     next frame starts from all-zero history without touching 80 registers) and keep 4-row TMA boxes
//     aligned with frame boundaries (516 = 4 * 129).
//   * row chains: feeding x[c] yields the output of column c - 2.  Lane l is fed pixels 16 l + 2 .. 16 l + 17
//     (pixels 512, 513 are zeros: the drain of the row), so its P1 outputs are columns 16 l .. 16 l + 15,
//     aligned with what it feeds into P3, whose decimated outputs 8 j + 4 (j = 2 l, 2 l + 1) appear at k = 6, 14.
//     The row prologue (pixels 0, 1 fed without output) is computed by lane 31 one step ahead in the two slots
//     where its own row has only zeros left, and rides the rotate-shuffle into lane 0.
//   * column chains: feeding row r yields output row r - 2; history slot = step & 3 (static after unrolling
//     the step loop by 4).  P2 outputs rows 0 and 510 (divisor 3) are fixed up at r = 2 and r = 512; P4 is fed
//     P3 rows r - 2 for 2 <= r <= 512 (zeros otherwise) and emits decimated row i at r = 8 i + 8.
//   * deferred power-of-two scaling as in pdq_fused2_core.h: planes stay unscaled (x4 per pass), divisor-3
//     outputs become 4 * div3(s), the single multiply by 2^-8 happens on the 4096 emitted values.
//
// Raw staging.  Stream row s of lane group G' = l >> 2 (4 lanes, 192 + 16 bytes of the row) lives in slot s & 15
// of that group's ring; TMA boxes of 4 stream rows x kSegPitch bytes, one per group per EVENT E covering the
// group's stream rows 4 E - 4 G' .. + 3 (time-shifted per group: a group only ever holds the rows its own lanes
// still need, which is what lets a 16-row ring absorb the 32-row skew).  ISSUE(E) at step 4 E - 9, WAIT(E) at step
// 4 E - 1: eight steps of lead, two events in flight.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define VPDQS_HD __host__ __device__ __forceinline__
#else
#define VPDQS_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define VPDQS_UNROLL _Pragma("unroll")
#else
#define VPDQS_UNROLL
#endif

namespace vpdq_sys {

#if defined(__CUDA_ARCH__)
VPDQS_HD float fadd(float a, float b) { return __fadd_rn(a, b); }
VPDQS_HD float fsub(float a, float b) { return __fsub_rn(a, b); }
VPDQS_HD float fmul(float a, float b) { return __fmul_rn(a, b); }
VPDQS_HD float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
VPDQS_HD float bits_to_float(uint32_t u) { return __uint_as_float(u); }
VPDQS_HD uint32_t byte_splice(uint32_t word, int k) { return __byte_perm(word, 0x4B000000u, 0x7540u + k); }
#else
}  // namespace vpdq_sys
#include <math.h>
namespace vpdq_sys {
// host build (emulator): compile with -ffp-contract=off; fmaf() is a correctly rounded fused op
VPDQS_HD float fadd(float a, float b) { volatile float r = a + b; return r; }
VPDQS_HD float fsub(float a, float b) { volatile float r = a - b; return r; }
VPDQS_HD float fmul(float a, float b) { volatile float r = a * b; return r; }
VPDQS_HD float ffma(float a, float b, float c) { return fmaf(a, b, c); }
VPDQS_HD float bits_to_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
VPDQS_HD uint32_t byte_splice(uint32_t word, int k) { return 0x4B000000u | ((word >> (8 * k)) & 0xFFu); }
#endif

// v / 3.0f, correctly rounded, branch free (Markstein; equal to IEEE division for every finite positive float,
// tests/emu/div3_check.c), and invariant under power-of-two scaling of v
VPDQS_HD float div3(float v) {
    const float c = 0.333333343267440796f;  // 0x3EAAAAAB
    const float q = fmul(v, c);
    const float r = ffma(-3.0f, q, v);
    return ffma(r, c, q);
}
// a divisor-3 output in the deferred-scale representation
VPDQS_HD float edge3(float v) { return fmul(div3(v), 4.0f); }

constexpr int kCols = 16;            // image columns per lane
constexpr int kStepsPerFrame = 516;  // 512 image rows + 4 zero rows
constexpr int kImageRows = 512;
constexpr int kRing = 16;            // stream rows per group ring
constexpr int kBoxRows = 4;          // stream rows per TMA box
constexpr int kGroupLanes = 4;
constexpr int kGroups = 32 / kGroupLanes;  // 8
constexpr int kFirstStep = -4;       // the step loop starts here (a multiple of 4; steps < 0 only run lane 31's prologue)
constexpr int kIssueLead = 9;        // ISSUE(E) at step 4 E - 9
constexpr int kWaitLead = 1;         // WAIT(E)  at step 4 E - 1

template <int CH>  // 3: RGB24, 1: 8-bit gray (== R = G = B)
struct Raw {
    static constexpr int kLaneBytes = kCols * CH;                  // 48 / 16: the lane's 16 pixels
    static constexpr int kChunks = CH == 3 ? 4 : 2;                // 16-byte chunks of the lane's window (pixels 16l .. 16l+17 -> 54 / 18 bytes)
    static constexpr int kWords = 4 * kChunks;
    static constexpr int kSkip = 2 * CH;                           // pixel 16 l + 2 starts here in the window
    static constexpr int kSegBytes = kGroupLanes * kLaneBytes;     // 192 / 64: a group's share of a row
    static constexpr int kSegPitch = CH == 3 ? 224 : 96;           // TMA box inner size: the share + the 16 bytes the last lane's
                                                                   // window reaches into the next group's, rounded so that a box
                                                                   // (4 rows) is a multiple of 128 bytes
    static constexpr int kBoxBytes = kSegPitch * kBoxRows;         // 896 / 384
    static constexpr int kGroupRingBytes = kSegPitch * kRing;      // 3584 / 1536
    static constexpr int kWarpRingBytes = kGroups * kGroupRingBytes;  // 28672 / 12288
    static constexpr int kRowBytes = 512 * CH;
};

// ---- stream bookkeeping ----------------------------------------------------------------------------------
// byte offset (inside the warp's ring) of the window of `lane` for stream row s
template <int CH>
VPDQS_HD int ring_offset(int lane, int s) {
    return (lane >> 2) * Raw<CH>::kGroupRingBytes + (s & (kRing - 1)) * Raw<CH>::kSegPitch + (lane & 3) * Raw<CH>::kLaneBytes;
}
// the TMA box of event E for group g: first stream row (may be negative = nothing to load)
VPDQS_HD int box_first_row(int E, int g) { return kBoxRows * E - kGroupLanes * g; }
template <int CH>
VPDQS_HD int box_ring_offset(int g, int s0) { return g * Raw<CH>::kGroupRingBytes + (s0 & (kRing - 1)) * Raw<CH>::kSegPitch; }
template <int CH>
VPDQS_HD int box_x(int g) { return g * Raw<CH>::kSegBytes; }
// last step of a warp that owns F frames: lane 31 feeds stream row 516 (F - 1) + 512 (the one drain row the last
// frame needs)
VPDQS_HD int last_step(int F) { return kStepsPerFrame * (F - 1) + kImageRows + 31; }
// events issued before the step loop starts
VPDQS_HD int first_loop_event() { return (kFirstStep + kIssueLead + 3) / 4; }  // ISSUE(E) step 4E-9 >= kFirstStep  ->  E >= 2

struct RowChain {  // running sum over a row, window 4: s + the last four inputs (h0 oldest)
    float s, h0, h1, h2, h3;
};
VPDQS_HD RowChain row_zero() { return RowChain{0.0f, 0.0f, 0.0f, 0.0f, 0.0f}; }
VPDQS_HD float row_feed(RowChain& c, float v) {
    c.s = fadd(c.s, v);
    c.s = fsub(c.s, c.h0);
    c.h0 = c.h1; c.h1 = c.h2; c.h2 = c.h3; c.h3 = v;
    return c.s;
}

struct LaneState {
    float s2[kCols];      // P2 running sums, one per owned column
    float h2[4][kCols];   // P2 histories, slot = step & 3
    float s4[2];          // P4 running sums of the two decimated columns 2l, 2l+1
    float h4[4][2];
    RowChain in1, in3;    // chain states handed over by lane l - 1 for THIS step's row (P1; P3)
    int r, f;             // stream position of this step: row 0 .. 515 of frame f (relative to the warp's first frame)
    VPDQS_HD void init(int lane) {
        VPDQS_UNROLL
        for (int k = 0; k < kCols; ++k) {
            s2[k] = 0.0f;
            h2[0][k] = h2[1][k] = h2[2][k] = h2[3][k] = 0.0f;
        }
        s4[0] = s4[1] = 0.0f;
        VPDQS_UNROLL
        for (int j = 0; j < 4; ++j) h4[j][0] = h4[j][1] = 0.0f;
        in1 = row_zero();
        in3 = row_zero();
        // stream row of lane l at the first step = kFirstStep - l < 0: rows of the virtual frame -1 (never live)
        f = -1;
        r = kStepsPerFrame + kFirstStep - lane;
    }
    VPDQS_HD bool reads_image(int n_frames) const { return (unsigned)f < (unsigned)n_frames && r < kImageRows; }
    VPDQS_HD void advance() {
        if (++r == kStepsPerFrame) {
            r = 0;
            ++f;
        }
    }
};

// luma of the pixel whose first byte sits at byte offset b0 of the little-endian word array w.  u8 -> fp32 product
// without an I2F: the byte is spliced into the mantissa of 2^23 (PRMT): M = 2^23 + b exactly;
// fma(c, M, -c * 2^23) = RN(c * b), bit-identical to __fmul_rn(c, (float)b) (c * 2^23 is exact).  CH == 1: the
// same three-term expression on one byte (SURVEY.md 8 note a-1).
template <int CH, int N>
VPDQS_HD float luma_at(const uint32_t (&w)[N], int b0) {
    const float cr = 0.299f, cg = 0.587f, cb = 0.114f, two23 = 8388608.0f;
    const int b1 = CH == 3 ? b0 + 1 : b0, b2 = CH == 3 ? b0 + 2 : b0;
    const float mr = bits_to_float(byte_splice(w[b0 >> 2], b0 & 3));
    const float mg = CH == 3 ? bits_to_float(byte_splice(w[b1 >> 2], b1 & 3)) : mr;
    const float mb = CH == 3 ? bits_to_float(byte_splice(w[b2 >> 2], b2 & 3)) : mr;
    const float r = ffma(cr, mr, -(cr * two23));
    const float g = ffma(cg, mg, -(cg * two23));
    const float b = ffma(cb, mb, -(cb * two23));
    return fadd(fadd(r, g), b);  // (0.299 R + 0.587 G) + 0.114 B
}

// One lane, one step.  J = step & 3 (history slot).  w = the lane's raw window of its stream row (zeros when the
// row is not an image row); for lane 31 the LAST chunk is instead the first 16 bytes of the row lane 0 works on in
// the NEXT step (zeros if that is not an image row).
// out1 / out3: the chain states to hand to lane l + 1 (lane 31 -> lane 0: the next row's initial states).
// emit(frame, i, v0, v1): decimated row i of columns 2l, 2l+1 is final.
template <int CH, int J, typename Emit>
VPDQS_HD void lane_step(LaneState& L, const uint32_t (&w)[Raw<CH>::kWords], int lane, int n_frames, RowChain& out1,
                        RowChain& out3, Emit emit) {
    const int r = L.r;
    if (r == 0) {  // new frame: the histories were flushed by the four zero rows, the sums hold rounding residue
        VPDQS_UNROLL
        for (int k = 0; k < kCols; ++k) L.s2[k] = 0.0f;
    }
    // ---- P1 (row pass 1, along the lanes) + P2 (column pass 1, private) ----
    RowChain c1 = L.in1;
    float y2[kCols];
    float xa = 0.0f, xb = 0.0f;
    VPDQS_UNROLL
    for (int k = 0; k < kCols; ++k) {
        float x = luma_at<CH>(w, Raw<CH>::kSkip + CH * k);  // pixel 16 l + 2 + k
        if (k == 14) xa = x;
        if (k == 15) xb = x;
        if (k >= 14 && lane == 31) x = 0.0f;  // pixels 512, 513: the drain of the row
        float v = row_feed(c1, x);            // output column 16 l + k, unscaled (x4)
        if (k == 0 && lane == 0) v = edge3(v);    // column 0: divisor 3
        if (k == 14 && lane == 31) v = edge3(v);  // column 510: divisor 3 (column 511 feeds no decimated output)
        const float old = L.h2[J][k];
        float s = fadd(L.s2[k], v);
        s = fsub(s, old);
        L.h2[J][k] = v;
        L.s2[k] = s;
        y2[k] = s;  // output row r - 2 of column 16 l + k, unscaled (x16)
    }
    if (r == 2 || r == kImageRows) {  // output rows 0 and 510: divisor 3
        VPDQS_UNROLL
        for (int k = 0; k < kCols; ++k) y2[k] = edge3(y2[k]);
        if (r == 2) L.s4[0] = L.s4[1] = 0.0f;  // P4 starts here; its histories were flushed by 5 zero feeds
    }
    // ---- P3 (row pass 2, along the lanes): only the decimated columns 8 j + 4 are kept ----
    RowChain c3 = L.in3;
    float z0 = 0.0f, z1 = 0.0f;
    VPDQS_UNROLL
    for (int k = 0; k < kCols; ++k) {
        const float v = row_feed(c3, y2[k]);  // output column 16 l + k - 2
        if (k == 6) z0 = v;
        if (k == 14) z1 = v;
    }
    // ---- P4 (column pass 2, private), fed P3 row r - 2 ----
    const bool zvalid = r >= 2 && r <= kImageRows;
    if (!zvalid) z0 = z1 = 0.0f;
    float o0, o1;
    {
        const float old0 = L.h4[J][0], old1 = L.h4[J][1];
        float s0 = fadd(L.s4[0], z0), s1 = fadd(L.s4[1], z1);
        s0 = fsub(s0, old0);
        s1 = fsub(s1, old1);
        L.h4[J][0] = z0;
        L.h4[J][1] = z1;
        L.s4[0] = o0 = s0;
        L.s4[1] = o1 = s1;
    }
    if ((unsigned)L.f < (unsigned)n_frames && r >= 8 && (r & 7) == 0)  // output row r - 4 = 8 i + 4
        emit(L.f, (r >> 3) - 1, fmul(o0, 0.00390625f), fmul(o1, 0.00390625f));  // the deferred 4^-4

    // ---- hand-over ----
    out1 = c1;
    out3 = c3;
    if (lane == 31) {  // -> lane 0, next row: the chain after the prologue pixels 0, 1 (fed without output) / a fresh chain
        out1 = RowChain{fadd(xa, xb), 0.0f, 0.0f, xa, xb};
        out3 = row_zero();
    }
    L.advance();
}

}  // namespace vpdq_sys
