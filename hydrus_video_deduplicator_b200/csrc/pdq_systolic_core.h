// pdq_systolic_core.h -- per-lane arithmetic and the schedule of the WARP-PER-FRAME systolic PDQ kernel
// (kx_systolic_jarosz, pdq_systolic.cu): luma + the four Jarosz box-filter passes + the 64x64 decimation with
// NO shared-memory transposition between the passes.
//
// Why.  The Jarosz filter is a running sum (s += x[r]; s -= x[l]; y = s / n), so every output carries the
// rounding history of its whole line prefix and each line is an inherently serial fp32 chain.
// The round-1 tiled kernels (tests/legacy/pdq_fused*.cu) gave every lane whole lines and therefore had to transpose the plane
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
//   * column chains: feeding row r yields output row r - 2; history slot = step & 7 (static after unrolling
//     the step loop by 8).  P2 output rows 0 and 510 have divisor 3: the lane applies it to what it feeds into P3
//     (at r = 3 and r = 513), the running sums themselves are never touched.  P3 and P4 run ONE STEP BEHIND P2 (so
//     that the two serial row chains of a step, P1 of row r and P3 of P2-row r - 3, are independent and interleave
//     in one basic block): P4 is fed P3 rows r - 3 for 3 <= r <= 513 (zeros otherwise) and emits decimated row i
//     at r = 8 i + 9.
//   * the step loop has two bodies, both branch free: the PLAIN one for iterations in which every lane is on a row
//     4 <= r <= 509 (90 % of a frame; no row tests at all) and the general one, in which the row number selects
//     restarts, divisor-3 inputs and zero feeds per lane (see iteration_is_plain and lane_step).
//   * deferred power-of-two scaling: planes stay unscaled (x4 per pass), divisor-3 outputs become 4 * div3(s) = edge3(s),
//     the single multiply by 2^-8 happens on the 4096 emitted values.
//
// Raw staging.  Stream row s of lane group G' = l >> 2 (4 lanes, 192 + 16 bytes of the row) lives in slot s & 15
// of that group's ring; TMA boxes of 4 stream rows x kSegPitch bytes, one per group per EVENT E covering the
// group's stream rows 4 E - 4 G' .. + 3 (time-shifted per group: a group only ever holds the rows its own lanes
// still need, which is what lets a 16-row ring absorb the 32-row skew).  ISSUE(E) at step 4 E - 10, WAIT(E) at step
// 4 E - 2: eight steps of lead, two events in flight.  (A step reads the windows of the NEXT step's rows.)
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define VPDQS_HD __host__ __device__ __forceinline__
#else
#define VPDQS_HD inline
#endif
#define VPDQS_UNLIKELY(x) __builtin_expect(!!(x), 0)
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

// packed fp32 pairs (sm_100a FADD2 / FFMA2: one instruction, two independent IEEE-RN results -- bit-identical to two
// scalar operations)
struct alignas(8) F2 {
    float x, y;
};
#if defined(__CUDA_ARCH__)
VPDQS_HD F2 f2_add(F2 a, F2 b) {
    const float2 r = __fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
    return F2{r.x, r.y};
}
VPDQS_HD F2 f2_sub(F2 a, F2 b) {
    const float2 r = __fadd2_rn(make_float2(a.x, a.y), make_float2(-b.x, -b.y));
    return F2{r.x, r.y};
}
VPDQS_HD F2 f2_fma(F2 a, F2 b, F2 c) {
    const float2 r = __ffma2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y), make_float2(c.x, c.y));
    return F2{r.x, r.y};
}
#else
VPDQS_HD F2 f2_add(F2 a, F2 b) { return F2{fadd(a.x, b.x), fadd(a.y, b.y)}; }
VPDQS_HD F2 f2_sub(F2 a, F2 b) { return F2{fsub(a.x, b.x), fsub(a.y, b.y)}; }
VPDQS_HD F2 f2_fma(F2 a, F2 b, F2 c) { return F2{ffma(a.x, b.x, c.x), ffma(a.y, b.y, c.y)}; }
#endif
VPDQS_HD F2 f2_splat(float v) { return F2{v, v}; }
VPDQS_HD F2 f2_mul(F2 a, F2 b) {
#if defined(__CUDA_ARCH__)
    const float2 r = __fmul2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
    return F2{r.x, r.y};
#else
    return F2{fmul(a.x, b.x), fmul(a.y, b.y)};
#endif
}

// branch-free selects on the bit patterns (one LOP3 each): m = all ones -> a, m = 0 -> b
VPDQS_HD uint32_t float_bits(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
#endif
}
VPDQS_HD float bitsel(uint32_t m, float a, float b) { return bits_to_float((float_bits(a) & m) | (float_bits(b) & ~m)); }
VPDQS_HD float bitkeep(uint32_t m, float a) { return bits_to_float(float_bits(a) & m); }  // m ? a : +0

// v / 3.0f, correctly rounded, branch free (Markstein; equal to IEEE division for every finite positive float,
// tests/emu/div3_check.c), and invariant under power-of-two scaling of v
VPDQS_HD float div3(float v) {
    const float c = 0.333333343267440796f;  // 0x3EAAAAAB
    const float q = fmul(v, c);
    const float r = ffma(-3.0f, q, v);
    return ffma(r, c, q);
}
// a divisor-3 output in the deferred-scale representation: 4 * RN(v / 3) = RN(4 v / 3), the same Markstein sequence run
// on 4 v with the power of two folded into the constants (q = v * 4c, r / 4 = v - 0.75 q, q + (r / 4) * 4c) -- three
// operations, bit-identical to fmul(div3(v), 4) for every normal float (tests/emu/div3_check.c)
VPDQS_HD float edge3(float v) {
    const float c4 = 1.33333337306976318f;  // 0x3FAAAAAB = 4 * 0x3EAAAAAB
    const float q = fmul(v, c4);
    const float r = ffma(-0.75f, q, v);
    return ffma(r, c4, q);
}
constexpr float kEdge3C = 1.33333337306976318f;
// edge3 with its two constants as arguments: (kEdge3C, -0.75) -> edge3(v); (1, -1) -> v itself, exactly (q = v * 1,
// the residual v - 1 * q = +0, q + 0 * c = v)
VPDQS_HD float edge3_if(float v, float c, float k) {
    const float q = fmul(v, c);
    const float r = ffma(k, q, v);
    return ffma(r, kEdge3C, q);
}
// the same on a packed pair, with the constants as arguments: (kEdge3C, -0.75) -> edge3(v); (1, -1) -> v itself, exactly
// (no multiply feeds an add here: nothing for ptxas to contract)
VPDQS_HD F2 edge3_if(F2 v, float c, float k) {
    const F2 q = f2_mul(v, f2_splat(c));
    const F2 r = f2_fma(f2_splat(k), q, v);
    return f2_fma(r, f2_splat(kEdge3C), q);
}

constexpr int kCols = 16;            // image columns per lane
constexpr int kStepsPerFrame = 516;  // 512 image rows + 4 zero rows
constexpr int kImageRows = 512;
#ifndef VPDQS_RING_ROWS
#define VPDQS_RING_ROWS 16
#endif
constexpr int kRing = VPDQS_RING_ROWS;  // stream rows per group ring: 16 (4 box slots, 8 steps of TMA lead) or 12 (3, 4)
constexpr int kBoxSlots = kRing / 4;
// box slot of ring row index a (= stream row + 4 * group): floor-mod; a mask when the slot count is a power of two (a may
// be negative in the first steps, where the result is not used)
VPDQS_HD constexpr int box_slot_of(int a4) {
    return (kBoxSlots & (kBoxSlots - 1)) == 0 ? (a4 & (kBoxSlots - 1)) : ((a4 % kBoxSlots) + kBoxSlots) % kBoxSlots;
}
constexpr int kBoxRows = 4;          // stream rows per TMA box
constexpr int kGroupLanes = 4;
constexpr int kGroups = 32 / kGroupLanes;  // 8
#ifndef VPDQS_BODY
#define VPDQS_BODY 8
#endif
constexpr int kBody = VPDQS_BODY;    // steps per iteration of the step loop: 8 (4 is kept for A/B runs; see LaneState)
static_assert(kBody == 4 || kBody == 8, "the TMA events are placed every 4 steps, statically");
constexpr int kHistSlots = kBody;    // slots of the 4-deep column-pass histories (8: the slot written differs from the one read)
constexpr int kFirstStep = -8;       // the step loop starts here (a multiple of kBody; steps < 0 only prepare lane 0's first row)
constexpr int kIssueLead = kRing - 6; // ISSUE(E) at step 4 E - 10 (ring of 16 rows; 4 E - 6 with 12): the earliest step at
                                     // which no lane still reads the box slot being refilled
constexpr int kWaitLead = 2;         // WAIT(E)  at step 4 E - 2 (a step reads the raw rows of the NEXT step: its lumas are
                                     // computed one step ahead, as filler work for the serial chains)
constexpr int kEventPhase = 2;       // both happen in the steps with (step & 3) == 2

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
// Ring layout per warp: [4 box slots][8 groups, stored in REVERSE order g' = 7 - g][4 rows][kSegPitch bytes].  The
// box of event E of EVERY group sits in box slot E & 3 (group g's rows of event E are stream rows 4 E - 4 g .. + 3),
// so that in the common case ONE 3-D TMA box (x, row, g') fills a whole slot; stream row s of group g is in box slot
// ((s >> 2) + g) & 3, row s & 3.
template <int CH>
VPDQS_HD int ring_group_offset(int g) { return (kGroups - 1 - g) * Raw<CH>::kBoxBytes; }
template <int CH>
VPDQS_HD int ring_row_offset(int g, int s) {
    const int a = s + kGroupLanes * g;
    return box_slot_of(a >> 2) * (kGroups * Raw<CH>::kBoxBytes) + (a & 3) * Raw<CH>::kSegPitch;
}
// byte offset (inside the warp's ring) of the window of `lane` for stream row s
template <int CH>
VPDQS_HD int ring_offset(int lane, int s) {
    return ring_row_offset<CH>(lane >> 2, s) + ring_group_offset<CH>(lane >> 2) + (lane & 3) * Raw<CH>::kLaneBytes;
}
// the TMA box of event E for group g: first stream row (may be negative = nothing to load)
VPDQS_HD int box_first_row(int E, int g) { return kBoxRows * E - kGroupLanes * g; }
template <int CH>
VPDQS_HD int box_ring_offset(int g, int E) { return box_slot_of(E) * (kGroups * Raw<CH>::kBoxBytes) + ring_group_offset<CH>(g); }
template <int CH>
VPDQS_HD int box_x(int g) { return g * Raw<CH>::kSegBytes; }
// The 3-D view of the batch that makes all 8 boxes of an event one TMA box: element (x, y, g') lives at byte
//   kBase3 + x + kRowBytes * y + kStride3 * g'   of the frame buffer,   g' = 7 - g,
// i.e. group g's share (byte 192 g .. of the row) of image row  y + 28 - 4 g.  An event whose group-0 box starts at
// image row r0 of a frame (global row Y0) is the box at (0, Y0 - 28, 0) of size (kSegPitch, 4, 8) -- valid when all
// eight boxes lie in that one frame: 28 <= r0 <= 508.
template <int CH>
struct View3 {
    static constexpr int kBackRows = kGroupLanes * (kGroups - 1);                                   // 28
    static constexpr long long kStride3 = (long long)kGroupLanes * Raw<CH>::kRowBytes - Raw<CH>::kSegBytes;  // 5952 / 1984
    static constexpr long long kBase3 = (long long)(kGroups - 1) * Raw<CH>::kSegBytes;              // 1344 / 448
};
VPDQS_HD bool event_is_one_box(int f, int r0, int n_frames) {
    return f < n_frames && r0 >= kGroupLanes * (kGroups - 1) && r0 <= kImageRows - kBoxRows;
}
// last step of a warp that owns F frames: lane 31 at stream row 516 (F - 1) + 513 (P3 / P4 run one step behind P2,
// whose last output row 510 appears at row 512)
VPDQS_HD int last_step(int F) { return kStepsPerFrame * (F - 1) + kImageRows + 1 + 31; }
// events issued before the step loop starts
VPDQS_HD int first_loop_event() {  // smallest E whose ISSUE step 4 E - kIssueLead falls inside the loop
    return (kFirstStep + kIssueLead + 3 + 400) / 4 - 100;
}

// The step loop has TWO bodies.  An iteration (kBody steps from step t0, lane 0 on stream row r0 of frame f0, both
// uniform) is PLAIN when during all of its steps every lane is on a row 4 <= r <= 509 of a live frame: no divisor-3
// row, no frame boundary, every window staged, P4 fed a real row -- and its TMA event is one 3-D box.  Plain
// iterations (90 % of a frame) run lane_step<.., true>: one basic block per kBody steps, no per-lane row tests at
// all.  The others run the general lane_step<.., false>, which tracks (r, f) per lane and branches into the tail for
// the rare rows.
VPDQS_HD constexpr int plain_event_row(int r0) { return r0 + ((kEventPhase + kIssueLead) & ~3); }  // group 0's first row of
                                                                                                // the event issued in the iteration
constexpr int kPlainFirst = 4 + 31;  // lane 31 is on row >= 4 (r0 itself is a multiple of 4: 516 = 4 * 129)
constexpr int kPlainLastA = kImageRows - 3 - (kBody - 1);  // lane 0 stays on rows <= 509 (and lane 31 sees an image row two
                                                           // steps ahead of it)
constexpr int kPlainLastB = kImageRows - kBoxRows - plain_event_row(kBody - 4);  // the LAST event's group-0 rows are <= 511
constexpr int kPlainLast = kPlainLastA < kPlainLastB ? kPlainLastA : kPlainLastB;  // 496 (body of 4 steps, ring of 16 rows)
VPDQS_HD bool iteration_is_plain(int f0, int r0, int n_frames) {
    return (unsigned)f0 < (unsigned)n_frames && r0 >= kPlainFirst && r0 <= kPlainLast;
}
// consecutive plain iterations from r0 on (the kernel runs them as one inner loop)
VPDQS_HD int plain_run_length(int r0) { return (kPlainLast - r0) / kBody + 1; }

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
// The same with parts of the state switched off by multipliers that are 1 or 0 -- exact: x * 1 + y rounds once like
// x + y, and x * 0 + y = y for finite x (every value here is finite and >= +0) -- instead of a select per value:
//   row_feed_kh: the oldest history value counts kh times.  The first feeds of a lane subtract what the previous lane
//                handed over; lane 0 must see zeros there, not what lane 31 rotates into it (nkh = -kh);
//   row_feed_ks: also the incoming running sum counts ks times (P3: lane 0 starts a fresh chain);
//   row_feed_kv: the new value counts kv times in the sum (lane 31's last two inputs are the drain zeros of its row;
//                the registers hold the NEXT row's pixels 0, 1, which go into the history unmasked: exactly the state
//                lane 0 needs next).
VPDQS_HD float row_feed_kh(RowChain& c, float v, float nkh) {
    c.s = fadd(c.s, v);
    c.s = ffma(c.h0, nkh, c.s);
    c.h0 = c.h1; c.h1 = c.h2; c.h2 = c.h3; c.h3 = v;
    return c.s;
}
VPDQS_HD float row_feed_ks(RowChain& c, float v, float ks, float nkh) {
    c.s = ffma(c.s, ks, v);
    c.s = ffma(c.h0, nkh, c.s);
    c.h0 = c.h1; c.h1 = c.h2; c.h2 = c.h3; c.h3 = v;
    return c.s;
}
VPDQS_HD float row_feed_kv(RowChain& c, float v, float kv) {
    c.s = ffma(v, kv, c.s);
    c.s = fsub(c.s, c.h0);
    c.h0 = c.h1; c.h1 = c.h2; c.h2 = c.h3; c.h3 = v;
    return c.s;
}

// The step loop is unrolled by kBody and a step is compiled per position T in the body, so that every index below is
// static.  The 4-deep histories of the column passes live in kHistSlots = kBody slots: the value fed at step u sits in
// slot u mod kHistSlots, a step reads slot (u - 4) and writes slot u.  With 4 slots those are the same registers -- the
// new value is produced while the old one is still needed, and ptxas has to park it and copy (20 MOVs per step);
// with 8 slots it is computed straight into its final register.  The 8-step plain loop is 29 KB of code, just inside
// the 32 KB instruction cache (an earlier 35 KB version of it was not: "no_instructions" stalls 3 % -> 19 %).
struct LaneState {
    F2 s2[kCols / 2];         // P2 running sums: columns (2p, 2p+1) of the lane packed in one register pair.  At the start
                              // of a step they are ALSO the P2 outputs of the previous step, which P3 consumes in this one
                              // (one value for both: see the tail of lane_step for the divisor-3 rows)
    F2 h2[kHistSlots][kCols / 2];  // P2 histories (the last four inputs): the value fed at step u is h2[u mod kHistSlots]
    F2 x[2][kCols / 2];       // lumas of THIS step's row (pixels 16 l + 2 ..), computed during the previous step:
                              // alternating sets x[step & 1]
    F2 keep;                  // the decimated pair a lane produces during an iteration of 8 steps (exactly one in a plain
    bool have;                // iteration: its rows advance by 8; at most one otherwise), stored once after the last step
    F2 s4;                    // P4 running sums of the two decimated columns 2l, 2l+1
    F2 h4[kHistSlots];
    RowChain in1, in3;        // chain states handed over by lane l - 1 for THIS step (P1: row r; P3: P2-row r - 3)
    int r, f;                 // stream position of this step: row 0 .. 515 of frame f (relative to the warp's first frame).
                              // Plain iterations only advance r, once, after their last step
    // the NEXT step's row is an image row of one of the warp's frames (its window is staged; else the lane reads zeros)
    VPDQS_HD bool img_next(int n_frames) const {
        return (r <= kImageRows - 2 && (unsigned)f < (unsigned)n_frames) ||
               (r == kStepsPerFrame - 1 && (unsigned)(f + 1) < (unsigned)n_frames);
    }
    VPDQS_HD void init(int lane) {
        VPDQS_UNROLL
        for (int p = 0; p < kCols / 2; ++p) {
            s2[p] = f2_splat(0.0f);
            VPDQS_UNROLL
            for (int j = 0; j < kHistSlots; ++j) h2[j][p] = f2_splat(0.0f);
            x[0][p] = x[1][p] = f2_splat(0.0f);
        }
        s4 = keep = f2_splat(0.0f);
        VPDQS_UNROLL
        for (int j = 0; j < kHistSlots; ++j) h4[j] = f2_splat(0.0f);
        in1 = row_zero();
        in3 = row_zero();
        // stream row of lane l at the first step = kFirstStep - l < 0: rows of the virtual frame -1 (never live)
        f = -1;
        r = kStepsPerFrame + kFirstStep - lane;
        have = false;
    }
};
// slot of the history value fed at body position U (U may be negative: earlier steps)
VPDQS_HD constexpr int hist_slot(int U) { return ((U % kHistSlots) + kHistSlots) % kHistSlots; }

// u8 -> fp32 product without an I2F: the byte is spliced into the mantissa of a power of two, M = 2^k + byte exactly, and
// fma(c, M, -c * 2^k) = RN(c * byte) -- bit-identical to __fmul_rn(c, (float)byte) (c * 2^k is exact).  Bytes 2 and 3 of a
// word go to the low mantissa byte of 2^23 with one PRMT (half-rate pipe); bytes 0 and 1 stay where they are under a
// mask -- one LOP3 (full rate): byte 0 in the mantissa of 2^23, byte 1 in that of 2^15 (its bits then weigh 2^0..2^7).
// (Measured on B200: the LOP3 form is 8 % SLOWER end to end -- the per-position offsets stop the FFMA2 constants
// from being shared -- so the product uses PRMT for all four positions; the switch stays for the record.)
#ifndef VPDQS_LOP3_SPLICE
#define VPDQS_LOP3_SPLICE 0
#endif
VPDQS_HD constexpr float magic_base(int b) { return (VPDQS_LOP3_SPLICE && (b & 3) == 1) ? 32768.0f : 8388608.0f; }
template <int N>
VPDQS_HD float magic_at(const uint32_t (&w)[N], int b) {
    const uint32_t word = w[b >> 2];
    if (VPDQS_LOP3_SPLICE && (b & 3) == 0) return bits_to_float((word & 0x000000FFu) | 0x4B000000u);  // 2^23 + byte
    if (VPDQS_LOP3_SPLICE && (b & 3) == 1) return bits_to_float((word & 0x0000FF00u) | 0x47000000u);  // 2^15 + byte
    return bits_to_float(byte_splice(word, b & 3));                                                    // 2^23 + byte
}

// luma of the two pixels whose first bytes sit at byte offsets b0 and b0 + CH of w, as one packed pair.
// CH == 1: the same three-term expression on one byte (SURVEY.md 8 note a-1).
template <int CH, int N>
VPDQS_HD F2 luma_pair_at(const uint32_t (&w)[N], int b0) {
    const float cr = 0.299f, cg = 0.587f, cb = 0.114f;
    const int p0 = b0, p1 = b0 + CH;
    const int g0 = CH == 3 ? p0 + 1 : p0, g1 = CH == 3 ? p1 + 1 : p1, u0 = CH == 3 ? p0 + 2 : p0, u1 = CH == 3 ? p1 + 2 : p1;
    const F2 mr{magic_at(w, p0), magic_at(w, p1)};
    const F2 mg = CH == 3 ? F2{magic_at(w, g0), magic_at(w, g1)} : mr;
    const F2 mb = CH == 3 ? F2{magic_at(w, u0), magic_at(w, u1)} : mr;
    const F2 r = f2_fma(f2_splat(cr), mr, F2{-(cr * magic_base(p0)), -(cr * magic_base(p1))});
    const F2 g = f2_fma(f2_splat(cg), mg, F2{-(cg * magic_base(g0)), -(cg * magic_base(g1))});
    const F2 b = f2_fma(f2_splat(cb), mb, F2{-(cb * magic_base(u0)), -(cb * magic_base(u1))});
    return f2_add(f2_add(r, g), b);  // (0.299 R + 0.587 G) + 0.114 B
}

// One lane, one step.  T = position of the step in the loop body (step mod kBody).  w = the raw window of the lane's
// NEXT stream row (zeros when that row is not an image row: L.img_next()); for lane 31 the LAST chunk is instead the
// first 16 bytes of the row lane 0 works on TWO steps ahead (zeros if that is not an image row).
//
// The step is ONE branch-free block in which the two serial chains -- P1 over this step's row r and P3 over the P2
// outputs of the PREVIOUS step (output row r - 3) -- start with every input ready and run side by side with the
// independent work (P2, and the lumas of the NEXT step's row) that fills their latency.
// PLAIN: the step belongs to a plain iteration (iteration_is_plain): every lane is on a row 4 <= r <= 509 of a live
// frame and L.r is the row of the iteration's FIRST step.  Otherwise the rare rows are handled per lane -- still
// without a branch, by what the row number selects:
//   r == 0            the P2 sums restart (they hold rounding residue after the four zero rows): the sum's first
//                     operation is fma(prev, 0, v) instead of fma(prev, 1, v) == prev + v
//   r == 3, r == 513  P3 is fed P2 output rows 0 / 510, whose divisor is 3: edge3(prev) instead of prev (the P2 sums
//                     themselves stay untouched and keep running)
//   r == 3            the P4 sums restart the same way (their histories were flushed by five zero feeds)
//   r < 3, r > 513, or a frame that is not live: P4 is fed zeros and nothing is emitted
// out1 / out3: the chain states to hand to lane l + 1 (lane 31 -> lane 0: the next row's initial states).
// emit(v0, v1): the next decimated row (in order: rows 0..63 of frame 0, 1, ...) of columns 2l, 2l+1 is final.
template <int CH, int T, bool PLAIN, typename Emit>
VPDQS_HD void lane_step(LaneState& L, const uint32_t (&w)[Raw<CH>::kWords], int lane, int n_frames, RowChain& out1,
                        RowChain& out3, Emit emit) {
    constexpr int JW = hist_slot(T), JR = hist_slot(T - 4);            // history slot written / read by this step
    constexpr int XR = T & 1, XW = XR ^ 1;                             // luma set read / written
    const int r = PLAIN ? L.r + T : L.r;
    const bool live = PLAIN || (unsigned)L.f < (unsigned)n_frames;
    const bool edge_row = !PLAIN && (r == 3 || r == kImageRows + 1);
    const bool fed = PLAIN || (live && r >= 3 && r <= kImageRows + 1);   // this step's P3 row is a real P2 output row
    const float keep2 = (PLAIN || r != 0) ? 1.0f : 0.0f, keep4 = (PLAIN || r != 3) ? 1.0f : 0.0f;
    const uint32_t fedmask = fed ? 0xFFFFFFFFu : 0u;
    const float edge_c = edge_row ? kEdge3C : 1.0f, edge_k = edge_row ? -0.75f : -1.0f;
    RowChain c1 = L.in1, c3 = L.in3;
    float z0 = 0.0f, z1 = 0.0f;
    const uint32_t last = lane == 31 ? 0xFFFFFFFFu : 0u;
    // per-lane constants (hoisted out of the step loop): lane 0 ignores the chain history rotated into it, lane 31
    // feeds the drain zeros; columns 0 (lane 0) and 510 (lane 31) of P1 have divisor 3
    const float kf = lane == 0 ? 0.0f : 1.0f, nkf = -kf, k31 = lane == 31 ? 0.0f : 1.0f;
    const float e0c = lane == 0 ? kEdge3C : 1.0f, e0k = lane == 0 ? -0.75f : -1.0f;
    const float e31c = lane == 31 ? kEdge3C : 1.0f, e31k = lane == 31 ? -0.75f : -1.0f;
    // lane 31: its last two lumas are the NEXT row's pixels 0, 1 (its own row has only the drain zeros left there)
    const float xa = L.x[XR][7].x, xb = L.x[XR][7].y;
    VPDQS_UNROLL
    for (int p = 0; p < kCols / 2; ++p) {
        const int k = 2 * p;
        const F2 x = L.x[XR][p];
        L.x[XW][p] = luma_pair_at<CH>(w, Raw<CH>::kSkip + CH * k);  // next step's pixels 16 l + 2 + k, + 1
        // P1: row pass 1 along the lanes -> output columns 16 l + k, + 1 (unscaled, x4)
        float v0 = p == 0 ? row_feed_kh(c1, x.x, nkf) : (p == 7 ? row_feed_kv(c1, x.x, k31) : row_feed(c1, x.x));
        if (p == 0) v0 = edge3_if(v0, e0c, e0k);    // column 0: divisor 3
        if (p == 7) v0 = edge3_if(v0, e31c, e31k);  // column 510: divisor 3 (column 511 feeds no decimated output)
        const float v1 = p == 0 ? row_feed_kh(c1, x.y, nkf) : (p == 7 ? row_feed_kv(c1, x.y, k31) : row_feed(c1, x.y));
        // P2: column pass 1, private -> output row r - 2 (unscaled, x16)
        const F2 v{v0, v1};
        const F2 old = L.h2[JR][p], prev = L.s2[p];
        F2 s = PLAIN ? f2_add(prev, v) : f2_fma(prev, f2_splat(keep2), v);  // (prev * 1 + v == prev + v, one rounding)
        s = f2_sub(s, old);
        L.h2[JW][p] = v;
        L.s2[p] = s;
        // P3: row pass 2 along the lanes over the previous step's P2 outputs -> output column 16 l + k - 2; only the
        // decimated columns 8 j + 4 are kept
        // (rare rows: edge3 with its two constants switched to 1 on every other row, which makes it the identity --
        // q = prev * 1, the residual prev - 1 * q = +0, q + 0 * c = prev -- without a select per value)
        const F2 pin = PLAIN ? prev : edge3_if(prev, edge_c, edge_k);
        const float u0 = p == 0 ? row_feed_ks(c3, pin.x, kf, nkf) : (p == 1 ? row_feed_kh(c3, pin.x, nkf) : row_feed(c3, pin.x));
        if (k == 6) z0 = u0;
        if (k == 14) z1 = u0;
        if (p <= 1)
            row_feed_kh(c3, pin.y, nkf);
        else
            row_feed(c3, pin.y);
    }
    // P4: column pass 2, private, fed P3 row r - 3 (zeros unless real) -> output row r - 5
    {
        const F2 z = PLAIN ? F2{z0, z1} : F2{bitkeep(fedmask, z0), bitkeep(fedmask, z1)};
        const F2 old = L.h4[JR];
        F2 s = PLAIN ? f2_add(L.s4, z) : f2_fma(L.s4, f2_splat(keep4), z);
        s = f2_sub(s, old);
        L.h4[JW] = z;
        L.s4 = s;
        // output row r - 5 = 8 i + 4 at r = 9, 17, .., 513
        const bool now = fed && (r & 7) == 1;
        if (kBody == 8) {  // one store per lane and iteration: a lane meets at most one such row in 8 steps
            const uint32_t m = now ? 0xFFFFFFFFu : 0u;
            L.keep = T == 0 ? s : F2{bitsel(m, s.x, L.keep.x), bitsel(m, s.y, L.keep.y)};  // (T == 0: any value will do)
            L.have = PLAIN ? true : (T == 0 ? now : (L.have || now));
            if (T == kBody - 1 && L.have) emit(fmul(L.keep.x, 0.00390625f), fmul(L.keep.y, 0.00390625f));  // the deferred 4^-4
        } else if (now) {
            emit(fmul(s.x, 0.00390625f), fmul(s.y, 0.00390625f));
        }
    }
    // hand-over.  Lane 31 -> lane 0 is the next row: its P1 chain starts after the prologue pixels 0, 1 (fed without
    // output): sum xa + xb, history (0, 0, xa, xb) -- lane 31's own history already ends in (xa, xb), and lane 0 takes
    // the two older values as zeros (row_feed_kh); its P3 chain starts fresh (row_feed_ks / row_feed_kh).
    out1 = RowChain{bitsel(last, fadd(xa, xb), c1.s), c1.h0, c1.h1, c1.h2, c1.h3};
    out3 = c3;
    if (PLAIN) {
        if (T == kBody - 1) L.r += kBody;
    } else {
        const bool wrap = r == kStepsPerFrame - 1;
        L.r = wrap ? 0 : r + 1;
        L.f += wrap ? 1 : 0;
    }
}

}  // namespace vpdq_sys
