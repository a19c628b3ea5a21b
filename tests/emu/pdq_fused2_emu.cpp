// tests/emu/pdq_fused2_emu.cpp -- CPU emulator of kx_fused_jarosz2 (TEST INFRASTRUCTURE).
// Executes the schedule of hydrus_video_deduplicator_b200/csrc/pdq_fused2_core.h step by step, warp by warp,
// lane by lane -- the same main_step() / p4_step() the CUDA kernel runs -- with the TMA box loads (incl.
// out-of-bounds zero fill) and the one-stage staging ring modelled explicitly.  Also checks that within a step
// no two warps touch the same tile slot / t3 strip and that every staged box is the one the step expects.
// Build: g++ -O1 -ffp-contract=off -shared -fPIC -o libpdq_fused2_emu.so pdq_fused2_emu.cpp
#include <stdlib.h>

#include "../legacy/pdq_fused2_core.h"

using namespace vpdq_core2;

namespace {
struct Cta {
    LaneState st[kMainWarps][32];
    P4State p4[32];
    alignas(16) F2 slot[2][kMainWarps][kSlotF2];
    alignas(16) F2 t3[2][kMainWarps * kT3Strip];
    uint8_t raw[kMainWarps][2][kRawBoxBytesMax];
    int raw_holds[kMainWarps];  // which tile index u is staged (checks the ring protocol)
    uint8_t regs[kMainWarps][2][kRawBoxBytesMax];  // the staged rows as held in the lanes' registers
    int regs_hold[kMainWarps];
};

template <int CH>
void tma_box(const uint8_t* frames, long long total_rows, int x, long long y, uint8_t* dst) {
    constexpr int kRawPitch = Raw<CH>::kPitch, kRowBytes = 512 * CH;
    for (int r = 0; r < kTile; ++r)
        for (int b = 0; b < kRawPitch; ++b) {
            const long long row = y + r;
            const int col = x + b;
            dst[r * kRawPitch + b] =
                (row >= 0 && row < total_rows && col >= 0 && col < kRowBytes) ? frames[row * kRowBytes + col] : 0;
        }
}

template <int CH>
int emu_run(const uint8_t* frames, long long n_frames, int grid, float* a64) {
    constexpr int kRawPitch = Raw<CH>::kPitch, kRawWords = Raw<CH>::kWords;
    const long long total_rows = n_frames * 512;
    int errors = 0;
    for (int cta = 0; cta < grid; ++cta) {
        const long long f_begin = n_frames * cta / grid, f_end = n_frames * (cta + 1) / grid;
        const int F = (int)(f_end - f_begin);
        if (F == 0) continue;
        const int FA = (F + 1) >> 1, FB = F - FA;
        const long long half_a = f_begin, half_b = f_begin + FA;
        Cta* s = (Cta*)aligned_alloc(64, (sizeof(Cta) + 63) / 64 * 64);
        memset(s, 0, sizeof(Cta));
        // poison the hand-over buffers: nothing may be consumed before it was produced
        for (int p = 0; p < 2; ++p) {
            for (int w = 0; w < kMainWarps; ++w)
                for (int e = 0; e < kSlotF2; ++e) s->slot[p][w][e] = F2{__builtin_nanf(""), __builtin_nanf("")};
            for (int e = 0; e < kMainWarps * kT3Strip; ++e) s->t3[p][e] = F2{__builtin_nanf(""), __builtin_nanf("")};
        }
        for (int w = 0; w < kMainWarps; ++w) {
            for (int l = 0; l < 32; ++l) s->st[w][l].init();
            s->raw_holds[w] = -1000;
            s->regs_hold[w] = -1000;
        }
        for (int l = 0; l < 32; ++l) s->p4[l].init();
        auto issue = [&](int u, int w) {
            tma_box<CH>(frames, total_rows, p1_box_x<CH>(u), p1_row0(half_a, u, w), &s->raw[w][0][0]);
            tma_box<CH>(frames, total_rows, p1_box_x<CH>(u), p1_row0(half_b, u, w), &s->raw[w][1][0]);
            s->raw_holds[w] = u;
        };
        // pull the staged rows of tile u into "registers" one step ahead, then hand the stage back for tile u + 1
        auto load_stage = [&](int u, int w) {
            if (s->raw_holds[w] != u) ++errors;  // ring protocol violated
            memcpy(s->regs[w], s->raw[w], sizeof s->regs[w]);  // every lane reads its rows BEFORE the refill
            s->regs_hold[w] = u;
            if (p1_live(u + 1, w, FA)) issue(u + 1, w);
        };
        for (int w = 0; w < kMainWarps; ++w) {
            if (p1_live(p1_first(w), w, FA)) issue(p1_first(w), w);
            if (sched_u(kTStart, 1, w) == p1_first(w) && p1_live(p1_first(w), w, FA)) load_stage(p1_first(w), w);
        }
        const int t_last = last_step(FA);
        for (int T = kTStart; T <= t_last; ++T) {
            bool b_touched[kMainWarps] = {}, t3_touched[kMainWarps] = {};
            for (int w = 0; w < kMainWarps; ++w) {
                const int u1 = sched_u(T, 1, w), u2 = sched_u(T, 2, w), u3 = sched_u(T, 3, w);
                StepArgs a;
                a.live1 = p1_live(u1, w, FA);
                a.live2 = p2_live(u2, FA);
                a.live3 = p34_live(u3, FA);
                a.swap2 = col_swap(u2);
                a.s1 = row_strip(u1);
                a.b2 = col_band(u2);
                a.s3 = row_strip(u3);
                a.tile_a = s->slot[T & 1][w];
                a.tile_b = s->slot[(T - 1) & 1][a.b2 & 7];
                a.t3_w = s->t3[T & 1] + (a.s3 & 7) * kT3Strip;
                if (a.live2) { if (b_touched[a.b2 & 7]) ++errors; b_touched[a.b2 & 7] = true; }
                if (a.live3) { if (t3_touched[a.s3 & 7]) ++errors; t3_touched[a.s3 & 7] = true; }
                if (a.live1 && s->regs_hold[w] != u1) ++errors;  // the registers must hold this step's rows
                const uint8_t (&staged)[2][kRawBoxBytesMax] = s->regs[w];
                for (int lane = 0; lane < 32; ++lane) {
                    uint32_t raw_a[kRawWords], raw_b[kRawWords];
                    memset(raw_a, 0, sizeof raw_a);
                    memset(raw_b, 0, sizeof raw_b);
                    if (a.live1) {
                        memcpy(raw_a, &staged[0][lane * kRawPitch], kRawPitch);
                        memcpy(raw_b, &staged[1][lane * kRawPitch], kRawPitch);
                    }
                    main_step<CH>(s->st[w][lane], a, raw_a, raw_b, lane, [] {});
                }
                if (p1_live(u1 + 1, w, FA)) load_stage(u1 + 1, w);
            }
            for (int lane = 0; lane < 32; ++lane) {  // the P4 warp
                const int g = lane >> 2, q = lane & 3;
                const int u4 = sched_u(T, 4, g);
                const bool live = p34_live(u4, FA);
                const int n = live ? u_pair(u4) : 0;
                P4Args a;
                a.live_a = live;
                a.live_b = live && n < FB;
                a.swap4 = col_swap(u4);
                a.b4 = col_band(u4);
                a.t3_r = s->t3[(T - 1) & 1] + g * kT3Strip + q * kT3Pitch;
                const int col = 4 * col_strip(u4, g) + q;
                a.out_a = a64 + (size_t)(half_a + n) * 4096 + col;
                a.out_b = a64 + (size_t)(a.live_b ? half_b + n : half_a + n) * 4096 + col;
                p4_step(s->p4[lane], a);
            }
        }
        free(s);
    }
    return errors;
}
}  // namespace

// channels = 3: frames [n][512][512][3] RGB24; channels = 1: [n][512][512] 8-bit gray
extern "C" __attribute__((visibility("default"))) int emu_fused2_a64(const uint8_t* frames, long long n_frames, int grid,
                                                                     float* a64, int channels) {
    return channels == 3 ? emu_run<3>(frames, n_frames, grid, a64) : emu_run<1>(frames, n_frames, grid, a64);
}
