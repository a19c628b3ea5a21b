// tests/emu/pdq_fused_emu.cpp -- CPU emulator of kx_fused_jarosz (TEST INFRASTRUCTURE).
// Executes the schedule of hydrus_video_deduplicator_b200/csrc/pdq_fused_core.h step by step, warp by warp,
// lane by lane -- the same fused_step() the CUDA kernel runs -- with the TMA box loads (incl. out-of-bounds
// zero fill) and the one-stage staging ring modelled explicitly.  Also checks that within a step no two
// warps touch the same tile slot / t3 strip.
// Build: g++ -O1 -ffp-contract=off -shared -fPIC -o libpdq_fused_emu.so pdq_fused_emu.cpp
#include <stdlib.h>

#include "../legacy/pdq_fused_core.h"

using namespace vpdq_core;

namespace {
struct Cta {
    LaneState st[16][32];
    float slot[2][16][1024];
    float t3[2][16 * kT3Strip];
    uint8_t raw[16][kRawBoxBytes];
    int raw_holds[16];  // which tile index u is staged (checks the ring protocol)
};

void tma_box(const uint8_t* frames, long long total_rows, int x, long long y, uint8_t* dst) {
    for (int r = 0; r < kTile; ++r)
        for (int b = 0; b < kRawPitch; ++b) {
            const long long row = y + r;
            const int col = x + b;
            dst[r * kRawPitch + b] =
                (row >= 0 && row < total_rows && col >= 0 && col < 1536) ? frames[row * 1536 + col] : 0;
        }
}
}  // namespace

extern "C" __attribute__((visibility("default"))) int emu_fused_a64(const uint8_t* frames, long long n_frames, int grid,
                                                                    float* a64) {
    const long long total_rows = n_frames * 512;
    int errors = 0;
    for (int cta = 0; cta < grid; ++cta) {
        const long long f_begin = n_frames * cta / grid, f_end = n_frames * (cta + 1) / grid;
        const int F = (int)(f_end - f_begin);
        if (F == 0) continue;
        Cta* s = (Cta*)calloc(1, sizeof(Cta));
        for (int w = 0; w < 16; ++w) {
            for (int l = 0; l < 32; ++l) s->st[w][l].init();
            s->raw_holds[w] = -1000;
        }
        auto issue = [&](int u, int w) {
            tma_box(frames, total_rows, p1_box_x(u & 15), p1_row0(f_begin, floor_div16(u), w), &s->raw[w][0]);
            s->raw_holds[w] = u;
        };
        for (int w = 0; w < 16; ++w)
            if (p1_live(p1_first(w), w, F)) issue(p1_first(w), w);
        const int steps = num_steps(F);
        for (int T = 0; T < steps; ++T) {
            bool b_touched[16] = {}, t3_touched[16] = {};
            for (int w = 0; w < 16; ++w) {
                const int u1 = sched_u(T, 1, w), u2 = sched_u(T, 2, w), u3 = sched_u(T, 3, w), u4 = sched_u(T, 4, w);
                StepArgs a;
                a.live1 = p1_live(u1, w, F);
                a.live2 = p2_live(u2, F);
                a.live3 = p34_live(u3, F);
                a.live4 = p34_live(u4, F);
                a.s1 = u1 & 15; a.b2 = u2 & 15; a.s3 = u3 & 15; a.b4 = u4 & 15;
                a.tile_a = s->slot[T & 1][w];
                a.tile_b = s->slot[(T - 1) & 1][a.b2];
                a.t3_w = s->t3[T & 1] + a.s3 * kT3Strip;
                a.t3_r = s->t3[(T - 1) & 1] + w * kT3Strip;
                if (a.live2) { if (b_touched[a.b2]) ++errors; b_touched[a.b2] = true; }
                if (a.live3) { if (t3_touched[a.s3]) ++errors; t3_touched[a.s3] = true; }
                if (a.live1 && s->raw_holds[w] != u1) ++errors;  // ring protocol violated
                uint8_t staged[kRawBoxBytes];
                memcpy(staged, s->raw[w], kRawBoxBytes);  // every lane reads its row BEFORE the refill
                if (a.live1 && p1_live(u1 + 1, w, F)) issue(u1 + 1, w);
                for (int lane = 0; lane < 32; ++lane) {
                    a.a_out = a64 + (size_t)(a.live4 ? (f_begin + (u4 >> 4)) : f_begin) * 4096 + 4 * w + (lane & 3);
                    uint32_t first2[2] = {0, 0};
                    uint32_t raw[kRawWords];
                    memset(raw, 0, sizeof raw);
                    if (a.live1) {
                        if (a.s1 == 0) {
                            const long long R = p1_row0(f_begin, floor_div16(u1), w) + lane;
                            if (R >= 0 && R < total_rows) memcpy(first2, frames + R * 1536, 8);
                        }
                        memcpy(raw, &staged[lane * kRawPitch], kRawPitch);
                    }
                    fused_step(s->st[w][lane], a, raw, first2, lane, [] {});
                }
            }
        }
        free(s);
    }
    return errors;
}
