// tests/emu/pdq_fused_emu.cpp -- CPU emulator of kx_fused_p123 (TEST INFRASTRUCTURE).
// Executes the schedule of hydrus_video_deduplicator_b200/csrc/pdq_fused_core.h step by step, phase by
// phase, warp by warp, lane by lane -- the same per-lane functions the CUDA kernel runs -- with the TMA box
// loads (incl. out-of-bounds zero fill) and the 2-deep staging ring modelled explicitly.  Also checks that
// within a phase no two warps touch the same tile slot.
// Build: g++ -O1 -ffp-contract=off -shared -fPIC -o libpdq_fused_emu.so pdq_fused_emu.cpp
#include <stdlib.h>
#include <vector>

#include "../../hydrus_video_deduplicator_b200/csrc/pdq_fused_core.h"

using namespace vpdq_core;

namespace {
struct Cta {
    Chain c1[16][32], c2[16][32], c3[16][32];
    float p0[16][32], p1[16][32];
    float slot[16][1024];
    uint8_t raw[2][16][kRawBoxBytes];
    int raw_holds[2][16];  // which tile index u is staged (for checking the ring protocol)
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

extern "C" __attribute__((visibility("default"))) int emu_fused_p3t(const uint8_t* frames, long long n_frames, int grid,
                                                                    float* p3t) {
    const long long total_rows = n_frames * 512;
    int errors = 0;
    for (int cta = 0; cta < grid; ++cta) {
        const long long f_begin = n_frames * cta / grid, f_end = n_frames * (cta + 1) / grid;
        const int F = (int)(f_end - f_begin);
        if (F == 0) continue;
        Cta* s = (Cta*)calloc(1, sizeof(Cta));
        for (int w = 0; w < 16; ++w)
            for (int l = 0; l < 32; ++l) {
                s->c1[w][l].init(); s->c2[w][l].init(); s->c3[w][l].init();
                s->p0[w][l] = s->p1[w][l] = 0.0f;
            }
        for (int st = 0; st < 2; ++st)
            for (int w = 0; w < 16; ++w) s->raw_holds[st][w] = -1000;
        auto issue = [&](int u, int w) {
            const int u_first = (w == 15) ? -16 : 0;
            const int st = (u - u_first) & 1;
            tma_box(frames, total_rows, p1_box_x(u & 15), p1_row0(f_begin, floor_div16(u), w), &s->raw[st][w][0]);
            s->raw_holds[st][w] = u;
        };
        for (int w = 0; w < 16; ++w) {
            const int u_first = (w == 15) ? -16 : 0;
            if (p1_live(u_first, w, F)) issue(u_first, w);
            if (p1_live(u_first + 1, w, F)) issue(u_first + 1, w);
        }
        const int steps = num_steps(F);
        for (int T = 0; T < steps; ++T) {
            // phase A
            for (int w = 0; w < 16; ++w) {
                const int u3 = sched_u3(T, w);
                if (p3_live(u3, F))
                    for (int l = 0; l < 32; ++l)
                        p3_lane(s->c3[w][l], s->slot[w], l, u3 & 15,
                                p3t + (size_t)(f_begin + (u3 >> 4)) * (64 * 512) + 32 * w + l);
                const int u = sched_u12(T, w);
                if (p1_live(u, w, F)) {
                    const int u_first = (w == 15) ? -16 : 0;
                    const int k = u - u_first, strip = u & 15;
                    if (s->raw_holds[k & 1][w] != u) ++errors;  // ring protocol violated
                    for (int l = 0; l < 32; ++l) {
                        uint32_t first2[2] = {0, 0};
                        if (strip == 0) {
                            const long long R = p1_row0(f_begin, floor_div16(u), w) + l;
                            if (R >= 0 && R < total_rows) memcpy(first2, frames + R * 1536, 8);
                        }
                        uint32_t raw[kRawWords];
                        memcpy(raw, &s->raw[k & 1][w][l * kRawPitch], kRawPitch);
                        p1_lane(s->c1[w][l], raw, first2, s->slot[w], l, strip);
                    }
                    if (p1_live(u + 2, w, F)) issue(u + 2, w);
                }
            }
            // phase B
            bool touched[16] = {};
            for (int w = 0; w < 16; ++w) {
                const int u = sched_u12(T, w);
                if (p2_live(u, F)) {
                    const int band = u & 15;
                    if (touched[band]) ++errors;  // two column roles on one slot in the same phase
                    touched[band] = true;
                    for (int c = 0; c < 32; ++c)
                        p2_lane(s->c2[w][c], s->p0[w][c], s->p1[w][c], s->slot[band], c, band, u < 0);
                }
            }
        }
        free(s);
    }
    return errors;
}
