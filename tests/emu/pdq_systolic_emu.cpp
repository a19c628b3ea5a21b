// tests/emu/pdq_systolic_emu.cpp -- CPU emulator of kx_systolic_jarosz (TEST INFRASTRUCTURE).
// Executes the schedule of hydrus_video_deduplicator_b200/csrc/pdq_systolic_core.h step by step, lane by lane --
// the same lane_step() the CUDA kernel runs -- with the rotate-shuffle hand-over, the per-group raw rings and
// the TMA events modelled explicitly:
//   * ISSUE(E) invalidates (and poisons) the destination boxes at once and the data only lands at WAIT(E): a read
//     between the two, or a read of a slot that holds another stream row than the reader expects (a box
//     overwritten too early), is counted as an error;
//   * out-of-bounds box parts are zero filled like TMA does.
// Build: g++ -O1 -ffp-contract=off -shared -fPIC -o libpdq_systolic_emu.so pdq_systolic_emu.cpp
#include <stdlib.h>

#include <vector>

#include "../../hydrus_video_deduplicator_b200/csrc/pdq_systolic_core.h"

using namespace vpdq_sys;

namespace {

bool g_general_only = false;  // model VPDQ_B200_SYSTOLIC_3D=0: no plain iterations, every event as per-group boxes

template <int CH>
struct WarpSim {
    using R = Raw<CH>;
    uint8_t ring[R::kWarpRingBytes];
    int slot_row[kGroups][kRing];   // stream row held by each ring row of a group, or a sentinel while a copy is in flight
    struct Pending {
        int E, off, bytes;
        std::vector<uint8_t> data;
        std::vector<int> g, s0;     // which (group, first stream row) boxes it carries
    };
    std::vector<Pending> pending;
    const uint8_t* frames;
    long long total_rows, first_row;  // of the whole batch / of this warp's first frame
    int F;
    int errors = 0;
    int one_box_events = 0, split_events = 0;

    static int row_index(int g, int s) { return box_slot_of((s + kGroupLanes * g) >> 2) * 4 + (s & 3); }
    void invalidate(int g, int s0) {
        for (int i = 0; i < kBoxRows; ++i) slot_row[g][row_index(g, s0 + i)] = -2000000000;
    }
    void issue(int E) {
        const int s00 = box_first_row(E, 0);
        const int f0 = s00 / kStepsPerFrame, r00 = s00 % kStepsPerFrame;
        if (!g_general_only && s00 >= 0 && event_is_one_box(f0, r00, F)) {
            // ONE 3-D box (x, row, g') at (0, Y0 - 28, 0): element address = kBase3 + x + kRowBytes*y + kStride3*g'
            ++one_box_events;
            Pending p{E, box_ring_offset<CH>(kGroups - 1, E), kGroups * R::kBoxBytes, std::vector<uint8_t>(kGroups * R::kBoxBytes), {}, {}};
            const long long y0 = first_row + (long long)f0 * 512 + r00 - View3<CH>::kBackRows;
            for (int gp = 0; gp < kGroups; ++gp)
                for (int i = 0; i < kBoxRows; ++i)
                    for (int x = 0; x < R::kSegPitch; ++x) {
                        const long long addr = View3<CH>::kBase3 + x + (long long)R::kRowBytes * (y0 + i) + View3<CH>::kStride3 * gp;
                        if (addr < 0 || addr >= total_rows * R::kRowBytes) {  // the one-box path must never leave the buffer
                            ++errors;
                            continue;
                        }
                        p.data[(gp * kBoxRows + i) * R::kSegPitch + x] = frames[addr];
                    }
            for (int g = 0; g < kGroups; ++g) {
                p.g.push_back(g);
                p.s0.push_back(box_first_row(E, g));
                invalidate(g, box_first_row(E, g));
            }
            memset(ring + p.off, 0xCD, p.bytes);
            pending.push_back(std::move(p));
            return;
        }
        ++split_events;
        for (int g = 0; g < kGroups; ++g) {
            const int s0 = box_first_row(E, g);
            if (s0 < 0) continue;
            const int f = s0 / kStepsPerFrame, r0 = s0 % kStepsPerFrame;
            if (f >= F || r0 >= kImageRows) continue;
            Pending p{E, box_ring_offset<CH>(g, E), R::kBoxBytes, std::vector<uint8_t>(R::kBoxBytes), {g}, {s0}};
            for (int i = 0; i < kBoxRows; ++i)
                for (int b = 0; b < R::kSegPitch; ++b) {
                    const long long row = first_row + (long long)f * 512 + r0 + i;
                    const int col = box_x<CH>(g) + b;
                    p.data[i * R::kSegPitch + b] =
                        (row >= 0 && row < total_rows && col < R::kRowBytes) ? frames[row * R::kRowBytes + col] : 0;
                }
            memset(ring + p.off, 0xCD, p.bytes);
            invalidate(g, s0);
            pending.push_back(std::move(p));
        }
    }
    void wait(int E) {
        for (size_t i = 0; i < pending.size();) {
            if (pending[i].E == E) {
                const Pending& p = pending[i];
                memcpy(ring + p.off, p.data.data(), p.bytes);
                for (size_t b = 0; b < p.g.size(); ++b)
                    for (int k = 0; k < kBoxRows; ++k) slot_row[p.g[b]][row_index(p.g[b], p.s0[b] + k)] = p.s0[b] + k;
                pending.erase(pending.begin() + i);
            } else {
                ++i;
            }
        }
    }
    // the 16-byte chunk q of `lane`'s window of stream row s
    void read_chunk(int lane, int s, int q, uint32_t* dst) {
        const int g = lane >> 2;
        // a window's last chunk may reach into the next group's share of the row: still this group's box
        if (slot_row[g][row_index(g, s)] != s) ++errors;
        memcpy(dst, ring + ring_offset<CH>(lane, s) + 16 * q, 16);
    }
};

template <int CH, int T, bool PLAIN>
void step_all(WarpSim<CH>& W, LaneState* st, float** optr, int t) {
    using R = Raw<CH>;
    RowChain out1[32], out3[32];
    const int next0 = t + 2;  // stream row of lane 0 two steps ahead (lane 31 prepares its prologue pixels)
    const bool next0_image = next0 >= 0 && next0 / kStepsPerFrame < W.F && next0 % kStepsPerFrame < kImageRows;
    for (int lane = 0; lane < 32; ++lane) {
        uint32_t w[R::kWords];
        memset(w, 0, sizeof w);
        const int s = t + 1 - lane;  // the NEXT step's row of this lane
        const bool want = s >= 0 && s / kStepsPerFrame < W.F && s % kStepsPerFrame < kImageRows;
        if (want != st[lane].img_next(W.F)) ++W.errors;  // the lane's own view must agree with the stream position
        if (PLAIN) {  // what a plain iteration takes for granted, per lane
            const int row = t - lane;
            if (!want || !next0_image || row < 0 || st[lane].r + T != row % kStepsPerFrame ||
                st[lane].f != row / kStepsPerFrame || row % kStepsPerFrame < 4 || row % kStepsPerFrame > kImageRows - 3)
                ++W.errors;
        }
        if (st[lane].img_next(W.F))
            for (int q = 0; q < R::kChunks - (lane == 31 ? 1 : 0); ++q) W.read_chunk(lane, s, q, w + 4 * q);
        if (lane == 31 && next0_image) W.read_chunk(0, next0, 0, w + 4 * (R::kChunks - 1));
        lane_step<CH, T, PLAIN>(st[lane], w, lane, W.F, out1[lane], out3[lane], [&](float v0, float v1) {
            optr[lane][0] = v0;
            optr[lane][1] = v1;
            optr[lane] += 64;
        });
    }
    for (int lane = 0; lane < 32; ++lane) {
        st[lane].in1 = out1[(lane + 31) & 31];
        st[lane].in3 = out3[(lane + 31) & 31];
    }
}

template <int CH>
int emu_run(const uint8_t* frames, long long n_frames, int n_warps, float* a64) {
    int errors = 0;
    for (int wi = 0; wi < n_warps; ++wi) {
        const long long f_begin = n_frames * wi / n_warps, f_end = n_frames * (wi + 1) / n_warps;
        const int F = (int)(f_end - f_begin);
        if (F == 0) continue;
        WarpSim<CH>* W = new WarpSim<CH>;
        memset(W->ring, 0xCD, sizeof W->ring);
        for (int g = 0; g < kGroups; ++g)
            for (int s = 0; s < kRing; ++s) W->slot_row[g][s] = -2000000000;
        W->frames = frames;
        W->total_rows = n_frames * 512;
        W->first_row = f_begin * 512;
        W->F = F;
        LaneState st[32];
        float* optr[32];
        for (int l = 0; l < 32; ++l) {
            st[l].init(l);
            optr[l] = a64 + (size_t)f_begin * 4096 + 2 * l;
        }
        for (int E = 0; E < first_loop_event(); ++E) W->issue(E);
        int issued = first_loop_event() - 1, waited = -1;
        const int t_last = last_step(F);
        int f0 = -1, r0 = kStepsPerFrame + kFirstStep;  // lane 0's stream position at the first step of the iteration
        bool plain = false;
        int plain_iterations = 0, iterations = 0;
        for (int t = kFirstStep; t <= t_last || (t - kFirstStep) % kBody != 0; ++t) {  // (the kernel finishes its last body)
            const int T = (t - kFirstStep) % kBody;  // position in the loop body (kFirstStep is a multiple of kBody)
            if (T == 0) {
                plain = !g_general_only && iteration_is_plain(f0, r0, F);
                plain_iterations += plain;
                ++iterations;
            }
            if ((t & 3) == kEventPhase) {
                const int Ew = (t + kWaitLead) / 4, Ei = (t + kIssueLead) / 4;
                if (Ew >= 0) { W->wait(Ew); waited = Ew; }
                if (plain) {  // the kernel issues the 3-D box at row plain_event_row(r0) of frame f0 without looking
                    const int s00 = box_first_row(Ei, 0), before = W->one_box_events;
                    if (Ew < 0 || s00 != f0 * kStepsPerFrame + plain_event_row(r0 + T - kEventPhase)) ++W->errors;
                    W->issue(Ei);
                    if (W->one_box_events != before + 1) ++W->errors;
                } else {
                    W->issue(Ei);
                }
                issued = Ei;
            }
            switch (T + (plain ? 8 : 0)) {
                case 0: step_all<CH, 0, false>(*W, st, optr, t); break;
                case 1: step_all<CH, 1, false>(*W, st, optr, t); break;
                case 2: step_all<CH, 2, false>(*W, st, optr, t); break;
                case 3: step_all<CH, 3, false>(*W, st, optr, t); break;
                case 4: step_all<CH, 4 % kBody, false>(*W, st, optr, t); break;
                case 5: step_all<CH, 5 % kBody, false>(*W, st, optr, t); break;
                case 6: step_all<CH, 6 % kBody, false>(*W, st, optr, t); break;
                case 7: step_all<CH, 7 % kBody, false>(*W, st, optr, t); break;
                case 8: step_all<CH, 0, true>(*W, st, optr, t); break;
                case 9: step_all<CH, 1, true>(*W, st, optr, t); break;
                case 10: step_all<CH, 2, true>(*W, st, optr, t); break;
                case 11: step_all<CH, 3, true>(*W, st, optr, t); break;
                case 12: step_all<CH, 4 % kBody, true>(*W, st, optr, t); break;  // (plain iterations: kBody 4 or 8)
                case 13: step_all<CH, 5 % kBody, true>(*W, st, optr, t); break;
                case 14: step_all<CH, 6 % kBody, true>(*W, st, optr, t); break;
                default: step_all<CH, 7 % kBody, true>(*W, st, optr, t); break;
            }
            if (T == kBody - 1) {
                r0 += kBody;
                if (r0 >= kStepsPerFrame) {
                    r0 -= kStepsPerFrame;
                    ++f0;
                }
            }
        }
        if (!g_general_only && F > 1 && plain_iterations * 10 < iterations * 8) ++errors;  // the plain body must be the common one
        for (int E = waited + 1; E <= issued; ++E) W->wait(E);
        if (!W->pending.empty()) ++errors;
        for (int l = 0; l < 32; ++l)
            if (optr[l] != a64 + (size_t)(f_begin + F) * 4096 + 2 * l) ++errors;  // every decimated row was emitted, in order
        if (!g_general_only && F > 1 && W->one_box_events < 100 * F) ++errors;  // the one-box path must be the common one
        errors += W->errors;
        delete W;
    }
    return errors;
}
}  // namespace

// channels = 3: frames [n][512][512][3] RGB24; channels = 1: [n][512][512] 8-bit gray.  n_warps = warps the batch is
// split over (the kernel: grid x 8).
extern "C" __attribute__((visibility("default"))) void emu_systolic_general_only(int on) { g_general_only = on != 0; }

extern "C" __attribute__((visibility("default"))) int emu_systolic_a64(const uint8_t* frames, long long n_frames,
                                                                       int n_warps, float* a64, int channels) {
    return channels == 3 ? emu_run<3>(frames, n_frames, n_warps, a64) : emu_run<1>(frames, n_frames, n_warps, a64);
}
