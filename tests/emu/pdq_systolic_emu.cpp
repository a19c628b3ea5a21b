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

template <int CH>
struct WarpSim {
    using R = Raw<CH>;
    uint8_t ring[R::kWarpRingBytes];
    int slot_row[kGroups][kRing];   // stream row held by each ring slot, or INT_MIN while a copy is in flight / never loaded
    struct Pending {
        int E, g, s0;
        std::vector<uint8_t> data;
    };
    std::vector<Pending> pending;
    const uint8_t* frames;
    long long total_rows, first_row;  // of the whole batch / of this warp's first frame
    int F;
    int errors = 0;

    void issue(int E) {
        for (int g = 0; g < kGroups; ++g) {
            const int s0 = box_first_row(E, g);
            if (s0 < 0) continue;
            const int f = s0 / kStepsPerFrame, r0 = s0 % kStepsPerFrame;
            if (f >= F || r0 >= kImageRows) continue;
            Pending p{E, g, s0, std::vector<uint8_t>(R::kBoxBytes)};
            for (int i = 0; i < kBoxRows; ++i)
                for (int b = 0; b < R::kSegPitch; ++b) {
                    const long long row = first_row + (long long)f * 512 + r0 + i;
                    const int col = box_x<CH>(g) + b;
                    p.data[i * R::kSegPitch + b] =
                        (row >= 0 && row < total_rows && col < R::kRowBytes) ? frames[row * R::kRowBytes + col] : 0;
                }
            const int off = box_ring_offset<CH>(g, s0);
            memset(ring + off, 0xCD, R::kBoxBytes);
            for (int i = 0; i < kBoxRows; ++i) slot_row[g][(s0 + i) & (kRing - 1)] = -2000000000;
            pending.push_back(std::move(p));
        }
    }
    void wait(int E) {
        for (size_t i = 0; i < pending.size();) {
            if (pending[i].E == E) {
                const Pending& p = pending[i];
                memcpy(ring + box_ring_offset<CH>(p.g, p.s0), p.data.data(), R::kBoxBytes);
                for (int k = 0; k < kBoxRows; ++k) slot_row[p.g][(p.s0 + k) & (kRing - 1)] = p.s0 + k;
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
        if (slot_row[g][s & (kRing - 1)] != s) ++errors;
        memcpy(dst, ring + ring_offset<CH>(lane, s) + 16 * q, 16);
    }
};

template <int CH, int J>
void step_all(WarpSim<CH>& W, LaneState* st, int t, float* a64, long long f_first) {
    using R = Raw<CH>;
    RowChain out1[32], out3[32];
    const int next0 = t + 2;  // stream row of lane 0 two steps ahead (lane 31 prepares its prologue pixels)
    const bool next0_image = next0 >= 0 && next0 / kStepsPerFrame < W.F && next0 % kStepsPerFrame < kImageRows;
    for (int lane = 0; lane < 32; ++lane) {
        uint32_t w[R::kWords];
        memset(w, 0, sizeof w);
        if (st[lane].next_reads_image(W.F)) {  // the window of the NEXT step's row
            const int s = t + 1 - lane;
            if (s != st[lane].f * kStepsPerFrame + st[lane].r + 1) ++W.errors;
            for (int q = 0; q < R::kChunks - (lane == 31 ? 1 : 0); ++q) W.read_chunk(lane, s, q, w + 4 * q);
        }
        if (lane == 31 && next0_image) W.read_chunk(0, next0, 0, w + 4 * (R::kChunks - 1));
        lane_step<CH, J>(st[lane], w, lane, W.F, out1[lane], out3[lane], [&](int f, int i, float v0, float v1) {
            float* o = a64 + (size_t)(f_first + f) * 4096 + i * 64 + 2 * lane;
            o[0] = v0;
            o[1] = v1;
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
        for (int l = 0; l < 32; ++l) st[l].init(l);
        for (int E = 0; E < first_loop_event(); ++E) W->issue(E);
        int issued = first_loop_event() - 1, waited = -1;
        const int t_last = last_step(F);
        for (int t = kFirstStep; t <= t_last; ++t) {
            if ((t & 3) == kEventPhase) {
                const int Ew = (t + kWaitLead) / 4, Ei = (t + kIssueLead) / 4;
                if (Ew >= 0) { W->wait(Ew); waited = Ew; }
                W->issue(Ei);
                issued = Ei;
            }
            switch (t & 3) {
                case 0: step_all<CH, 0>(*W, st, t, a64, f_begin); break;
                case 1: step_all<CH, 1>(*W, st, t, a64, f_begin); break;
                case 2: step_all<CH, 2>(*W, st, t, a64, f_begin); break;
                default: step_all<CH, 3>(*W, st, t, a64, f_begin); break;
            }
        }
        for (int E = waited + 1; E <= issued; ++E) W->wait(E);
        if (!W->pending.empty()) ++errors;
        errors += W->errors;
        delete W;
    }
    return errors;
}
}  // namespace

// channels = 3: frames [n][512][512][3] RGB24; channels = 1: [n][512][512] 8-bit gray.  n_warps = warps the batch is
// split over (the kernel: grid x 8).
extern "C" __attribute__((visibility("default"))) int emu_systolic_a64(const uint8_t* frames, long long n_frames,
                                                                       int n_warps, float* a64, int channels) {
    return channels == 3 ? emu_run<3>(frames, n_frames, n_warps, a64) : emu_run<1>(frames, n_frames, n_warps, a64);
}
