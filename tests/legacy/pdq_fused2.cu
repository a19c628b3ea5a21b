// pdq_fused2.cu -- kx_fused_jarosz2: luma + the four Jarosz box-filter passes + 64x64 decimation of the PDQ hash
// in ONE persistent kernel, TWO frames per lane in packed fp32 pairs (FADD2 / FFMA2 / FMUL2, sm_100a), fp32
// intermediates never leaving the SM.  Same arithmetic, same order, bit-identical results as pdq_fused.cu /
// pdq_kernels.cu / the oracle; about half the issue slots per frame.
//
// Structure (index algebra and per-lane code: pdq_fused2_core.h, which the CPU emulator also compiles):
//   * persistent grid, one CTA of 8 main warps + 1 P4 warp per SM, each CTA owns a contiguous frame range cut
//     in two halves that are processed pairwise (lane component x = half A, y = half B);
//   * RGB rows are staged by TMA (cp.async.bulk.tensor.2d, two 32-row x 112-byte boxes per main warp per step,
//     one per frame, per-warp mbarrier; SASS: UTMALDG).  The staged rows are pulled into registers at the top
//     of the step and the next boxes are requested at once, so one stage per warp suffices.  Out-of-bounds box
//     parts come back as zeros, which is exactly what the drain steps of every running sum need;
//   * 32x32 tiles of float2 in shared memory (pitch 34 float2: conflict free for lane=row 128-bit and
//     lane=column 64-bit access), double buffered by step parity, hand the data from row to column roles;
//   * one CTA barrier per step keeps the wavefront (bulk-synchronous; no other inter-warp signalling).
#include <cuda.h>

#include <mutex>

#include "legacy.cuh"
#include "pdq_fused2_core.h"

namespace vpdq {
using namespace vpdq_core2;


constexpr int kFused2Threads = 32 * kWarps;  // 288

struct Fused2Smem {
    alignas(128) uint8_t raw[kMainWarps][2][kRawBoxBytesMax];  //  57 344 B  TMA destinations (one stage per warp, 2 frames)
    alignas(16) F2 slot[2][kMainWarps][kSlotF2];             // 139 264 B  tile hand-over, by step parity
    alignas(16) F2 t3[2][kMainWarps * kT3Strip];             //  18 432 B  P3 -> P4 hand-over, by step parity
    alignas(8) unsigned long long bar[kMainWarps];           //      64 B  mbarriers
};

__device__ int g_fused2_timeout = 0;  // set if an mbarrier wait gave up (never expected)

namespace {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_test_wait(unsigned long long* bar, uint32_t parity) {  // non-blocking
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    // bounded: a mis-programmed copy must not hang the GPU; ~1 s worth of polling, then flag and go on
#pragma unroll 1
    for (int spin = 0; spin < (1 << 24); ++spin)
        if (mbar_try_wait(bar, parity)) return;
    g_fused2_timeout = 1;
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
        : "memory");
}
// a 128-bit shared load the compiler may not narrow (the last one of a staged row only needs its low half, but a
// 64-bit access at the 112-byte row pitch is 2-way bank conflicted; the full 128-bit one is conflict free)
__device__ __forceinline__ uint4 lds128(const void* p) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(p)));
    return v;
}
__device__ __forceinline__ void cta_barrier() { asm volatile("bar.sync 0, %0;" ::"n"(kFused2Threads) : "memory"); }
}  // namespace

template <int CH>  // 3: RGB24 frames, 1: 8-bit gray frames (== R = G = B)
__global__ void __launch_bounds__(kFused2Threads, 1)
    kx_fused_jarosz2(const __grid_constant__ CUtensorMap tmap, long long n_frames_total, float* __restrict__ a64) {
    extern __shared__ __align__(128) uint8_t smem_bytes2[];
    Fused2Smem& sm = *reinterpret_cast<Fused2Smem*>(smem_bytes2);
    // (the shuffle tells the compiler that w is warp-uniform: index arithmetic and TMA operands go to uniform registers)
    const int w = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

    const long long f_begin = n_frames_total * blockIdx.x / gridDim.x;
    const long long f_end = n_frames_total * (blockIdx.x + 1) / gridDim.x;
    const int F = (int)(f_end - f_begin);
    if (F == 0) return;
    const int FA = (F + 1) >> 1, FB = F - FA;
    const long long half_a = f_begin, half_b = f_begin + FA;

    if (w < kMainWarps && lane == 0) mbar_init(&sm.bar[w], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    const int t_last = last_step(FA);

    if (w == kMainWarps) {
        // ---------------- the P4 warp ----------------
        const int g = lane >> 2, q = lane & 3;
        P4State st;
        st.init();
        for (int T = kTStart; T <= t_last; ++T) {
            const int u4 = sched_u(T, 4, g);
            const bool live = p34_live(u4, FA);
            const int n = live ? u_pair(u4) : 0;
            P4Args a;
            a.live_a = live;
            a.live_b = live && n < FB;
            a.swap4 = col_swap(u4);
            a.b4 = col_band(u4);
            a.t3_r = sm.t3[(T - 1) & 1] + g * kT3Strip + q * kT3Pitch;
            const int col = 4 * col_strip(u4, g) + q;
            a.out_a = a64 + (size_t)(half_a + n) * 4096 + col;
            a.out_b = a64 + (size_t)(a.live_b ? half_b + n : half_a + n) * 4096 + col;
            p4_step(st, a);
            cta_barrier();
        }
        return;
    }

    // ---------------- main warps ----------------
    const int u_first = p1_first(w);
    constexpr int kRawPitch = Raw<CH>::kPitch, kRawWords = Raw<CH>::kWords;
    auto issue = [&](int u) {  // lane 0: stage the raw boxes (both frames) of P1 tile u of this warp
        mbar_expect_tx(&sm.bar[w], 2 * Raw<CH>::kBoxBytes);
        tma_load_2d(&sm.raw[w][0][0], &tmap, p1_box_x<CH>(u), (int)p1_row0(half_a, u, w), &sm.bar[w]);
        tma_load_2d(&sm.raw[w][1][0], &tmap, p1_box_x<CH>(u), (int)p1_row0(half_b, u, w), &sm.bar[w]);
    };
    if (lane == 0 && p1_live(u_first, w, FA)) issue(u_first);

    // The staged rows of tile u are pulled into registers one step ahead (at the end of the step before the one
    // that consumes them) and the stage goes straight back to TMA for tile u + 1, which then has a whole step to
    // land.  LDS results arrive asynchronously: issuing the loads is not enough, every lane must HOLD its rows
    // before the async proxy may overwrite them (see pdq_fused.cu).  The vote consumes one word of each LDS.128
    // -> the scoreboard wait happens there, and the ballot doubles as the warp-wide rendezvous.  (The magic value
    // never matches on all lanes.)
    uint32_t raw_a[kRawWords], raw_b[kRawWords];
#pragma unroll
    for (int q = 0; q < kRawWords; ++q) raw_a[q] = raw_b[q] = 0u;
    auto load_stage = [&](int u, bool landed) {
        if (!landed) mbar_wait(&sm.bar[w], (uint32_t)((u - u_first) & 1));
        const uint8_t* ra = &sm.raw[w][0][lane * kRawPitch];
        const uint8_t* rb = &sm.raw[w][1][lane * kRawPitch];
        uint32_t dep = 0;
#pragma unroll
        for (int q = 0; q < kRawWords / 4; ++q) {
            const uint4 va = lds128(ra + 16 * q), vb = lds128(rb + 16 * q);
            raw_a[4 * q + 0] = va.x; raw_a[4 * q + 1] = va.y; raw_a[4 * q + 2] = va.z; raw_a[4 * q + 3] = va.w;
            raw_b[4 * q + 0] = vb.x; raw_b[4 * q + 1] = vb.y; raw_b[4 * q + 2] = vb.z; raw_b[4 * q + 3] = vb.w;
            dep ^= va.x ^ vb.x;
        }
        const unsigned held = __ballot_sync(0xffffffffu, dep != 0x5bd1e995u);
        if (lane == 0 && held != 0u && p1_live(u + 1, w, FA)) issue(u + 1);
    };
    if (sched_u(kTStart, 1, w) == u_first && p1_live(u_first, w, FA)) load_stage(u_first, false);  // warp 7

    LaneState st;
    st.init();

    for (int T = kTStart; T <= t_last; ++T) {
        const int u1 = sched_u(T, 1, w), u2 = sched_u(T, 2, w), u3 = sched_u(T, 3, w);
        StepArgs a;
        a.live1 = p1_live(u1, w, FA);
        a.live2 = p2_live(u2, FA);
        a.live3 = p34_live(u3, FA);
        a.swap2 = col_swap(u2);
        a.s1 = row_strip(u1);
        a.b2 = col_band(u2);
        a.s3 = row_strip(u3);
        a.tile_a = sm.slot[T & 1][w];
        a.tile_b = sm.slot[(T - 1) & 1][a.b2 & 7];
        a.t3_w = sm.t3[T & 1] + (a.s3 & 7) * kT3Strip;

        // (when P1 is not live the raw registers are stale; its results are never stored)
        const bool next_live = p1_live(u1 + 1, w, FA);
        bool landed = false;  // the next step's boxes, polled late in this step (hides the poll's latency)
        main_step<CH>(st, a, raw_a, raw_b, lane, [&]() {
            if (next_live) landed = mbar_test_wait(&sm.bar[w], (uint32_t)((u1 + 1 - u_first) & 1));
        });
        if (next_live) load_stage(u1 + 1, landed);
        cta_barrier();
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
int fused2_debug_flags(int* flags) {
    int v = 0;
    VPDQ_CUDA(cudaMemcpyFromSymbol(&v, g_fused2_timeout, sizeof v));
    *flags = v;
    return VPDQ_B200_OK;
}

// RGB24 (channels = 3) or 8-bit gray (channels = 1) frames -> a64 [n][64][64]: the Jarosz-filtered, decimated luma plane
int fused2_jarosz_launch(const uint8_t* d_frames, int channels, int64_t n_frames, float* d_a64, cudaStream_t stream) {
    CUtensorMap tmap;
    int rc = fused_make_tensor_map(d_frames, n_frames, channels, &tmap);
    if (rc) return rc;
    int dev = 0, sms = 148;
    VPDQ_CUDA(cudaGetDevice(&dev));
    VPDQ_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    static std::mutex mu;
    static bool attr_done[64] = {};
    {
        std::lock_guard<std::mutex> lk(mu);
        if (dev < 0 || dev >= 64 || !attr_done[dev]) {
            VPDQ_CUDA(cudaFuncSetAttribute(kx_fused_jarosz2<3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)sizeof(Fused2Smem)));
            VPDQ_CUDA(cudaFuncSetAttribute(kx_fused_jarosz2<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)sizeof(Fused2Smem)));
            if (dev >= 0 && dev < 64) attr_done[dev] = true;
        }
    }
    const unsigned grid = (unsigned)(n_frames < sms ? n_frames : sms);  // persistent: one CTA per SM
    if (channels == 3)
        kx_fused_jarosz2<3><<<grid, kFused2Threads, sizeof(Fused2Smem), stream>>>(tmap, (long long)n_frames, d_a64);
    else
        kx_fused_jarosz2<1><<<grid, kFused2Threads, sizeof(Fused2Smem), stream>>>(tmap, (long long)n_frames, d_a64);
    g_launches += 1;
    VPDQ_CUDA(cudaGetLastError());
    return VPDQ_B200_OK;
}

}  // namespace vpdq
