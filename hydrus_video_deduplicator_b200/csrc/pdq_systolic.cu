// pdq_systolic.cu -- kx_systolic_jarosz: luma + the four Jarosz box-filter passes + the 64x64 decimation of the PDQ
// hash, ONE WARP PER FRAME, no shared-memory transposition between the passes (the design and the index algebra:
// pdq_systolic_core.h, which the CPU emulator tests/emu/pdq_systolic_emu.cpp also compiles).
//
//   * persistent grid: one CTA of 8 warps per SM; every warp owns a contiguous range of frames and streams them
//     row by row.  Lane l owns image columns 16 l .. 16 l + 15: the column passes are private running sums in
//     registers, the row passes run along the lanes -- the chain state (5 floats per pass) moves to the next
//     lane with one rotate-shuffle per step, so lane l works on stream row t - l at step t;
//   * raw rows are staged by TMA (cp.async.bulk.tensor.3d / .2d; SASS UTMALDG.3D / .2D): per warp 8 lane-group rings
//     of 16 stream rows x 224 bytes, filled by boxes of 4 rows -- in the common case all eight group boxes of an event
//     are ONE 3-D box -- one mbarrier per event, 8 steps of lead, two events in flight.  A lane reads its 54-byte
//     window of its row with 4 LDS.128;
//   * the step loop has two bodies of 8 steps, both branch free: runs of PLAIN iterations (every lane on a row
//     4..509 of a live frame, one-box events: 88 % of a frame) are an inner loop of their own without a single row
//     test; the general body handles the rare rows per lane by what the row number selects (pdq_systolic_core.h);
//   * results (the decimated plane a64 [n][64][64] f32) go straight from registers to global memory, one store per
//     lane and iteration.
//
// Same arithmetic, same order, bit-identical results as the oracle (and as the tiled kernels it replaces).
#include <cuda.h>
#include <stdlib.h>

#include <mutex>

#include "common.cuh"
#include "pdq_systolic_core.h"

namespace vpdq {
using namespace vpdq_sys;

#ifndef VPDQS_WARPS
#define VPDQS_WARPS 8
#endif
constexpr int kSysWarps = VPDQS_WARPS;  // warps (= frames in flight) per SM: bounded by the raw rings in shared memory
constexpr int kSysThreads = 32 * kSysWarps;

template <int CH>
struct SysSmem {
    alignas(128) uint8_t ring[kSysWarps][Raw<CH>::kWarpRingBytes];
    alignas(128) uint8_t zeros[128];                    // the window of a stream row that is not an image row
    alignas(8) unsigned long long bar[kSysWarps][2];    // one mbarrier per in-flight event
};

__device__ int g_systolic_timeout = 0;  // set if an mbarrier wait gave up (never expected); read by the host at its sync points

namespace {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    // bounded: a mis-programmed copy must not hang the GPU; ~1 s worth of polling, then flag (the host turns the
    // flag into VPDQ_B200_ERR_CUDA at its next synchronisation point)
#pragma unroll 1
    for (int spin = 0; spin < (1 << 24); ++spin)
        if (mbar_try_wait(bar, parity)) return;
    g_systolic_timeout = 1;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(bar)
        : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
}  // namespace

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int x, int y, int z, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(z), "r"(bar)
        : "memory");
}

template <int CH>  // 3: RGB24 frames, 1: 8-bit gray frames (== R = G = B)
__global__ void __launch_bounds__(kSysThreads, 1)
    kx_systolic_jarosz(const __grid_constant__ CUtensorMap tmap2, const __grid_constant__ CUtensorMap tmap3, int use3d,
                       long long n_frames_total, float* __restrict__ a64) {
    using R = Raw<CH>;
    extern __shared__ __align__(128) uint8_t smem_sys[];
    SysSmem<CH>& sm = *reinterpret_cast<SysSmem<CH>*>(smem_sys);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

    if (threadIdx.x < 32) reinterpret_cast<uint32_t*>(sm.zeros)[threadIdx.x] = 0u;
    if (lane == 0) {
        mbar_init(smem_u32(&sm.bar[warp][0]), 1);
        mbar_init(smem_u32(&sm.bar[warp][1]), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    // warp index interleaved over the CTAs: a small batch spreads over the SMs first
    const long long wi = (long long)warp * gridDim.x + blockIdx.x, n_warps = (long long)kSysWarps * gridDim.x;
    // balanced split, the extra frames on the lowest warp indices (= spread over the CTAs)
    const long long base = n_frames_total / n_warps, rem = n_frames_total % n_warps;
    const long long f_begin = wi * base + (wi < rem ? wi : rem);
    const int F = (int)(base + (wi < rem ? 1 : 0));
    if (F == 0) return;
    const int first_row = (int)(f_begin * 512);

    const uint32_t ring = smem_u32(&sm.ring[warp][0]), zeros = smem_u32(sm.zeros);
    const uint32_t bar0 = smem_u32(&sm.bar[warp][0]);
    const int grp = lane >> 2;
    const uint32_t lane_base = ring + ring_group_offset<CH>(grp) + (lane & 3) * R::kLaneBytes;
    const int lane_a = 1 - lane + kGroupLanes * grp;  // window row of step t: a = t + lane_a (see ring_row_offset)
    const uint32_t grp0_base = ring + ring_group_offset<CH>(0);

    // generic form of an event: lanes 0..7 each stage the 4-row box of their group (2-D map), if it exists
    auto issue_split = [&](int E) {
        bool has = false;
        int y = 0;
        uint32_t dst = 0;
        if (lane < kGroups) {
            const int s0 = box_first_row(E, lane);
            if (s0 >= 0) {
                const int f = s0 / kStepsPerFrame, r0 = s0 - f * kStepsPerFrame;
                if (f < F && r0 < kImageRows) {
                    has = true;
                    y = first_row + f * 512 + r0;
                    dst = ring + box_ring_offset<CH>(lane, E);
                }
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, has);
        const uint32_t bar = bar0 + 8 * (E & 1);
        if (lane == 0) {
            if (m)
                mbar_expect_tx(bar, __popc(m) * R::kBoxBytes);
            else
                mbar_arrive(bar);
        }
        __syncwarp();
        if (has) tma_load_2d(dst, &tmap2, box_x<CH>(lane), y, bar);
    };
    // an event in the general form: ONE 3-D box when all eight group boxes lie inside one live frame, else per group
    auto issue = [&](int E) {
        const int s00 = box_first_row(E, 0);
        const int ev_f = s00 >= 0 ? s00 / kStepsPerFrame : -1, ev_r = s00 - ev_f * kStepsPerFrame;
        if (use3d && s00 >= 0 && event_is_one_box(ev_f, ev_r, F)) {
            if (lane == 0) {
                const uint32_t bar = bar0 + 8 * (E & 1);
                mbar_expect_tx(bar, kGroups * R::kBoxBytes);
                tma_load_3d(ring + box_slot_of(E) * (kGroups * R::kBoxBytes), &tmap3, 0,
                            first_row + ev_f * 512 + ev_r - View3<CH>::kBackRows, 0, bar);
            }
            __syncwarp();
        } else {
            issue_split(E);
        }
    };
    auto wait = [&](int E) { mbar_wait(bar0 + 8 * (E & 1), (uint32_t)((E >> 1) & 1)); };

    LaneState L;
    L.init(lane);

    for (int E = 0; E < first_loop_event(); ++E) issue_split(E);
    int issued = first_loop_event() - 1, waited = -1;

    float* optr = a64 + (size_t)f_begin * 4096 + 2 * lane;  // decimated rows are emitted in order
    auto emit = [&](float v0, float v1) {
        *reinterpret_cast<float2*>(optr) = make_float2(v0, v1);
        optr += 64;
    };

    // stream position of lane 0 at the first step of the iteration (uniform)
    int f0 = -1, r0 = kStepsPerFrame + kFirstStep;
    int plain_y = 0;  // see the plain run below
    bool landed = false;

    // One step.  PLAIN (pdq_systolic_core.h, iteration_is_plain): every window is staged, the event is one 3-D box.
    auto step = [&](int t, auto jtag, auto ptag) {
        constexpr int T = decltype(jtag)::value;
        constexpr bool PLAIN = decltype(ptag)::value;
        // events every 4 steps, at a compile-time position in the body.  One step before an event: probe its barrier,
        // so that the (normal) answer "landed" costs no latency there
        if ((T & 3) == kEventPhase - 1) {
            const int Ew = (t + 1 + kWaitLead) >> 2;
            landed = (PLAIN || Ew >= 0) ? mbar_try_wait(bar0 + 8 * (Ew & 1), (uint32_t)((Ew >> 1) & 1)) : true;
        }
        if ((T & 3) == kEventPhase) {
            const int Ew = (t + kWaitLead) >> 2, Ei = (t + kIssueLead) >> 2;
            if (PLAIN || Ew >= 0) {
                if (!landed) wait(Ew);
                waited = Ew;
            }
            if (PLAIN) {
                if (lane == 0) {
                    const uint32_t bar = bar0 + 8 * (Ei & 1);
                    mbar_expect_tx(bar, kGroups * R::kBoxBytes);
                    tma_load_3d(ring + box_slot_of(Ei) * (kGroups * R::kBoxBytes), &tmap3, 0, plain_y + t, 0, bar);
                }
                __syncwarp();  // (measured: leaving it out is 1 % slower)
            } else {
                issue(Ei);
            }
            issued = Ei;
        }
        // the raw window of the NEXT step's row (its lumas are this step's filler work)
        uint32_t w[R::kWords];
        {
            const int a = t + lane_a;
            const uint32_t win = lane_base + box_slot_of(a >> 2) * (kGroups * R::kBoxBytes) + (a & 3) * R::kSegPitch;
            const uint32_t off = PLAIN || L.img_next(F) ? win : zeros;
            const int s2 = t + 2;  // the stream row of lane 0 two steps ahead (lane 31 prepares its prologue pixels)
            const int rr = r0 + T + 2 >= kStepsPerFrame ? r0 + T + 2 - kStepsPerFrame : r0 + T + 2;
            const int ff = r0 + T + 2 >= kStepsPerFrame ? f0 + 1 : f0;
            const bool nimg = PLAIN || (rr < kImageRows && (unsigned)ff < (unsigned)F);
            const uint32_t last31 = nimg ? grp0_base + ring_row_offset<CH>(0, s2) : zeros;
            const uint32_t last = lane == 31 ? last31 : off + 16 * (R::kChunks - 1);
#pragma unroll
            for (int q = 0; q < R::kChunks - 1; ++q) {
                const uint4 v = lds128(off + 16 * q);
                w[4 * q + 0] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
            }
            const uint4 v = lds128(last);
            constexpr int q = R::kChunks - 1;
            w[4 * q + 0] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
        }
        RowChain o1, o3;
        lane_step<CH, T, PLAIN>(L, w, lane, F, o1, o3, emit);
        const int src = (lane + 31) & 31;
        L.in1.s = __shfl_sync(0xffffffffu, o1.s, src);
        L.in1.h0 = __shfl_sync(0xffffffffu, o1.h0, src);
        L.in1.h1 = __shfl_sync(0xffffffffu, o1.h1, src);
        L.in1.h2 = __shfl_sync(0xffffffffu, o1.h2, src);
        L.in1.h3 = __shfl_sync(0xffffffffu, o1.h3, src);
        L.in3.s = __shfl_sync(0xffffffffu, o3.s, src);
        L.in3.h0 = __shfl_sync(0xffffffffu, o3.h0, src);
        L.in3.h1 = __shfl_sync(0xffffffffu, o3.h1, src);
        L.in3.h2 = __shfl_sync(0xffffffffu, o3.h2, src);
        L.in3.h3 = __shfl_sync(0xffffffffu, o3.h3, src);
    };
    auto body = [&](int t, auto ptag) {
        step(t, std::integral_constant<int, 0>{}, ptag);
        step(t + 1, std::integral_constant<int, 1>{}, ptag);
        step(t + 2, std::integral_constant<int, 2>{}, ptag);
        step(t + 3, std::integral_constant<int, 3>{}, ptag);
        if (kBody == 8) {
            step(t + 4, std::integral_constant<int, 4 % kBody>{}, ptag);
            step(t + 5, std::integral_constant<int, 5 % kBody>{}, ptag);
            step(t + 6, std::integral_constant<int, 6 % kBody>{}, ptag);
            step(t + 7, std::integral_constant<int, 7 % kBody>{}, ptag);
        }
    };

    const int t_last = last_step(F);
    int t = kFirstStep;
#pragma unroll 1
    while (t <= t_last) {  // (steps past t_last only see rows that are not live)
        if (use3d && iteration_is_plain(f0, r0, F)) {
            // a run of plain iterations: its own loop, so that its registers are allocated on their own.  The event
            // issued at step t (T == kEventPhase) is the 3-D box at global row
            //   first_row + 512 f0 + plain_event_row(r0) - 28,  r0 = (t - kEventPhase) - 516 f0   ==   plain_y + t
            const int n = plain_run_length(r0);
            plain_y = first_row - (kStepsPerFrame - 512) * f0 + plain_event_row(-kEventPhase) - View3<CH>::kBackRows;
#pragma unroll 1
            for (int k = 0; k < n; ++k, t += kBody) body(t, std::true_type{});
            r0 += kBody * n;
        } else {
            body(t, std::false_type{});
            t += kBody;
            r0 += kBody;
            if (r0 >= kStepsPerFrame) {
                r0 -= kStepsPerFrame;
                ++f0;
            }
        }
    }
    for (int E = waited + 1; E <= issued; ++E) wait(E);  // no copy may be in flight when the CTA retires
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn sys_get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

int systolic_debug_flags(int* flags) {
    int v = 0;
    VPDQ_CUDA(cudaMemcpyFromSymbol(&v, g_systolic_timeout, sizeof v));
    *flags = v;
    return VPDQ_B200_OK;
}

int systolic_debug_force_timeout(int value) {
    VPDQ_CUDA(cudaMemcpyToSymbol(g_systolic_timeout, &value, sizeof value));
    return VPDQ_B200_OK;
}

int systolic_timeout_flag_async(int* h_flag, cudaStream_t stream) {
    VPDQ_CUDA(cudaMemcpyFromSymbolAsync(h_flag, g_systolic_timeout, sizeof(int), 0, cudaMemcpyDeviceToHost, stream));
    return VPDQ_B200_OK;
}

// The 3-D view that turns the eight time-shifted group boxes of an event into ONE box (pdq_systolic_core.h View3):
// element (x, y, g') = byte kBase3 + x + row_bytes * y + kStride3 * g' of the frame buffer.  The views of different
// g' overlap in memory on purpose; only boxes that lie inside the buffer are ever requested.
static int systolic_make_tensor_map3(const uint8_t* d_frames, int64_t n_frames, int channels, CUtensorMap* tmap) {
    EncodeTiledFn encode = sys_get_encode();
    if (!encode) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return VPDQ_B200_ERR_CUDA;
    }
    const int pitch = channels == 3 ? Raw<3>::kSegPitch : Raw<1>::kSegPitch;
    const long long stride3 = channels == 3 ? View3<3>::kStride3 : View3<1>::kStride3;
    const long long base3 = channels == 3 ? View3<3>::kBase3 : View3<1>::kBase3;
    const cuuint64_t gdim[3] = {(cuuint64_t)pitch, (cuuint64_t)n_frames * 512, (cuuint64_t)kGroups};
    const cuuint64_t gstride[2] = {(cuuint64_t)512 * channels, (cuuint64_t)stride3};
    const cuuint32_t box[3] = {(cuuint32_t)pitch, (cuuint32_t)kBoxRows, (cuuint32_t)kGroups};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = encode(tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(d_frames) + base3, gdim, gstride,
                              box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (3-D view) failed with CUresult %d", (int)r);
        return VPDQ_B200_ERR_CUDA;
    }
    return VPDQ_B200_OK;
}

// the batch seen as [n * 512 rows][512 * channels bytes]; box = 4 rows x one lane group's share (+ overlap)
static int systolic_make_tensor_map(const uint8_t* d_frames, int64_t n_frames, int channels, CUtensorMap* tmap) {
    EncodeTiledFn encode = sys_get_encode();
    if (!encode) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return VPDQ_B200_ERR_CUDA;
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)512 * channels, (cuuint64_t)n_frames * 512};
    const cuuint64_t gstride[1] = {(cuuint64_t)512 * channels};
    const cuuint32_t box[2] = {(cuuint32_t)(channels == 3 ? Raw<3>::kSegPitch : Raw<1>::kSegPitch), (cuuint32_t)kBoxRows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(d_frames), gdim, gstride, box,
                              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return VPDQ_B200_ERR_CUDA;
    }
    return VPDQ_B200_OK;
}

// RGB24 (channels = 3) or 8-bit gray (channels = 1) frames -> a64 [n][64][64]: the Jarosz-filtered, decimated luma plane
int systolic_jarosz_launch(const uint8_t* d_frames, int channels, int64_t n_frames, float* d_a64, cudaStream_t stream) {
    if (n_frames > (int64_t)(0x7fffffff / 512) - 8) {
        set_error("pdq: at most %d frames per launch", 0x7fffffff / 512 - 8);
        return VPDQ_B200_ERR_INVALID;
    }
    CUtensorMap tmap, tmap3;
    int rc = systolic_make_tensor_map(d_frames, n_frames, channels, &tmap);
    if (rc) return rc;
    // (VPDQ_B200_SYSTOLIC_3D=0: every event as eight 2-D boxes -- for A/B runs)
    static const int use3d = [] {
        const char* e = getenv("VPDQ_B200_SYSTOLIC_3D");
        return (e && e[0] == '0') ? 0 : 1;
    }();
    rc = use3d ? systolic_make_tensor_map3(d_frames, n_frames, channels, &tmap3) : systolic_make_tensor_map(d_frames, n_frames, channels, &tmap3);
    if (rc) return rc;
    int dev = 0, sms = 148;
    VPDQ_CUDA(cudaGetDevice(&dev));
    VPDQ_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    static std::mutex mu;
    static bool attr_done[64] = {};
    {
        std::lock_guard<std::mutex> lk(mu);
        if (dev < 0 || dev >= 64 || !attr_done[dev]) {
            VPDQ_CUDA(cudaFuncSetAttribute(kx_systolic_jarosz<3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)sizeof(SysSmem<3>)));
            VPDQ_CUDA(cudaFuncSetAttribute(kx_systolic_jarosz<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)sizeof(SysSmem<1>)));
            if (dev >= 0 && dev < 64) attr_done[dev] = true;
        }
    }
    // persistent: one CTA per SM; fewer when the batch has fewer frames than the grid has warps
    const unsigned grid = (unsigned)(n_frames < sms ? n_frames : sms);
    if (channels == 3)
        kx_systolic_jarosz<3><<<grid, kSysThreads, sizeof(SysSmem<3>), stream>>>(tmap, tmap3, use3d, (long long)n_frames, d_a64);
    else
        kx_systolic_jarosz<1><<<grid, kSysThreads, sizeof(SysSmem<1>), stream>>>(tmap, tmap3, use3d, (long long)n_frames, d_a64);
    g_launches += 1;
    VPDQ_CUDA(cudaGetLastError());
    return VPDQ_B200_OK;
}

}  // namespace vpdq
