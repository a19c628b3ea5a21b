// pdq_fused.cu -- kx_fused_jarosz: luma + the four Jarosz box-filter passes + 64x64 decimation of the PDQ
// hash in ONE persistent kernel, fp32 intermediates never leaving the SM (sm_100a).
//
// Replaces k1/k2/k3 and the column-pass half of k4 of pdq_kernels.cu (same arithmetic, same order,
// bit-identical results).  HBM traffic per frame drops from ~5.2 MB (v1: the P1/P2 planes round-trip
// through L2/HBM) to the algorithmic 786 KB in + 16 KB out (the decimated plane k5_finalize consumes).
//
// Structure (details and the index algebra: pdq_fused_core.h, which the CPU emulator also compiles):
//   * persistent grid, one CTA of 16 warps per SM, each CTA owns a contiguous range of frames;
//   * RGB rows are staged by TMA (cp.async.bulk.tensor.2d, one 32-row x 112-byte box per warp per step,
//     per-warp mbarrier; SASS: UTMALDG).  The staged row is pulled into registers at the top of the step
//     and the next box is requested at once, so one stage per warp suffices.  Out-of-bounds box parts
//     come back as zeros, which is exactly what the two drain steps of every running sum need;
//   * 32x32 fp32 tiles in shared memory (XOR-swizzled: conflict free for lane=row float4 and lane=column
//     scalar access), double buffered by step parity, are the hand-over between row and column roles;
//   * every warp runs its four roles (P1..P4) in the same step, interleaved, and ONE CTA barrier per step
//     keeps the wavefront (bulk-synchronous; no other inter-warp signalling).
#include <cuda.h>

#include <mutex>

#include "legacy.cuh"
#include "pdq_fused_core.h"

namespace vpdq {
using namespace vpdq_core;

constexpr int kFusedThreads = 512;

struct FusedSmem {
    alignas(128) uint8_t raw[kBands][kRawBoxBytes];   //  57 344 B  TMA destinations (one stage per warp)
    alignas(16) float slot[2][kBands][kTile * kTile];  // 131 072 B  tile hand-over, by step parity
    alignas(16) float t3[2][kBands * kT3Strip];        //  18 432 B  P3 -> P4 hand-over, by step parity
    alignas(8) unsigned long long bar[kBands];         //     128 B  mbarriers
};

__device__ int g_fused_timeout = 0;  // set if an mbarrier wait gave up (never expected)

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
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    // bounded: a mis-programmed copy must not hang the GPU; ~1 s worth of polling, then flag and go on
#pragma unroll 1
    for (int spin = 0; spin < (1 << 24); ++spin)
        if (mbar_try_wait(bar, parity)) return;
    g_fused_timeout = 1;
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
        : "memory");
}

__global__ void __launch_bounds__(kFusedThreads, 1)
    kx_fused_jarosz(const __grid_constant__ CUtensorMap tmap, const uint8_t* __restrict__ frames,
                    long long n_frames_total, float* __restrict__ a64) {
    extern __shared__ __align__(128) uint8_t smem_bytes[];
    FusedSmem& sm = *reinterpret_cast<FusedSmem*>(smem_bytes);
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const long long f_begin = n_frames_total * blockIdx.x / gridDim.x;
    const long long f_end = n_frames_total * (blockIdx.x + 1) / gridDim.x;
    const int F = (int)(f_end - f_begin);
    if (F == 0) return;
    const long long total_rows = n_frames_total * 512;

    if (lane == 0) mbar_init(&sm.bar[w], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    const int u_first = p1_first(w);
    auto issue = [&](int u) {  // lane 0: stage the raw RGB box of P1 tile u of this warp's band
        mbar_expect_tx(&sm.bar[w], kRawBoxBytes);
        tma_load_2d(&sm.raw[w][0], &tmap, p1_box_x(u & 15), (int)p1_row0(f_begin, floor_div16(u), w), &sm.bar[w]);
    };
    if (lane == 0 && p1_live(u_first, w, F)) issue(u_first);

    LaneState st;
    st.init();

    const int steps = num_steps(F);
    for (int T = 0; T < steps; ++T) {
        const int u1 = sched_u(T, 1, w), u2 = sched_u(T, 2, w), u3 = sched_u(T, 3, w), u4 = sched_u(T, 4, w);
        StepArgs a;
        a.live1 = p1_live(u1, w, F);
        a.live2 = p2_live(u2, F);
        a.live3 = p34_live(u3, F);
        a.live4 = p34_live(u4, F);
        a.s1 = u1 & 15; a.b2 = u2 & 15; a.s3 = u3 & 15; a.b4 = u4 & 15;
        a.tile_a = sm.slot[T & 1][w];
        a.tile_b = sm.slot[(T - 1) & 1][a.b2];
        a.t3_w = sm.t3[T & 1] + a.s3 * kT3Strip;
        a.t3_r = sm.t3[(T - 1) & 1] + w * kT3Strip;
        a.a_out = a64 + (size_t)(a.live4 ? (f_begin + (u4 >> 4)) : f_begin) * 4096 + 4 * w + (lane & 3);

        uint32_t first2[2] = {0u, 0u};
        if (a.live1 && a.s1 == 0) {
            const long long R = p1_row0(f_begin, floor_div16(u1), w) + lane;
            if (R >= 0 && R < total_rows) {
                const uint2 v = __ldg(reinterpret_cast<const uint2*>(frames + (size_t)R * 1536));
                first2[0] = v.x;
                first2[1] = v.y;
            }
        }
        if (a.live1) mbar_wait(&sm.bar[w], (uint32_t)((u1 - u_first) & 1));
        // (when P1 is not live the staged bytes are stale; its results are never stored)
        uint32_t raw[kRawWords];
        {
            const uint4* rr = reinterpret_cast<const uint4*>(&sm.raw[w][lane * kRawPitch]);
#pragma unroll
            for (int q = 0; q < kRawWords / 4; ++q) {
                const uint4 v = rr[q];
                raw[4 * q + 0] = v.x; raw[4 * q + 1] = v.y; raw[4 * q + 2] = v.z; raw[4 * q + 3] = v.w;
            }
        }
        // The stage is refilled by the next TMA.  LDS results arrive asynchronously: issuing them is not enough,
        // every lane must HOLD its row before the async proxy may overwrite it (observed otherwise: ~1e-3 of frames
        // corrupted under load).  The vote consumes one word of each LDS.128 -> the scoreboard wait happens there,
        // and the ballot doubles as the warp-wide rendezvous.  (The magic value never matches on all lanes.)
        auto refill = [&]() {
            const uint32_t dep = (raw[0] ^ raw[4] ^ raw[8]) ^ (raw[12] ^ raw[16] ^ raw[20]) ^ raw[24];
            const unsigned held = __ballot_sync(0xffffffffu, dep != 0x5bd1e995u);
            if (lane == 0 && a.live1 && held != 0u && p1_live(u1 + 1, w, F)) issue(u1 + 1);
        };
        fused_step(st, a, raw, first2, lane, refill);
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
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

// bit 0: an mbarrier wait inside kx_fused_jarosz gave up (a TMA copy never completed) -- never expected
int fused_debug_flags(int* flags) {
    int v = 0;
    VPDQ_CUDA(cudaMemcpyFromSymbol(&v, g_fused_timeout, sizeof v));
    *flags = v;
    return VPDQ_B200_OK;
}

// tensor map over the whole batch seen as [n*512 rows][512*channels B], box = 32 rows x 112 B (RGB24) or 48 B (gray)
// (shared with pdq_fused2.cu)
int fused_make_tensor_map(const uint8_t* d_frames, int64_t n_frames, int channels, CUtensorMap* tmap) {
    EncodeTiledFn encode = get_encode();
    if (!encode) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return VPDQ_B200_ERR_CUDA;
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)512 * channels, (cuuint64_t)n_frames * 512};
    const cuuint64_t gstride[1] = {(cuuint64_t)512 * channels};
    const cuuint32_t box[2] = {(cuuint32_t)(channels == 3 ? kRawPitch : 48), (cuuint32_t)kTile};
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

// RGB24 frames -> a64 [n][64][64]: the Jarosz-filtered, decimated luma plane
int fused_jarosz_launch(const uint8_t* d_frames, int64_t n_frames, float* d_a64, cudaStream_t stream) {
    CUtensorMap tmap;
    const int rc_map = fused_make_tensor_map(d_frames, n_frames, 3, &tmap);
    if (rc_map) return rc_map;
    int dev = 0, sms = 148;
    VPDQ_CUDA(cudaGetDevice(&dev));
    VPDQ_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    static std::mutex mu;
    static bool attr_done[64] = {};
    {
        std::lock_guard<std::mutex> lk(mu);
        if (dev < 0 || dev >= 64 || !attr_done[dev]) {
            VPDQ_CUDA(cudaFuncSetAttribute(kx_fused_jarosz, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)sizeof(FusedSmem)));
            if (dev >= 0 && dev < 64) attr_done[dev] = true;
        }
    }
    const unsigned grid = (unsigned)(n_frames < sms ? n_frames : sms);  // persistent: one CTA per SM
    kx_fused_jarosz<<<grid, kFusedThreads, sizeof(FusedSmem), stream>>>(tmap, d_frames, (long long)n_frames, d_a64);
    g_launches += 1;
    VPDQ_CUDA(cudaGetLastError());
    return VPDQ_B200_OK;
}

}  // namespace vpdq
