// pdq_fused.cu -- kx_fused_p123: luma + row pass 1 + column pass 1 + row pass 2 of the PDQ Jarosz filter
// in ONE persistent kernel, fp32 intermediates never leaving the SM (sm_100a).
//
// Replaces k1/k2/k3 of pdq_kernels.cu (same arithmetic, same order, bit-identical results): HBM traffic per
// frame drops from ~5.2 MB (v1: P1 and P2 planes round-trip through L2/HBM) to the algorithmic 786 KB in +
// 128 KB of decimated row-pass-2 output that k4_colpass_finalize consumes.
//
// Structure (details and the index algebra: pdq_fused_core.h, which the CPU emulator also compiles):
//   * persistent grid, one CTA of 16 warps per SM, each CTA owns a contiguous range of frames;
//   * RGB rows are staged by TMA (cp.async.bulk.tensor.2d, one 32-row x 112-byte box per warp per step,
//     2-deep ring with per-warp mbarriers; SASS: UTMALDG) -- out-of-bounds box parts come back as zeros,
//     which is exactly what the two drain steps of every running sum need;
//   * a 32x32 fp32 tile per warp in shared memory (XOR-swizzled, conflict free for lane=row float4 and
//     lane=column scalar access) is the hand-over between the row roles and the column role; two CTA
//     barriers per step separate them (bulk-synchronous wavefront, no other inter-warp signalling).
#include <cuda.h>

#include <mutex>

#include "common.cuh"
#include "pdq_fused_core.h"

namespace vpdq {
using namespace vpdq_core;

constexpr int kFusedThreads = 512;

struct FusedSmem {
    alignas(128) uint8_t raw[2][kBands][kRawBoxBytes];  // 114 688 B  TMA destinations
    alignas(16) float slot[kBands][kTile * kTile];      //  65 536 B  tile hand-over
    alignas(8) unsigned long long bar[2][kBands];       //     256 B  mbarriers
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
    kx_fused_p123(const __grid_constant__ CUtensorMap tmap, const uint8_t* __restrict__ frames,
                  long long n_frames_total, float* __restrict__ p3t) {
    extern __shared__ __align__(128) uint8_t smem_bytes[];
    FusedSmem& sm = *reinterpret_cast<FusedSmem*>(smem_bytes);
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const long long f_begin = n_frames_total * blockIdx.x / gridDim.x;
    const long long f_end = n_frames_total * (blockIdx.x + 1) / gridDim.x;
    const int F = (int)(f_end - f_begin);
    if (F == 0) return;
    const long long total_rows = n_frames_total * 512;

    if (lane == 0) {
        mbar_init(&sm.bar[0][w], 1);
        mbar_init(&sm.bar[1][w], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    const int u_first = (w == 15) ? -16 : 0;
    auto issue = [&](int u) {  // lane 0: stage the raw RGB box of P1 tile u of this warp's band
        const int st = (u - u_first) & 1;
        mbar_expect_tx(&sm.bar[st][w], kRawBoxBytes);
        tma_load_2d(&sm.raw[st][w][0], &tmap, p1_box_x(u & 15), (int)p1_row0(f_begin, floor_div16(u), w),
                    &sm.bar[st][w]);
    };
    if (lane == 0) {
        if (p1_live(u_first, w, F)) issue(u_first);
        if (p1_live(u_first + 1, w, F)) issue(u_first + 1);
    }

    Chain c1, c2, c3;
    c1.init();
    c2.init();
    c3.init();
    float p0 = 0.0f, p1 = 0.0f;

    const int steps = num_steps(F);
    for (int T = 0; T < steps; ++T) {
        // ---------------- phase A: row roles (lane = row) ----------------
        {
            const int u = sched_u3(T, w);
            if (p3_live(u, F)) {
                float* out = p3t + (size_t)(f_begin + (u >> 4)) * (64 * 512) + 32 * w + lane;
                p3_lane(c3, sm.slot[w], lane, u & 15, out);
            }
        }
        {
            const int u = sched_u12(T, w);
            if (p1_live(u, w, F)) {
                const int strip = u & 15;
                const int k = u - u_first;
                uint32_t first2[2] = {0u, 0u};
                if (strip == 0) {
                    const long long R = p1_row0(f_begin, floor_div16(u), w) + lane;
                    if (R >= 0 && R < total_rows) {
                        const uint2 v = __ldg(reinterpret_cast<const uint2*>(frames + (size_t)R * 1536));
                        first2[0] = v.x;
                        first2[1] = v.y;
                    }
                }
                mbar_wait(&sm.bar[k & 1][w], (uint32_t)((k >> 1) & 1));
                const uint4* rr = reinterpret_cast<const uint4*>(&sm.raw[k & 1][w][lane * kRawPitch]);
                uint32_t raw[kRawWords];
#pragma unroll
                for (int q = 0; q < kRawWords / 4; ++q) {
                    const uint4 v = rr[q];
                    raw[4 * q + 0] = v.x; raw[4 * q + 1] = v.y; raw[4 * q + 2] = v.z; raw[4 * q + 3] = v.w;
                }
                p1_lane(c1, raw, first2, sm.slot[w], lane, strip);
                __syncwarp();  // every lane has read its staged row: the stage may be refilled
                if (lane == 0 && p1_live(u + 2, w, F)) issue(u + 2);
            }
        }
        __syncthreads();
        // ---------------- phase B: column role (lane = column), in place ----------------
        {
            const int u = sched_u12(T, w);
            if (p2_live(u, F)) {
                const int band = u & 15;
                p2_lane(c2, p0, p1, sm.slot[band], lane, band, u < 0);
            }
        }
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

size_t fused_scratch_per_frame() { return (size_t)64 * 512 * sizeof(float); }

// RGB24 frames -> p3t [n][64][512] (row pass 2 at the 64 decimated columns, transposed)
int fused_p123_launch(const uint8_t* d_frames, int64_t n_frames, float* d_p3t, cudaStream_t stream) {
    EncodeTiledFn encode = get_encode();
    if (!encode) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return VPDQ_B200_ERR_CUDA;
    }
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {1536, (cuuint64_t)n_frames * 512};
    const cuuint64_t gstride[1] = {1536};
    const cuuint32_t box[2] = {(cuuint32_t)kRawPitch, (cuuint32_t)kTile};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(d_frames), gdim, gstride, box,
                              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return VPDQ_B200_ERR_CUDA;
    }
    int dev = 0, sms = 148;
    VPDQ_CUDA(cudaGetDevice(&dev));
    VPDQ_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    static std::mutex mu;
    static bool attr_done[64] = {};
    {
        std::lock_guard<std::mutex> lk(mu);
        if (dev < 0 || dev >= 64 || !attr_done[dev]) {
            VPDQ_CUDA(cudaFuncSetAttribute(kx_fused_p123, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)sizeof(FusedSmem)));
            if (dev >= 0 && dev < 64) attr_done[dev] = true;
        }
    }
    const unsigned grid = (unsigned)(n_frames < sms ? n_frames : sms);  // persistent: one CTA per SM
    kx_fused_p123<<<grid, kFusedThreads, sizeof(FusedSmem), stream>>>(tmap, d_frames, (long long)n_frames, d_p3t);
    g_launches += 1;
    VPDQ_CUDA(cudaGetLastError());
    return VPDQ_B200_OK;
}

}  // namespace vpdq
