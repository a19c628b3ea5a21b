// pdq_kernels.cu -- the second half of the PDQ frame hash on sm_100a and the host-side launch of the whole pipeline.
//
// Replaces the native work behind hvdaccelerators.vpdq.VideoHasher.hash_frame
// (reference call site: src/hydrusvideodeduplicator/vpdqpy/vpdqpy.py:118); algorithm and op order are
// SURVEY.md Appendix A (Meta ThreatExchange PDQ).  Every fp32 operation that the CPU path performs is
// performed here in the same order with the same single rounding; nothing is re-associated.
//
//   kx_systolic_jarosz (pdq_systolic.cu)   RGB24 / gray frames -> the Jarosz-filtered, decimated 64x64 luma plane
//   k5_finalize        (here)               64x64 plane -> quality, 64->16 DCT, median, 256 bits; one CTA per frame
//
// (The round-1 pipelines -- the v1 line kernels and the two tiled fused kernels -- live in tests/legacy/ as a
// test-only library that the parity tests cross-check against; they are not part of libvpdq_b200.so.)
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "common.cuh"
#include "pdq_finalize.cuh"

namespace vpdq {

__device__ float g_dct[16 * 64];  // the 16 x 64 DCT table (coalesced fill of k5's shared copies)

// ---------------------------------------------------------------------------------------------------
// K5: the finalize step (pdq_finalize.cuh): persistent CTAs of 256 threads walk over frames.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k5_finalize(const float* __restrict__ a64, long long n_frames,
                                                   uint8_t* __restrict__ hashes, int32_t* __restrict__ quality,
                                                   float* __restrict__ a64_dbg, float* __restrict__ b16_dbg) {
    __shared__ FinalizeSmem sm;
    const int t = threadIdx.x;
    finalize_load_tables(sm, g_dct, t);
#pragma unroll 1
    for (long long f = blockIdx.x; f < n_frames; f += gridDim.x) {
        finalize_frame(sm, a64 + (size_t)f * (kDec * kDec), hashes + (size_t)f * 32, quality + f,
                       a64_dbg ? a64_dbg + (size_t)f * 4096 : nullptr, b16_dbg ? b16_dbg + (size_t)f * 256 : nullptr, t,
                       [] { __syncthreads(); });
        __syncthreads();  // the shared buffers are reused by the next frame
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
static float h_dct[16 * 64];
static std::once_flag h_dct_once;

const float* pdq_host_dct() {
    std::call_once(h_dct_once, [] {
        // Appendix A step 6: scale rounded to fp32 first, product in double, stored as fp32
        const float scale = (float)sqrt(2.0 / 64.0);
        for (int i = 0; i < 16; i++)
            for (int j = 0; j < 64; j++)
                h_dct[i * 64 + j] = (float)(scale * cos((M_PI / 2 / 64.0) * (i + 1) * (2 * j + 1)));
    });
    return h_dct;
}

int pdq_upload_tables() {
    static std::mutex mu;
    static bool done[64] = {};
    int dev = 0;
    VPDQ_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    if (dev >= 0 && dev < 64 && done[dev]) return VPDQ_B200_OK;
    VPDQ_CUDA(cudaMemcpyToSymbol(g_dct, pdq_host_dct(), sizeof(h_dct)));
    if (dev >= 0 && dev < 64) done[dev] = true;
    return VPDQ_B200_OK;
}

// the only intermediate that touches HBM: the decimated plane, 64 x 64 fp32 per frame
constexpr size_t kScratchPerFrame = (size_t)kDec * kDec * sizeof(float);

size_t pdq_scratch_per_frame() { return kScratchPerFrame; }
size_t pdq_scratch_bytes(int64_t n_frames) { return (size_t)(n_frames < 1 ? 1 : n_frames) * kScratchPerFrame; }

int pdq_timeout_flags_async(int* h_flags, cudaStream_t stream) { return systolic_timeout_flag_async(h_flags, stream); }
int pdq_force_timeout_flags(int value) { return systolic_debug_force_timeout(value); }

int pdq_finalize_launch(const float* d_a64, int64_t n_frames, uint8_t* d_hashes, int32_t* d_quality, float* d_a64_dbg,
                        float* d_b16_dbg, cudaStream_t stream) {
    if (n_frames == 0) return VPDQ_B200_OK;
    const int rc = pdq_upload_tables();
    if (rc) return rc;
    int dev = 0, sms = 148;
    VPDQ_CUDA(cudaGetDevice(&dev));
    VPDQ_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int64_t resident = (int64_t)sms * 7;  // 7 CTAs of 256 threads / 33 KB of shared memory fit an SM
    const unsigned grid = (unsigned)(n_frames < resident ? n_frames : resident);
    k5_finalize<<<grid, 256, 0, stream>>>(d_a64, (long long)n_frames, d_hashes, d_quality, d_a64_dbg, d_b16_dbg);
    g_launches += 1;
    VPDQ_CUDA(cudaGetLastError());
    return VPDQ_B200_OK;
}

// frames -> hashes + quality; the scratch bounds how many frames are in flight per launch pair
int pdq_launch(const uint8_t* d_frames, int channels, int64_t n_frames, uint8_t* d_hashes, int32_t* d_quality,
               float* d_a64, float* d_b16, void* d_scratch, size_t scratch_bytes, cudaStream_t stream) {
    if (n_frames == 0) return VPDQ_B200_OK;
    int64_t chunk = (int64_t)(scratch_bytes / kScratchPerFrame);
    if (chunk < 1) {
        set_error("pdq: scratch too small (%zu bytes, need >= %zu)", scratch_bytes, kScratchPerFrame);
        return VPDQ_B200_ERR_INVALID;
    }
    if (chunk > (1 << 21)) chunk = 1 << 21;  // TMA coordinates are 32-bit row indices
    const size_t frame_bytes = (size_t)kPlane * channels;
    for (int64_t f0 = 0; f0 < n_frames; f0 += chunk) {
        const int64_t nf = (n_frames - f0 < chunk) ? (n_frames - f0) : chunk;
        float* a64 = static_cast<float*>(d_scratch);
        int rc = systolic_jarosz_launch(d_frames + (size_t)f0 * frame_bytes, channels, nf, a64, stream);
        if (rc) return rc;
        rc = pdq_finalize_launch(a64, nf, d_hashes + (size_t)f0 * 32, d_quality + f0,
                                 d_a64 ? d_a64 + (size_t)f0 * 4096 : nullptr, d_b16 ? d_b16 + (size_t)f0 * 256 : nullptr,
                                 stream);
        if (rc) return rc;
    }
    return VPDQ_B200_OK;
}

}  // namespace vpdq
