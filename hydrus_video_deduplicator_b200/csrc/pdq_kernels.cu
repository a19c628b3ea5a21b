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
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "common.cuh"

namespace vpdq {

__device__ float g_dct[16 * 64];  // the 16 x 64 DCT table (coalesced fill of k5's shared copies)

// ---------------------------------------------------------------------------------------------------
// K5: quality + 64->16 DCT + median + bits from the decimated plane a64 [n][64][64] that kx_systolic_jarosz emits.
// One CTA (256 threads) per frame; same arithmetic and order as the straightforward k4_colpass_finalize<true> kept in
// tests/legacy/pdq_lines.cu (cross-checked by the parity tests), about half the instructions:
//   * T = D*A with two accumulator pairs per thread in packed FMUL2 / FADD2 (4 D values per 128-bit load
//     from a transposed table), B = T*D^T with 128-bit loads along k;
//   * quality: (u - down, u - right) as one packed pair, exact /255, and |trunc(x)| taken as the mantissa of
//     RZ(|x| + 2^23) -- the integer bit patterns are summed as they are (256 threads x 32 terms: the 2^23
//     exponent offsets cancel mod 2^32);
//   * median: #{B < v} only: the 128-th smallest value is the largest v with #{B < v} <= 127.
// ---------------------------------------------------------------------------------------------------
constexpr int kDJ = 68;  // pitch of the D / T rows: 128-bit k-chunks of 8 consecutive rows fall in different banks

__device__ __forceinline__ float2 fadd2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) { return __fmul2_rn(a, b); }

__global__ void __launch_bounds__(256) k5_finalize(const float* __restrict__ a64, long long n_frames,
                                                   uint8_t* __restrict__ hashes, int32_t* __restrict__ quality,
                                                   float* __restrict__ a64_dbg, float* __restrict__ b16_dbg) {
    __shared__ __align__(16) float A[kDec][kDec];  // decimated 64x64 plane
    __shared__ __align__(16) float Dt[kDec][16];   // Dt[k][i] = D[i][k]
    __shared__ __align__(16) float Dj[16][kDJ];    // D[j][k]
    __shared__ __align__(16) float T[16][kDJ];     // D * A
    __shared__ __align__(16) float B[256];
    __shared__ unsigned g_sum;
    __shared__ int med_key;

    const int t = threadIdx.x;

    // the two copies of the DCT table are filled once per CTA; a CTA then walks over frames (persistent grid)
#pragma unroll
    for (int e = t; e < 16 * 64; e += 256) {
        Dj[e >> 6][e & 63] = __ldg(g_dct + e);
        Dt[e >> 4][e & 15] = __ldg(g_dct + (e & 15) * 64 + (e >> 4));  // consecutive lanes -> consecutive words of Dt
    }
#pragma unroll 1
    for (long long f = blockIdx.x; f < n_frames; f += gridDim.x) {
    if (t == 0) {
        g_sum = 0u;
        med_key = INT_MIN;
    }
    {
        const float4* src = reinterpret_cast<const float4*>(a64 + (size_t)f * (kDec * kDec));
        float4* dst = reinterpret_cast<float4*>(&A[0][0]);
#pragma unroll
        for (int e = 0; e < 4; ++e) dst[t + 256 * e] = __ldg(src + t + 256 * e);
    }
    __syncthreads();

    if (a64_dbg)
        for (int e = t; e < kDec * kDec; e += 256) a64_dbg[(size_t)f * 4096 + e] = A[e >> 6][e & 63];

    // quality: sum of |trunc((u - v) * 100 / 255)| over vertical and horizontal neighbours; thread = column j,
    // 16 consecutive rows (an absent neighbour is replaced by u itself: difference 0, term 0)
    {
        const int j = t & 63, i0 = (t >> 6) * 16;
        const int jr = j < 63 ? j + 1 : j;
        const float c255 = 0.00392156886f;  // RN(1/255), see div255()
        unsigned g = 0u;
        float u = A[i0][j];
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const int i = i0 + r;
            const float down = A[i < 63 ? i + 1 : i][j];
            const float right = A[i][jr];
            const float2 x = fmul2(fadd2(make_float2(u, u), make_float2(-down, -right)), make_float2(100.0f, 100.0f));
            const float2 q0 = fmul2(x, make_float2(c255, c255));
            const float2 rr = __ffma2_rn(make_float2(-255.0f, -255.0f), q0, x);
            const float2 q = __ffma2_rn(rr, make_float2(c255, c255), q0);  // = x / 255.0f, correctly rounded
            g += __float_as_uint(__fadd_rz(fabsf(q.x), 8388608.0f)) + __float_as_uint(__fadd_rz(fabsf(q.y), 8388608.0f));
            u = down;
        }
        g = __reduce_add_sync(0xffffffffu, g);
        if ((t & 31) == 0) atomicAdd(&g_sum, g);
    }

    // T = D * A : thread -> column j, four rows i0..i0+3; sequential in k, separate multiply and add
    {
        const int j = t & 63, i0 = (t >> 6) * 4;
        // (products packed, sums scalar: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under
        //  --fmad=false, which would change the bits; it leaves this form alone -- checked in the SASS and by the
        //  parity tests)
        float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 16
        for (int k = 0; k < 64; ++k) {
            const float a = A[k][j];
            const float4 d = *reinterpret_cast<const float4*>(&Dt[k][i0]);
            const float2 p01 = fmul2(make_float2(d.x, d.y), make_float2(a, a));
            const float2 p23 = fmul2(make_float2(d.z, d.w), make_float2(a, a));
            acc[0] = fadd(acc[0], p01.x);
            acc[1] = fadd(acc[1], p01.y);
            acc[2] = fadd(acc[2], p23.x);
            acc[3] = fadd(acc[3], p23.y);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) T[i0 + u][j] = acc[u];
    }
    __syncthreads();

    // B = T * D^T : thread -> (i, j)
    float bv;
    {
        const int i = t >> 4, j = t & 15;
        float acc = 0.0f;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const float4 tv = *reinterpret_cast<const float4*>(&T[i][4 * q]);
            const float4 dv = *reinterpret_cast<const float4*>(&Dj[j][4 * q]);
            acc = fadd(acc, fmul(tv.x, dv.x));
            acc = fadd(acc, fmul(tv.y, dv.y));
            acc = fadd(acc, fmul(tv.z, dv.z));
            acc = fadd(acc, fmul(tv.w, dv.w));
        }
        bv = acc;
        B[t] = acc;
    }
    __syncthreads();
    if (b16_dbg) b16_dbg[(size_t)f * 256 + t] = bv;

    // median = 128-th smallest of the 256 values (what Torben's method returns for n = 256) = the largest value v
    // with #{B < v} <= 127.  Keys: an order-preserving map float -> int (of v + 0.0f, so that -0 == +0).
    int key = __float_as_int(fadd(bv, 0.0f));
    key ^= (key >> 31) & 0x7fffffff;
    {
        int lt = 0;
        const float4* b4 = reinterpret_cast<const float4*>(B);
#pragma unroll 16
        for (int u = 0; u < 64; ++u) {
            const float4 b = b4[u];
            lt += (b.x < bv) + (b.y < bv) + (b.z < bv) + (b.w < bv);
        }
        const int cand = __reduce_max_sync(0xffffffffu, lt <= 127 ? key : INT_MIN);
        if ((t & 31) == 0) atomicMax(&med_key, cand);
    }
    __syncthreads();

    // bit k = 16 i + j = t  ->  byte t >> 3, bit t & 7: eight little-endian 32-bit ballots
    const unsigned word = __ballot_sync(0xffffffffu, key > med_key);
    if ((t & 31) == 0) reinterpret_cast<uint32_t*>(hashes)[f * 8 + (t >> 5)] = word;
    if (t == 0) {
        const int q = (int)(g_sum / 90u);
        quality[f] = q > 100 ? 100 : q;
    }
    __syncthreads();  // A, T, B, g_sum, med_key are reused by the next frame
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
