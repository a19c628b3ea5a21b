// common.cuh -- shared declarations for libvpdq_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <stddef.h>
#include <stdint.h>

#include "../../include/vpdq_b200.h"

namespace vpdq {

constexpr int kDim = VPDQ_B200_FRAME_DIM;   // 512
constexpr int kPlane = kDim * kDim;         // 262144 pixels
constexpr int kDec = 64;                    // decimated side

// ---- exactly-rounded fp32 primitives: the compiler may never contract these into FMAs -------
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

// every kernel launch of this library bumps this (bench.py reports it as gpu_launches)
extern std::atomic<uint64_t> g_launches;

// Error plumbing shared by the host-side translation units.
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
#define VPDQ_CUDA(call)                                                   \
    do {                                                                  \
        cudaError_t e__ = (call);                                         \
        if (e__ != cudaSuccess) return ::vpdq::cuda_fail(e__, #call);      \
    } while (0)

// ---- PDQ (pdq_kernels.cu) ---------------------------------------------------------------------
size_t pdq_scratch_bytes(int64_t n_frames);
// chunk = number of frames whose intermediates are live at once (bounded by the scratch supplied)
int pdq_launch(const uint8_t* d_frames, int channels, int64_t n_frames, uint8_t* d_hashes, int32_t* d_quality,
               float* d_a64, float* d_b16, void* d_scratch, size_t scratch_bytes, cudaStream_t stream);
// fused luma + Jarosz passes + decimation (pdq_fused.cu): RGB24 frames -> a64 [n][64][64]
size_t fused_scratch_per_frame();
int fused_jarosz_launch(const uint8_t* d_frames, int64_t n_frames, float* d_a64, cudaStream_t stream);
int fused_debug_flags(int* flags);
int fused_timeout_flag_async(int* h_flag, cudaStream_t stream);
int fused_force_timeout(int value);
// the same, two frames per lane in packed fp32 pairs (pdq_fused2.cu) -- the default
int fused2_jarosz_launch(const uint8_t* d_frames, int channels, int64_t n_frames, float* d_a64, cudaStream_t stream);
int fused2_debug_flags(int* flags);
int fused2_timeout_flag_async(int* h_flag, cudaStream_t stream);
int fused2_force_timeout(int value);
// the warp-per-frame systolic kernel (pdq_systolic.cu): no shared-memory transposition between the passes
int systolic_jarosz_launch(const uint8_t* d_frames, int channels, int64_t n_frames, float* d_a64, cudaStream_t stream);
int systolic_debug_flags(int* flags);
int systolic_debug_force_timeout(int value);
int systolic_timeout_flag_async(int* h_flag, cudaStream_t stream);
// which Jarosz pipeline hashes RGB24 frames: 3 = systolic, 2 = frame-pair fused kernel (default), 1 = fused (VPDQ_B200_PDQ_IMPL=fused),
// 0 = v1 line kernels (=lines).  All are CUDA and bit-identical; the switch exists for A/B measurements.
int pdq_impl();
int pdq_set_impl(int impl);
int pdq_upload_tables();  // DCT matrix -> device (once per device)
// the kernels' "a bounded TMA wait gave up" flags -> h_flags[0..2] (pinned), stream-ordered; nonzero = results of the
// launches before it on this device are invalid.  Every host-pointer entry checks them at its synchronisation point.
int pdq_timeout_flags_async(int* h_flags, cudaStream_t stream);
int pdq_force_timeout_flags(int value);  // test hook
const float* pdq_host_dct();

// ---- POINT resize (resize_kernels.cu) ----------------------------------------------------------------
int point_resize_launch(const uint8_t* d_src, int64_t n_frames, int src_h, int src_w, uint8_t* d_dst,
                        cudaStream_t stream);

// ---- Hamming (hamming_kernels.cu) ---------------------------------------------------------------
int hamming_scan_launch(const uint64_t* d_db, int64_t n_db, const int64_t* d_offsets, int64_t n_videos,
                        const uint64_t* d_query, int n_query, int tol, uint64_t* d_qmask, int32_t* d_tcount,
                        cudaStream_t stream);
int hamming_pairs_launch(const uint64_t* d_q, int64_t n_q, const uint64_t* d_t, int64_t n_t, int tol,
                         int skip_diagonal, uint32_t* d_any, uint64_t* d_pairs, int64_t cap,
                         unsigned long long* d_count, cudaStream_t stream);

}  // namespace vpdq
