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

// ---- PDQ (pdq_systolic.cu + pdq_kernels.cu) ---------------------------------------------------------------
size_t pdq_scratch_per_frame();            // 16 KB: the decimated 64 x 64 fp32 plane
size_t pdq_scratch_bytes(int64_t n_frames);
// frames -> hashes + quality (optionally the 64x64 plane and the 16x16 DCT); the scratch supplied bounds how many
// frames are in flight per launch pair
int pdq_launch(const uint8_t* d_frames, int channels, int64_t n_frames, uint8_t* d_hashes, int32_t* d_quality,
               float* d_a64, float* d_b16, void* d_scratch, size_t scratch_bytes, cudaStream_t stream);
// first half: luma + the four Jarosz passes + 64x64 decimation, one warp per frame (pdq_systolic.cu)
int systolic_jarosz_launch(const uint8_t* d_frames, int channels, int64_t n_frames, float* d_a64, cudaStream_t stream);
int systolic_debug_flags(int* flags);
int systolic_debug_force_timeout(int value);
int systolic_timeout_flag_async(int* h_flag, cudaStream_t stream);
// second half: quality + DCT + median + bits (k5_finalize, pdq_kernels.cu)
int pdq_finalize_launch(const float* d_a64, int64_t n_frames, uint8_t* d_hashes, int32_t* d_quality, float* d_a64_dbg,
                        float* d_b16_dbg, cudaStream_t stream);
int pdq_upload_tables();  // DCT matrix -> device (once per device)
// the kernels' "a bounded TMA wait gave up" flag -> h_flags[0] (pinned, >= 4 ints), stream-ordered; nonzero = results
// of the launches before it on this device are invalid.  Every host-pointer entry checks it at its synchronisation
// point.
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
int hamming_scan_multi_launch(const uint64_t* d_db, int64_t n_db, const int64_t* d_offsets, int64_t n_videos,
                              const uint64_t* d_query, const int32_t* d_chunk_rows, int n_chunks, int tol,
                              uint64_t* d_qmask, cudaStream_t stream);
int video_reduce_launch(const uint64_t* d_qmask, int64_t n_videos, const int32_t* d_qv_chunks, const int32_t* d_qv_frames,
                        int n_qvideos, int max_distance, int32_t* d_matched_dense, int32_t* d_rows, int64_t cap,
                        unsigned long long* d_count, cudaStream_t stream);
int hamming_pairs_launch(const uint64_t* d_q, int64_t n_q, const uint64_t* d_t, int64_t n_t, int tol,
                         int skip_diagonal, uint32_t* d_any, uint64_t* d_pairs, int64_t cap,
                         unsigned long long* d_count, cudaStream_t stream);

}  // namespace vpdq
