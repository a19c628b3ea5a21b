// tests/legacy/legacy.cuh -- TEST-ONLY declarations shared by the round-1 PDQ pipelines (see pdq_lines.cu).
#pragma once
#include <cuda.h>

#include "../../hydrus_video_deduplicator_b200/csrc/common.cuh"

namespace vpdq {
const char* legacy_error();
// one-frame fused kernel (pdq_fused.cu): RGB24 frames -> a64 [n][64][64]
int fused_jarosz_launch(const uint8_t* d_frames, int64_t n_frames, float* d_a64, cudaStream_t stream);
int fused_debug_flags(int* flags);
int fused_make_tensor_map(const uint8_t* d_frames, int64_t n_frames, int channels, CUtensorMap* tmap);
// frame-pair fused kernel (pdq_fused2.cu)
int fused2_jarosz_launch(const uint8_t* d_frames, int channels, int64_t n_frames, float* d_a64, cudaStream_t stream);
int fused2_debug_flags(int* flags);
}  // namespace vpdq
