// resize_kernels.cu -- SURVEY.md 8f-2: the reference's `frame.reformat(512, 512, "rgb24", POINT)`
// (vpdqpy/vpdqpy.py:90-95) on the device, so that natively-sized decoded frames can stay in HBM.
//
// swscale's POINT scaler is a centre-based nearest neighbour in 16.16 fixed point:
//     xInc = ((src << 16) + (dst >> 1)) / dst ;   src_index(i) = ((xInc >> 1) + i * xInc) >> 16
// (this is the rule that reproduces the reference's golden hashes bit-exactly).  Pure byte
// gather: HBM-bound, algorithmic bytes = 786 432 written + the sampled source bytes per frame.
#include "common.cuh"

namespace vpdq {

// one thread = 4 output pixels (12 bytes = three 32-bit stores); a warp writes 384 contiguous bytes
__global__ void __launch_bounds__(256)
    k_point_resize_rgb(const uint8_t* __restrict__ src, int64_t n_frames, int src_h, int src_w, uint32_t x_inc,
                       uint32_t y_inc, uint8_t* __restrict__ dst) {
    const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;  // over n * 512 * 128 pixel quads
    if (t >= n_frames * (kDim * (kDim / 4))) return;
    const int q = (int)(t & 127);
    const int y = (int)((t >> 7) & 511);
    const int64_t f = t >> 16;
    int sy = (int)(((uint64_t)(y_inc >> 1) + (uint64_t)y * y_inc) >> 16);
    if (sy > src_h - 1) sy = src_h - 1;
    const uint8_t* row = src + ((size_t)f * src_h + sy) * (size_t)src_w * 3;
    uint8_t px[12];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int x = 4 * q + j;
        int sx = (int)(((uint64_t)(x_inc >> 1) + (uint64_t)x * x_inc) >> 16);
        if (sx > src_w - 1) sx = src_w - 1;
        const uint8_t* p = row + (size_t)sx * 3;
        px[3 * j + 0] = __ldg(p);
        px[3 * j + 1] = __ldg(p + 1);
        px[3 * j + 2] = __ldg(p + 2);
    }
    uint32_t* out = reinterpret_cast<uint32_t*>(dst + ((size_t)f * kPlane + (size_t)y * kDim + 4 * q) * 3);
#pragma unroll
    for (int wd = 0; wd < 3; ++wd)
        out[wd] = (uint32_t)px[4 * wd] | ((uint32_t)px[4 * wd + 1] << 8) | ((uint32_t)px[4 * wd + 2] << 16) |
                  ((uint32_t)px[4 * wd + 3] << 24);
}

int point_resize_launch(const uint8_t* d_src, int64_t n_frames, int src_h, int src_w, uint8_t* d_dst,
                        cudaStream_t stream) {
    if (n_frames == 0) return VPDQ_B200_OK;
    const uint32_t x_inc = (uint32_t)((((uint64_t)src_w << 16) + (kDim >> 1)) / kDim);
    const uint32_t y_inc = (uint32_t)((((uint64_t)src_h << 16) + (kDim >> 1)) / kDim);
    const int64_t quads = n_frames * (int64_t)(kDim * (kDim / 4));
    const int64_t blocks = (quads + 255) / 256;
    if (blocks > 0x7fffffffLL) {
        set_error("point_resize: too many frames in one call");
        return VPDQ_B200_ERR_INVALID;
    }
    k_point_resize_rgb<<<(unsigned)blocks, 256, 0, stream>>>(d_src, n_frames, src_h, src_w, x_inc, y_inc, d_dst);
    g_launches += 1;
    VPDQ_CUDA(cudaGetLastError());
    return VPDQ_B200_OK;
}

}  // namespace vpdq
