// tests/legacy/legacy_support.cu -- TEST-ONLY: the few host helpers the round-1 sources expect from the library.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>

#include <mutex>

#include "legacy.cuh"

namespace vpdq {
static thread_local char t_err[512] = "";
std::atomic<uint64_t> g_launches{0};
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof t_err, fmt, ap);
    va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return VPDQ_B200_ERR_CUDA;
}
const char* legacy_error() { return t_err; }

static float h_dct[16 * 64];
static std::once_flag h_dct_once;
const float* pdq_host_dct() {
    std::call_once(h_dct_once, [] {
        const float scale = (float)sqrt(2.0 / 64.0);
        for (int i = 0; i < 16; i++)
            for (int j = 0; j < 64; j++)
                h_dct[i * 64 + j] = (float)(scale * cos((M_PI / 2 / 64.0) * (i + 1) * (2 * j + 1)));
    });
    return h_dct;
}
}  // namespace vpdq
