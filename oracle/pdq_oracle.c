/*
 * oracle/pdq_oracle.c -- CPU restatement of the PDQ frame hash.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file's shared object.  The product (hydrus_video_deduplicator_b200/) never does.
 *
 * What it restates.  The reference (hydrus-video-deduplicator v0.11.2) does no arithmetic itself:
 *   vpdqpy.py:113  hasher = vpdq.VideoHasher(average_fps, 512, 512, num_threads)
 *   vpdqpy.py:118  hasher.hash_frame(bytes(frame.planes[0]))      # 512x512 RGB24, 786432 bytes
 *   vpdqpy.py:119  return hasher.finish()                         # -> VpdqHash (32 B / kept frame)
 * all land in the third-party wheel hvdaccelerators==0.4.0 (pyproject.toml:36, uv.lock:186-189), a
 * C++ extension around Meta ThreatExchange pdq/vpdq (docs/credits.md:7,9).  That wheel is not in
 * /root/reference and cannot be installed here (no network), so this file restates the PUBLISHED
 * PDQ algorithm (ThreatExchange pdq/cpp: pdqhashing.cpp, downscaling.cpp, torben.cpp,
 * pdqhashtypes.cpp) in the order written down in SURVEY.md Appendix A, and is pinned against the
 * reference's own golden vectors (the .txt files under tests/testdb/"video hashes", consumed by
 * tests/unit_tests/test_vpdqpy.py:105-128): see tests/test_oracle_golden.py.
 *
 * Arithmetic contract: IEEE binary32, round-to-nearest-even, one rounding per
 * written operation, NO fused multiply-add.  Build with -ffp-contract=off and without
 * -ffast-math / -march=native (oracle/Makefile).  What the goldens pin: structure, constants,
 * median rule, bit order, sampling.  What they cannot pin (60 frames are too few): the ulp-level
 * op order of the real binary -- the oracle is Appendix A by definition ("op order unpinned").
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PDQ_API __attribute__((visibility("default")))

/* ---- DCT matrix: 16 x 64, rows 1..16 of the 64-point DCT-II (Appendix A step 6) ---------- */
static float g_dct[16 * 64];
static pthread_once_t g_dct_once = PTHREAD_ONCE_INIT;

static void fill_dct(void) {
    const float scale = (float)sqrt(2.0 / 64.0);
    for (int i = 0; i < 16; i++)
        for (int j = 0; j < 64; j++)
            g_dct[i * 64 + j] = (float)(scale * cos((M_PI / 2 / 64.0) * (i + 1) * (2 * j + 1)));
}

PDQ_API const float* oracle_dct_matrix(void) {
    pthread_once(&g_dct_once, fill_dct);
    return g_dct;
}

/* ---- 1-D box filter with a running sum (Appendix A step 3; upstream downscaling.cpp) ------ */
static void box1d(const float* in, float* out, int n, int stride, int w) {
    const int half = (w + 2) / 2;
    const int p1 = half - 1;
    const int p2 = w - half + 1;
    const int p3 = n - w;
    const int p4 = half - 1;
    int li = 0, ri = 0, oi = 0;
    float sum = 0.0f;
    int cur = 0;
    for (int i = 0; i < p1; i++) { /* accumulate, no writes */
        sum += in[ri];
        cur++;
        ri += stride;
    }
    for (int i = 0; i < p2; i++) { /* growing window */
        sum += in[ri];
        cur++;
        out[oi] = sum / (float)cur;
        ri += stride;
        oi += stride;
    }
    for (int i = 0; i < p3; i++) { /* full window: add the right edge, THEN drop the left */
        sum += in[ri];
        sum -= in[li];
        out[oi] = sum / (float)cur;
        li += stride;
        ri += stride;
        oi += stride;
    }
    for (int i = 0; i < p4; i++) { /* shrinking window */
        sum -= in[li];
        cur--;
        out[oi] = sum / (float)cur;
        li += stride;
        oi += stride;
    }
}

static int jarosz_window(int dim) { return (dim + 2 * 64 - 1) / (2 * 64); }

/* two repetitions of [rows: a->b, cols: b->a] */
static void jarosz(float* a, float* b, int rows, int cols) {
    const int wr = jarosz_window(cols); /* window along a row  */
    const int wc = jarosz_window(rows); /* window along a col  */
    for (int rep = 0; rep < 2; rep++) {
        for (int i = 0; i < rows; i++) box1d(a + (size_t)i * cols, b + (size_t)i * cols, cols, 1, wr);
        for (int j = 0; j < cols; j++) box1d(b + j, a + j, rows, cols, wc);
    }
}

static void decimate64(const float* in, int rows, int cols, float* out) {
    for (int i = 0; i < 64; i++) {
        const int ini = (int)(((i + 0.5) * rows) / 64);
        for (int j = 0; j < 64; j++) {
            const int inj = (int)(((j + 0.5) * cols) / 64);
            out[i * 64 + j] = in[(size_t)ini * cols + inj];
        }
    }
}

static int quality64(const float* a) {
    int g = 0;
    for (int i = 0; i < 63; i++)
        for (int j = 0; j < 64; j++) {
            const float u = a[i * 64 + j], v = a[(i + 1) * 64 + j];
            const int d = (int)(((u - v) * 100.0f) / 255.0f);
            g += abs(d);
        }
    for (int i = 0; i < 64; i++)
        for (int j = 0; j < 63; j++) {
            const float u = a[i * 64 + j], v = a[i * 64 + j + 1];
            const int d = (int)(((u - v) * 100.0f) / 255.0f);
            g += abs(d);
        }
    int q = g / 90;
    return q > 100 ? 100 : q;
}

static void dct64to16(const float* a, float* t, float* b) {
    const float* d = oracle_dct_matrix();
    for (int i = 0; i < 16; i++)
        for (int j = 0; j < 64; j++) {
            float s = 0.0f;
            for (int k = 0; k < 64; k++) s += d[i * 64 + k] * a[k * 64 + j];
            t[i * 64 + j] = s;
        }
    for (int i = 0; i < 16; i++)
        for (int j = 0; j < 16; j++) {
            float s = 0.0f;
            for (int k = 0; k < 64; k++) s += t[i * 64 + k] * d[j * 64 + k];
            b[i * 16 + j] = s;
        }
}

/* Torben's median (upstream torben.cpp): for n = 256 returns the 128-th smallest value. */
static float torben(const float* m, int n) {
    float min = m[0], max = m[0];
    for (int i = 1; i < n; i++) {
        if (m[i] < min) min = m[i];
        if (m[i] > max) max = m[i];
    }
    int less, greater, equal;
    float guess, maxlt, mingt;
    for (;;) {
        guess = (min + max) / 2;
        less = greater = equal = 0;
        maxlt = min;
        mingt = max;
        for (int i = 0; i < n; i++) {
            if (m[i] < guess) {
                less++;
                if (m[i] > maxlt) maxlt = m[i];
            } else if (m[i] > guess) {
                greater++;
                if (m[i] < mingt) mingt = m[i];
            } else
                equal++;
        }
        if (less <= (n + 1) / 2 && greater <= (n + 1) / 2) break;
        if (less > greater)
            max = maxlt;
        else
            min = mingt;
    }
    if (less >= (n + 1) / 2) return maxlt;
    if (less + equal >= (n + 1) / 2) return guess;
    return mingt;
}

/* bit k = 16*i + j  ->  uint16 word k>>4, bit k&15, little-endian  ==  byte k>>3, bit k&7 */
static void bits256(const float* b, uint8_t out[32]) {
    const float med = torben(b, 256);
    memset(out, 0, 32);
    for (int k = 0; k < 256; k++)
        if (b[k] > med) out[k >> 3] |= (uint8_t)(1u << (k & 7));
}

/* ---- one frame, from a float luma plane that this call may overwrite ---------------------- */
typedef struct {
    float* a; /* rows*cols */
    float* b; /* rows*cols */
} scratch_t;

static int hash_from_luma(scratch_t* s, int rows, int cols, uint8_t hash[32], int* quality, float* a64_out,
                          float* b16_out) {
    float a64[64 * 64], t[16 * 64], b16[16 * 16];
    if (rows == 64 && cols == 64) {
        memcpy(a64, s->a, sizeof a64);
    } else {
        jarosz(s->a, s->b, rows, cols);
        decimate64(s->a, rows, cols, a64);
    }
    if (quality) *quality = quality64(a64);
    dct64to16(a64, t, b16);
    bits256(b16, hash);
    if (a64_out) memcpy(a64_out, a64, sizeof a64);
    if (b16_out) memcpy(b16_out, b16, sizeof b16);
    return 0;
}

static void luma_from_rgb(const uint8_t* rgb, size_t npix, float* luma) {
    for (size_t p = 0; p < npix; p++) {
        const float r = (float)rgb[3 * p], g = (float)rgb[3 * p + 1], b = (float)rgb[3 * p + 2];
        luma[p] = (0.299f * r + 0.587f * g) + 0.114f * b;
    }
}

/* BASELINE "luma frames": u8 gray L is DEFINED as the RGB frame R=G=B=L (SURVEY.md note a-1) */
static void luma_from_gray(const uint8_t* gray, size_t npix, float* luma) {
    for (size_t p = 0; p < npix; p++) {
        const float l = (float)gray[p];
        luma[p] = (0.299f * l + 0.587f * l) + 0.114f * l;
    }
}

static int hash_one(const uint8_t* px, int channels, int rows, int cols, uint8_t hash[32], int* quality,
                    float* a64_out, float* b16_out) {
    if (rows < 64 || cols < 64 || (channels != 1 && channels != 3)) return -1;
    const size_t npix = (size_t)rows * cols;
    /* per-thread scratch kept between frames (a fair baseline does not page-fault 2 MB per frame) */
    static __thread float* tl_buf = NULL;
    static __thread size_t tl_n = 0;
    if (tl_n < 2 * npix) {
        free(tl_buf);
        tl_buf = (float*)malloc(2 * npix * sizeof(float));
        tl_n = tl_buf ? 2 * npix : 0;
        if (!tl_buf) return -2;
    }
    scratch_t s;
    s.a = tl_buf;
    s.b = tl_buf + npix;
    if (channels == 3)
        luma_from_rgb(px, npix, s.a);
    else
        luma_from_gray(px, npix, s.a);
    return hash_from_luma(&s, rows, cols, hash, quality, a64_out, b16_out);
}

/* replaces VideoHasher.hash_frame's per-frame work (vpdqpy.py:118); rgb = rows x cols x 3 */
PDQ_API int oracle_pdq_hash_rgb(const uint8_t* rgb, int rows, int cols, uint8_t hash[32], int* quality) {
    return hash_one(rgb, 3, rows, cols, hash, quality, NULL, NULL);
}

PDQ_API int oracle_pdq_hash_gray(const uint8_t* gray, int rows, int cols, uint8_t hash[32], int* quality) {
    return hash_one(gray, 1, rows, cols, hash, quality, NULL, NULL);
}

/* debugging aid for the CUDA parity tests: also returns the 64x64 decimated plane and the DCT */
PDQ_API int oracle_pdq_stages_rgb(const uint8_t* rgb, int rows, int cols, uint8_t hash[32], int* quality,
                                  float* a64 /*4096*/, float* b16 /*256*/) {
    return hash_one(rgb, 3, rows, cols, hash, quality, a64, b16);
}

/* ---- batch, with a plain thread pool (mirrors VideoHasher(num_threads), vpdqpy.py:113) ----- */
typedef struct {
    const uint8_t* px;
    int channels, rows, cols;
    long n;
    uint8_t* hashes;
    int* qualities;
    long next;
    pthread_mutex_t mu;
    int err;
} batch_t;

static void* batch_worker(void* arg) {
    batch_t* b = (batch_t*)arg;
    const size_t fbytes = (size_t)b->rows * b->cols * b->channels;
    for (;;) {
        pthread_mutex_lock(&b->mu);
        const long i = b->next++;
        pthread_mutex_unlock(&b->mu);
        if (i >= b->n) break;
        const int rc = hash_one(b->px + (size_t)i * fbytes, b->channels, b->rows, b->cols, b->hashes + 32 * i,
                                b->qualities + i, NULL, NULL);
        if (rc) b->err = rc;
    }
    return NULL;
}

PDQ_API int oracle_pdq_hash_batch(const uint8_t* px, int channels, long n, int rows, int cols, uint8_t* hashes,
                                  int* qualities, int nthreads) {
    oracle_dct_matrix(); /* init once before threads race for it */
    batch_t b = {px, channels, rows, cols, n, hashes, qualities, 0, PTHREAD_MUTEX_INITIALIZER, 0};
    if (nthreads <= 1) {
        batch_worker(&b);
        return b.err;
    }
    if (nthreads > 1024) nthreads = 1024;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
    int started = 0;
    for (int t = 0; t < nthreads; t++)
        if (pthread_create(&th[started], NULL, batch_worker, &b) == 0) started++;
    if (started == 0) batch_worker(&b);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
    free(th);
    return b.err;
}
