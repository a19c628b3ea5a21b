/*
 * oracle/match_oracle.c -- CPU restatement of the similarity side.  TEST INFRASTRUCTURE ONLY
 * (same rule as pdq_oracle.c: the product never loads this).
 *
 * Restates, from the reference's call sites (the arithmetic is in the absent hvdaccelerators wheel):
 *   vpdqpy.py:50-56    Vpdq.match_hash(q, t, tol=31.0) -> vpdq.matchHash(q, t, int(tol))
 *   vpdqpy.py:122-131  Vpdq.is_similar(a, b, threshold=75.0) -> (sim >= threshold, sim)
 *   vptree.py:22-31    calculate_distance(a, b) = (100 - int(matchHashBytes(a, b, 31))) + 1
 *   vptree.py:865-902  search_file(hash_id, radius): every DB video with distance <= radius
 *   DedupeDB.py:555-557  an empty hash is similar to nothing, itself included
 *
 * Semantics fixed here ("parity unpinned" items 1-2; no reference test discriminates):
 *   frame match  : popcount(q_i ^ t_j) <= tol        (all genuine PDQ hashes are at even distances, F5)
 *   video score  : 100 * #{i : exists j, match(q_i, t_j)} / n_q   (upstream vpdq's qMatch), as a double
 *   either side empty -> 0.0
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PDQ_API __attribute__((visibility("default")))

static inline int hamming256(const uint8_t* a, const uint8_t* b) {
    uint64_t x[4], y[4];
    memcpy(x, a, 32);
    memcpy(y, b, 32);
    return __builtin_popcountll(x[0] ^ y[0]) + __builtin_popcountll(x[1] ^ y[1]) +
           __builtin_popcountll(x[2] ^ y[2]) + __builtin_popcountll(x[3] ^ y[3]);
}

PDQ_API int oracle_hamming256(const uint8_t* a, const uint8_t* b) { return hamming256(a, b); }

/* number of query frames with at least one target frame within tol */
PDQ_API long oracle_matched_frames(const uint8_t* q, long nq, const uint8_t* t, long nt, int tol) {
    long m = 0;
    for (long i = 0; i < nq; i++)
        for (long j = 0; j < nt; j++)
            if (hamming256(q + 32 * i, t + 32 * j) <= tol) {
                m++;
                break;
            }
    return m;
}

/* vpdq.matchHash / matchHashBytes (vpdqpy.py:56, vptree.py:31) */
PDQ_API double oracle_match_hash(const uint8_t* q, long nq, const uint8_t* t, long nt, int tol) {
    if (nq <= 0 || nt <= 0) return 0.0;
    return (100.0 * (double)oracle_matched_frames(q, nq, t, nt, tol)) / (double)nq;
}

/* vptree.py:22-25 fix_vpdq_similarity / :29-31 calculate_distance */
PDQ_API int oracle_calculate_distance(const uint8_t* a, long na, const uint8_t* b, long nb) {
    return (100 - (int)oracle_match_hash(a, na, b, nb, 31)) + 1;
}

/* Frame-level brute force: every ordered (i, j) with popcount(q_i ^ t_j) <= tol, row-major order.
 * Returns the total count; writes at most cap pairs. */
PDQ_API long oracle_hamming_pairs(const uint8_t* q, long nq, const uint8_t* t, long nt, int tol, int64_t* pairs,
                                  long cap) {
    long n = 0;
    for (long i = 0; i < nq; i++)
        for (long j = 0; j < nt; j++)
            if (hamming256(q + 32 * i, t + 32 * j) <= tol) {
                if (n < cap) {
                    pairs[2 * n] = i;
                    pairs[2 * n + 1] = j;
                }
                n++;
            }
    return n;
}

/* Video-level brute force over a CSR database: for query video (frames q) against every video v of the
 * target DB (frames t[off[v]..off[v+1])), matched[v] = # query frames with a match inside v. */
PDQ_API void oracle_video_matched(const uint8_t* q, long nq, const uint8_t* t, const int64_t* off, long nvideos,
                                  int tol, int32_t* matched) {
    for (long v = 0; v < nvideos; v++)
        matched[v] = (int32_t)oracle_matched_frames(q, nq, t + 32 * off[v], off[v + 1] - off[v], tol);
}

/* ---- threaded all-pairs count, for the CPU baseline (pair-comparisons / s) ------------------ */
typedef struct {
    const uint8_t *q, *t;
    long nq, nt;
    int tol;
    long next, chunk;
    long count;
    pthread_mutex_t mu;
} ap_t;

static void* ap_worker(void* arg) {
    ap_t* a = (ap_t*)arg;
    long local = 0;
    for (;;) {
        pthread_mutex_lock(&a->mu);
        const long i0 = a->next;
        a->next += a->chunk;
        pthread_mutex_unlock(&a->mu);
        if (i0 >= a->nq) break;
        const long i1 = i0 + a->chunk < a->nq ? i0 + a->chunk : a->nq;
        for (long i = i0; i < i1; i++)
            for (long j = 0; j < a->nt; j++) local += hamming256(a->q + 32 * i, a->t + 32 * j) <= a->tol;
    }
    pthread_mutex_lock(&a->mu);
    a->count += local;
    pthread_mutex_unlock(&a->mu);
    return NULL;
}

PDQ_API long oracle_hamming_count_mt(const uint8_t* q, long nq, const uint8_t* t, long nt, int tol, int nthreads) {
    ap_t a = {q, t, nq, nt, tol, 0, 64, 0, PTHREAD_MUTEX_INITIALIZER};
    if (nthreads <= 1) {
        ap_worker(&a);
        return a.count;
    }
    if (nthreads > 1024) nthreads = 1024;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
    int started = 0;
    for (int k = 0; k < nthreads; k++)
        if (pthread_create(&th[started], NULL, ap_worker, &a) == 0) started++;
    if (started == 0) ap_worker(&a);
    for (int k = 0; k < started; k++) pthread_join(th[k], NULL);
    free(th);
    return a.count;
}
