// hamming_kernels.cu -- brute-force Hamming similarity over packed 256-bit PDQ hashes (sm_100a).
//
// Replaces hvdaccelerators.vpdq.matchHash / matchHashBytes (reference call sites
// vpdqpy/vpdqpy.py:56 and db/vptree.py:31) and, as ONE pass over the hash database, the
// calculate_distance calls the vp-tree issues per visited node (db/vptree.py:737).
//
// Frame match:  popcount(q ^ t) <= tol  on 4 x u64.  Integer work only: no
// tensor cores, no floating point.
//
//   k_hamming_scan   streaming regime (the reference's real call pattern: one query video against the
//                    whole DB).  One thread per DB hash, one 256-bit load (LDG.E.256) per hash, the
//                    <= 64 query hashes broadcast from shared memory, 96-bit carry-save prefilter
//                    (2 POPC per pair).  HBM-bound up to ~10 resident queries: algorithmic traffic =
//                    32 B per DB hash.
//   k_hamming_pairs  all-pairs regime (1M x 1M).  The DB is L2-resident there and the kernel is bound
//                    by POPC issue (XU pipe, measured 16 POPC/clk/SM), so the inner loop is a 96-bit
//                    prefilter with a carry-save step: for the three xor words x0,x1,x2,
//                    popc(x0)+popc(x1)+popc(x2) = popc(x0^x1^x2) + 2*popc(maj(x0,x1,x2)) -- 2 POPC per
//                    pair instead of 3, the rest on the LOP3 pipe.  A pair whose first 96 bits already
//                    differ in > tol places cannot match; survivors (~4e-4 of random pairs) take the
//                    exact 256-bit path.
#include <cuda_pipeline.h>

#include "common.cuh"

namespace vpdq {

__device__ __forceinline__ void ld256(const uint64_t* p, uint32_t (&w)[8]) {
    uint64_t a, b, c, d;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(a), "=l"(b), "=l"(c), "=l"(d)
                 : "l"(p));
    w[0] = (uint32_t)a; w[1] = (uint32_t)(a >> 32);
    w[2] = (uint32_t)b; w[3] = (uint32_t)(b >> 32);
    w[4] = (uint32_t)c; w[5] = (uint32_t)(c >> 32);
    w[6] = (uint32_t)d; w[7] = (uint32_t)(d >> 32);
}

// video owning frame idx: the v with offsets[v] <= idx < offsets[v+1]
__device__ __forceinline__ int64_t video_of(const int64_t* __restrict__ offsets, int64_t n_videos, int64_t idx) {
    int64_t lo = 0, hi = n_videos;
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(offsets + mid) <= idx)
            lo = mid;
        else
            hi = mid;
    }
    return lo;
}

// distance over the first 96 bits with 2 POPC (carry-save adder over the three xor words).  The two 3-input
// functions are pinned to one LOP3 each (xor3 = 0x96, majority = 0xE8) and the final  ps + 2*pc  goes to the
// otherwise idle FMA pipe (IMAD), so the ALU pipe carries 5 ops per pair and the XU pipe 2.
__device__ __forceinline__ int prefix96_distance(const uint32_t (&q)[3], const uint4& t) {
    const uint32_t x0 = q[0] ^ t.x, x1 = q[1] ^ t.y, x2 = q[2] ^ t.z;
    uint32_t sum, carry, d;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(sum) : "r"(x0), "r"(x1), "r"(x2));
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(carry) : "r"(x0), "r"(x1), "r"(x2));
    const uint32_t ps = __popc(sum), pc = __popc(carry);
    asm("mad.lo.u32 %0, %1, 2, %2;" : "=r"(d) : "r"(pc), "r"(ps));
    return (int)d;
}

// ---------------------------------------------------------------------------------------------------
// streaming scan
// ---------------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanUnroll = 2;

// chunk_rows == nullptr: ONE chunk = query rows [0, n_query) (the single-video call).  Otherwise blockIdx.y = chunk c:
// query rows chunk_rows[c] .. chunk_rows[c+1] - 1 (<= 64 of them), results in qmask[c][n_videos] -- many query videos
// (or the pieces of a long one) against the database in one launch.
__global__ void __launch_bounds__(kScanThreads)
    k_hamming_scan(const uint64_t* __restrict__ db, int64_t n_db, const int64_t* __restrict__ offsets,
                   int64_t n_videos, const uint64_t* __restrict__ query, int n_query,
                   const int32_t* __restrict__ chunk_rows, int tol, unsigned long long* __restrict__ qmask,
                   int32_t* __restrict__ tcount) {
    __shared__ uint4 q_s[64 * 2];
    if (chunk_rows) {
        const int r0 = __ldg(chunk_rows + blockIdx.y);
        n_query = __ldg(chunk_rows + blockIdx.y + 1) - r0;
        query += 4 * (int64_t)r0;
        qmask += (int64_t)blockIdx.y * n_videos;
    }
    for (int e = threadIdx.x; e < n_query * 2; e += kScanThreads)
        q_s[e] = __ldg(reinterpret_cast<const uint4*>(query) + e);
    __syncthreads();

    constexpr int64_t kTile = (int64_t)kScanThreads * kScanUnroll;
    const int64_t n_tiles = (n_db + kTile - 1) / kTile;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t base = tile * kTile + threadIdx.x;
        uint32_t h[kScanUnroll][8];
#pragma unroll
        for (int u = 0; u < kScanUnroll; ++u) {
            const int64_t idx = base + (int64_t)u * kScanThreads;
            if (idx < n_db) ld256(db + 4 * idx, h[u]);
        }
#pragma unroll
        for (int u = 0; u < kScanUnroll; ++u) {
            const int64_t idx = base + (int64_t)u * kScanThreads;
            if (idx >= n_db) continue;
            unsigned long long mask = 0ull;
            const uint32_t h3[3] = {h[u][0], h[u][1], h[u][2]};
            auto exact = [&](int qi, int d96) {  // rare: the 96-bit prefix is within tol
                const uint4 a = q_s[2 * qi], b = q_s[2 * qi + 1];
                const int d = d96 + __popc(h[u][3] ^ a.w) + __popc(h[u][4] ^ b.x) + __popc(h[u][5] ^ b.y) +
                              __popc(h[u][6] ^ b.z) + __popc(h[u][7] ^ b.w);
                if (d <= tol) mask |= 1ull << qi;
            };
            int qi = 0;
            for (; qi + 4 <= n_query; qi += 4) {  // four independent prefilters in flight (ILP)
                int d[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) d[j] = prefix96_distance(h3, q_s[2 * (qi + j)]);  // 2 POPC each
                if (min(min(d[0], d[1]), min(d[2], d[3])) <= tol) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (d[j] <= tol) exact(qi + j, d[j]);
                }
            }
            for (; qi < n_query; ++qi) {
                const int d = prefix96_distance(h3, q_s[2 * qi]);
                if (d <= tol) exact(qi, d);
            }
            if (mask) {
                const int64_t v = offsets ? video_of(offsets, n_videos, idx) : idx;
                atomicOr(qmask + v, mask);
                if (tcount) atomicAdd(tcount + v, 1);
            }
        }
    }
}

int hamming_scan_launch(const uint64_t* d_db, int64_t n_db, const int64_t* d_offsets, int64_t n_videos,
                        const uint64_t* d_query, int n_query, int tol, uint64_t* d_qmask, int32_t* d_tcount,
                        cudaStream_t stream) {
    if (n_db == 0 || n_query == 0) return VPDQ_B200_OK;
    int dev = 0, sms = 148;
    VPDQ_CUDA(cudaGetDevice(&dev));
    VPDQ_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int64_t kTile = (int64_t)kScanThreads * kScanUnroll;
    const int64_t n_tiles = (n_db + kTile - 1) / kTile;
    const int64_t max_grid = (int64_t)sms * 8;  // 8 resident CTAs of 256 threads per SM, persistent
    const unsigned grid = (unsigned)(n_tiles < max_grid ? n_tiles : max_grid);
    k_hamming_scan<<<grid, kScanThreads, 0, stream>>>(d_db, n_db, d_offsets, n_videos, d_query, n_query, nullptr, tol,
                                                      reinterpret_cast<unsigned long long*>(d_qmask), d_tcount);
    g_launches += 1;
    VPDQ_CUDA(cudaGetLastError());
    return VPDQ_B200_OK;
}

int hamming_scan_multi_launch(const uint64_t* d_db, int64_t n_db, const int64_t* d_offsets, int64_t n_videos,
                              const uint64_t* d_query, const int32_t* d_chunk_rows, int n_chunks, int tol,
                              uint64_t* d_qmask, cudaStream_t stream) {
    if (n_db == 0 || n_chunks == 0) return VPDQ_B200_OK;
    int dev = 0, sms = 148;
    VPDQ_CUDA(cudaGetDevice(&dev));
    VPDQ_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int64_t kTile = (int64_t)kScanThreads * kScanUnroll;
    const int64_t n_tiles = (n_db + kTile - 1) / kTile;
    // the whole grid (all chunks) fills the machine a few times over; a chunk's CTAs stride over the DB tiles
    int64_t per_chunk = ((int64_t)sms * 8 + n_chunks - 1) / n_chunks;
    if (per_chunk > n_tiles) per_chunk = n_tiles;
    if (per_chunk < 1) per_chunk = 1;
    for (int c0 = 0; c0 < n_chunks; c0 += 65535) {  // grid.y limit
        const int nc = n_chunks - c0 < 65535 ? n_chunks - c0 : 65535;
        dim3 grid((unsigned)per_chunk, (unsigned)nc);
        k_hamming_scan<<<grid, kScanThreads, 0, stream>>>(d_db, n_db, d_offsets, n_videos, d_query, 0, d_chunk_rows + c0, tol,
                                                          reinterpret_cast<unsigned long long*>(d_qmask) + (int64_t)c0 * n_videos,
                                                          nullptr);
        g_launches += 1;
    }
    VPDQ_CUDA(cudaGetLastError());
    return VPDQ_B200_OK;
}

// ---------------------------------------------------------------------------------------------------
// video-level reduce: per-chunk frame masks -> matchHash numerators -> distances -> compact list
// ---------------------------------------------------------------------------------------------------
// The reference scores a pair of videos as  similarity = 100 * |{query frames with >= 1 match}| / n_query_frames  and
// distance = (100 - int(similarity)) + 1  (vpdqpy.py:56, db/vptree.py:22-31).  After a (multi-chunk) scan, bit i of
// qmask[c][v] says that query frame i of chunk c has a match in target video v; query video q owns the chunks
// qv_chunks[q] .. qv_chunks[q+1] - 1 and has qv_frames[q] frames.  One thread per target video:
//   matched(q, v) = sum over q's chunks of popcount(qmask[c][v])      (distinct query frames by construction)
// written densely (matched_dense[q][v]) and / or appended as (q, v, matched, distance) rows when matched > 0 and
// distance <= max_distance (max_distance <= 0: no distance filter).  floor(100 m / n) in integers equals the
// reference's int(100.0 * m / n): a quotient that is not an integer is at least 1 / n away from one.
__global__ void __launch_bounds__(256)
    k_video_reduce(const unsigned long long* __restrict__ qmask, int64_t n_videos, const int32_t* __restrict__ qv_chunks,
                   const int32_t* __restrict__ qv_frames, int n_qvideos, int max_distance,
                   int32_t* __restrict__ matched_dense, int4* __restrict__ rows, long long cap,
                   unsigned long long* __restrict__ count) {
    const int64_t v = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (v >= n_videos) return;
    for (int q = blockIdx.y; q < n_qvideos; q += gridDim.y) {
        const int c0 = __ldg(qv_chunks + q), c1 = __ldg(qv_chunks + q + 1);
        int m = 0;
        for (int c = c0; c < c1; ++c) m += __popcll(__ldg(qmask + (int64_t)c * n_videos + v));
        if (matched_dense) matched_dense[(int64_t)q * n_videos + v] = m;
        if (rows && m > 0) {
            const int n = __ldg(qv_frames + q);
            const int dist = n > 0 ? (100 - (100 * m) / n) + 1 : 101;
            if (max_distance <= 0 || dist <= max_distance) {
                const unsigned long long pos = atomicAdd(count, 1ull);
                if ((long long)pos < cap) rows[pos] = make_int4(q, (int)v, m, dist);
            }
        }
    }
}

int video_reduce_launch(const uint64_t* d_qmask, int64_t n_videos, const int32_t* d_qv_chunks, const int32_t* d_qv_frames,
                        int n_qvideos, int max_distance, int32_t* d_matched_dense, int32_t* d_rows, int64_t cap,
                        unsigned long long* d_count, cudaStream_t stream) {
    if (n_videos == 0 || n_qvideos == 0) return VPDQ_B200_OK;
    const int64_t gx = (n_videos + 255) / 256;
    if (gx > 0x7fffffff) {
        set_error("video_reduce: too many videos");
        return VPDQ_B200_ERR_INVALID;
    }
    dim3 grid((unsigned)gx, (unsigned)(n_qvideos < 65535 ? n_qvideos : 65535));
    k_video_reduce<<<grid, 256, 0, stream>>>(reinterpret_cast<const unsigned long long*>(d_qmask), n_videos, d_qv_chunks,
                                             d_qv_frames, n_qvideos, max_distance, d_matched_dense,
                                             reinterpret_cast<int4*>(d_rows), (long long)cap, d_count);
    g_launches += 1;
    VPDQ_CUDA(cudaGetLastError());
    return VPDQ_B200_OK;
}

// ---------------------------------------------------------------------------------------------------
// all pairs
// ---------------------------------------------------------------------------------------------------
constexpr int kPairThreads = 256;
constexpr int kPairQR = 8;                            // queries per thread (96-bit prefixes in registers)
constexpr int kPairQTile = kPairThreads * kPairQR;    // 2048 queries per CTA
constexpr int kPairTTile = 1024;                      // target prefixes per shared-memory stage (16 KB)

__device__ __forceinline__ int full_distance(const uint64_t* __restrict__ a, const uint64_t* __restrict__ b) {
    return __popcll(__ldg(a) ^ __ldg(b)) + __popcll(__ldg(a + 1) ^ __ldg(b + 1)) +
           __popcll(__ldg(a + 2) ^ __ldg(b + 2)) + __popcll(__ldg(a + 3) ^ __ldg(b + 3));
}

__global__ void __launch_bounds__(kPairThreads)
    k_hamming_pairs(const uint64_t* __restrict__ qs, int64_t n_q, const uint64_t* __restrict__ ts, int64_t n_t,
                    int64_t t_chunk, int tol, int skip_diagonal, uint32_t* __restrict__ any,
                    unsigned long long* __restrict__ pairs, int64_t cap, unsigned long long* __restrict__ count) {
    __shared__ uint4 tile[2][kPairTTile];  // first 128 bits of each target hash

    const int64_t q0 = (int64_t)blockIdx.x * kPairQTile + threadIdx.x;
    uint32_t q[kPairQR][3];
#pragma unroll
    for (int r = 0; r < kPairQR; ++r) {
        const int64_t i = q0 + (int64_t)r * kPairThreads;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (i < n_q) v = __ldg(reinterpret_cast<const uint4*>(qs + 4 * i));
        q[r][0] = v.x; q[r][1] = v.y; q[r][2] = v.z;
    }

    const int64_t t_begin = (int64_t)blockIdx.y * t_chunk;
    const int64_t t_end = (t_begin + t_chunk < n_t) ? (t_begin + t_chunk) : n_t;
    if (t_begin >= t_end) return;
    const int n_stage = (int)((t_end - t_begin + kPairTTile - 1) / kPairTTile);

    auto prefetch = [&](int s) {
        const int64_t base = t_begin + (int64_t)s * kPairTTile;
#pragma unroll
        for (int e = 0; e < kPairTTile / kPairThreads; ++e) {
            const int slot = e * kPairThreads + threadIdx.x;
            const int64_t j = base + slot;
            if (j < t_end) __pipeline_memcpy_async(&tile[s & 1][slot], ts + 4 * j, 16);
        }
        __pipeline_commit();
    };

    prefetch(0);
    for (int s = 0; s < n_stage; ++s) {
        if (s + 1 < n_stage) prefetch(s + 1);
        else __pipeline_commit();
        __pipeline_wait_prior(1);
        __syncthreads();

        const int64_t base = t_begin + (int64_t)s * kPairTTile;
        const int cnt = (int)((t_end - base < kPairTTile) ? (t_end - base) : kPairTTile);
        const uint4* tl = tile[s & 1];
#pragma unroll 4
        for (int tt = 0; tt < cnt; ++tt) {
            const uint4 tp = tl[tt];  // broadcast read
            int best = 1 << 20;
#pragma unroll
            for (int r = 0; r < kPairQR; ++r) {
                best = min(best, prefix96_distance(q[r], tp));
            }
            if (best <= tol) {  // rare: some query's 96-bit prefix is within tol of this target
                const int64_t j = base + tt;
#pragma unroll
                for (int r = 0; r < kPairQR; ++r) {
                    const int d = prefix96_distance(q[r], tp);
                    const int64_t i = q0 + (int64_t)r * kPairThreads;
                    if (d <= tol && i < n_q && !(skip_diagonal && i == j) &&
                        full_distance(qs + 4 * i, ts + 4 * j) <= tol) {
                        if (any) atomicOr(any + (i >> 5), 1u << (i & 31));
                        const unsigned long long pos = atomicAdd(count, 1ull);
                        if (pairs && (int64_t)pos < cap) pairs[pos] = ((unsigned long long)i << 32) | (unsigned long long)j;
                    }
                }
            }
        }
        __syncthreads();  // everyone is done with tile[s & 1] before it is refilled at s + 2
    }
}

int hamming_pairs_launch(const uint64_t* d_q, int64_t n_q, const uint64_t* d_t, int64_t n_t, int tol,
                         int skip_diagonal, uint32_t* d_any, uint64_t* d_pairs, int64_t cap,
                         unsigned long long* d_count, cudaStream_t stream) {
    if (n_q == 0 || n_t == 0) return VPDQ_B200_OK;
    if (n_q >= (1ll << 32) || n_t >= (1ll << 32)) {
        set_error("hamming_pairs: at most 2^32-1 hashes per side");
        return VPDQ_B200_ERR_INVALID;
    }
    int dev = 0, sms = 148;
    VPDQ_CUDA(cudaGetDevice(&dev));
    VPDQ_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int64_t q_tiles = (n_q + kPairQTile - 1) / kPairQTile;
    // split the targets so that the grid holds >= 4 CTAs per SM, in whole shared-memory stages
    int64_t want = ((int64_t)sms * 4 + q_tiles - 1) / q_tiles;
    const int64_t t_stages = (n_t + kPairTTile - 1) / kPairTTile;
    if (want > t_stages) want = t_stages;
    if (want < 1) want = 1;
    if (want > 65535) want = 65535;
    const int64_t t_chunk = ((t_stages + want - 1) / want) * kPairTTile;
    const int64_t t_splits = (n_t + t_chunk - 1) / t_chunk;
    dim3 grid((unsigned)q_tiles, (unsigned)t_splits);
    k_hamming_pairs<<<grid, kPairThreads, 0, stream>>>(d_q, n_q, d_t, n_t, t_chunk, tol, skip_diagonal, d_any,
                                                       reinterpret_cast<unsigned long long*>(d_pairs), cap, d_count);
    g_launches += 1;
    VPDQ_CUDA(cudaGetLastError());
    return VPDQ_B200_OK;
}

}  // namespace vpdq
