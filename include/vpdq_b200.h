/*
 * vpdq_b200.h -- C ABI of libvpdq_b200.so: the B200 (sm_100a) replacement for the native half of
 * hydrus-video-deduplicator's hot path.
 *
 * In the reference every entry below is a call into the third-party C++ wheel
 * `hvdaccelerators.vpdq` (pyproject.toml:36); the citations name the reference-side call site each
 * entry point replaces (paths relative to /root/reference/src/hydrusvideodeduplicator/).
 *
 * Conventions
 *   - plain C, no exceptions across the boundary; every function returns VPDQ_B200_OK (0) or a
 *     negative vpdq_b200_status; vpdq_b200_last_error() returns a thread-local message.
 *   - "_dev" entry points take DEVICE pointers, are stream-ordered on `stream` (a cudaStream_t passed
 *     as void*; NULL = the legacy default stream), never allocate and never synchronise.
 *   - "_host" entry points and the hasher handle take HOST pointers, own their staging buffers and
 *     include the host<->device copies (this is what the reference's Python binds to).
 *   - a PDQ hash is 32 bytes in "native PDQ order": bit k = 16*i + j of the 16x16 DCT sign matrix
 *     lives in byte k>>3, bit k&7 (DedupeDB.py:538-544); hash matrices are row-major [n][32] bytes,
 *     equivalently [n][4] little-endian uint64 -- exactly the reference's phash BLOB (dedup.py:77).
 *   - frames are 512 x 512, row-major, RGB24 interleaved (channels = 3; vpdqpy.py:90-95,118) or
 *     8-bit gray (channels = 1, DEFINED as the RGB frame R=G=B=L).
 *   - there is NO CPU implementation behind any of these: without a CUDA device they fail with
 *     VPDQ_B200_ERR_CUDA.
 */
#ifndef VPDQ_B200_H
#define VPDQ_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define VPDQ_B200_API __attribute__((visibility("default")))
#else
#define VPDQ_B200_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef enum vpdq_b200_status {
    VPDQ_B200_OK = 0,
    VPDQ_B200_ERR_INVALID = -1,     /* bad argument (NULL, negative size, wrong frame size, ...)  */
    VPDQ_B200_ERR_CUDA = -2,        /* a CUDA runtime call or kernel failed / no device           */
    VPDQ_B200_ERR_NOMEM = -3,       /* host or device allocation failed                           */
    VPDQ_B200_ERR_UNSUPPORTED = -4, /* e.g. frame dimensions other than 512 x 512                 */
    VPDQ_B200_ERR_OVERFLOW = -5     /* an output list was larger than the capacity supplied       */
} vpdq_b200_status;

#define VPDQ_B200_FRAME_DIM 512
#define VPDQ_B200_HASH_BYTES 32          /* VpdqHash.bytesPerPdqHash, dedup.py:83-84               */
#define VPDQ_B200_QUALITY_KEEP 31        /* finish() keeps quality >= 31, DedupeDB.py:550-553      */
#define VPDQ_B200_DEFAULT_TOLERANCE 31   /* vpdqpy.py:53, vptree.py:31                             */

VPDQ_B200_API const char* vpdq_b200_last_error(void);
VPDQ_B200_API int vpdq_b200_abi_version(void);
VPDQ_B200_API int vpdq_b200_device_count(int* count);
/* Self-check after a device synchronise: 0 = healthy; bit 0 = a TMA copy inside a PDQ kernel never completed (its
 * bounded wait gave up) -- results of that launch are then invalid.  Every host-pointer entry point (hash_frames_host,
 * the hasher handle) reads the same flags at its own synchronisation point and fails with VPDQ_B200_ERR_CUDA; callers
 * of the stream-ordered *_dev entry points check here after they synchronise. */
VPDQ_B200_API int vpdq_b200_debug_flags(int device, int* flags);
/* test hook: set (value != 0) or clear the flags above, to exercise the failure path */
VPDQ_B200_API int vpdq_b200_debug_force_timeout(int device, int value);
/* number of CUDA kernels this library has launched in this process (monotonic) */
VPDQ_B200_API int vpdq_b200_kernel_launches(uint64_t* count);
/* the 16 x 64 fp32 DCT table the kernels use (host copy; bit-identical to the oracle's) */
VPDQ_B200_API int vpdq_b200_dct_matrix(float* out /* [16*64] */);

/* ------------------------------------------------------------------------------------------------
 * PDQ frame hashing.  Replaces the work behind VideoHasher.hash_frame (vpdqpy.py:118):
 * RGB->luma, 2x Jarosz box blur, decimate to 64x64, quality metric, 64->16 DCT, median, 256 bits.
 * ---------------------------------------------------------------------------------------------- */

/* bytes of device scratch vpdq_b200_pdq_hash_frames_dev needs for `n_frames` frames */
VPDQ_B200_API int vpdq_b200_pdq_scratch_bytes(int64_t n_frames, size_t* bytes);

/* d_frames [n][512][512][channels] u8 -> d_hashes [n][32] u8, d_quality [n] i32 (0..100).
 * No quality filtering here: the caller (finish()) drops frames. */
VPDQ_B200_API int vpdq_b200_pdq_hash_frames_dev(const uint8_t* d_frames, int channels, int64_t n_frames, int width, int height,
                                  uint8_t* d_hashes, int32_t* d_quality, void* d_scratch, size_t scratch_bytes,
                                  void* stream);

/* Debug/parity aid: additionally returns the decimated 64x64 plane ([n][64][64] f32) and the
 * 16x16 DCT ([n][16][16] f32); either may be NULL. */
VPDQ_B200_API int vpdq_b200_pdq_stages_dev(const uint8_t* d_frames, int channels, int64_t n_frames, int width, int height,
                             uint8_t* d_hashes, int32_t* d_quality, float* d_a64, float* d_b16, void* d_scratch,
                             size_t scratch_bytes, void* stream);

/* First half of the frame hash on its own: RGB24 frames -> the Jarosz-filtered, decimated 64x64 luma plane
 * d_a64 [n][64][64] f32 (PDQ's "buffer64x64", the input of the quality metric and the DCT).  This is exactly the
 * kernel kx_systolic_jarosz; bench.py times it alone for the roofline of the dominant kernel. */
VPDQ_B200_API int vpdq_b200_pdq_jarosz_dev(const uint8_t* d_frames, int64_t n_frames, int width, int height,
                                           float* d_a64, void* stream);

/* The reference's frame.reformat(width=512, height=512, format="rgb24", interpolation=POINT)
 * (vpdqpy.py:90-95) on the device: d_src [n][src_height][src_width][3] u8 -> d_dst [n][512][512][3] u8,
 * swscale's centre-based nearest neighbour.  Stream-ordered, no allocation. */
VPDQ_B200_API int vpdq_b200_point_resize_dev(const uint8_t* d_src, int64_t n_frames, int src_height, int src_width,
                                             uint8_t* d_dst, void* stream);

/* One-shot host call: h_frames -> h_hashes/h_quality, copies included (pinned or pageable memory). */
VPDQ_B200_API int vpdq_b200_pdq_hash_frames_host(const uint8_t* h_frames, int channels, int64_t n_frames, int width, int height,
                                   uint8_t* h_hashes, int32_t* h_quality, int device);

/* Streaming hasher handle == hvdaccelerators.vpdq.VideoHasher
 *   create  <- vpdq.VideoHasher(average_fps, width, height, num_threads)        vpdqpy.py:113
 *   push    <- hasher.hash_frame(bytes)   (blocks while the staging ring is full) vpdqpy.py:115-118
 *   finish  <- hasher.finish(): hashes of the frames with quality >= 31, in push order  vpdqpy.py:119
 * `num_threads` is accepted for signature compatibility and ignored (the GPU is the pool).
 *
 * A handle owns no device or pinned memory: all handles of a (device, channels) pair feed ONE submission service
 * (created on first use: a pinned + a device ring of VPDQ_B200_ARENA_FRAMES [256] frame slots, VPDQ_B200_COPY_THREADS
 * [min(8, cores/2)] copy workers, one pump thread).  Frames of different handles share uploads and kernel launches;
 * finish() waits for the handle's own frames only.  Handles may be used from different threads concurrently (one
 * thread per handle at a time). */
typedef struct vpdq_b200_hasher vpdq_b200_hasher;
VPDQ_B200_API int vpdq_b200_hasher_create(int device, int width, int height, int channels, int num_threads,
                            vpdq_b200_hasher** out);
/* copies the frames out of h_frames before returning (the caller may reuse the buffer at once) */
VPDQ_B200_API int vpdq_b200_hasher_push(vpdq_b200_hasher* h, const uint8_t* h_frames, int64_t n_frames);
/* Same, but returns as soon as the frames are queued: the copy workers read h_frames LATER.  The caller must keep the
 * memory alive and unchanged until vpdq_b200_hasher_consumed() has passed these frames or finish() / destroy()
 * returned.  (The Python binding uses it for immutable `bytes` frames, which it simply keeps a reference to:
 * hash_frame(bytes(frame.planes[0])), vpdqpy.py:118, then costs one queue insertion on the caller's thread.) */
VPDQ_B200_API int vpdq_b200_hasher_push_nocopy(vpdq_b200_hasher* h, const uint8_t* h_frames, int64_t n_frames);
/* number of leading frames (in push order, since the last finish) whose source memory is no longer needed */
VPDQ_B200_API int vpdq_b200_hasher_consumed(vpdq_b200_hasher* h, int64_t* n);
/* counters of the submission service of (device, channels) since it started, for tuning and reports:
 * out[0] frames pushed, [1] upload (H2D) calls, [2] kernel launch groups, [3] frames launched, [4] ns pushes spent blocked on
 * a full ring, [5] ns finish() calls spent waiting, [6] largest launch, [7] reserved */
VPDQ_B200_API int vpdq_b200_service_stats(int device, int channels, int64_t* out /* [8] */);
/* total frames pushed so far (upper bound for finish's capacity) */
VPDQ_B200_API int vpdq_b200_hasher_pushed(vpdq_b200_hasher* h, int64_t* n);
/* Writes kept hashes (quality >= quality_keep) to h_hashes [cap][32]; *n_kept = number kept.
 * h_all_hashes [pushed][32] / h_all_quality [pushed] (optional, may be NULL) receive the unfiltered
 * results.  The hasher is reset and reusable afterwards. */
VPDQ_B200_API int vpdq_b200_hasher_finish(vpdq_b200_hasher* h, int quality_keep, uint8_t* h_hashes, int64_t cap, int64_t* n_kept,
                            uint8_t* h_all_hashes, int32_t* h_all_quality);
VPDQ_B200_API int vpdq_b200_hasher_destroy(vpdq_b200_hasher* h);

/* ------------------------------------------------------------------------------------------------
 * Hamming similarity.  Replaces vpdq.matchHash (vpdqpy.py:56), vpdq.matchHashBytes (vptree.py:31)
 * and, as one brute-force pass, the calculate_distance storm under VpTreeManager.search_file
 * (vptree.py:664-815, 865-902).  Frame match: popcount(q ^ t) <= tolerance.
 * ---------------------------------------------------------------------------------------------- */

/* Streaming scan of a hash database against ONE query video (<= 64 frames per call).
 *   d_db      [n_db][4] u64        database frame hashes, videos stored contiguously
 *   d_offsets [n_videos + 1] i64   CSR: video v owns frames d_offsets[v] .. d_offsets[v+1]-1
 *                                  (NULL: every frame is its own video, n_videos must equal n_db)
 *   d_query   [n_query][4] u64
 *   d_qmask   [n_videos] u64  OUT  bit i set  <=>  query frame i matches >= 1 frame of video v
 *                                  (popcount = numerator of matchHash(query, video v)); the call ORs
 *                                  into it, zero it first (cudaMemsetAsync) for a fresh scan
 *   d_tcount  [n_videos] i32  OUT  (optional) # frames of video v matching >= 1 query frame
 *                                  (numerator of the reverse direction matchHash(video v, query)); adds */
VPDQ_B200_API int vpdq_b200_hamming_scan_dev(const uint64_t* d_db, int64_t n_db, const int64_t* d_offsets, int64_t n_videos,
                               const uint64_t* d_query, int n_query, int tolerance, uint64_t* d_qmask,
                               int32_t* d_tcount, void* stream);

/* The same scan for MANY query chunks in one launch (several query videos, or the 64-frame pieces of a long one):
 *   d_query      [n_rows][4] u64      all query hashes, 32-byte aligned
 *   d_chunk_rows [n_chunks + 1] i32   chunk c = query rows d_chunk_rows[c] .. d_chunk_rows[c+1]-1 (at most 64)
 *   d_qmask      [n_chunks][n_videos] u64  OUT  as above, one row of masks per chunk; ORs, zero it first */
VPDQ_B200_API int vpdq_b200_hamming_scan_multi_dev(const uint64_t* d_db, int64_t n_db, const int64_t* d_offsets,
                                     int64_t n_videos, const uint64_t* d_query, const int32_t* d_chunk_rows, int n_chunks,
                                     int tolerance, uint64_t* d_qmask, void* stream);

/* Video-level reduce of scan masks == vpdq.matchHash + calculate_distance for every (query video, target video)
 * (vpdqpy.py:56, db/vptree.py:22-31), on the device:
 *   query video q owns chunks d_qv_chunks[q] .. d_qv_chunks[q+1]-1 of d_qmask and has d_qv_frames[q] frames;
 *   matched(q, v) = # query frames of q with >= 1 match in target video v; distance = (100 - 100*matched/n_q) + 1.
 *   d_matched [n_qvideos][n_videos] i32  OUT (optional) matched(q, v), dense
 *   d_rows    [cap][4] i32               OUT (optional, 16-byte aligned) one row (q, v, matched, distance) per pair
 *                                        with matched > 0 and (max_distance <= 0 or distance <= max_distance), unordered
 *   d_count   [1] u64                    OUT number of rows (may exceed cap; adds) */
VPDQ_B200_API int vpdq_b200_video_match_dev(const uint64_t* d_qmask, int64_t n_videos, const int32_t* d_qv_chunks,
                              const int32_t* d_qv_frames, int n_qvideos, int max_distance, int32_t* d_matched,
                              int32_t* d_rows, int64_t cap, unsigned long long* d_count, void* stream);

/* Brute-force all pairs between two hash sets (self-join when both are the same buffer).
 *   d_any   [(n_q + 31) / 32] u32  OUT (optional) bit i: query i has >= 1 match ("candidate bitmap"); ORs
 *   d_pairs [cap] u64              OUT (optional) (i << 32) | j for every match, unordered
 *   d_count [1] u64                OUT total number of matches found (may exceed cap; adds)
 *   skip_diagonal != 0 ignores i == j (self-join). */
VPDQ_B200_API int vpdq_b200_hamming_pairs_dev(const uint64_t* d_q, int64_t n_q, const uint64_t* d_t, int64_t n_t, int tolerance,
                                int skip_diagonal, uint32_t* d_any, uint64_t* d_pairs, int64_t cap,
                                unsigned long long* d_count, void* stream);

/* vpdq.matchHash / matchHashBytes with host buffers: *similarity = 100 * matched_q / n_q as a double;
 * 0.0 when either side is empty (DedupeDB.py:555-557).  n_q, n_t in frames. */
VPDQ_B200_API int vpdq_b200_match_hash_host(const uint8_t* h_q, int64_t n_q, const uint8_t* h_t, int64_t n_t, int tolerance,
                              double* similarity, int device);

/* Resident hash database: the brute-force replacement of the vp-tree index (db/vptree.py).  The DB
 * (the phash BLOBs of shape_perceptual_hashes, DedupeDB.py:159-180, concatenated; video v owns frames
 * h_offsets[v] .. h_offsets[v+1]-1) is copied to HBM once; each search is one streaming pass.
 *   search  <- the calculate_distance storm under VpTreeManager.search_file (vptree.py:865-902):
 *              h_matched[v] = # query frames with >= 1 match in video v, so that
 *              matchHashBytes(query, video v) = 100.0 * h_matched[v] / n_query  (vptree.py:31). */
typedef struct vpdq_b200_db vpdq_b200_db;
VPDQ_B200_API int vpdq_b200_db_create(int device, const uint8_t* h_db, int64_t n_db, const int64_t* h_offsets, int64_t n_videos,
                        vpdq_b200_db** out);
VPDQ_B200_API int vpdq_b200_db_search(vpdq_b200_db* db, const uint8_t* h_query, int64_t n_query, int tolerance,
                        int32_t* h_matched /* [n_videos] */);
/* The same search, answered compactly: rows (0, video, matched, distance) for the stored videos with matched > 0 and
 * distance <= max_distance (= VpTreeManager.search_file's radius, vptree.py:865-902; <= 0: every video with a
 * match).  *n_rows = number found (VPDQ_B200_ERR_OVERFLOW if larger than cap).  One upload, one scan launch over all
 * 64-frame chunks of the query, the reduce on the device, one small read-back. */
VPDQ_B200_API int vpdq_b200_db_search_radius(vpdq_b200_db* db, const uint8_t* h_query, int64_t n_query, int tolerance,
                               int max_distance, int32_t* h_rows /* [cap][4] */, int64_t cap, int64_t* n_rows);
VPDQ_B200_API int vpdq_b200_db_destroy(vpdq_b200_db* db);

/* One-shot form of the above (creates, searches, destroys).  Any n_query (chunks of 64 inside). */
VPDQ_B200_API int vpdq_b200_search_host(const uint8_t* h_db, int64_t n_db, const int64_t* h_offsets, int64_t n_videos,
                          const uint8_t* h_query, int64_t n_query, int tolerance, int32_t* h_matched, int device);

#ifdef __cplusplus
}
#endif
#endif /* VPDQ_B200_H */
