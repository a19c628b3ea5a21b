"""Device-resident entry points: torch CUDA tensors in, torch CUDA tensors out.

torch is used here only as the device-memory container and stream provider (and by ``dist.py`` for
NCCL); every operation is one call through the C ABI (``*_dev`` functions of include/vpdq_b200.h) on the
caller's current CUDA stream.  Nothing here synchronises unless it has to return a Python number.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _ffi


def _stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(t: torch.Tensor, name: str, dtype=None) -> torch.Tensor:
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (there is no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise ValueError(f"{name} must have dtype {dtype}, got {t.dtype}")
    return t.contiguous()


def pdq_scratch(n_frames: int, device: torch.device) -> torch.Tensor:
    """Scratch for one hash_frames call: the decimated 64x64 fp32 plane, 16 KB per frame.  Allocated per call through
    torch's stream-ordered caching allocator (no cudaMalloc in steady state), so that calls on different streams or
    threads never share a buffer and a buffer is not reused before the kernels that read it have run."""
    need = C.c_size_t(0)
    _ffi.check(_ffi.lib().vpdq_b200_pdq_scratch_bytes(int(n_frames), C.byref(need)))
    return torch.empty(need.value, dtype=torch.uint8, device=device)


def hash_frames(frames: torch.Tensor, *, stages: bool = False):
    """frames: [n, 512, 512, 3] (RGB24) or [n, 512, 512] (gray == R=G=B) uint8 CUDA tensor ->
    (hashes [n, 32] uint8, quality [n] int32) on the same device, unfiltered.
    stages=True also returns (A [n, 64, 64] f32 decimated plane, B [n, 16, 16] f32 DCT)."""
    frames = _need_cuda(frames, "frames", torch.uint8)
    if frames.dim() == 4 and frames.shape[3] == 3:
        ch = 3
    elif frames.dim() == 3:
        ch = 1
    else:
        raise ValueError("frames must be [n, 512, 512, 3] or [n, 512, 512]")
    n, h, w = frames.shape[:3]
    dev = frames.device
    hashes = torch.empty((n, 32), dtype=torch.uint8, device=dev)
    quality = torch.empty((n,), dtype=torch.int32, device=dev)
    a64 = torch.empty((n, 64, 64), dtype=torch.float32, device=dev) if stages else None
    b16 = torch.empty((n, 16, 16), dtype=torch.float32, device=dev) if stages else None
    with torch.cuda.device(dev):
        scratch = pdq_scratch(n, dev)
        _ffi.check(_ffi.lib().vpdq_b200_pdq_stages_dev(
            frames.data_ptr(), ch, n, w, h, hashes.data_ptr(), quality.data_ptr(),
            a64.data_ptr() if stages else None, b16.data_ptr() if stages else None,
            scratch.data_ptr(), scratch.numel(), _stream_ptr()))
    if stages:
        return hashes, quality, a64, b16
    return hashes, quality


def jarosz_planes(frames: torch.Tensor) -> torch.Tensor:
    """[n, 512, 512, 3] uint8 CUDA -> [n, 64, 64] f32: the Jarosz-filtered, decimated luma plane (the fused kernel
    on its own; hash_frames = this + the finalize kernel)."""
    frames = _need_cuda(frames, "frames", torch.uint8)
    if frames.dim() != 4 or frames.shape[1:] != (512, 512, 3):
        raise ValueError("frames must be [n, 512, 512, 3]")
    out = torch.empty((frames.shape[0], 64, 64), dtype=torch.float32, device=frames.device)
    with torch.cuda.device(frames.device):
        _ffi.check(_ffi.lib().vpdq_b200_pdq_jarosz_dev(frames.data_ptr(), frames.shape[0], 512, 512, out.data_ptr(),
                                                       _stream_ptr()))
    return out


def point_resize(frames: torch.Tensor) -> torch.Tensor:
    """[n, H, W, 3] uint8 CUDA (natively sized decoded frames) -> [n, 512, 512, 3]: the reference's
    frame.reformat(512, 512, "rgb24", POINT) (vpdqpy.py:90-95) on the device."""
    frames = _need_cuda(frames, "frames", torch.uint8)
    if frames.dim() != 4 or frames.shape[3] != 3:
        raise ValueError("frames must be [n, H, W, 3]")
    n, h, w = frames.shape[:3]
    out = torch.empty((n, 512, 512, 3), dtype=torch.uint8, device=frames.device)
    with torch.cuda.device(frames.device):
        _ffi.check(_ffi.lib().vpdq_b200_point_resize_dev(frames.data_ptr(), n, h, w, out.data_ptr(), _stream_ptr()))
    return out


def hash_native_frames(frames: torch.Tensor):
    """POINT-resize + hash natively sized RGB frames without leaving the device."""
    return hash_frames(point_resize(frames))


def _as_hash_matrix(t: torch.Tensor, name: str) -> torch.Tensor:
    """[n, 32] uint8 or [n, 4] int64 -> contiguous CUDA tensor, 32-byte rows"""
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (there is no CPU path)")
    if t.dtype == torch.uint8 and t.dim() == 2 and t.shape[1] == 32:
        return t.contiguous()
    if t.dtype == torch.int64 and t.dim() == 2 and t.shape[1] == 4:
        return t.contiguous()
    raise ValueError(f"{name} must be [n, 32] uint8 or [n, 4] int64")


def hamming_scan(db: torch.Tensor, query: torch.Tensor, offsets: torch.Tensor | None = None, tolerance: int = 31,
                 *, reverse_counts: bool = False):
    """One streaming pass of <= 64 query frames over the database.
    -> qmask [n_videos] int64 (bit i = query frame i matched in video v) and, if reverse_counts,
       tcount [n_videos] int32 (# frames of video v that matched some query frame)."""
    db = _as_hash_matrix(db, "db")
    query = _as_hash_matrix(query, "query")
    n_db, n_q = db.shape[0], query.shape[0]
    if n_q > 64:
        raise ValueError("at most 64 query frames per scan; chunk the query")
    if offsets is not None:
        offsets = _need_cuda(offsets, "offsets", torch.int64)
        n_videos = offsets.numel() - 1
    else:
        n_videos = n_db
    dev = db.device
    qmask = torch.zeros((n_videos,), dtype=torch.int64, device=dev)
    tcount = torch.zeros((n_videos,), dtype=torch.int32, device=dev) if reverse_counts else None
    with torch.cuda.device(dev):
        _ffi.check(_ffi.lib().vpdq_b200_hamming_scan_dev(
            db.data_ptr(), n_db, offsets.data_ptr() if offsets is not None else None, n_videos, query.data_ptr(),
            n_q, int(tolerance), qmask.data_ptr(), tcount.data_ptr() if reverse_counts else None, _stream_ptr()))
    return (qmask, tcount) if reverse_counts else qmask


def chunk_rows(q_offsets):
    """Cut query videos (CSR q_offsets, host int array) into scan chunks of <= 64 frames.
    -> (chunk_rows int32 [n_chunks + 1], qv_chunks int32 [n_qvideos + 1]): chunk c = query rows chunk_rows[c] ..
    chunk_rows[c+1]-1; video q owns chunks qv_chunks[q] .. qv_chunks[q+1]-1 (an empty video owns none)."""
    import numpy as np

    q_offsets = np.asarray(q_offsets, dtype=np.int64)
    n = np.diff(q_offsets)
    per_video = (n + 63) // 64
    qv_chunks = np.zeros(len(n) + 1, dtype=np.int64)
    np.cumsum(per_video, out=qv_chunks[1:])
    video_of_chunk = np.repeat(np.arange(len(n)), per_video)
    k = np.arange(int(qv_chunks[-1])) - qv_chunks[video_of_chunk]           # chunk index inside its video
    starts = q_offsets[video_of_chunk] + 64 * k
    rows = np.empty(len(starts) + 1, dtype=np.int64)
    rows[:-1] = starts
    rows[-1] = q_offsets[-1]
    return rows.astype(np.int32), qv_chunks.astype(np.int32)


def video_matches(db: torch.Tensor, offsets: torch.Tensor, queries: torch.Tensor, q_offsets, tolerance: int = 31, *,
                  max_distance: int = 0, dense: bool = False, capacity: int | None = None):
    """Score MANY query videos against every video of a database in one scan launch + one reduce launch:
    matchHash / calculate_distance (vpdqpy.py:56, db/vptree.py:22-31) for all (query video, target video) pairs.
      db [n, 32] uint8 CUDA + offsets [V + 1] int64 CUDA (CSR); queries [m, 32] uint8 CUDA + q_offsets (host ints, CSR,
      videos contiguous from row 0)
    -> rows [k, 4] int32 CUDA (query video, target video, matched query frames, distance) for the pairs with a match
       (and distance <= max_distance if > 0), unordered; with dense=True instead matched [n_qvideos, V] int32."""
    import numpy as np

    db = _as_hash_matrix(db, "db")
    queries = _as_hash_matrix(queries, "queries")
    offsets = _need_cuda(offsets, "offsets", torch.int64)
    dev = db.device
    n_videos = offsets.numel() - 1
    rows_h, qv_chunks_h = chunk_rows(q_offsets)
    n_chunks, n_qv = len(rows_h) - 1, len(qv_chunks_h) - 1
    frames_h = np.diff(np.asarray(q_offsets, dtype=np.int64)).astype(np.int32)
    if n_videos == 0 or n_qv == 0 or n_chunks == 0 or db.shape[0] == 0:
        if dense:
            return torch.zeros((n_qv, n_videos), dtype=torch.int32, device=dev)
        return torch.zeros((0, 4), dtype=torch.int32, device=dev)
    meta = torch.from_numpy(np.concatenate([rows_h, qv_chunks_h, frames_h])).to(dev)
    d_rows, d_qv_chunks, d_frames = meta[: n_chunks + 1], meta[n_chunks + 1: n_chunks + 2 + n_qv], meta[n_chunks + 2 + n_qv:]
    qmask = torch.zeros((n_chunks, n_videos), dtype=torch.int64, device=dev)
    L = _ffi.lib()
    with torch.cuda.device(dev):
        _ffi.check(L.vpdq_b200_hamming_scan_multi_dev(db.data_ptr(), db.shape[0], offsets.data_ptr(), n_videos,
                                                      queries.data_ptr(), d_rows.data_ptr(), n_chunks, int(tolerance),
                                                      qmask.data_ptr(), _stream_ptr()))
        if dense:
            matched = torch.empty((n_qv, n_videos), dtype=torch.int32, device=dev)
            _ffi.check(L.vpdq_b200_video_match_dev(qmask.data_ptr(), n_videos, d_qv_chunks.data_ptr(), d_frames.data_ptr(),
                                                   n_qv, 0, matched.data_ptr(), None, 0, None, _stream_ptr()))
            return matched
        cap = int(capacity) if capacity is not None else max(1024, 4 * n_qv)
        while True:
            out = torch.empty((cap, 4), dtype=torch.int32, device=dev)
            count = torch.zeros((1,), dtype=torch.int64, device=dev)
            _ffi.check(L.vpdq_b200_video_match_dev(qmask.data_ptr(), n_videos, d_qv_chunks.data_ptr(), d_frames.data_ptr(),
                                                   n_qv, int(max_distance), None, out.data_ptr(), cap, count.data_ptr(),
                                                   _stream_ptr()))
            n = int(count.item())
            if n <= cap:
                return out[:n]
            cap = n  # the list overflowed: rerun the (cheap) reduce with room for everything


def hamming_pairs(q: torch.Tensor, t: torch.Tensor, tolerance: int = 31, *, skip_diagonal: bool = False,
                  capacity: int = 1 << 20, want_bitmap: bool = True):
    """Brute-force all pairs.  -> (count: int, pairs [min(count, capacity), 2] int64 (unordered),
    any_bitmap [(n_q+31)//32] int32 or None).  Synchronises to read the count."""
    q = _as_hash_matrix(q, "q")
    t = _as_hash_matrix(t, "t")
    dev = q.device
    n_q, n_t = q.shape[0], t.shape[0]
    count = torch.zeros((1,), dtype=torch.int64, device=dev)
    pairs = torch.empty((max(1, capacity),), dtype=torch.int64, device=dev)
    bitmap = torch.zeros(((n_q + 31) // 32,), dtype=torch.int32, device=dev) if want_bitmap else None
    with torch.cuda.device(dev):
        _ffi.check(_ffi.lib().vpdq_b200_hamming_pairs_dev(
            q.data_ptr(), n_q, t.data_ptr(), n_t, int(tolerance), int(bool(skip_diagonal)),
            bitmap.data_ptr() if want_bitmap else None, pairs.data_ptr(), int(capacity), count.data_ptr(),
            _stream_ptr()))
    n = int(count.item())
    packed = pairs[: min(n, capacity)]
    out = torch.stack(((packed >> 32) & 0xFFFFFFFF, packed & 0xFFFFFFFF), dim=1)
    return n, out, bitmap
