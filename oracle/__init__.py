"""ctypes front-end of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY.

Importers allowed by the project rules: tests/, __graft_entry__.smoke(), and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.

Function-by-function reference citations are in pdq_oracle.c / match_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "liboracle.so"
_lib = None


def build(force: bool = False) -> Path:
    """Compile oracle/*.c with gcc (oracle/Makefile)."""
    if force or not _LIB_PATH.exists() or any(
        (_HERE / s).stat().st_mtime > _LIB_PATH.stat().st_mtime for s in ("pdq_oracle.c", "match_oracle.c", "Makefile")
    ):
        subprocess.run(["make", "-C", str(_HERE), "-B", "liboracle.so"], check=True, capture_output=True)
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_LIB_PATH))
        u8p, i32p, f32p, i64p = (C.POINTER(t) for t in (C.c_uint8, C.c_int32, C.c_float, C.c_int64))
        L.oracle_dct_matrix.restype = f32p
        L.oracle_pdq_hash_rgb.argtypes = [u8p, C.c_int, C.c_int, u8p, i32p]
        L.oracle_pdq_hash_gray.argtypes = [u8p, C.c_int, C.c_int, u8p, i32p]
        L.oracle_pdq_stages_rgb.argtypes = [u8p, C.c_int, C.c_int, u8p, i32p, f32p, f32p]
        L.oracle_pdq_hash_batch.argtypes = [u8p, C.c_int, C.c_long, C.c_int, C.c_int, u8p, i32p, C.c_int]
        L.oracle_hamming256.argtypes = [u8p, u8p]
        L.oracle_matched_frames.argtypes = [u8p, C.c_long, u8p, C.c_long, C.c_int]
        L.oracle_matched_frames.restype = C.c_long
        L.oracle_match_hash.argtypes = [u8p, C.c_long, u8p, C.c_long, C.c_int]
        L.oracle_match_hash.restype = C.c_double
        L.oracle_calculate_distance.argtypes = [u8p, C.c_long, u8p, C.c_long]
        L.oracle_hamming_pairs.argtypes = [u8p, C.c_long, u8p, C.c_long, C.c_int, i64p, C.c_long]
        L.oracle_hamming_pairs.restype = C.c_long
        L.oracle_video_matched.argtypes = [u8p, C.c_long, u8p, i64p, C.c_long, C.c_int, i32p]
        L.oracle_video_matched.restype = None
        L.oracle_hamming_count_mt.argtypes = [u8p, C.c_long, u8p, C.c_long, C.c_int, C.c_int]
        L.oracle_hamming_count_mt.restype = C.c_long
        _lib = L
    return _lib


def _u8(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _as_hashes(x) -> np.ndarray:
    """bytes | ndarray -> contiguous [n, 32] uint8"""
    if isinstance(x, (bytes, bytearray, memoryview)):
        x = np.frombuffer(bytes(x), dtype=np.uint8)
    a = np.ascontiguousarray(x, dtype=np.uint8).reshape(-1, 32)
    return a


def dct_matrix() -> np.ndarray:
    p = lib().oracle_dct_matrix()
    return np.ctypeslib.as_array(p, shape=(16, 64)).copy()


def pdq_hash_frames(frames: np.ndarray, nthreads: int = 1) -> tuple[np.ndarray, np.ndarray]:
    """frames: [n, H, W, 3] (RGB24) or [n, H, W] (gray, defined as R=G=B) uint8 ->
    (hashes [n, 32] uint8, quality [n] int32).  Restates VideoHasher.hash_frame (vpdqpy.py:118)."""
    frames = np.ascontiguousarray(frames, dtype=np.uint8)
    if frames.ndim == 4 and frames.shape[3] == 3:
        ch = 3
    elif frames.ndim == 3:
        ch = 1
    else:
        raise ValueError("frames must be [n,H,W,3] or [n,H,W]")
    n, h, w = frames.shape[:3]
    hashes = np.zeros((n, 32), dtype=np.uint8)
    quality = np.zeros(n, dtype=np.int32)
    rc = lib().oracle_pdq_hash_batch(_u8(frames), ch, n, h, w, _u8(hashes),
                                     quality.ctypes.data_as(C.POINTER(C.c_int32)), int(nthreads))
    if rc:
        raise RuntimeError(f"oracle_pdq_hash_batch rc={rc}")
    return hashes, quality


def pdq_stages(frame: np.ndarray):
    """one RGB frame -> (hash[32], quality, A[64,64] decimated plane, B[16,16] DCT)"""
    frame = np.ascontiguousarray(frame, dtype=np.uint8)
    h, w = frame.shape[:2]
    hsh = np.zeros(32, np.uint8)
    q = C.c_int32(0)
    a = np.zeros((64, 64), np.float32)
    b = np.zeros((16, 16), np.float32)
    f32p = C.POINTER(C.c_float)
    rc = lib().oracle_pdq_stages_rgb(_u8(frame), h, w, _u8(hsh), C.byref(q), a.ctypes.data_as(f32p),
                                     b.ctypes.data_as(f32p))
    if rc:
        raise RuntimeError(f"oracle_pdq_stages_rgb rc={rc}")
    return hsh, int(q.value), a, b


QUALITY_THRESHOLD = 31  # DedupeDB.py:550-553: keep a frame iff quality >= 31 (unpinned)


def video_hash(frames: np.ndarray, nthreads: int = 1) -> bytes:
    """VideoHasher(...).hash_frame()* .finish() (vpdqpy.py:113-119): kept frames' hashes, in order."""
    hashes, quality = pdq_hash_frames(frames, nthreads)
    return hashes[quality >= QUALITY_THRESHOLD].tobytes()


def match_hash(q, t, tol: int = 31) -> float:
    """vpdq.matchHash / matchHashBytes (vpdqpy.py:56, vptree.py:31)."""
    qa, ta = _as_hashes(q), _as_hashes(t)
    return float(lib().oracle_match_hash(_u8(qa), len(qa), _u8(ta), len(ta), int(tol)))


def is_similar(a, b, threshold: float = 75.0) -> tuple[bool, float]:
    """Vpdq.is_similar (vpdqpy.py:122-131)."""
    s = match_hash(a, b, 31)
    return s >= threshold, s


def calculate_distance(a, b) -> int:
    """vptree.calculate_distance (vptree.py:29-31)."""
    qa, ta = _as_hashes(a), _as_hashes(b)
    return int(lib().oracle_calculate_distance(_u8(qa), len(qa), _u8(ta), len(ta)))


def hamming_pairs(q, t, tol: int = 31, cap: int | None = None) -> np.ndarray:
    """all ordered (i, j) with popcount(q_i ^ t_j) <= tol -> [npairs, 2] int64, row-major order"""
    qa, ta = _as_hashes(q), _as_hashes(t)
    cap = int(cap if cap is not None else max(1024, 4 * (len(qa) + len(ta))))
    while True:
        out = np.zeros((cap, 2), np.int64)
        n = lib().oracle_hamming_pairs(_u8(qa), len(qa), _u8(ta), len(ta), int(tol),
                                       out.ctypes.data_as(C.POINTER(C.c_int64)), cap)
        if n <= cap:
            return out[:n]
        cap = int(n)


def video_matched(q, t, offsets, tol: int = 31) -> np.ndarray:
    """per DB video: number of query frames with a match in it (the numerator of matchHash)."""
    qa, ta = _as_hashes(q), _as_hashes(t)
    off = np.ascontiguousarray(offsets, dtype=np.int64)
    out = np.zeros(len(off) - 1, np.int32)
    lib().oracle_video_matched(_u8(qa), len(qa), _u8(ta), off.ctypes.data_as(C.POINTER(C.c_int64)), len(off) - 1,
                               int(tol), out.ctypes.data_as(C.POINTER(C.c_int32)))
    return out


def search_file(db_videos: list[bytes], query_index: int, radius: int) -> list[tuple[int, int]]:
    """Brute-force statement of VpTreeManager.search_file (vptree.py:865-902): (self, 0) first, then every
    other video whose calculate_distance to the query is <= radius; the vp-tree returns a subset of this."""
    out = [(query_index, 0)]
    q = db_videos[query_index]
    for v, t in enumerate(db_videos):
        if v == query_index:
            continue
        d = calculate_distance(q, t)
        if d <= radius:
            out.append((v, d))
    return out


def hamming_count_mt(q, t, tol: int = 31, nthreads: int | None = None) -> int:
    qa, ta = _as_hashes(q), _as_hashes(t)
    return int(lib().oracle_hamming_count_mt(_u8(qa), len(qa), _u8(ta), len(ta), int(tol),
                                             int(nthreads or os.cpu_count() or 1)))
