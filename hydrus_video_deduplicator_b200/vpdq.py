"""Drop-in for ``hvdaccelerators.vpdq`` -- the native module the reference imports at
vpdqpy/vpdqpy.py:9, db/vptree.py:9 and dedup.py:26 -- backed by libvpdq_b200.so (sm_100a CUDA).

Same names, argument meaning and error behaviour as the reference's call sites use
(SURVEY.md Appendix B):

    hasher = VideoHasher(average_fps, 512, 512, num_threads)     # vpdqpy.py:113
    hasher.hash_frame(rgb24_bytes)                                # vpdqpy.py:118 (may block)
    phash  = hasher.finish()                                      # vpdqpy.py:119 -> VpdqHash
    VpdqHash.from_string(s); str(phash); phash.bytes; len(phash)  # hashing.py:30,40; dedup.py:77
    VpdqHash.bytesPerPdqHash                                      # dedup.py:83-84
    matchHash(q, t, 31); matchHashBytes(a, b, 31)                 # vpdqpy.py:56; vptree.py:31
"""
from __future__ import annotations

import ctypes as C
from collections import deque

from . import _ffi

__all__ = ["VpdqHash", "VideoHasher", "matchHash", "matchHashBytes"]


class VpdqHash:
    """The per-video perceptual hash: the concatenated 32-byte PDQ hashes of the kept frames, in frame
    order, "native PDQ byte order" (DedupeDB.py:538-544).  Immutable value type."""

    bytesPerPdqHash = _ffi.HASH_BYTES  # dedup.py:83-84
    __slots__ = ("_b",)

    def __init__(self, data: bytes = b""):
        data = bytes(data)
        if len(data) % self.bytesPerPdqHash:
            raise ValueError(f"VpdqHash needs a multiple of {self.bytesPerPdqHash} bytes, got {len(data)}")
        self._b = data

    @staticmethod
    def from_string(s: str) -> "VpdqHash":
        """Inverse of str(): 64 hex characters per frame, no separators (hashing.py:40)."""
        s = s.strip()
        if len(s) % (2 * VpdqHash.bytesPerPdqHash):
            raise ValueError(f"VpdqHash string needs a multiple of 64 hex characters, got {len(s)}")
        return VpdqHash(bytes.fromhex(s))

    @staticmethod
    def from_bytes(b: bytes) -> "VpdqHash":
        return VpdqHash(b)

    @property
    def bytes(self) -> bytes:  # dedup.py:77
        return self._b

    def __str__(self) -> str:  # hashing.py:30
        return self._b.hex()

    def __repr__(self) -> str:
        return f"VpdqHash(frames={len(self)})"

    def __len__(self) -> int:  # number of frame hashes; only ever used as `> 0` (test_vpdqpy.py:95)
        return len(self._b) // self.bytesPerPdqHash

    def __eq__(self, other) -> bool:
        return isinstance(other, VpdqHash) and self._b == other._b

    def __ne__(self, other) -> bool:
        return not self.__eq__(other)

    def __hash__(self) -> int:
        return hash(self._b)


def _PUSH_NOCOPY(handle, frame: bytes, n: int) -> int:
    """vpdq_b200_hasher_push_nocopy with a bytes object as the source (ctypes passes the object's own buffer)"""
    return _ffi.lib().vpdq_b200_hasher_push_nocopy(handle, frame, n)


def _ptr(buf):
    """-> (void*, nbytes, keep-alive) for bytes / bytearray / memoryview / numpy; no copy when the
    buffer is contiguous (bytes are passed by pointer)."""
    if isinstance(buf, bytes):
        return C.cast(C.c_char_p(buf), C.c_void_p), len(buf), buf
    mv = memoryview(buf)
    if not mv.c_contiguous or mv.readonly:
        b = mv.tobytes()
        return C.cast(C.c_char_p(b), C.c_void_p), len(b), b
    arr = (C.c_uint8 * mv.nbytes).from_buffer(mv)
    return C.cast(arr, C.c_void_p), mv.nbytes, arr


class VideoHasher:
    """One video's streaming hasher (vpdqpy.py:113-119).  A handle is pure bookkeeping: every hasher of a device
    feeds that device's submission service (csrc/hash_service.h), which copies the frames into a shared pinned
    ring with a pool of copy threads, uploads them and hashes frames of ALL live hashers together;
    ``hash_frame`` blocks only when the ring is full (the reference's back-pressure, vpdqpy.py:115-117) and
    ``finish`` waits for this video's frames only.

    ``hash_frame(bytes)`` does not copy on the caller's thread: ``bytes`` is immutable, so the hasher keeps a
    reference until the copy workers are done with it (at most a ring's worth of frames) and returns at once.
    Mutable buffers (bytearray, numpy) are copied before the call returns."""

    def __init__(self, average_fps: int, width: int, height: int, num_threads: int = 0, *, device: int | None = None,
                 channels: int = 3):
        self._h = None
        self._held: deque = deque()  # (frame index, bytes object) still being read by the copy workers
        self._n = 0
        self._frame_bytes = int(width) * int(height) * int(channels)
        self.average_fps = int(average_fps)  # unused, as in the reference (vpdqpy.py:110-112)
        dev = _ffi.default_device() if device is None else int(device)
        h = C.c_void_p()
        _ffi.check(_ffi.lib().vpdq_b200_hasher_create(dev, int(width), int(height), int(channels), int(num_threads),
                                                      C.byref(h)))
        self._h = h

    def _handle(self):
        if self._h is None or not self._h.value:
            raise RuntimeError("VideoHasher is closed")
        return self._h

    def hash_frame(self, frame) -> None:
        """frame: ``width*height*3`` bytes, RGB24 row-major (vpdqpy.py:118)."""
        if type(frame) is bytes:  # immutable: hand the pointer over, keep the object alive, return at once
            if len(frame) != self._frame_bytes:
                raise ValueError(f"frame has {len(frame)} bytes, expected {self._frame_bytes}")
            rc = _PUSH_NOCOPY(self._handle(), frame, 1)
            if rc:
                _ffi.check(rc)
            self._n += 1
            self._held.append((self._n, frame))
            if not (self._n & 15):
                self._release_consumed()
            return
        p, n, _keep = _ptr(frame)
        if n != self._frame_bytes:
            raise ValueError(f"frame has {n} bytes, expected {self._frame_bytes}")
        _ffi.check(_ffi.lib().vpdq_b200_hasher_push(self._handle(), p, 1))
        self._n += 1

    def _release_consumed(self) -> None:
        done = C.c_int64(0)
        _ffi.lib().vpdq_b200_hasher_consumed(self._h, C.byref(done))
        held = self._held
        while held and held[0][0] <= done.value:
            held.popleft()

    def hash_frames(self, frames) -> None:
        """Batch form of hash_frame: one contiguous buffer holding a whole number of frames."""
        p, n, _keep = _ptr(frames)
        if n % self._frame_bytes:
            raise ValueError(f"buffer of {n} bytes is not a whole number of {self._frame_bytes}-byte frames")
        _ffi.check(_ffi.lib().vpdq_b200_hasher_push(self._handle(), p, n // self._frame_bytes))
        self._n += n // self._frame_bytes

    def finish(self, *, return_all: bool = False):
        """-> VpdqHash of the frames with quality >= 31, in push order (vpdqpy.py:119, DedupeDB.py:550-553).
        With return_all=True returns (VpdqHash, all_hashes: bytes, all_quality: list[int]) as well."""
        L = _ffi.lib()
        n = C.c_int64(0)
        _ffi.check(L.vpdq_b200_hasher_pushed(self._handle(), C.byref(n)))
        cap = max(1, n.value)
        out = C.create_string_buffer(cap * _ffi.HASH_BYTES)
        kept = C.c_int64(0)
        all_h = C.create_string_buffer(cap * _ffi.HASH_BYTES) if return_all else None
        all_q = (C.c_int32 * cap)() if return_all else None
        try:
            _ffi.check(L.vpdq_b200_hasher_finish(self._handle(), _ffi.QUALITY_KEEP, out, cap, C.byref(kept), all_h, all_q))
        finally:
            self._held.clear()  # finish() returns only when every frame has been consumed
            self._n = 0
        phash = VpdqHash(out.raw[: kept.value * _ffi.HASH_BYTES])
        if return_all:
            return phash, all_h.raw[: n.value * _ffi.HASH_BYTES], list(all_q[: n.value])
        return phash

    def close(self) -> None:
        h, self._h = self._h, None
        if h is not None and h.value:
            _ffi.lib().vpdq_b200_hasher_destroy(h)  # waits for frames still in flight
        self._held.clear()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def matchHashBytes(query: bytes, target: bytes, distance_tolerance: int = _ffi.DEFAULT_TOLERANCE, *,
                   device: int | None = None) -> float:
    """Percent (0.0..100.0) of query frames with at least one target frame at Hamming distance <=
    distance_tolerance; 0.0 if either side is empty (vptree.py:31; DedupeDB.py:555-557)."""
    query, target = bytes(query), bytes(target)
    if len(query) % _ffi.HASH_BYTES or len(target) % _ffi.HASH_BYTES:
        raise ValueError("hash blobs must be multiples of 32 bytes")
    sim = C.c_double(0.0)
    dev = _ffi.default_device() if device is None else int(device)
    _ffi.check(_ffi.lib().vpdq_b200_match_hash_host(query, len(query) // _ffi.HASH_BYTES, target,
                                                    len(target) // _ffi.HASH_BYTES, int(distance_tolerance),
                                                    C.byref(sim), dev))
    return float(sim.value)


def matchHash(query: VpdqHash, target: VpdqHash, distance_tolerance: int = _ffi.DEFAULT_TOLERANCE, *,
              device: int | None = None) -> float:
    """vpdq.matchHash (vpdqpy.py:56)."""
    return matchHashBytes(query.bytes, target.bytes, distance_tolerance, device=device)
