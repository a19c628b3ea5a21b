"""Brute-force replacement of the reference's vp-tree similarity search (``db/vptree.py``).

The reference walks a vp-tree and calls ``calculate_distance`` (= one Python->native ``matchHashBytes``
hop, vptree.py:29-31) per visited node; its "distance" ``101 - int(similarity)`` is not a metric, so that
traversal is lossy and randomised.  Here the whole phash table lives in HBM and one
streaming CUDA pass scores the query video against EVERY stored video, so the result is the exact set the
tree search approximates.

Reference interface mirrored (same names, argument meaning, result shape):
    fix_vpdq_similarity(similarity)                    vptree.py:22-25
    calculate_distance(phash_a, phash_b)               vptree.py:29-31
    HashIndex.search_file(hash_id, max_hamming_distance)             vptree.py:865-902
    HashIndex.search_perceptual_hashes(phashes, max_hamming_distance) vptree.py:664-815
"""
from __future__ import annotations

import ctypes as C
from typing import Collection, Iterable, Sequence

import numpy as np

from . import _ffi, vpdq


def fix_vpdq_similarity(similarity: float) -> int:
    """Turn [100.0, 0.0] similarity to [1, 101] (vptree.py:22-25)."""
    return (100 - int(similarity)) + 1


def calculate_distance(phash_a: bytes, phash_b: bytes) -> int:
    """Distance between two perceptual hashes, in [1, 101] (vptree.py:29-31)."""
    return fix_vpdq_similarity(vpdq.matchHashBytes(phash_a, phash_b, 31))


def dedupe_list(xs: Iterable) -> list:
    """Order-preserving de-duplication (vptree.py:107-123)."""
    seen, out = set(), []
    for x in xs:
        if x not in seen:
            out.append(x)
            seen.add(x)
    return out


class HashIndex:
    """The phash table (``shape_perceptual_hashes`` x ``shape_perceptual_hash_map``, DedupeDB.py:159-172)
    resident on one GPU: all frame hashes concatenated as [n_frames][32] bytes plus CSR video offsets."""

    def __init__(self, hash_ids: Sequence[int], phashes: Sequence[bytes], *, device: int | None = None):
        if len(hash_ids) != len(phashes):
            raise ValueError("hash_ids and phashes must have the same length")
        self.hash_ids = [int(h) for h in hash_ids]
        self._row = {h: i for i, h in enumerate(self.hash_ids)}
        if len(self._row) != len(self.hash_ids):
            raise ValueError("hash_ids must be unique")
        self._phashes = [bytes(p) for p in phashes]
        for p in self._phashes:
            if len(p) % _ffi.HASH_BYTES:
                raise ValueError("every phash must be a multiple of 32 bytes")
        self.offsets = np.zeros(len(self._phashes) + 1, dtype=np.int64)
        np.cumsum([len(p) // _ffi.HASH_BYTES for p in self._phashes], out=self.offsets[1:])
        self.n_frames = int(self.offsets[-1])
        blob = b"".join(self._phashes)
        self._db = C.c_void_p()
        dev = _ffi.default_device() if device is None else int(device)
        _ffi.check(_ffi.lib().vpdq_b200_db_create(dev, blob, self.n_frames, self.offsets.ctypes.data_as(C.c_void_p),
                                                  len(self._phashes), C.byref(self._db)))

    def __len__(self) -> int:
        return len(self.hash_ids)

    def close(self) -> None:
        db, self._db = self._db, None
        if db is not None and db.value:
            _ffi.lib().vpdq_b200_db_destroy(db)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------------
    def matched_frames(self, phash: bytes, distance_tolerance: int = 31) -> np.ndarray:
        """[n_videos] int32: number of the query's frames that have a match in each stored video."""
        phash = bytes(phash)
        if len(phash) % _ffi.HASH_BYTES:
            raise ValueError("phash must be a multiple of 32 bytes")
        out = np.zeros(max(1, len(self)), dtype=np.int32)
        _ffi.check(_ffi.lib().vpdq_b200_db_search(self._db, phash, len(phash) // _ffi.HASH_BYTES,
                                                  int(distance_tolerance), out.ctypes.data_as(C.c_void_p)))
        return out[: len(self)]

    def similarities(self, phash: bytes, distance_tolerance: int = 31) -> np.ndarray:
        """[n_videos] float64 == matchHashBytes(phash, stored_video, tol) for every stored video."""
        n_q = len(phash) // _ffi.HASH_BYTES
        m = self.matched_frames(phash, distance_tolerance).astype(np.float64)
        if n_q == 0:
            return np.zeros(len(self))
        sim = (100.0 * m) / float(n_q)
        sim[np.diff(self.offsets) == 0] = 0.0  # empty stored hash: similar to nothing
        return sim

    def within_radius(self, phash: bytes, max_distance: int, distance_tolerance: int = 31) -> list[tuple[int, int]]:
        """[(row, distance)] of the stored videos with calculate_distance(phash, video) <= max_distance (<= 100), by
        ascending row.  The scan over all 64-frame chunks of the query, the video-level reduce and the radius filter run
        on the device; only the hits come back."""
        phash = bytes(phash)
        if len(phash) % _ffi.HASH_BYTES:
            raise ValueError("phash must be a multiple of 32 bytes")
        n_q = len(phash) // _ffi.HASH_BYTES
        if n_q == 0 or len(self) == 0:
            return []
        rows = np.zeros((len(self), 4), dtype=np.int32)
        n = C.c_int64(0)
        _ffi.check(_ffi.lib().vpdq_b200_db_search_radius(self._db, phash, n_q, int(distance_tolerance), int(max_distance),
                                                         rows.ctypes.data_as(C.c_void_p), len(self), C.byref(n)))
        rows = rows[: n.value]
        order = np.argsort(rows[:, 1], kind="stable")
        return [(int(r[1]), int(r[3])) for r in rows[order]]

    def distances(self, phash: bytes) -> np.ndarray:
        """[n_videos] int64 == calculate_distance(phash, stored_video) (vptree.py:29-31)."""
        return (100 - self.similarities(phash).astype(np.int64)) + 1

    def search_perceptual_hashes(self, search_perceptual_hashes: Collection[bytes],
                                 max_hamming_distance: int) -> list[tuple[int, int]]:
        """All (hash_id, distance) with distance <= max_hamming_distance, smallest distance per file
        (vptree.py:664-815).  max_hamming_distance == 0 means byte-identical phashes (vptree.py:671-690)."""
        best: dict[int, int] = {}
        for phash in search_perceptual_hashes:
            phash = bytes(phash)
            if max_hamming_distance == 0:
                for hid, stored in zip(self.hash_ids, self._phashes):
                    if stored == phash:
                        best[hid] = 0
                continue
            if max_hamming_distance >= 101:  # even a video without a single matching frame (distance 101) is inside
                d = self.distances(phash)
                hits = [(int(row), int(d[row])) for row in np.nonzero(d <= max_hamming_distance)[0]]
            else:
                hits = self.within_radius(phash, max_hamming_distance)
            for row, dist in hits:
                hid = self.hash_ids[row]
                if hid not in best or dist < best[hid]:
                    best[hid] = dist
        return dedupe_list(best.items())

    def search_file(self, hash_id: int, max_hamming_distance: int) -> list[tuple[int, int]]:
        """vptree.py:865-902: (hash_id, 0) first, then everything within the radius of this file's phash."""
        result = [(hash_id, 0)]
        phash = self._phashes[self._row[hash_id]]
        result.extend(self.search_perceptual_hashes([phash], max_hamming_distance))
        return dedupe_list(result)

    def find_potential_duplicates(self, threshold: float = 50.0) -> list[tuple[int, int, int]]:
        """The search loop of dedup.py:445-502 without the Hydrus POSTs: every directed (a, b, distance)
        with a != b and distance <= fix_vpdq_similarity(threshold).  len(result) // 2 is the reference's
        return value when the relation is symmetric."""
        radius = fix_vpdq_similarity(threshold)
        out = []
        for hid in self.hash_ids:
            for other, dist in self.search_file(hid, radius):
                if other != hid:
                    out.append((hid, other, dist))
        return out
