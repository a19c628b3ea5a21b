"""Seeded synthetic inputs shared by the parity tests, bench.py and smoke().

Frames: one third uniform noise, one third smooth (8x8 random grid, bilinear upsample -- the
order-sensitive case for the running-sum box filter), one third blocks (32x32 random, nearest).
Hashes: random 256-bit words of popcount exactly 128 (like genuine PDQ hashes) with planted
near-duplicates at even distances 0..40 so the tolerance boundary (31) is exercised from both sides.
"""
from __future__ import annotations

import numpy as np

DIM = 512


def _bilinear_up(grid: np.ndarray, size: int) -> np.ndarray:
    """[g, g, c] float -> [size, size, c] float, separable linear interpolation"""
    g = grid.shape[0]
    pos = (np.arange(size) + 0.5) * g / size - 0.5
    i0 = np.clip(np.floor(pos).astype(int), 0, g - 1)
    i1 = np.clip(i0 + 1, 0, g - 1)
    w = np.clip(pos - np.floor(pos), 0.0, 1.0)
    rows = grid[i0] * (1 - w)[:, None, None] + grid[i1] * w[:, None, None]
    return rows[:, i0] * (1 - w)[None, :, None] + rows[:, i1] * w[None, :, None]


def synth_frame(rng: np.random.Generator, kind: int, channels: int = 3) -> np.ndarray:
    if kind == 0:
        f = rng.integers(0, 256, size=(DIM, DIM, channels), dtype=np.uint8)
    elif kind == 1:
        grid = rng.uniform(0, 255, size=(8, 8, channels))
        f = np.clip(np.rint(_bilinear_up(grid, DIM)), 0, 255).astype(np.uint8)
    else:
        small = rng.integers(0, 256, size=(32, 32, channels), dtype=np.uint8)
        f = np.repeat(np.repeat(small, DIM // 32, axis=0), DIM // 32, axis=1)
    return f if channels == 3 else f[:, :, 0]


def synth_frames(n: int, seed: int = 0, channels: int = 3) -> np.ndarray:
    """[n, 512, 512, 3] (or [n, 512, 512] for channels=1) uint8; frame k is of kind k % 3"""
    rng = np.random.default_rng(seed)
    shape = (n, DIM, DIM, 3) if channels == 3 else (n, DIM, DIM)
    out = np.empty(shape, dtype=np.uint8)
    for k in range(n):
        out[k] = synth_frame(rng, k % 3, channels)
    return out


def noisy_copy(frames: np.ndarray, seed: int, amp: int = 2) -> np.ndarray:
    rng = np.random.default_rng(seed)
    d = rng.integers(-amp, amp + 1, size=frames.shape, dtype=np.int16)
    return np.clip(frames.astype(np.int16) + d, 0, 255).astype(np.uint8)


def random_balanced_hashes(n: int, rng: np.random.Generator) -> np.ndarray:
    """[n, 32] uint8, every row has exactly 128 bits set"""
    keys = rng.random((n, 256))
    idx = np.argsort(keys, axis=1)[:, :128]
    bits = np.zeros((n, 256), dtype=np.uint8)
    np.put_along_axis(bits, idx, 1, axis=1)
    return np.packbits(bits, axis=1, bitorder="little")


def flip_bits(h: np.ndarray, d: int, rng: np.random.Generator) -> np.ndarray:
    """copy of one hash [32] u8 at Hamming distance exactly d, popcount preserved when d is even"""
    bits = np.unpackbits(h, bitorder="little").copy()
    ones = np.flatnonzero(bits == 1)
    zeros = np.flatnonzero(bits == 0)
    k1 = d // 2
    k0 = d - k1
    bits[rng.choice(ones, size=k1, replace=False)] = 0
    bits[rng.choice(zeros, size=k0, replace=False)] = 1
    return np.packbits(bits, bitorder="little")


def synth_hashes(n: int, seed: int = 1, planted_frac: float = 0.01) -> np.ndarray:
    """[n, 32] uint8: random balanced hashes, a planted_frac share of rows being near-duplicates of an
    earlier row at distances cycling through 0, 2, ..., 40."""
    rng = np.random.default_rng(seed)
    h = random_balanced_hashes(n, rng)
    n_plant = int(n * planted_frac)
    if n_plant and n > 1:
        dst = rng.choice(np.arange(1, n), size=min(n_plant, n - 1), replace=False)
        for k, j in enumerate(dst):
            src = int(rng.integers(0, j))
            h[j] = flip_bits(h[src], (2 * k) % 42, rng)
    return h


def synth_video_db(n_videos: int, frames_per_video: int, seed: int = 2, dup_frac: float = 0.2):
    """A CSR hash DB: list of phash blobs where a dup_frac share of videos are noisy copies (per-frame
    distances 0..34) of an earlier video.  -> (list[bytes], offsets int64[n_videos+1])"""
    rng = np.random.default_rng(seed)
    vids: list[np.ndarray] = []
    for v in range(n_videos):
        nf = frames_per_video if frames_per_video > 0 else int(rng.integers(0, 12))
        if v > 0 and rng.random() < dup_frac:
            src = vids[int(rng.integers(0, v))]
            if len(src):
                take = src[rng.integers(0, len(src), size=nf)] if nf else src[:0]
                vid = np.stack([flip_bits(f, int(rng.integers(0, 18)) * 2, rng) for f in take]) if nf else take
                vids.append(vid.reshape(-1, 32))
                continue
        vids.append(random_balanced_hashes(nf, rng).reshape(-1, 32))
    offsets = np.zeros(n_videos + 1, dtype=np.int64)
    np.cumsum([len(v) for v in vids], out=offsets[1:])
    return [v.tobytes() for v in vids], offsets
