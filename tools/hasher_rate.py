"""Frames/s through the reference-shaped streaming API: VideoHasher.hash_frame(bytes) per frame, then finish()."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np

from hydrus_video_deduplicator_b200 import vpdq

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
rng = np.random.default_rng(0)
frames = [rng.integers(0, 256, 512 * 512 * 3, dtype=np.uint8).tobytes() for _ in range(64)]
h = vpdq.VideoHasher(1, 512, 512, 0)
for k in range(64):
    h.hash_frame(frames[k])
h.finish()
h = vpdq.VideoHasher(1, 512, 512, 0)
t0 = time.perf_counter()
for k in range(n):
    h.hash_frame(frames[k & 63])
t1 = time.perf_counter()
ph = h.finish()
t2 = time.perf_counter()
print(f"hash_frame x {n}: {n / (t1 - t0):.0f} frames/s pushing, {n / (t2 - t0):.0f} frames/s incl. finish(); kept {len(ph)}")
