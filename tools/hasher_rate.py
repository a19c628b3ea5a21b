"""Frames/s through the reference-shaped streaming API (vpdq.VideoHasher.hash_frame(bytes) per frame, one hasher per
video, finish() per video) -- the same measurement bench.py reports under e2e.video_hasher_api, on its own:
  [VPDQ_B200_COPY_THREADS=..] [VPDQ_B200_LAUNCH_MIN=..] python tools/hasher_rate.py"""
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import bench
from hydrus_video_deduplicator_b200 import device as dev_api

dev = torch.device("cuda", 0)
pool0 = bench.device_frames(torch, 128, dev, seed=1000)


class A:
    pass


out = bench.hasher_api_section(torch, dev_api, pool0, 0, torch.cuda.synchronize, lambda x: x, 1, A())
env = {k: v for k, v in os.environ.items() if k.startswith("VPDQ_B200_")}
print(json.dumps({"env": env, **{k: (round(v["frames/s"]), v["bit_identical_to_device_path"],
                                     {a: round(b, 3) for a, b in v["service"].items()}) for k, v in out.items()
                                 if isinstance(v, dict)}}))
