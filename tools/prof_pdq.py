"""Launch the PDQ pipeline a few times on synthetic frames (for ncu): python tools/prof_pdq.py [n_frames]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from bench import device_frames
from hydrus_video_deduplicator_b200 import device

n = int(sys.argv[-1]) if len(sys.argv) > 1 and sys.argv[-1].isdigit() else 4096
dev = torch.device("cuda", 0)
frames = device_frames(torch, n, dev, seed=5)
for _ in range(3):
    h, q = device.hash_frames(frames)
torch.cuda.synchronize()
print("ok", n, int(h.sum()))
