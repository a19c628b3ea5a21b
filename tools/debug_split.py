import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import oracle
from bench import device_frames
from hydrus_video_deduplicator_b200 import device
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dev = torch.device("cuda", 0)
frames = device_frames(torch, n, dev, seed=0)
h1, q1, a1, _ = device.hash_frames(frames, stages=True)
h2, q2, a2, _ = device.hash_frames(frames, stages=True)
print("repeat identical:", torch.equal(h1, h2), torch.equal(a1, a2))
ha, qa, aa, _ = device.hash_frames(frames[:1], stages=True)
hb, qb, ab, _ = device.hash_frames(frames[1:], stages=True)
hs = torch.cat([ha, hb]); as_ = torch.cat([aa, ab])
bad = torch.nonzero((h1 != hs).any(dim=1)).flatten().cpu().numpy()
badA = torch.nonzero((a1 != as_).flatten(1).any(dim=1)).flatten().cpu().numpy()
print("frames differing (hash):", len(bad), bad[:20], " (A plane):", len(badA), badA[:20])
F = n / 148
print("frames per CTA", F, "bad mod:", [(int(b), round(b / F, 2)) for b in badA[:12]])
for f in list(badA[:3]):
    _, _, ra, _ = oracle.pdq_stages(frames[f].cpu().numpy())
    d1 = np.argwhere(a1[f].cpu().numpy() != ra); d2 = np.argwhere(as_[f].cpu().numpy() != ra)
    print("frame", f, "one-launch vs oracle mismatches:", len(d1), d1[:6].tolist(), " split vs oracle:", len(d2), d2[:6].tolist())
