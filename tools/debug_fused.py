import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import oracle
from hydrus_video_deduplicator_b200 import device
from tests import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
frames = synth.synth_frames(n, seed=21)
d = torch.from_numpy(frames).cuda()
try:
    h, q, a, b = device.hash_frames(d, stages=True)
    torch.cuda.synchronize()
except Exception as e:
    print("ERROR:", repr(e)[:2000]); sys.exit(1)
rh, rq = oracle.pdq_hash_frames(frames)
print("hash equal:", (h.cpu().numpy() == rh).all(), "quality equal:", (q.cpu().numpy() == rq).all())
for k in range(n):
    _, _, ra, rb = oracle.pdq_stages(frames[k])
    diff = (a[k].cpu().numpy() != ra)
    print(k, "A mismatches:", int(diff.sum()), "first:", np.argwhere(diff)[:5].tolist())
