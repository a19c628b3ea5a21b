import sys; sys.path.insert(0, '/root/repo')
import numpy as np, torch, oracle
from tests import synth
from hydrus_video_deduplicator_b200 import device, _ffi
for n, ch in ((3, 3), (300, 3), (20, 1)):
    frames = synth.synth_frames(min(n, 6), seed=5 + n, channels=ch)
    reps = -(-n // frames.shape[0]); batch = np.concatenate([frames] * reps)[:n]
    h, q = device.hash_frames(torch.from_numpy(batch).cuda()); torch.cuda.synchronize()
    rgb = frames if ch == 3 else np.repeat(frames[..., None], 3, axis=3)
    ref_h, _ = oracle.pdq_hash_frames(rgb, nthreads=4)
    print(n, ch, "ok" if h.cpu().numpy().tobytes() == np.concatenate([ref_h] * reps)[:n].tobytes() else "MISMATCH", "flags", _ffi.debug_flags(0))
