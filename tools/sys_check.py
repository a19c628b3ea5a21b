"""Dev check of the systolic Jarosz kernel on the GPU: bit-identity with the round-1 tiled kernel (tests/legacy) and the
oracle for many batch shapes (RGB24 and gray), then timing of the Jarosz kernel alone and of small batches:
python tools/sys_check.py [n_frames] [reps]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import oracle
from bench import device_frames
from hydrus_video_deduplicator_b200 import _ffi, device
from tests import legacy

dev = torch.device("cuda", 0)
bad = 0
for ch in (3, 1):
    for n in (1, 2, 3, 7, 8, 9, 147, 149, 1000, 1184, 1185, 1200, 2500):
        frames = device_frames(torch, n, dev, seed=900 + n)
        if ch == 1:
            frames = frames[..., 1].contiguous()
        ours = device.hash_frames(frames, stages=True)
        theirs = legacy.hash_frames(frames, "fused2")
        same = all(torch.equal(a, b) for a, b in zip(ours, theirs))
        idx = list(range(0, n, max(1, n // 6)))
        rgb = frames[idx] if ch == 3 else frames[idx].unsqueeze(-1).expand(-1, -1, -1, 3).contiguous()
        ref_h, ref_q = oracle.pdq_hash_frames(rgb.cpu().numpy(), nthreads=8)
        ok = ours[0][idx].cpu().numpy().tobytes() == ref_h.tobytes()
        if not (same and ok):
            bad += 1
            print(f"MISMATCH ch={ch} n={n}: same={same} oracle={ok}")
print("shapes checked, mismatches:", bad, "flags", _ffi.debug_flags(0))

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
pool = [device_frames(torch, n, dev, seed=5 + p) for p in range(2)]
stream = torch.cuda.current_stream().cuda_stream


def time_jarosz(frames, count, reps):
    a64 = torch.empty((count, 64, 64), dtype=torch.float32, device=dev)
    ms = []
    for k in range(3 + reps):
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
        _ffi.check(_ffi.lib().vpdq_b200_pdq_jarosz_dev(frames[k % len(frames)].data_ptr(), count, 512, 512, a64.data_ptr(), stream))
        eb.record()
        torch.cuda.synchronize()
        if k >= 3:
            ms.append(ea.elapsed_time(eb))
    return ms


ms = time_jarosz(pool, n, reps)
m = sum(ms) / len(ms)
print(f"systolic n={n}: {m:.4f} ms/launch (min {min(ms):.4f})  {n / m / 1e3:.3f} M frames/s  {n * 786468 / m / 1e6:.0f} GB/s")
for small in (1, 10, 32, 64, 148, 300, 1184):
    ms = time_jarosz([pool[0][:small]], small, 5)
    print(f"  small batch {small:4d}: {1e3 * min(ms):.1f} us")
print("debug_flags:", _ffi.debug_flags(0))
