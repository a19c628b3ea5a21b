"""Dev check of the systolic Jarosz kernel on the GPU: bit-identity with the tiled kernel and the oracle for many
batch shapes (RGB24 and gray), then timing of both: python tools/sys_check.py [n_frames] [reps]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

import oracle
from bench import device_frames
from hydrus_video_deduplicator_b200 import _ffi, device

dev = torch.device("cuda", 0)
bad = 0
for ch in (3, 1):
    for n in (1, 2, 3, 7, 8, 9, 147, 149, 1000, 1184, 1185, 1200, 2500):
        frames = device_frames(torch, n, dev, seed=900 + n)
        if ch == 1:
            frames = frames[..., 1].contiguous()
        out = {}
        for impl in ("fused2", "systolic"):
            _ffi.set_pdq_impl(impl)
            out[impl] = device.hash_frames(frames, stages=True)
        torch.cuda.synchronize()
        same = all(torch.equal(a, b) for a, b in zip(out["fused2"][:3], out["systolic"][:3]))
        idx = list(range(0, n, max(1, n // 6)))
        ref_h, ref_q = oracle.pdq_hash_frames(
            (frames[idx] if ch == 3 else frames[idx].unsqueeze(-1).expand(-1, -1, -1, 3).contiguous()).cpu().numpy(), nthreads=8)
        ok = out["systolic"][0][idx].cpu().numpy().tobytes() == ref_h.tobytes()
        if not (same and ok):
            bad += 1
            d = (out["fused2"][2] != out["systolic"][2]).flatten(1).any(1).nonzero().flatten()
            print(f"MISMATCH ch={ch} n={n}: same={same} oracle={ok} differing frames {d[:10].tolist()} of {d.numel()}")
print("shapes checked, mismatches:", bad, "flags", _ffi.debug_flags(0))

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
pool = [device_frames(torch, n, dev, seed=5 + p) for p in range(2)]
stream = torch.cuda.current_stream().cuda_stream
outs = {}
for impl in ("fused2", "systolic"):
    _ffi.set_pdq_impl(impl)
    a64 = torch.empty((n, 64, 64), dtype=torch.float32, device=dev)
    ms = []
    for k in range(3 + reps):
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
        _ffi.check(_ffi.lib().vpdq_b200_pdq_jarosz_dev(pool[k % 2].data_ptr(), n, 512, 512, a64.data_ptr(), stream))
        eb.record()
        torch.cuda.synchronize()
        if k >= 3:
            ms.append(ea.elapsed_time(eb))
    _ffi.check(_ffi.lib().vpdq_b200_pdq_jarosz_dev(pool[0].data_ptr(), n, 512, 512, a64.data_ptr(), stream))
    torch.cuda.synchronize()
    outs[impl] = a64
    m = sum(ms) / len(ms)
    print(f"{impl:9s} n={n}: {m:.4f} ms/launch (min {min(ms):.4f})  {n / m / 1e3:.3f} M frames/s  "
          f"{n * 786468 / m / 1e6:.0f} GB/s")
print("bit-identical planes:", bool(torch.equal(outs["fused2"], outs["systolic"])), " debug_flags:", _ffi.debug_flags(0))
for small in (1, 10, 32, 64, 148, 300):
    fr = pool[0][:small]
    for impl in ("fused2", "systolic"):
        _ffi.set_pdq_impl(impl)
        a64 = torch.empty((small, 64, 64), dtype=torch.float32, device=dev)
        ms = []
        for k in range(8):
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ea.record()
            _ffi.check(_ffi.lib().vpdq_b200_pdq_jarosz_dev(fr.data_ptr(), small, 512, 512, a64.data_ptr(), stream))
            eb.record()
            torch.cuda.synchronize()
            if k >= 3:
                ms.append(ea.elapsed_time(eb))
        print(f"  small batch {small:4d} {impl:9s}: {1e3 * min(ms):.1f} us")
