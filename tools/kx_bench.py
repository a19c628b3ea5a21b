"""Time the Jarosz kernel alone (vpdq_b200_pdq_jarosz_dev) per implementation and cross-check the decimated
planes bit for bit: python tools/kx_bench.py [n_frames] [reps]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from bench import device_frames
from hydrus_video_deduplicator_b200 import _ffi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda", 0)
pool = [device_frames(torch, n, dev, seed=5 + p) for p in range(2)]
stream = torch.cuda.current_stream().cuda_stream
outs = {}
for impl in ("fused2", "fused"):
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
    print(f"{impl:7s} n={n}: {m:.4f} ms/launch (min {min(ms):.4f})  {n / m / 1e3:.3f} M frames/s  "
          f"{n * 786468 / m / 1e6:.0f} GB/s")
print("bit-identical planes:", bool(torch.equal(outs["fused2"], outs["fused"])), " debug_flags:", _ffi.debug_flags(0))
