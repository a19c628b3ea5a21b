"""Dev: time the Jarosz kernel of the library named by VPDQ_B200_LIB (default: the product) on 8192 device-resident
frames and print a checksum of the planes (equal checksums = identical results): python tools/kx_time.py [n] [reps]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from bench import device_frames
from hydrus_video_deduplicator_b200 import _ffi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda", 0)
pool = [device_frames(torch, n, dev, seed=5 + p) for p in range(2)]
a64 = torch.empty((n, 64, 64), dtype=torch.float32, device=dev)
stream = torch.cuda.current_stream().cuda_stream
ms = []
for k in range(3 + reps):
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ea.record()
    _ffi.check(_ffi.lib().vpdq_b200_pdq_jarosz_dev(pool[k % 2].data_ptr(), n, 512, 512, a64.data_ptr(), stream))
    eb.record()
    torch.cuda.synchronize()
    if k >= 3:
        ms.append(ea.elapsed_time(eb))
m = sum(ms) / len(ms)
chk = int(a64.view(torch.int32).to(torch.int64).sum().item())
print(f"n={n}: {m:.4f} ms (min {min(ms):.4f})  {n / m / 1e3:.3f} M frames/s  frac {n * 786468 / m / 1e6 / 6531:.3f}  checksum {chk}  flags {_ffi.debug_flags(0)}")
