"""Dev-only: time kx_fused_jarosz2 variants (VPDQ_B200_FUSED2_VARIANT bit flags) in subprocesses."""
import os, subprocess, sys
names = {0: "baseline", 1: "no luma", 2: "no P2", 4: "no P3 loads", 8: "no barrier", 16: "no TMA", 32: "no P4",
         64: "no P1 stores", 70: "no P1st/P2/P3ld", 71: "no luma/P1st/P2/P3ld", 24: "no barrier/TMA", 6: "no P2/P3ld",
         17: "no luma/TMA"}
code = r'''
import sys, torch
sys.path.insert(0, ".")
from bench import device_frames
from hydrus_video_deduplicator_b200 import _ffi
n = 4096
dev = torch.device("cuda", 0)
fr = device_frames(torch, n, dev, seed=5)
a64 = torch.empty((n, 64, 64), dtype=torch.float32, device=dev)
st = torch.cuda.current_stream().cuda_stream
ms = []
for k in range(8):
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ea.record()
    _ffi.check(_ffi.lib().vpdq_b200_pdq_jarosz_dev(fr.data_ptr(), n, 512, 512, a64.data_ptr(), st))
    eb.record(); torch.cuda.synchronize()
    if k >= 3: ms.append(ea.elapsed_time(eb))
print("%.4f" % (sum(ms) / len(ms)))
'''
sel = [int(x) for x in sys.argv[1:]] or list(names)
for v, name in ((v, names[v]) for v in sel):
    env = dict(os.environ, VPDQ_B200_FUSED2_VARIANT=str(v))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(f"variant {v:3d} {name:24s}: {r.stdout.strip()} ms  {r.stderr.strip()[-200:] if r.returncode else ''}", flush=True)
