"""Dev-only: time the Jarosz kernel of alternative builds of the library (tools/variants/lib_*.so, made with
VPDQ_OUT=... VPDQ_EXTRA_DEFS=... csrc/build.sh) and print a checksum of the decimated planes."""
import glob
import subprocess
import sys

code = r'''
import sys, torch
sys.path.insert(0, ".")
from pathlib import Path
from hydrus_video_deduplicator_b200 import _ffi
_ffi.LIB_PATH = Path(sys.argv[1])
from bench import device_frames
n = 4096
dev = torch.device("cuda", 0)
fr = device_frames(torch, n, dev, seed=5)
a64 = torch.empty((n, 64, 64), dtype=torch.float32, device=dev)
st = torch.cuda.current_stream().cuda_stream
ms = []
for k in range(10):
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ea.record()
    _ffi.check(_ffi.lib().vpdq_b200_pdq_jarosz_dev(fr.data_ptr(), n, 512, 512, a64.data_ptr(), st))
    eb.record(); torch.cuda.synchronize()
    if k >= 3: ms.append(ea.elapsed_time(eb))
print("%.4f ms  checksum %.9e  flags %d" % (sum(ms) / len(ms), float(a64.double().sum()), _ffi.debug_flags(0)))
'''
for so in sorted(glob.glob("tools/variants/lib_*.so")):
    r = subprocess.run([sys.executable, "-c", code, so], capture_output=True, text=True)
    print(so, r.stdout.strip(), r.stderr.strip()[-300:] if r.returncode else "", flush=True)
