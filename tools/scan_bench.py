"""one streaming scan launch for profiling: python tools/scan_bench.py [n_db] [n_query]"""
import sys
sys.path.insert(0, "/root/repo")
import torch
from hydrus_video_deduplicator_b200 import device
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1
g = torch.Generator(device="cuda").manual_seed(1)
db = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g)
off = torch.arange(0, n + 1, 300, dtype=torch.int64, device="cuda")
if int(off[-1]) != n:
    off = torch.cat([off, torch.tensor([n], dtype=torch.int64, device="cuda")])
q = db[:nq].clone()
for _ in range(4):
    m = device.hamming_scan(db, q, off, 31)
torch.cuda.synchronize()
print("ok", int((m != 0).sum()))
