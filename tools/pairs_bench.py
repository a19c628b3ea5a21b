"""quick timing of the all-pairs Hamming kernel: python tools/pairs_bench.py [n] [iters]"""
import sys, time
sys.path.insert(0, "/root/repo")
import torch
from hydrus_video_deduplicator_b200 import device
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
g = torch.Generator(device="cuda").manual_seed(1)
h = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g)
h[n // 2 : n // 2 + 1000] = h[:1000] ^ torch.tensor([0x0F] * 5 + [0] * 27, dtype=torch.uint8, device="cuda")
device.hamming_pairs(h[:4096], h, 31, skip_diagonal=True)
for _ in range(iters):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    cnt, pairs, _ = device.hamming_pairs(h, h, 31, skip_diagonal=True)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    print(f"n={n} {ms:.2f} ms  {n * n / ms / 1e9:.3f} T pairs/s  matches={cnt}")
