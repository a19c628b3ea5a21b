"""Static mix + stall-count sum of an address range of kx_systolic_jarosz<3>: python tools/sass_region.py lib.so 0xLO 0xHI steps"""
import re, subprocess, sys, collections
lib, lo, hi, steps = sys.argv[1], int(sys.argv[2], 16), int(sys.argv[3], 16), int(sys.argv[4])
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
ins = []; grab = False; pend = None
for line in out.splitlines():
    if "Function :" in line:
        grab = "kx_systolic_jaroszILi3" in line
        continue
    if not grab: continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/", line)
    if m:
        pend = [int(m.group(1), 16), m.group(2).strip(), int(m.group(3), 16)]; continue
    m = re.match(r"\s+/\* (0x[0-9a-f]+) \*/", line)
    if m and pend:
        pend.append(int(m.group(1), 16)); ins.append(pend); pend = None
tot = 0; n = 0; by = collections.Counter(); cnt = collections.Counter()
for a, t, w1, w2 in ins:
    if not (lo <= a <= hi): continue
    stall = (w2 >> 41) & 0xF
    op = re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0]
    tot += max(1, stall); n += 1; by[op] += max(1, stall); cnt[op] += 1
print(f"{n} instr = {n / steps:.1f}/step; stall-sum {tot} = {tot / steps:.1f} cycles/step")
for op, c in sorted(cnt.items(), key=lambda kv: -kv[1])[:40]:
    print(f"   {op:8s} n {c / steps:6.2f}  cycles {by[op] / steps:6.2f}")
