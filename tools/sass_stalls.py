"""Static issue-cycle estimate of the step loop of kx_systolic_jarosz<3>: sums the stall counts in the SASS control
words over the loop (optionally only over the instructions an ncu source page says are hot).
python tools/sass_stalls.py lib.so body [hot.txt]"""
import re, subprocess, sys, collections
lib, body = sys.argv[1], int(sys.argv[2])
hot = None
if len(sys.argv) > 3:
    hot = [float(l.split()[1]) for l in open(sys.argv[3])]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
ins = []; grab = False; pend = None
for line in out.splitlines():
    if "Function :" in line:
        grab = "kx_systolic_jaroszILi3" in line
        continue
    if not grab: continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/", line)
    if m:
        pend = [int(m.group(1), 16), m.group(2).strip(), int(m.group(3), 16)]
        continue
    m = re.match(r"\s+/\* (0x[0-9a-f]+) \*/", line)
    if m and pend:
        w2 = int(m.group(1), 16)
        pend.append(w2); ins.append(pend); pend = None
print("instructions", len(ins))
# loop = biggest backward branch
best = None
amap = {a: i for i, (a, *_ ) in enumerate(ins)}
for i, (a, t, w1, w2) in enumerate(ins):
    if t.split()[0].startswith("BRA") or " BRA " in " " + t:
        m = re.search(r"0x([0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in amap and (best is None or a - tgt > best[1] - best[0]): best = (tgt, a)
lo, hi = amap[best[0]], amap[best[1]]
tot = 0; n = 0; by = collections.Counter(); cnt = collections.Counter()
for i in range(lo, hi + 1):
    a, t, w1, w2 = ins[i]
    if hot is not None and not (0.4 < hot[i] < 0.7): continue
    stall = (w2 >> 41) & 0xF
    op = re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0]
    tot += max(1, stall); n += 1; by[op] += max(1, stall); cnt[op] += 1
print(f"loop {hex(best[0])}..{hex(best[1])}: {n} instr = {n / body:.1f}/step; stall-sum {tot} = {tot / body:.1f} cycles/step (one warp alone, no scoreboard waits)")
for op, c in by.most_common(14):
    print(f"   {op:8s} n {cnt[op] / body:6.1f}  cycles {c / body:6.1f}  avg {c / cnt[op]:.2f}")
