"""Instruction mix of the step loop of kx_systolic_jarosz<3> in a built library (cuobjdump -sass): finds the biggest
backward branch (the step loop), counts the instructions in it by mnemonic: python tools/sass_mix.py lib.so [body]"""
import re, subprocess, sys, collections
lib = sys.argv[1]; body = int(sys.argv[2]) if len(sys.argv) > 2 else 4
out = subprocess.run(["cuobjdump", "-sass", "-fun", "_ZN4vpdq18kx_systolic_jaroszILi3EEEv14CUtensorMap_stS1_ixPf", lib],
                     capture_output=True, text=True).stdout
ins = []
for line in out.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr = {a: i for i, (a, _) in enumerate(ins)}
best = None
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA\S*\s+(?:\S+,\s*)?`\(\.L_x_\d+\)|BRA\S*\s+.*0x([0-9a-f]+)", t)
    if "BRA" in t:
        m2 = re.search(r"0x([0-9a-f]+)", t)
        if m2:
            tgt = int(m2.group(1), 16)
            if tgt < a and tgt in addr and (best is None or a - tgt > best[1] - best[0]):
                best = (tgt, a)
print("instructions total", len(ins), "loop", best and (hex(best[0]), hex(best[1])))
lo, hi = addr[best[0]], addr[best[1]]
loop = ins[lo:hi + 1]
c = collections.Counter()
for a, t in loop:
    t = re.sub(r"^@!?U?P\d+\s+", "", t)
    c[t.split()[0].split(".")[0]] += 1
n = len(loop)
print(f"loop instructions {n} = {n / body:.1f} per step, {n * 16 / 1024:.1f} KB")
for k, v in c.most_common(30):
    print(f"  {k:10s} {v:5d}  {v / body:6.1f}")
