"""Loops (backward branches spanning > 300 instructions) of kx_systolic_jarosz<3> in a library, with the static mix of
each: python tools/sass_loops.py lib.so steps_per_body"""
import re, subprocess, sys
lib, steps = sys.argv[1], int(sys.argv[2])
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
grab = False; ins = []
for line in out.splitlines():
    if "Function :" in line:
        grab = "kx_systolic_jaroszILi3" in line; continue
    if not grab: continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
for a, t in ins:
    if "BRA" in t:
        m = re.search(r"0x([0-9a-f]+)\s*$", t)
        if m and int(m.group(1), 16) < a and (a - int(m.group(1), 16)) // 16 > 300:
            lo = int(m.group(1), 16)
            print(f"loop {lo:#x}..{a:#x}: {(a - lo) // 16 + 1} instr, {(a - lo + 16) / 1024:.1f} KB")
            subprocess.run([sys.executable, "tools/sass_region.py", lib, hex(lo), hex(a), str(steps)])
