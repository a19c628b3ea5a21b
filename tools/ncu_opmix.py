"""Executed-instruction mix of a kernel from an ncu report's source page (per opcode: warp-instructions executed,
share, stall samples): python tools/ncu_opmix.py report.ncu-rep kernel_regex [units]   (units: divide counts, e.g.
warp-steps = frames * 516)"""
import csv, io, re, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
units = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
# first kernel instance only
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:end]))))
ex = collections.Counter(); st = collections.Counter(); tot = 0; tots = 0
for r in rows:
    t = re.sub(r"^@!?U?P\d+\s+", "", r["Source"].strip())
    op = t.split()[0].split(".")[0]
    if op in ("FADD2", "FFMA2", "FMUL2", "FADD", "FFMA", "FMUL", "PRMT", "LOP3", "SHFL", "LDS", "MOV"):
        pass
    n = int(r["Instructions Executed"]); s = int(r["# Samples"])
    ex[op] += n; st[op] += s; tot += n; tots += s
print(f"executed {tot}  = {tot / units:.1f} per unit; samples {tots}")
for op, n in ex.most_common(40):
    print(f"  {op:10s} {n / units:8.2f}  {100 * n / tot:5.1f}%   samples {100 * st[op] / max(1, tots):5.1f}%")
