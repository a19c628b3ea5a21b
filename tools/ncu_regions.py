"""Stall-reason breakdown of a kernel per code class (by how often an instruction executed relative to the hottest loop):
python tools/ncu_regions.py report.ncu-rep kernel_regex"""
import csv, io, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:end]))))
ex = [int(r["Instructions Executed"]) for r in rows]
# the plain loop = the most common execution count among the hot instructions
c = collections.Counter(e for e in ex if e > 0)
plain = max(c.items(), key=lambda kv: kv[1] * kv[0])[0]
reasons = [k for k in rows[0].keys() if k.startswith("stall_") and "Not Issued" not in k]
agg = collections.defaultdict(lambda: collections.Counter()); n = collections.Counter(); exs = collections.Counter()
for r, e in zip(rows, ex):
    w = e / plain
    k = "plain" if 0.95 < w < 1.05 else ("general" if 0.01 < w <= 0.95 else ("hotter" if w >= 1.05 else "cold"))
    n[k] += 1; exs[k] += e
    for q in reasons: agg[k][q] += int(r[q] or 0)
tot = sum(sum(a.values()) for a in agg.values())
for k in agg:
    s = sum(agg[k].values())
    top = ", ".join(f"{q[6:]} {100 * v / max(1, s):.0f}%" for q, v in agg[k].most_common(8))
    print(f"{k:8s} {n[k]:5d} instr, executed {exs[k] / plain:8.1f} x plain-count, samples {100 * s / tot:5.1f}%: {top}")
