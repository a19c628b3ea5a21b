"""Copy the artefacts of tools/gpu_round.sh (gpurun_out/) into profiles/ under this round's names and derive the small
files bench.py and DESIGN.md quote: python tools/make_profiles.py r02"""
import csv
import io
import json
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT, PROF = ROOT / "gpurun_out", ROOT / "profiles"
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"

shutil.copyfile(OUT / "bench_default.json", PROF / f"{tag}_bench_final.json")
shutil.copyfile(OUT / "bench_reference.json", PROF / f"{tag}_bench_reference.json")

# launch list: our kernels only, name + grid + duration
rows = list(csv.reader(io.StringIO((OUT / "launches_ncu.csv").read_text())))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hdr_i]
keep = [hdr] + [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
with open(PROF / f"{tag}_launches_final_ncu.csv", "w", newline="") as fh:
    csv.writer(fh).writerows(keep)

# condensed ncu --set full capture + DRAM traffic per launch
rep = OUT / "prof_pdq.ncu-rep"
summary = subprocess.run([sys.executable, str(ROOT / "tools" / "ncu_summary.py"), str(rep)], capture_output=True, text=True).stdout
(PROF / f"{tag}_ncu_summary_pdq.txt").write_text(summary)
raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
h, units = r[0], r[1]
traffic = {}
for row in r[2:]:
    name = row[h.index("Kernel Name")]
    key = "kx_systolic_jarosz" if "kx_systolic" in name else "k5_finalize" if "k5_finalize" in name else name[:40]

    def val(metric):
        v, u = float(row[h.index(metric)]), units[h.index(metric)]
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)

    traffic[key] = {"frames": 4096, "dram_read_bytes": val("dram__bytes_read.sum"), "dram_write_bytes": val("dram__bytes_write.sum"),
                    "duration_us_under_ncu": float(row[h.index("gpu__time_duration.sum")]) *
                    {"ms": 1e3, "us": 1.0, "ns": 1e-3, "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3}.get(units[h.index("gpu__time_duration.sum")], 1.0),
                    "source": "ncu --set full --clock-control none, tools/prof_pdq.py 4096"}
(PROF / f"{tag}_traffic.json").write_text(json.dumps(traffic, indent=1) + "\n")
print(json.dumps(traffic, indent=1))
