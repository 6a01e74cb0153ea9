"""Extract dram__bytes_read.sum + dram__bytes_write.sum of one kernel launch from an `ncu --set full` capture and
record it in profiles/ncu_traffic.json (what bench.py reports as roofline.traffic).

usage: python tools/ncu_traffic.py <key> <report.ncu-rep> <kernel-name substring> "<command the capture ran>"
"""
import csv, datetime, json, os, subprocess, sys

UNIT = {"byte": 1., "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
key, rep, kname, cmd = sys.argv[1:5]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
best = None
for r in rows[2:]:
    d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
    if kname not in d.get("Kernel Name", ""):
        continue
    tot = 0.
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(d[m].replace(",", "")) * UNIT[u[m]]
    best = {"bytes_per_launch": int(round(tot)), "kernel": d["Kernel Name"][:120],
            "duration_us_under_ncu": float(d["gpu__time_duration.sum"].replace(",", "")) *
            {"ns": 1e-3, "us": 1., "ms": 1e3, "s": 1e6}.get(u["gpu__time_duration.sum"], 1.),
            "source": os.path.basename(rep), "command": cmd,
            "captured": datetime.date.today().isoformat()}
    break
if best is None:
    sys.exit("kernel %s not found in %s" % (kname, rep))
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
allv = json.load(open(p)) if os.path.exists(p) else {}
allv[key] = best
json.dump(allv, open(p, "w"), indent=1, sort_keys=True)
print(key, best)
