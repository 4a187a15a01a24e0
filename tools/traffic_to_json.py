#!/usr/bin/env python
"""Fold gpurun_out/traffic_<workload>.csv (tools/measure_traffic.sh) into profiles/r01_traffic.json:
{workload: {"kernel": ..., "dram_bytes_read": ..., "dram_bytes_write": ..., "traffic": read + write, "time_us": ...}}."""
import csv, glob, json, os, sys
out = {}
for path in sorted(glob.glob("gpurun_out/traffic_*.csv")):
    wl = os.path.basename(path)[len("traffic_"):-4]
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    if len(rows) < 2:
        continue
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    d = {}
    for r in rows[1:]:
        name, unit, val = r[ix["Metric Name"]], r[ix["Metric Unit"]], float(r[ix["Metric Value"]].replace(",", ""))
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1)
        d[name] = val * scale
        d["kernel"] = r[ix["Kernel Name"]]
    out[wl] = {"kernel": d.get("kernel"), "dram_bytes_read": d.get("dram__bytes_read.sum"),
               "dram_bytes_write": d.get("dram__bytes_write.sum"),
               "traffic": (d.get("dram__bytes_read.sum") or 0) + (d.get("dram__bytes_write.sum") or 0),
               "time_us_under_ncu": d.get("gpu__time_duration.sum")}
dst = sys.argv[1] if len(sys.argv) > 1 else "profiles/r01_traffic.json"
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out, indent=1))
