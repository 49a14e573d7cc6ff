"""profiles/vsweep_traffic.json from an `ncu --set full` report of the v-sweep (bench.py reads it for roofline.traffic).
    python tools/ncu_traffic.py gpurun_out/x.ncu-rep "<the ncu command>" """
import csv
import io
import json
import os
import subprocess
import sys

rep, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
tscale = {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}
col = lambda k: hdr.index(k)
# one v-sweep over the batch = every launch in the report (a batch that does not fill its last round of teams is swept in two
# launches with different strip widths): sum them
r = w = t = 0.0
names = []
for vals in rows[2:]:
    if len(vals) < len(hdr):
        continue
    r += float(vals[col("dram__bytes_read.sum")]) * scale[units[col("dram__bytes_read.sum")]]
    w += float(vals[col("dram__bytes_write.sum")]) * scale[units[col("dram__bytes_write.sum")]]
    t += float(vals[col("gpu__time_duration.sum")]) * tscale[units[col("gpu__time_duration.sum")]]
    names.append(vals[col("Kernel Name")])
out = {"kernel": " + ".join(names), "launches": len(names), "dram_bytes_read": r, "dram_bytes_write": w,
       "gpu_time_ms_under_ncu": t, "source": f"{os.path.basename(rep)}: {cmd}"}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "vsweep_traffic.json")
json.dump(out, open(path, "w"), indent=1)
print(out)
