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
hdr, units, vals = rows[0], rows[1], rows[2]
get = lambda k: (float(vals[hdr.index(k)]), units[hdr.index(k)])
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
r, ru = get("dram__bytes_read.sum"); w, wu = get("dram__bytes_write.sum"); t, tu = get("gpu__time_duration.sum")
out = {"kernel": vals[hdr.index("Kernel Name")], "dram_bytes_read": r * scale[ru], "dram_bytes_write": w * scale[wu],
       "gpu_time_ms_under_ncu": t * {"ms": 1.0, "us": 1e-3, "s": 1e3}[tu], "source": f"{os.path.basename(rep)}: {cmd}"}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "vsweep_traffic.json")
json.dump(out, open(path, "w"), indent=1)
print(out)
