"""Per-kernel table from an ncu CSV with several metrics:  python tools/ncu_step_table.py launches.csv"""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); mi = hdr.index("Metric Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit"); ii = hdr.index("ID")
per = collections.OrderedDict()
for r in rows[1:]:
    key = (r[ii], r[ki][:48])
    v = float(r[vi].replace(",", "")); u = r[ui]
    if r[mi] == "gpu__time_duration.sum":
        v = v / 1e6 if u.startswith("n") else (v / 1e3 if u.startswith("u") else v)
    if r[mi].startswith("dram__bytes"):
        v = {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}.get(u, 1e-9) * v
    per.setdefault(key, {})[r[mi]] = v
agg = collections.OrderedDict()
for (_, k), m in per.items():
    a = agg.setdefault(k, collections.defaultdict(float)); a["n"] += 1
    for kk, vv in m.items(): a[kk] += vv
print(f"{'kernel':50s} {'n':>3s} {'ms/launch':>9s} {'GB/launch':>9s} {'TB/s':>6s} {'issue%':>6s} {'warps%':>6s}")
for k, a in agg.items():
    n = a["n"]; ms = a["gpu__time_duration.sum"] / n
    gb = (a.get("dram__bytes_read.sum", 0) + a.get("dram__bytes_write.sum", 0)) / n
    print(f"{k:50s} {int(n):3d} {ms:9.3f} {gb:9.3f} {gb / ms if ms else 0:6.2f} {a.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0) / n:6.1f} {a.get('sm__warps_active.avg.pct_of_peak_sustained_active', 0) / n:6.1f}")
