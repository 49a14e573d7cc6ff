"""Digest of one `ncu --set full --import-source on` report: headline counters, stall-reason shares and the hottest SASS lines.
    python tools/ncu_digest.py gpurun_out/x.ncu-rep [--top 40] [--sass out.txt]
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
sass_out = sys.argv[sys.argv.index("--sass") + 1] if "--sass" in sys.argv else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
print("kernel:", vals[hdr.index("Kernel Name")][:100])
for i, h in enumerate(hdr):
    if h in want or ("issue_stalled" in h and "per_issue_active" in h and float(vals[i] or 0) > 0.1):
        print(f"  {h:90s} {vals[i]:>16s} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix["# Samples"]]) for r in data)
texec = sum(int(r[ix["Instructions Executed"]]) for r in data)
print(f"samples {tot}, warp instructions {texec/1e9:.3f} G")
keys = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
print("  " + "  ".join(f"{k[6:]} {sum(int(r[ix[k]]) for r in data)/tot*100:.1f}" for k in keys if sum(int(r[ix[k]]) for r in data) / tot > 0.01))
lines = []
for n, r in enumerate(data):
    st = {k[6:]: int(r[ix[k]]) for k in keys if int(r[ix[k]]) > 0}
    lead = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    lines.append((int(r[ix["# Samples"]]), n, int(r[ix["Instructions Executed"]]), lead, r[ix["L1 Wavefronts Shared Excessive"]], r[ix["Source"]]))
if sass_out:
    with open(sass_out, "w") as f:
        for s, n, ex, lead, bc, srcl in lines:
            f.write(f"{n:5d} {ex/1e6:9.2f}M {s:6d} {str(lead):40s} bc={bc:>9s} {srcl}\n")
for s, n, ex, lead, bc, srcl in sorted(lines, reverse=True)[:top]:
    print(f"{n:5d} {ex/1e6:9.2f}M {s:6d} {100*s/tot:5.1f}% {str(lead):44s} {srcl[:70]}")
