"""Summarise one-kernel `ncu --set full` reports into the JSON kept under profiles/.
usage: python tools/ncu_summary.py report.ncu-rep out.json "<capture command>" "<note>" [pairs]"""
import csv, io, json, subprocess, sys

rep, out, cmd, note = sys.argv[1:5]
pairs = int(sys.argv[5]) if len(sys.argv) > 5 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active"]
m = {}
for h, u, v in zip(hdr, units, vals):
    if h in keep or (h.startswith("smsp__pcsamp_warps_issue_stalled") and "not_issued" not in h and v not in ("0", "")):
        m[h] = f"{v} {u}".strip()


def to_bytes(s):
    v, u = s.split()
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


d = {"capture": cmd, "note": note, "kernel": hdr and vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else None,
     "dram_traffic_bytes_per_launch": to_bytes(m["dram__bytes_read.sum"]) + to_bytes(m["dram__bytes_write.sum"]),
     "metrics": m}
if pairs:
    d["pairs"] = pairs
json.dump(d, open(out, "w"), indent=1)
print(json.dumps(d, indent=1)[:1500])
