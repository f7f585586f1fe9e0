"""Summarise the ncu evidence of one bench run for profiles/: (1) the launch list (`--metrics gpu__time_duration.sum`) as
per-kernel time shares of one step, (2) the `--set full` capture as a table, (3) profiles/ncu_traffic.json (DRAM bytes per
launch, read by bench.py for roofline.traffic).  Usage: python tools/ncu_summary.py <tag> [workload] [batch]"""
import csv, json, os, subprocess, sys, collections

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
workload = sys.argv[2] if len(sys.argv) > 2 else "c2"
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 64
rnd = os.environ.get("PN_ROUND", "r02")
out = ["# Round " + rnd[1:].lstrip("0") + ", %s -- ncu evidence, one forward + decode at C2 (model 101, 513x513, OS16, batch 64, bf16)" % tag, ""]

# ---- launch list
rows = [r for r in csv.reader(open(os.path.join(ROOT, "gpurun_out", "launches.csv"))) if r and r[0].isdigit()]
# columns: ID, Process ID, Process Name, Host Name, Kernel Name, Context, Stream, Block Size, Grid Size, Device, CC, Section, Metric Name, Unit, Value
names = [r[4] for r in rows]
vals = [float(r[-1].replace(",", "")) for r in rows]
unit = rows[0][-2]
scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1.0)
vals = [v * scale for v in vals]
def short(n):
    n = n.split("(")[0]
    return n.replace("void ", "").replace("pn::", "")
# last complete step = last 19-launch window that starts at the stem
idx = [i for i, n in enumerate(names) if "stem" in n]
start = idx[-2] if len(idx) >= 2 else idx[-1]
end = idx[-1] if len(idx) >= 2 else len(names)
dec = [i for i in range(start, end) if "decode_kernel" in names[i]]
if dec:
    end = dec[0] + 1                       # what follows the decode kernel belongs to the next phase of bench.py, not to the step
step = list(zip(names[start:end], vals[start:end]))
tot = sum(v for _, v in step)
out += ["## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, tools/gpu_profile.sh)", "",
        "%d launches captured; the last complete step (%d launches, %.1f us serialised, cold-cache per-launch times -- shares, not absolutes, are comparable with bench.py):" % (len(names), len(step), tot), "",
        "| # | kernel | us | share |", "|---|---|---|---|"]
for i, (n, v) in enumerate(step):
    out.append("| %d | `%s` | %.1f | %.1f %% |" % (i, short(n)[:60], v, 100 * v / tot))
out.append("")

# ---- full capture
rep = os.path.join(ROOT, "gpurun_out", "prof_full.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr, data = rr[0], rr[2:]
col = lambda n: hdr.index(n)
M = [("gpu__time_duration.sum", "time us"), ("dram__bytes_read.sum", "dram rd MB"), ("dram__bytes_write.sum", "dram wr MB"),
     ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram %"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
     ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem pipe (LSU) %"),
     ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem pipe (tensor operands) %"),
     ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
     ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"), ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid")]
out += ["## `ncu --set full --clock-control none --import-source on` (first forward of the process; %d kernels)" % len(data), "",
        "| # | kernel | " + " | ".join(l for _, l in M) + " |", "|---|---|" + "---|" * len(M)]
plan = ["stem", "sep1", "sep2", "sep3", "sep4", "sep5", "sep6", "sep7", "sep8", "sep9", "sep10", "sep11", "dw12", "pw12", "dw13", "pw13", "heads",
        "candidates", "decode"]
traffic = {}
def num(s):
    try: return float(s.replace(",", ""))
    except Exception: return float("nan")
for i, r in enumerate(data):
    cells = []
    for m, _ in M:
        v = r[col(m)] if m in hdr else ""
        u = rr[1][col(m)] if m in hdr else ""
        x = num(v)
        if u == "byte": x /= 1e6
        if u == "Kbyte": x /= 1e3
        if u == "Gbyte": x *= 1e3
        if u in ("ns", "nsecond"): x /= 1e3
        if u in ("ms", "msecond"): x *= 1e3
        cells.append("%.1f" % x if x == x else v)
    out.append("| %d | `%s` | " % (i, short(r[col("Kernel Name")])[:44]) + " | ".join(cells) + " |")
    if i < len(plan):
        def mb(m):
            x = num(r[col(m)]); u = rr[1][col(m)]
            return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        traffic[plan[i]] = int(mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum"))
traffic["candidates+decode"] = traffic.get("candidates", 0) + traffic.get("decode", 0)
out += ["", "Launch order: " + ", ".join(plan) + ".  DRAM traffic below the algorithmic bytes on the late layers = the 126 MB L2 keeps part of the",
        "previous layer's output resident (activations of blocks 7-13 are 71-143 MB).", ""]
open(os.path.join(ROOT, "profiles", "%s_%s_ncu_summary.md" % (rnd, tag)), "w").write("\n".join(out))
json.dump({"workload": workload, "batch": batch, "source": "profiles/%s_%s_ncu_summary.md" % (rnd, tag), "dram_bytes": traffic},
          open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
print("\n".join(out[:60]))
