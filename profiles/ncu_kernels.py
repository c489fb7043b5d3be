"""Per-kernel summary + hot source lines of an ncu report.
usage: python profiles/ncu_kernels.py report.ncu-rep [lines_per_kernel]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
M = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
     "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
     "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    print("==", r[idx["Kernel Name"]][:80])
    for m in M:
        if m in idx:
            print(f"   {m:62s} {r[idx[m]]} {rows[1][idx[m]]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
out, name, last_fp = {}, None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        last_fp = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        name = r[1]
        out.setdefault(name, [])
    elif r[0] == "Line No":
        hd = {h: i for i, h in enumerate(r)}
    elif name and r[0].isdigit():
        g = lambda k: int(r[hd[k]]) if r[hd[k]].isdigit() else 0
        out[name].append((g("# Samples"), g("Instructions Executed"), g("Thread Instructions Executed"), g("stall_long_sb"), last_fp, int(r[0]), r[1].strip()[:95]))
for k, v in out.items():
    tot = sum(x[0] for x in v) or 1
    toti = sum(x[1] for x in v) or 1
    print(f"\n===== {k[:80]}  samples {tot} warp-inst {toti}")
    for x in sorted(v, reverse=True)[:top]:
        print(f"{100*x[0]/tot:6.1f}% inst {100*x[1]/toti:5.1f}% lanes {x[2]/max(x[1],1):5.1f} longsb {100*x[3]/max(x[0],1):3.0f}% | {x[4]}:{x[5]} {x[6]}")
