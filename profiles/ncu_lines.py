"""Aggregate an ncu source page (`ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`) per CUDA source line.
usage: python profiles/ncu_lines.py file.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
fpath = None
hdr = None
out = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr and r[0].isdigit():
        def g(name):
            try:
                return int(r[hdr[name]])
            except Exception:
                return 0
        out.append((g("# Samples"), g("Instructions Executed"), g("Thread Instructions Executed"), g("stall_long_sb"), g("L2 Theoretical Sectors Global"),
                    fpath, int(r[0]), r[1].strip()[:100]))
tot = sum(o[0] for o in out) or 1
toti = sum(o[1] for o in out) or 1
tots = sum(o[4] for o in out) or 1
print(f"samples {tot}  warp-inst {toti}  L2 sectors(global) {tots}")
print(f"{'smp%':>6} {'inst%':>6} {'lanes':>5} {'longsb%':>7} {'sect%':>6}  line")
for o in sorted(out, reverse=True)[:top]:
    lanes = o[2] / o[1] if o[1] else 0
    print(f"{100*o[0]/tot:6.1f} {100*o[1]/toti:6.1f} {lanes:5.1f} {100*o[3]/max(o[0],1):7.0f} {100*o[4]/tots:6.1f}  {o[5]}:{o[6]} {o[7]}")
