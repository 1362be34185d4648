"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv; python ncu_by_line.py src.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = [i for i, r in enumerate(rows) if "Instructions Executed" in r][0]
hdr = rows[hi]
col = {n: hdr.index(n) for n in ("Instructions Executed", "# Samples", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal",
                                "Thread Instructions Executed")}
files = {}
cur = None
out = []
for r in rows[hi + 1:]:
    if len(r) <= col["L1 Wavefronts Shared Ideal"]:
        continue
    if r[0] != "":   # a CUDA source line with aggregated metrics
        def num(k):
            try:
                return float(r[col[k]])
            except ValueError:
                return 0.0
        out.append((num("Instructions Executed"), num("# Samples"), num("L1 Wavefronts Shared"),
                    num("L1 Wavefronts Shared Ideal"), num("Thread Instructions Executed"), r[0], r[1].strip()[:110]))
tot_i = sum(o[0] for o in out) or 1
tot_s = sum(o[1] for o in out) or 1
tot_w = sum(o[2] for o in out) or 1
print(f"total warp-inst {tot_i:.3e}  samples {tot_s:.0f}  smem wavefronts {tot_w:.3e}")
print(f"{'inst%':>6} {'smp%':>6} {'wave%':>6} {'w/ideal':>7} {'thr/inst':>8}  line  source")
for o in sorted(out, key=lambda o: -o[1])[:top]:
    print(f"{100*o[0]/tot_i:6.2f} {100*o[1]/tot_s:6.2f} {100*o[2]/tot_w:6.2f} {o[2]/o[3] if o[3] else 0:7.2f} {o[4]/o[0] if o[0] else 0:8.1f}  {o[5]:>4}  {o[6]}")
