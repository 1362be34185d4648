"""Bucket an `ncu --page source --csv` (SASS view) dump by opcode: executed warp instructions and stall samples.
usage: ncu -i X.ncu-rep --page source --csv > sass.csv; python ncu_by_opcode.py sass.csv [launches]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
nl = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hi = [i for i, r in enumerate(rows) if "Instructions Executed" in r][0]
hdr = rows[hi]
ci, cs, cw = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("L1 Wavefronts Shared")
ct = hdr.index("Thread Instructions Executed")
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= cw or not r[1].strip():
        continue
    toks = r[1].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LDS", "STS", "LDG", "STG")) and "." in op else "")
    try:
        a = agg[op]
        a[0] += float(r[ci]); a[1] += float(r[cs]); a[2] += float(r[cw]); a[3] += float(r[ct])
    except ValueError:
        pass
ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print(f"warp instructions per launch: {ti/nl:.4e}; stall samples {ts:.0f}")
print(f"{'opcode':<14}{'inst%':>7}{'smp%':>7}{'wavefronts/launch':>19}{'thr/inst':>9}")
for op, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
    print(f"{op:<14}{100*a[0]/ti:7.2f}{100*a[1]/ts:7.2f}{a[2]/nl:19.3e}{a[3]/a[0] if a[0] else 0:9.1f}")
