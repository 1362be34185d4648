"""Per-CUDA-line executed warp instructions per frame from an `ncu --page source --csv --print-source cuda,sass` dump.
usage: python ncu_inst_by_line.py src_cs.csv <launches> <frames_per_launch> [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
nl, frames = float(sys.argv[2]), float(sys.argv[3])
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
hi = [i for i, r in enumerate(rows) if "Instructions Executed" in r][0]
hdr = rows[hi]
ie, ws, sp = hdr.index("Instructions Executed"), hdr.index("L1 Wavefronts Shared"), hdr.index("# Samples")
out = {}
for r in rows[hi + 1:]:
    if len(r) > ws and r[0] != "":
        try:
            key = (r[0], r[1].strip()[:105])
            v = (float(r[ie]), float(r[ws]), float(r[sp]))
            out[key] = max(out.get(key, (0, 0, 0)), v)
        except ValueError:
            pass
tot = sum(v[0] for v in out.values())
print(f"sum over lines: {tot/nl/frames:.1f} warp-inst/frame (inlined callees are listed under their own line)")
for (ln, src), v in sorted(out.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{v[0]/nl/frames:7.2f} inst/frame {v[1]/nl/frames:6.2f} smem-wave/frame {v[2]:6.0f} smp  L{ln:>4} {src}")
