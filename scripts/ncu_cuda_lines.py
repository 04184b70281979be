"""Dev helper: `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > f.csv` -> warp instructions per unit and stall samples per CUDA source line.
Usage: python scripts/ncu_cuda_lines.py f.csv <units> [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]); top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
cur = None; hdr = None; agg = []
for r in rows:
    if r and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r and r[0] in ("Function Name",): continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and r and r[0].isdigit():
        try: n = int(r[hdr.index("Instructions Executed")])
        except ValueError: continue
        agg.append((cur, int(r[0]), n, (int(r[hdr.index("# Samples")]) if r[hdr.index("# Samples")].isdigit() else 0), r[1][:110]))
tot = sum(a[2] for a in agg); st = sum(a[3] for a in agg)
print(f"total {tot / units:.1f} warp instructions per unit, {st} stall samples")
for a in sorted(agg, key=lambda a: -a[2])[:top]:
    print(f"{a[0]:20s} {a[1]:5d} {a[2] / units:8.1f}  {100.0 * a[3] / st:5.1f}%  {a[4]}")
