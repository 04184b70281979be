"""Dev helper: bucket SASS lines of one kernel (ncu --page source --csv dump) by execution count relative to a unit
(e.g. warps launched), to see fixed vs per-candidate vs per-group instruction cost; also stall summary."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
unit = float(sys.argv[2])
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
first = None; seen = set(); lines = []
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
stalls = collections.Counter()
for r in rows[2:]:
    if len(r) < len(hdr): continue
    a = r[ix["Address"]]
    if a in seen: continue
    seen.add(a)
    try: n = int(r[ix["Instructions Executed"]])
    except ValueError: continue
    for s in stall_cols: stalls[s] += int(r[ix[s]] or 0)
    lines.append((n, int(r[ix["# Samples"]] or 0), r[ix["Source"]].strip()))
tot = sum(n for n, _, _ in lines)
print(len(lines), "sass lines; warp instructions", tot, f"= {tot/unit:.1f} per unit")
b = collections.defaultdict(lambda: [0, 0, 0])
for n, smp, s in lines:
    k = round(n / unit, 1)
    b[k][0] += n; b[k][1] += 1; b[k][2] += smp
ts = sum(v[2] for v in b.values()) or 1
for k, (n, c, smp) in sorted(b.items()):
    if n / unit >= 1.0: print(f"  exec/unit {k:7.1f}: {c:4d} lines, {n/unit:7.1f} instr/unit, {100*smp/ts:5.1f}% of stall samples")
tt = sum(stalls.values()) or 1
print("stalls:", ", ".join(f"{k[6:]} {100*v/tt:.1f}%" for k, v in stalls.most_common(9)))
if len(sys.argv) > 3:
    for n, smp, s in sorted(lines, key=lambda x: -x[1])[:int(sys.argv[3])]: print(f"  {smp:6d} {n/unit:7.2f}  {s[:100]}")
