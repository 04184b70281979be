"""Dev helper: summarise an `ncu --page source --csv` dump: instruction mix by opcode and stall totals, hottest SASS lines."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter(); tot = 0
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
stalls = collections.Counter()
lines = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    try: n = int(r[ix["Instructions Executed"]])
    except ValueError: continue
    src = r[ix["Source"]].strip()
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    op = op.split(".")[0]
    ops[op] += n; tot += n
    samp = int(r[ix["# Samples"]] or 0)
    for s in stall_cols:
        stalls[s] += int(r[ix[s]] or 0)
    lines.append((samp, n, src))
print("total warp instructions", tot)
for op, n in ops.most_common(28): print(f"  {op:12s} {n:12d} {100*n/tot:5.1f}%")
ts = sum(stalls.values())
print("stall samples:", ", ".join(f"{k[6:]} {100*v/ts:.1f}%" for k, v in stalls.most_common(10)))
if len(sys.argv) > 2:
    print("hottest lines:")
    for samp, n, src in sorted(lines, reverse=True)[:int(sys.argv[2])]: print(f"  {samp:7d} {n:10d}  {src[:110]}")
