"""Dev helper: per-kernel counts of the SASS mnemonics that show the Blackwell-native paths (TMA loads / stores, mbarrier ops, packed fp32,
special-function ops) in the built library.  Usage: python scripts/sass_summary.py > profiles/<tag>/sass_tma.txt"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "fireflies_b200/_lib/libffb200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WANT = ["UTMALDG", "UTMASTG", "UTMACMDFLUSH", "SYNCS", "ELECT", "LDGSTS", "FFMA2", "FMUL2", "FADD2", "MUFU.EX2", "MUFU.RCP", "SHFL", "LDS", "STS",
        "REDG", "ATOMG", "RED", "BAR"]
cur, counts, total = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", "-p", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        tmpl = re.search(r"_ZN\d+\w+?(\d+)([a-z_0-9]+?)I(.*?)EEv", m.group(1))
        name = m.group(1)
        d = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", d).replace("void ", "").replace("ffb::", "")
        counts[cur] = collections.Counter(); total[cur] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        total[cur] += 1
        op = m.group(1)
        for w in WANT:
            if op == w or op.startswith(w + ".") or (w.startswith("MUFU") and op.startswith(w)):
                counts[cur][w] += 1
print(f"# {lib}: SASS mnemonic counts per kernel (static instruction counts; cuobjdump -sass)")
print(f"{'kernel':78s} {'instr':>6s} " + " ".join(f"{w:>8s}" for w in WANT[:12]))
for k, c in counts.items():
    if total[k] < 50: continue
    print(f"{k[:78]:78s} {total[k]:6d} " + " ".join(f"{c[w]:8d}" for w in WANT[:12]))
