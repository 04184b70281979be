"""Dev helper: build A/B variants of libffb200.so with different -D settings into fireflies_b200/_lib/ab/<tag>.so.
Usage: python scripts/ab_build.py tag1:-DX=1,-DY=2 tag2:..."""
import os, subprocess, sys
sys.path.insert(0, ".")
from fireflies_b200 import _build as b
out = os.path.join(b.LIBDIR, "ab"); os.makedirs(out, exist_ok=True)
for spec in sys.argv[1:]:
    tag, _, defs = spec.partition(":")
    defs = [d for d in defs.split(",") if d]
    objs = []
    procs = []
    for src in b.SOURCES:
        obj = os.path.join(out, f"{tag}_{src.replace('.cu', '.o')}")
        procs.append(subprocess.Popen([b._nvcc(), *b.FLAGS, *defs, "-c", os.path.join(b.CSRC, src), "-o", obj]))
        objs.append(obj)
    assert all(p.wait() == 0 for p in procs)
    subprocess.check_call([b._nvcc(), "-shared", "-o", os.path.join(out, f"{tag}.so"), *objs, "-Xcompiler", "-fvisibility=hidden"])
    for o in objs: os.remove(o)
    print("built", tag, defs)
