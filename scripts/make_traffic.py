"""Dev helper: profiles/traffic.json from an `ncu --set full` capture of the splat kernels (bench.py --batch 64): DRAM bytes
(dram__bytes_read.sum + dram__bytes_write.sum) per launch divided by the samples of the launch.
Usage: python scripts/make_traffic.py gpurun_out/<tag>/prof_splat.ncu-rep <samples> <tag>"""
import csv, io, json, subprocess, sys
rep, samples, tag = sys.argv[1], int(sys.argv[2]), sys.argv[3]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
SC = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
per = {}
for r in rows[2:]:
    name = r[ix["Kernel Name"]]
    short = ("splat_fwd_tma" if "splat_fwd_tma" in name else "splat_bwd_stp" if "splat_bwd_stp" in name else "splat_bwd_st" if "splat_bwd_st<" in name
             else "splat_bwd_tma" if "splat_bwd_tma" in name else None)
    if not short:
        continue
    b = sum(float(r[ix[m]]) * SC[units[ix[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    per[short] = b / samples
json.dump({"source": f"ncu --set full capture of one launch at {samples} samples by scripts/gpu_round.sh (profiles/{tag}/ncu_splat_summary.txt): "
                     "dram__bytes_read.sum + dram__bytes_write.sum, divided by the samples of the launch",
           "per_sample_bytes": per}, open("profiles/traffic.json", "w"), indent=1)
print(per)
