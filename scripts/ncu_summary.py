"""Dev helper: compact per-kernel summary of an .ncu-rep (`ncu --page raw --csv`) -> text for profiles/.
Usage: python scripts/ncu_summary.py gpurun_out/x/prof.ncu-rep > profiles/x_summary.txt"""
import csv, io, subprocess, sys
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg.per_second"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print("kernel:", r[ix["Kernel Name"]])
    for w in WANT:
        if w in ix:
            print(f"  {w:70s} {r[ix[w]]:>16s} {units[ix[w]]}")
    SC = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    rd = float(r[ix["dram__bytes_read.sum"]]) * SC[units[ix["dram__bytes_read.sum"]]]
    wr = float(r[ix["dram__bytes_write.sum"]]) * SC[units[ix["dram__bytes_write.sum"]]]
    t, ut = float(r[ix["gpu__time_duration.sum"]]), units[ix["gpu__time_duration.sum"]]
    print(f"  traffic (read+write) = {(rd + wr) / 1e9:.4f} Gbyte per launch, duration {t} {ut} (under ncu: cold cache, serialised)")
