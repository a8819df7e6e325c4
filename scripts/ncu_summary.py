"""Summarise an .ncu-rep into the handful of counters DESIGN.md / bench.py quote (run in the build container)."""
import csv, subprocess, sys
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_bytes.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_active.avg", "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_uniform.sum", "smsp__cycles_active.avg"]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("==", rep.split("/")[-1], "|", r[hdr.index("Kernel Name")][:90])
        for w in WANT:
            for i, h in enumerate(hdr):
                if h == w:
                    print("   %-70s %s %s" % (h, r[i], units[i]))
