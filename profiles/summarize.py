#!/usr/bin/env python
"""Text summary of an ncu report for profiles/: key raw metrics + instructions per bin by source region.
    python profiles/summarize.py gpurun_out/prof.ncu-rep [bins_per_launch] > profiles/<name>.txt"""
import csv
import subprocess
import sys

rep = sys.argv[1]
bins = float(sys.argv[2]) if len(sys.argv) > 2 else 524288.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__icc_request_hit_rate.pct", "lts__t_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
for r in rows[2:]:
    print("== launch ==")
    for k in KEYS:
        if k in hdr:
            print("%-70s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
    try:
        rd = float(r[hdr.index("dram__bytes_read.sum")])
        wr = float(r[hdr.index("dram__bytes_write.sum")])
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot = rd * scale[units[hdr.index("dram__bytes_read.sum")]] + wr * scale[units[hdr.index("dram__bytes_write.sum")]]
        print("%-70s %.0f bytes = %.1f B per bin" % ("dram traffic per launch (read+write)", tot, tot / bins))
        n = float(r[hdr.index("smsp__inst_executed.sum")])
        print("%-70s %.1f" % ("warp-instructions per bin", n / bins))
    except Exception as e:
        print("(derived metrics unavailable: %s)" % e)
    print("warp stall reasons per issue (> 0.2):")
    for k in hdr:
        if "issue_stalled" in k and "per_issue_active" in k:
            v = float(r[hdr.index(k)] or 0)
            if v > 0.2:
                print("    %-40s %.2f" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
print()
print(subprocess.run([sys.executable, __file__.replace("summarize.py", "ncu_regions.py"), rep, str(bins)],
                     capture_output=True, text=True).stdout)
