"""Summarise an .ncu-rep (ncu --set full) into a small markdown table for profiles/."""
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_elapsed.avg.per_second"]
STALL = "smsp__average_warps_issue_stalled_"
rep, out = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else rep
raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
with open(out, "w") as f:
    f.write("# %s\n\n`ncu --set full --clock-control none` ; values per launch\n\n" % title)
    for r in rows[2:]:
        f.write("## %s\n\n| metric | value | unit |\n|---|---|---|\n" % r[hdr.index("Kernel Name")].split("(")[0])
        for k in KEYS:
            if k in hdr:
                f.write("| %s | %s | %s |\n" % (k, r[hdr.index(k)], units[hdr.index(k)]))
        stalls = [(h[len(STALL):-len("_per_issue_active.ratio")], float(r[i])) for i, h in enumerate(hdr)
                  if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and r[i]]
        stalls.sort(key=lambda kv: -kv[1])
        f.write("\nTop stall reasons (warps stalled per issue-active cycle): " +
                ", ".join("%s %.2f" % kv for kv in stalls[:7]) + "\n\n")
print(open(out).read()[:1500])
