"""Extract the metrics the roofline argument needs from .ncu-rep files into a markdown table."""
import csv
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "LSU pipe %"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long-sb"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short-sb"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]
print("| report | kernel | " + " | ".join(k[1] for k in KEYS) + " |")
print("|---|---|" + "---|" * len(KEYS))
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("alignn::", "")[:48]
        vals = []
        for k, _ in KEYS:
            if k in hdr:
                i = hdr.index(k)
                v = r[i]
                try:
                    v = f"{float(v.replace(',', '')):.4g}"
                except ValueError:
                    pass
                vals.append(f"{v} {units[i]}".strip())
            else:
                vals.append("-")
        print(f"| {rep.split('/')[-1]} | {name} | " + " | ".join(vals) + " |")
