"""Markdown summary of ncu reports for profiles/README.md.

    python tools/ncu_summary.py launches <launch_list.csv>      # per-kernel shares of a launch list
    python tools/ncu_summary.py kernel <report.ncu-rep>         # key metrics of the first kernel in a report
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            h, start = r, i + 1
            break
    kn, mv, mn = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[start:]:
        if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[kn]).replace("void ", "").replace("rmb::", "")
        name = re.sub(r"<.*", "", name) if name.startswith("at::") else name
        agg[name][0] += 1
        agg[name][1] += float(r[mv].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
        print(f"| `{k[:48]}` | {v[0]} | {v[1] / 1e3:.1f} | {v[1] / v[0] / 1e3:.1f} | {100 * v[1] / tot:.1f} % |")


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units, v = rows[0], rows[1], rows[2]
    print(f"kernel: `{v[h.index('Kernel Name')][:100]}`\n\n| metric | unit | value |\n|---|---|---|")
    for k in KEYS:
        if k in h:
            i = h.index(k)
            print(f"| {k} | {units[i]} | {v[i]} |")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
