#!/usr/bin/env python3
"""Turns ncu output into the text summaries committed under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches_r01h.csv          # per-kernel totals of a launch list
    python tools/ncu_summary.py kernel gpurun_out/resolve_streams_r01h.ncu-rep  # selected raw-page metrics of a capture
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

METRICS = """gpu__time_duration.sum launch__grid_size launch__block_size launch__registers_per_thread
launch__shared_mem_per_block_dynamic launch__occupancy_limit_registers launch__occupancy_limit_shared_mem
sm__warps_active.avg.pct_of_peak_sustained_active dram__bytes_read.sum dram__bytes_write.sum
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed lts__t_sector_hit_rate.pct smsp__inst_executed.sum
smsp__issue_active.avg.pct_of_peak_sustained_active sm__throughput.avg.pct_of_peak_sustained_elapsed
sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active
lts__throughput.avg.pct_of_peak_sustained_elapsed l1tex__throughput.avg.pct_of_peak_sustained_elapsed
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio sm__cycles_active.avg sm__cycles_elapsed.max""".split()


def launches(path):
    rows = [l for l in open(path) if l.startswith('"')]
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in csv.DictReader(io.StringIO("".join(rows))):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        ns = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] == "us":
            ns *= 1e3
        tot[r["Kernel Name"]] += ns
        cnt[r["Kernel Name"]] += 1
    all_ns = sum(tot.values())
    ours = sum(v for k, v in tot.items() if "apt::" in k)
    print(f"# total device time in capture: {all_ns / 1e3:.1f} us over {sum(cnt.values())} launches; ours {ours / 1e3:.1f} us")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        share = f"{100 * v / ours:5.1f}% of ours" if "apt::" in k else "  0.0% of ours"
        print(f"{v / 1e3:10.1f} us {cnt[k]:4d} launches {100 * v / all_ns:5.1f}% of all, {share}  {v / cnt[k] / 1e3:8.1f} us/launch  {k[:150]}")


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(head, r))
        print("Kernel Name:", d.get("Kernel Name"))
        for m in METRICS:
            if m in d:
                print(f"{m}: {d[m]} {units[head.index(m)]}")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
