#!/usr/bin/env python
"""Turn gpurun_out/*.csv / *.ncu-rep into the small text summaries committed under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_v0.csv > profiles/r1_v0_launches.txt
    python profiles/summarize.py kernel   gpurun_out/prof_site_v0.ncu-rep > profiles/r1_v0_site_kernel.txt
"""
import csv
import subprocess
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "lts__t_bytes.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__inst_executed_pipe_tensor_op_dmma.sum", "sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "sm__icc_request_hit_rate.pct",
        "sm__icc_requests.sum", "gcc__cache_requests_type_instruction.sum",
        "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed", "gcc__average_cache_request_hit_rate.pct",
        "sm__inst_executed.avg.per_cycle_elapsed"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ik, iv, iu = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[hdr + 1:]:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1.0)
        tot[r[ik]] += v
        cnt[r[ik]] += 1
    T = sum(tot.values())
    print(f"# ncu launch list ({path}): per-kernel device time, cold-cache / serialised -> compare SHARES")
    print(f"{'total us':>12} {'n':>5} {'avg us':>10} {'share':>8}  kernel")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"{v:12.1f} {cnt[k]:5d} {v / cnt[k]:10.1f} {100 * v / T:7.2f}%  {k[:110]}")


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H, U = rows[0], rows[1]
    print(f"# ncu --set full ({path})")
    for v in rows[2:]:
        print("## kernel:", v[H.index("Kernel Name")][:120])
        for i, h in enumerate(H):
            if h in KEYS:
                print(f"{h:95s} {U[i]:14s} {v[i]}")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
