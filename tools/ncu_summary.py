#!/usr/bin/env python
"""Summarise an ``ncu --set full`` report (read here, no GPU needed) into a small text table.

    python tools/ncu_summary.py gpurun_out/r01e_march.ncu-rep [more.ncu-rep ...] > profiles/r01_march.txt

One column per report / kernel launch, one row per metric of the list below (the ones DESIGN.md and
bench.py's roofline refer to).  ``--all`` dumps every metric instead.
"""

from __future__ import annotations

import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__thread_inst_executed_per_inst_executed.pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    # DRAM
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    # L2
    "lts__t_bytes.sum", "lts__t_sectors_op_read.sum", "lts__t_sector_hit_rate.pct",
    "lts__t_sectors_srcunit_tex_op_read.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    # L1
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum",
    "l1tex__data_pipe_lsu_wavefronts.sum",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed",
    "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum",
    "derived__l1tex__lsu_writeback_bytes_mem_lgds.sum.per_second",
    # stalls (warp cycles per issued instruction)
    "smsp__average_warp_latency_per_inst_issued.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio",
]


def read_report(path: str):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True)
    if out.returncode != 0:
        raise SystemExit(f"ncu -i {path} failed: {out.stderr[-400:]}")
    rows = list(csv.reader(io.StringIO(out.stdout)))
    header, units = rows[0], rows[1]
    launches = []
    for vals in rows[2:]:
        d = {}
        for h, u, v in zip(header, units, vals):
            d[h] = (v, u)
        launches.append(d)
    return launches


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    dump_all = "--all" in sys.argv
    cols = []
    for path in args:
        for i, launch in enumerate(read_report(path)):
            name = launch.get("Kernel Name", ("?", ""))[0]
            cols.append((f"{path.split('/')[-1]}#{i}", name, launch))
    for tag, name, _ in cols:
        print(f"# {tag}: {name}")
    keys = METRICS
    if dump_all:
        keys = sorted({k for _, _, l in cols for k in l})
    width = max(len(k) for k in keys) + 2
    print(f"{'metric':{width}s}" + "".join(f"{t[:34]:>36s}" for t, _, _ in cols))
    for k in keys:
        cells = []
        for _, _, launch in cols:
            v, u = launch.get(k, ("", ""))
            cells.append(f"{(v + (' ' + u if u and v else '')):>36s}")
        if any(c.strip() for c in cells):
            print(f"{k:{width}s}" + "".join(cells))


if __name__ == "__main__":
    main()
