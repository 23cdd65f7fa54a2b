#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small CSV of the metrics the design discussion uses: one row per profiled
launch.  Usage: ncu_summary.py report.ncu-rep out.csv"""
import csv
import subprocess
import sys

WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg", "sm__cycles_elapsed.max",
        "smsp__inst_executed.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum", "lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]

    def find(name):
        if name in hdr:
            return hdr.index(name)
        for i, h in enumerate(hdr):
            if h.endswith("." + name):
                return i
        return -1

    cols = [(w, find(w)) for w in WANT]
    with open(out, "w", newline="") as f:
        wr = csv.writer(f)
        wr.writerow(["launch"] + [w for w, i in cols if i >= 0])
        wr.writerow(["unit"] + [units[i] for w, i in cols if i >= 0])
        for k, r in enumerate(rows[2:]):
            wr.writerow([k] + [r[i] for w, i in cols if i >= 0])
    for k, r in enumerate(rows[2:]):
        g = lambda w: r[find(w)] if find(w) >= 0 else "?"   # noqa: E731
        print(f"[{k}] {g('Kernel Name')[:90]}\n     {g('gpu__time_duration.sum')} {units[find('gpu__time_duration.sum')]}  dram R {g('dram__bytes_read.sum')} W {g('dram__bytes_write.sum')} "
              f"{units[find('dram__bytes_read.sum')]}  L2hit {g('lts__t_sector_hit_rate.pct')}  regs {g('launch__registers_per_thread')}  "
              f"warps {g('sm__warps_active.avg.pct_of_peak_sustained_active')}%  l1tex {g('l1tex__throughput.avg.pct_of_peak_sustained_elapsed')}%")


if __name__ == "__main__":
    main()
