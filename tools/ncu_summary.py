#!/usr/bin/env python3
"""Condenses an .ncu-rep (ncu --set full) into the metrics DESIGN.md / bench.py cite.  usage: ncu_summary.py rep out.txt [traffic.json]"""
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__cycles_elapsed.max", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_set_accesses_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed", "l1tex__f_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_sectors.sum",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed_op_global_ld.sum"]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = []
    traffic = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        lines.append("== %s  grid %s block %s" % (d.get("Kernel Name"), d.get("Grid Size"), d.get("Block Size")))
        for i, h in enumerate(hdr):
            if h in KEYS or (h.startswith(STALL) and h.endswith("_per_issue_active.ratio")):
                lines.append("%-90s %-14s %s" % (h, units[i], r[i]))
        try:
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            rd = float(d["dram__bytes_read.sum"]) * mult[units[hdr.index("dram__bytes_read.sum")]]
            wr = float(d["dram__bytes_write.sum"]) * mult[units[hdr.index("dram__bytes_write.sum")]]
            traffic.append(rd + wr)
        except Exception:
            pass
    open(out, "w").write("\n".join(lines) + "\n")
    if len(sys.argv) > 3 and traffic:
        json.dump({"dram_bytes_per_launch": sum(traffic) / len(traffic), "launches": len(traffic), "source": rep}, open(sys.argv[3], "w"))


if __name__ == "__main__":
    main()
