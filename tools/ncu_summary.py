"""Summarise one kernel of an `ncu --page raw --csv` export into the text kept under profiles/.
   python tools/ncu_summary.py <raw.csv> [header line ...] > profiles/<name>.ncu.txt"""
import csv
import sys

KEYS = [
    "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__lts2xbar_cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex.avg.pct_of_peak_sustained_elapsed",
    "lts__t_tag_requests.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
    "sm__inst_executed_pipe_tensor_op_gmma.sum", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_uniform.sum",
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    for line in sys.argv[2:]:
        print("# " + line)
    for vals in rows[2:]:
        d = {h: (v, u) for h, v, u in zip(hdr, vals, units)}
        for k in KEYS:
            if k in d:
                print("%-88s %-16s %s" % (k, d[k][1], d[k][0]))
        tensor = [h for h in hdr if "tensor" in h and "pct_of_peak" in h and h.endswith("elapsed") and ".avg." in h]
        for h in tensor:
            if h not in KEYS and d[h][0] not in ("0", "", "n/a"):
                print("%-88s %-16s %s" % (h, d[h][1], d[h][0]))
        print()


if __name__ == "__main__":
    main()
