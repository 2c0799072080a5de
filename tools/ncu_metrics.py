"""Print the key metrics of an `ncu --page raw --csv` export (one column per profiled launch)."""
import csv
import sys

KEYS = ["Kernel Name", "Grid Size", "gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor",
        "sm__pipe_tensor_subpipe", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct",
        "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__throughput.avg.pct", "l1tex__throughput.avg.pct",
        "l1tex__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ldgsts", "l1tex__t_bytes",
        "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit", "launch__shared_mem_per_block", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "sm__inst_executed.sum", "smsp__warp_issue_stalled"]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    units = rows[1]
    data = rows[2:]
    for i, name in enumerate(hdr):
        if any(k in name for k in KEYS):
            vals = [r[i] for r in data]
            print("%-86s %-10s %s" % (name[:86], units[i][:10], "  ".join(v[:22] for v in vals)))


if __name__ == "__main__":
    main(sys.argv[1])
