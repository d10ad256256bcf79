"""Print selected raw metrics of an .ncu-rep: python scripts/ncu_raw.py file.ncu-rep [substring ...]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
subs = sys.argv[2:] or ["gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum ",
                        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum ", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum ",
                        "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum ", "sm__cycles_active.avg ", "sm__warps_active.avg.pct_of_peak_sustained_active",
                        "launch__registers_per_thread", "launch__grid_size", "smsp__average_warps_issue_stalled", "smsp__issue_active.avg.pct",
                        "l1tex__lsu_writeback_active.avg.pct", "l1tex__data_bank", "sm__inst_executed_pipe_lsu", "lts__throughput.avg.pct", "l1tex__throughput.avg.pct",
                        "sm__throughput.avg.pct", "smsp__warp_issue_stalled"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
for row in rows[2:]:
    print("==", row[hdr.index("Kernel Name")][:90], row[hdr.index("Grid Size")] if "Grid Size" in hdr else "")
    for i, h in enumerate(hdr):
        if any((h + " ").startswith(s) or s in h for s in subs):
            print("   %-90s %-14s %s" % (h, units[i], row[i]))
