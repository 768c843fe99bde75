"""Key metrics per captured launch from `ncu -i X.ncu-rep --page raw --csv > file` (argv[1] = that file)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("launch__cluster_dim_x", "cluster"),
        ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_pipe_pct"),
        ("smsp__sass_inst_executed_op_utcmma.sum", "utcmma_inst"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lsu_data_pipe_pct"),
        ("l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "lsu_writeback_pct"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1_pct"),
        ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct")]
for r in rows[2:]:
    out = []
    for k, short in want:
        if k in idx:
            v = r[idx[k]]
            if short == "kernel":
                v = v.split("(")[0][:60]
            out.append("%s=%s%s" % (short, v, ("" if units[idx[k]] in ("", "%") else " " + units[idx[k]])))
    print("  ".join(out))
