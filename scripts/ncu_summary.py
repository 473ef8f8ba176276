"""Print the handful of metrics that matter from an .ncu-rep (raw page)."""
import csv, subprocess, sys
keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg ", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct", "sm__ops_path_tensor_src_fp16_dst_fp32.sum ", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum ",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum ", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum ",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum ", "sm__inst_executed.sum.per_cycle_elapsed", "smsp__issue_active.avg.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread ", "launch__shared_mem_per_block_dynamic",
        "smsp__average_warps_issue_stalled", "lts__t_bytes.sum ", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "l1tex__lsu_writeback_active", "l1tex__data_pipe_lsu_wavefronts.sum ",
        "smsp__inst_executed.sum ", "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum ", "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum "]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for row in rows[2:]:
    print("===", row[hdr.index("Kernel Name")][:40], row[hdr.index("Grid Size")], row[hdr.index("Block Size")])
    for i, h in enumerate(hdr):
        if any(k in h + " " for k in keys) and "peak_sustained " not in h + " " and not h.endswith(".max") and not h.endswith(".min"):
            if "issue_stalled" in h and float(row[i].replace(",", "") or 0) < 0.3:
                continue
            print(f"  {h:95s} {row[i]:>18s} {units[i]}")
