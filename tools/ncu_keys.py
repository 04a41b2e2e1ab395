"""Key metrics of an .ncu-rep (one launch) as a markdown table: python tools/ncu_keys.py rep.ncu-rep [title]"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines()))
h, u = r[0], r[1]
pats = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "hmma_cycles_active_realtime.avg",
        "sm__cycles_elapsed.avg", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "sm__inst_executed_pipe_tma.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__t_sectors_op_red.sum", "lts__t_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__warp_issue_stalled", "lts__t_sectors_srcunit_tex_op_write.sum", "lts__d_sectors_fill", "dram__cycles_active.avg.pct"]
print(f"# {sys.argv[2] if len(sys.argv) > 2 else rep}\n")
for row in r[2:]:
    print("| metric | unit | value |\n|---|---|---|")
    for i, k in enumerate(h):
        if any(p == k or (p in k and len(p) > 12) for p in pats) and "stalled" not in k:
            print(f"| {k} | {u[i]} | {row[i]} |")
    st = [(float(row[i]), k) for i, k in enumerate(h) if "smsp__average_warp_latency_issue_stalled" in k or "smsp__average_warps_issue_stalled" in k and row[i]]
    for v, k in sorted(st, reverse=True)[:8]:
        print(f"| {k} |  | {v} |")
