"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the few numbers DESIGN.md cites.
    python profiles/ncu_summary.py gpurun_out/x.ncu-rep [--md]
"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
W = [("time_us", "gpu__time_duration.sum"), ("dram_rd_MB", "dram__bytes_read.sum"), ("dram_wr_MB", "dram__bytes_write.sum"),
     ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
     ("sm_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
     ("issue_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
     ("warp_inst", "smsp__inst_executed.sum"),
     ("occ_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"), ("regs", "launch__registers_per_thread"),
     ("l1_hit", "l1tex__t_sector_hit_rate.pct"), ("l2_hit", "lts__t_sector_hit_rate.pct"),
     ("ld_sectors", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"), ("ld_req", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"),
     ("smem_wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
     ("fp64_pct", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
     ("xu_pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
     ("fma_pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
     ("alu_pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
     ("lsu_pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
     ("l1tex_pct", "l1tex__throughput.avg.pct_of_peak_sustained_active"),
     ("lts_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
     ("grid", "launch__grid_size"), ("block", "launch__block_size")]
seen = set()
for r in data:
    name = r[idx["Kernel Name"]]
    key = name
    if key in seen:
        continue
    seen.add(key)
    print("## " + name[:110])
    for short, m in W:
        if m in idx:
            v = r[idx[m]]
            try:
                v = f"{float(v):,.3f}".rstrip("0").rstrip(".")
            except ValueError:
                pass
            print(f"  {short:16s} {v:>20s} {units[idx[m]]}")
