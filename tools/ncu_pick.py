#!/usr/bin/env python3
"""Developer tool: prints the metrics that matter for this repository's kernels from an .ncu-rep (raw page, csv).
    python tools/ncu_pick.py gpurun_out/x.ncu-rep [extra-regex]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
extra = sys.argv[2] if len(sys.argv) > 2 else None
KEYS = [r"^gpu__time_duration.sum$", r"^dram__bytes_read.sum$", r"^dram__bytes_write.sum$", r"lts__t_sector_hit_rate.pct$", r"^lts__t_sectors_srcunit_tex_op_read.sum$",
        r"^lts__t_sectors_srcunit_tex_op_red.sum$", r"^lts__t_sectors_srcunit_tex_op_write.sum$", r"l1tex__t_sector_hit_rate.pct$", r"^l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum$",
        r"^l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum$", r"^l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum$", r"^l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum$",
        r"^l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$", r"^sm__inst_executed.sum$", r"^smsp__issue_active.avg.pct_of_peak_sustained_active$", r"^sm__warps_active.avg.pct_of_peak_sustained_active$",
        r"sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active$", r"^launch__registers_per_thread$", r"^launch__grid_size$", r"^launch__block_size$", r"sm__throughput.avg.pct_of_peak_sustained_elapsed$",
        r"l1tex__throughput.avg.pct_of_peak_sustained_elapsed$", r"lts__throughput.avg.pct_of_peak_sustained_elapsed$", r"^gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed$",
        r"^lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed$", r"^smsp__average_warps_issue_stalled_.*_per_issue_active.ratio$",
        r"^l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed$", r"^l1tex__m_xbar2l1tex_read_sectors.sum$", r"^l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed$",
        r"^l1tex__m_xbar2l1tex_read_sectors.sum.pct_of_peak_sustained_elapsed$", r"^l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum$", r"^l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum$",
        r"^lts__t_sectors.sum$", r"^lts__t_sectors.avg.pct_of_peak_sustained_elapsed$", r"^l1tex__t_set_accesses_pipe_lsu_mem_global_op_ld.sum$", r"^l1tex__data_pipe_lsu_wavefronts.sum$",
        r"^l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed$", r"^l1tex__data_pipe_lsu_wavefronts_mem_shared.sum$", r"^l1tex__f_wavefronts.sum$"]
if extra:
    KEYS.append(extra)
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
for row in rows[2:]:
    name = row[hdr.index("Kernel Name")]
    print("==", name[:110])
    for i, h in enumerate(hdr):
        if any(re.search(k, h) for k in KEYS):
            v = row[i]
            try:
                f = float(v.replace(",", ""))
                if "stall" in h and f < 0.3:
                    continue
                v = f"{f:,.3f}".rstrip("0").rstrip(".")
            except ValueError:
                pass
            print(f"   {h:85s} {v} {units[i]}")
