#!/usr/bin/env python
"""Condensed summary of an .ncu-rep (raw page): duration, DRAM, occupancy, pipes, stall reasons.
usage: tools/ncu_summary.py file.ncu-rep [...]"""
import csv, io, subprocess, sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct2"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved_occ_pct"),
    ("sm__maximum_warps_per_active_cycle_pct", "theoretical_occ_pct"),
    ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem"), ("launch__occupancy_limit_registers", "occ_lim_regs"),
    ("launch__occupancy_limit_shared_mem", "occ_lim_smem"), ("launch__waves_per_multiprocessor", "waves"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_cycles_pct"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_pct"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed", "smem_wf_pct"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("smsp__inst_executed.sum", "inst_executed"),
    ("sm__cycles_elapsed.avg", "cycles"), ("sm__cycles_active.avg", "cycles_active"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("smsp__cycles_active.avg", "smsp_cycles_active"),
    ("sm__cycles_elapsed.avg.per_second", "sm_hz"),
    ("local_load_bytes", "x"),
]
STALL = "smsp__average_warps_issue_stalled_"

for f in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", f, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2:]
    for v in vals:
        d = dict(zip(hdr, v))
        u = dict(zip(hdr, units))
        print("==", f, "|", d.get("Kernel Name", "")[:90])
        for k, name in KEYS:
            if k in d:
                print(f"  {name:>22}: {d[k]} {u[k]}")
        st = sorted(((float(d[k].replace(',', '')), k[len(STALL):]) for k in d if k.startswith(STALL) and k.endswith("per_warp_active.pct") is False and d[k] not in ("", "n/a")), reverse=True)
        seen = 0
        for val, k in st:
            if "_not_issued" in k or "ratio" not in k:
                continue
            print(f"  stall {k:>48}: {val:.2f}")
            seen += 1
            if seen >= 8:
                break
