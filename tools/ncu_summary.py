#!/usr/bin/env python
"""Condense an ncu report (--set full) into the few numbers the roofline discussion uses.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [more.ncu-rep ...] > profiles/x.txt

One block per profiled launch: duration, registers, occupancy limiters, issue / FMA / LSU
pipe utilisation, shared-memory wavefronts and bank conflicts, DRAM and L2 traffic, and
the warp-stall breakdown (stall cycles per issued instruction).
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__occupancy_limit_registers", "occ limit regs (blocks)"),
    ("launch__occupancy_limit_shared_mem", "occ limit smem (blocks)"),
    ("launch__occupancy_limit_warps", "occ limit warps (blocks)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe inst %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "FMA pipe cycles active %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe inst %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe inst %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem wavefronts % of peak"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "FFMA thread inst"),
    ("smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "FADD thread inst"),
    ("smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "FMUL thread inst"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_sectors_op_red.sum", "L2 red sectors"),
    ("lts__t_sectors_op_atom.sum", "L2 atom sectors"),
    ("lts__t_sectors.sum", "L2 sectors"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print(f"# {rep}: no launches")
            continue
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            print(f"## {rep}  launch id {r[col['ID']]}: {r[col['Kernel Name']]}")
            for key, label in KEYS:
                if key in col and r[col[key]] != "":
                    print(f"  {label:32s} {r[col[key]]} {units[col[key]]}")
            stalls = []
            for h, i in col.items():
                if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and r[i] not in ("", "0"):
                    stalls.append((float(r[i]), h[len(STALL):-len("_per_issue_active.ratio")]))
            stalls.sort(reverse=True)
            print("  stalls (warp-cycles per issued inst): " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:8]))
            print()


if __name__ == "__main__":
    main()
