"""Markdown summary of the launches in an .ncu-rep (`ncu --set full` capture): the handful of metrics the roofline
discussion in profiles/ uses.  usage: ncu_summary.py file.ncu-rep [file2 ...]"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("duration us", "gpu__time_duration.sum"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("regs/thread", "launch__registers_per_thread"),
    ("dyn smem/block B", "launch__shared_mem_per_block_dynamic"),
    ("SM busy % (sm__throughput)", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("tensor pipe active % of elapsed", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
    ("SM cycles active avg", "sm__cycles_active.avg"),
    ("XU (MUFU) pipe % of elapsed", "sm__inst_executed_pipe_xu_realtime.avg.pct_of_peak_sustained_elapsed"),
    ("FMA pipe %", "sm__inst_executed_pipe_fma_realtime.avg.pct_of_peak_sustained_elapsed"),
    ("ALU pipe %", "sm__inst_executed_pipe_alu_realtime.avg.pct_of_peak_sustained_elapsed"),
    ("FP64 pipe %", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_elapsed"),
    ("issue slots busy %", "smsp__issue_active.avg.pct_of_peak_sustained_elapsed"),
    ("warps active % of peak", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("instructions executed", "smsp__inst_executed.sum"),
    ("DRAM read B", "dram__bytes_read.sum"),
    ("DRAM write B", "dram__bytes_write.sum"),
    ("L2 bytes (lts__t_bytes)", "lts__t_bytes.sum"),
    ("L1/TEX hit %", "l1tex__t_sector_hit_rate.pct"),
    ("smem bank conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    ("stall long_scoreboard (warp-cycles/issue)", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    ("stall short_scoreboard", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
    ("stall barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
    ("stall membar", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"),
    ("stall wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
    ("stall math pipe throttle", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
    ("stall mio throttle", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"),
    ("stall lg throttle", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"),
    ("stall sleeping / branch resolving", "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio"),
    ("stall not selected", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"),
    ("stall no instruction", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"),
]

for path in sys.argv[1:]:
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3:
        print(f"## {path}: no launches\n")
        continue
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    name_i = col.get("Kernel Name")
    print(f"## {path}\n")
    names = [r[name_i].split("(")[0].replace("<unnamed>::", "") for r in rows[2:]]
    print("| metric | " + " | ".join(f"launch {i}: `{n}`" for i, n in enumerate(names)) + " |")
    print("|---|" + "---|" * len(names))
    for label, key in KEYS:
        hit = [h for h in hdr if h.endswith(key)]
        if not hit:
            continue
        i = col[hit[0]]
        vals = []
        for r in rows[2:]:
            v = r[i]
            try:
                f = float(v.replace(",", ""))
                v = f"{f:,.0f}" if abs(f) >= 1000 else f"{f:.3g}"
            except ValueError:
                pass
            vals.append(v)
        print(f"| {label} [{units[i]}] | " + " | ".join(vals) + " |")
    print()
