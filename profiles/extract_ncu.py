"""Extract the numbers bench.py and DESIGN.md quote from an `ncu --set full` report (run here, no GPU needed):
    python profiles/extract_ncu.py gpurun_out/prof.ncu-rep frames_per_launch out.csv
Writes one row per kernel with duration, DRAM bytes, instruction counts, pipe utilisation and stall ratios, and prints the
JSON fragment for profiles/dram_traffic.json."""
import csv
import io
import json
import subprocess
import sys

rep, frames, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
KEEP = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__inst_executed.sum",
        "smsp__thread_inst_executed.sum", "sm__thread_inst_executed.sum", "smsp__inst_executed.sum", "smsp__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
idx = [hdr.index(k) for k in KEEP if k in hdr]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([hdr[i] for i in idx])
    w.writerow([units[i] for i in idx])
    for r in data:
        w.writerow([r[i] for i in idx])


def val(r, k):
    return float(r[hdr.index(k)].replace(",", "")) if k in hdr else None


def scale(unit):
    return {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)


frag = {}
for r in data:
    name = r[hdr.index("Kernel Name")]
    rd = val(r, "dram__bytes_read.sum") * scale(units[hdr.index("dram__bytes_read.sum")])
    wr = val(r, "dram__bytes_write.sum") * scale(units[hdr.index("dram__bytes_write.sum")])
    key = "cascade_kernel" if "cascade" in name else ("level_kernel" if "level" in name else name)
    frag[key] = {"bytes_per_frame": int((rd + wr) / frames),
                 "source": f"{name.split('(')[0]}: {rd / 1e6:.2f} MB read + {wr / 1e6:.2f} MB written for {frames} frames ({rep.split('/')[-1]})",
                 "ncu_issue_slots_busy_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                 "ncu_warp_instructions_per_frame": int((val(r, "sm__inst_executed.sum") or val(r, "smsp__inst_executed.sum") or 0) / frames)}
    ti = val(r, "smsp__thread_inst_executed.sum") or val(r, "sm__thread_inst_executed.sum")
    if ti:
        frag[key]["ncu_thread_instructions_per_frame"] = int(ti / frames)
    w_, c_ = val(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"), val(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")
    if w_:
        frag[key]["ncu_smem_wavefronts_conflict_pct"] = round(100 * c_ / w_, 1)
print(json.dumps(frag, indent=1))
