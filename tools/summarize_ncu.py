"""Summarise an `ncu --set full` report of one steady-state step (tools/profile_step.py) into
markdown + a small JSON of per-kernel DRAM traffic that bench.py reports as roofline.traffic.

    python tools/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r01_ncu_full_final.md profiles/ncu_traffic.json
"""
import csv
import io
import json
import subprocess
import sys

rep, out_md, out_json = sys.argv[1:4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units = rows[0], rows[1]
col = {n: i for i, n in enumerate(h)}
METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__block_size",
    "launch__grid_size", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


agg = {}
order = []
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("dabgpu::", "")
    key = (name, r[col["Grid Size"]])
    if key not in agg:
        agg[key] = []
        order.append(key)
    agg[key].append(r)

traffic = {}
with open(out_md, "w") as f:
    f.write(f"# ncu --set full --clock-control none, one steady-state step (2 TF, 1024 streams), source: {rep}\n")
    f.write("# per kernel and grid: first launch of the step; `launches` = launches of that kernel in the step\n\n")
    for key in order:
        rs = agg[key]
        r = rs[0]
        f.write(f"## {key[0]}  grid={key[1]}  launches={len(rs)}\n")
        for m in METRICS:
            if m in col:
                f.write(f"  {m:86s} {r[col[m]]:>16s} {units[col[m]]}\n")
        f.write("\n")
        rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
        wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
        def num(m):
            try:
                return float(r[col[m]].replace(",", ""))
            except (KeyError, ValueError):
                return None
        dur = num("gpu__time_duration.sum")
        if units[col["gpu__time_duration.sum"]] in ("us", "usecond"):
            dur = dur / 1e3
        elif units[col["gpu__time_duration.sum"]] in ("ns", "nsecond"):
            dur = dur / 1e6
        traffic.setdefault(key[0], []).append({
            "grid": key[1], "dram_bytes_per_launch": rd + wr, "duration_ms_under_ncu": dur,
            "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "alu_pipe_pct": num("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
            "fma_pipe_pct": num("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
            "lsu_pipe_pct": num("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
            "smem_data_pipe_pct": num("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
            "dram_pct": num("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "warps_active_pct": num("sm__warps_active.avg.pct_of_peak_sustained_active"),
            "registers": num("launch__registers_per_thread"),
            "smem_bank_conflicts": num("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
            "warp_instructions": num("smsp__inst_executed.sum")})
json.dump({"source": rep, "kernels": traffic}, open(out_json, "w"), indent=1)
print("wrote", out_md, out_json)
