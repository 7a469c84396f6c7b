"""Condense `ncu --set full` reports into (a) a text summary per kernel under profiles/ and (b) the JSON that
bench.py reads for `roofline.traffic` / `roofline.compute` (profiles/r02_kernel_metrics.json).
Usage: ncu_summarize.py <tag> name=report.ncu-rep:kernel-substr:workload ...   (reads with `ncu -i ... --page raw --csv`)"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__average_warp_latency_issue_stalled_barrier.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def read(rep, kern):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    kcol = names.index("Kernel Name")
    cand = [r for r in rows[hdr + 2:] if len(r) == len(names) and kern in r[kcol]]
    if not cand:
        raise SystemExit(f"no kernel matching {kern} in {rep}")
    r = cand[-1]  # the last captured launch (warm)
    vals = {}
    for n, u, v in zip(names, units, r):
        try:
            vals[n] = (float(v.replace(",", "")), u)
        except ValueError:
            vals[n] = (v, u)
    return vals, len(cand)


def to_bytes(v, u):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)


def to_ms(v, u):
    return v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}.get(u, 1)


tag = sys.argv[1]
metrics = {}
mpath = os.path.join(ROOT, "profiles", "r02_kernel_metrics.json")
if os.path.exists(mpath):
    metrics = json.load(open(mpath))
for spec in sys.argv[2:]:
    name, rest = spec.split("=")
    rep, kern, workload = rest.split(":")
    v, n = read(rep, kern)
    txt = os.path.join("profiles", f"{tag}_{name}_ncu_full.txt")
    with open(os.path.join(ROOT, txt), "w") as f:
        f.write(f"# {name}: `ncu --set full --clock-control none`, workload {workload}, {n} launch(es) captured, last one summarised.\n")
        f.write(f"# source report: {os.path.basename(rep)} (gpurun_out/, scratch); per-launch values under the profiler, never a bench value.\n")
        for k in KEEP:
            if k in v:
                f.write(f"{k:95s} {v[k][0]} {v[k][1]}\n")
    g = lambda k, d=None: v[k][0] if k in v and not isinstance(v[k][0], str) else d
    rd = to_bytes(*v["dram__bytes_read.sum"]) if "dram__bytes_read.sum" in v else None
    wr = to_bytes(*v["dram__bytes_write.sum"]) if "dram__bytes_write.sum" in v else None
    metrics[name] = {
        "workload": workload, "source": f"{txt} (ncu --set full, one launch)",
        "duration_ms_under_ncu": to_ms(*v["gpu__time_duration.sum"]) if "gpu__time_duration.sum" in v else None,
        "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": (rd + wr) if rd is not None else None,
        "compute": {
            "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "warps_active_per_sm": g("sm__warps_active.avg.per_cycle_active"),
            "pipe_fp64_pct": g("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", g("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")),
            "pipe_fma_pct": g("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", g("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active")),
            "pipe_alu_pct": g("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", g("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active")),
            "pipe_xu_pct": g("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
            "pipe_lsu_pct": g("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
            "registers_per_thread": g("launch__registers_per_thread"),
            "local_memory_insts": (g("smsp__inst_executed_op_local_ld.sum", 0) or 0) + (g("smsp__inst_executed_op_local_st.sum", 0) or 0),
            "warp_instructions": g("smsp__inst_executed.sum", g("sm__inst_executed.sum")),
            "active_lanes_per_instruction": g("smsp__thread_inst_executed_per_inst_executed.ratio"),
        },
    }
    print(name, json.dumps(metrics[name]["compute"]))
json.dump(metrics, open(mpath, "w"), indent=1)
