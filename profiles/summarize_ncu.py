#!/usr/bin/env python
"""Turn an `ncu --set full` report into the small JSON summaries committed under profiles/.

    python profiles/summarize_ncu.py gpurun_out/prof_c4_v11.ncu-rep profiles/r1_v11_c4_ncu_full_summary.json [--traffic c4@1]

Reads the report with `ncu -i <rep> --page raw --csv` (ncu is in the image; no GPU needed), keeps the metrics the roofline
discussion in DESIGN.md uses, and with --traffic KEY rewrites profiles/traffic.json[KEY] with
dram__bytes_read.sum + dram__bytes_write.sum per launch (what bench.py reports as roofline.traffic).
"""
import csv
import io
import json
import os
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__inst_executed_pipe_tex.sum",
]
# kernel name fragment -> role key of profiles/traffic.json (what bench.py sums per stage)
ROLES = [("march_kernel", "march_kernel"), ("classify", "classify"), ("scatter_kernel", "scatter"), ("shade_sorted", "shade_sorted"), ("shade_kernel", "shade_kernel"),
         ("blend_irradiance", "blend_irradiance"), ("blend_depth", "blend_depth")]
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    traffic_key = sys.argv[sys.argv.index("--traffic") + 1] if "--traffic" in sys.argv else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {name: i for i, name in enumerate(hdr)}
    summary, traffic = [], {}
    for r in rows[2:]:
        entry = {"Kernel Name": r[col["Kernel Name"]]}
        for k in KEEP:
            if k in col:
                entry[k] = f"{r[col[k]]} {units[col[k]]}".strip()
        summary.append(entry)
        if traffic_key:
            b = 0.0
            for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                b += float(r[col[k]].replace(",", "")) * UNIT_SCALE.get(units[col[k]], 1.0)
            name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0].replace("lux::", "")
            role = next((r for key, r in ROLES if key in name), name)
            traffic[role] = traffic.get(role, 0.0) + b
    json.dump(summary, open(out, "w"), indent=1)
    print("wrote", out, len(summary), "kernels")
    if traffic_key:
        p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic.json")
        t = json.load(open(p)) if os.path.exists(p) else {}
        t[traffic_key] = traffic
        t["_source"] = f"{os.path.basename(out)} (ncu --set full --clock-control none): dram__bytes_read.sum + dram__bytes_write.sum per launch"
        json.dump(t, open(p, "w"), indent=1)
        print("traffic", traffic_key, traffic)


if __name__ == "__main__":
    main()
