#!/usr/bin/env python
"""Key metrics of `ncu --set full` captures as a markdown table + JSON.
usage: python tools/ncu_metrics_table.py out.json label=path.ncu-rep [label=path.ncu-rep ...]
The labels are bench.py workload names; bench.py reads `dram_bytes_per_launch` from the JSON for its
roofline.traffic field."""
import csv
import json
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "kernel time under ncu"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue-slot utilisation %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA-pipe (FP32) instruction utilisation %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA-pipe cycles active %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "warp execution efficiency (threads/inst)"),
    ("sm__warps_active.avg.per_cycle_active", "resident warps per SM"),
    ("launch__registers_per_thread", "registers/thread"),
    ("sm__icc_request_hit_rate.pct", "instruction-cache hit rate %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.per_second", "shared-memory wavefronts/s"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory pipe % of peak"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "FADD thread instructions"),
    ("smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "FMUL thread instructions"),
    ("smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "FFMA thread instructions"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU-pipe instruction utilisation %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
]
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (vals[i], units[i]) for i, h in enumerate(hdr)}


RAYS = {}


def main():
    out_json = sys.argv[1]
    args = sys.argv[2:]
    # rays:<label>=<n>  — rays of the captured launch (from the plain run of the same command)
    for a in [a for a in args if a.startswith("rays:")]:
        k, v = a[5:].split("=")
        RAYS[k] = int(v)
    caps = [a.split("=", 1) for a in args if not a.startswith("rays:")]
    data, table = {}, {}
    for label, path in caps:
        r = raw(path)
        table[label] = r
        rd, wr = r["dram__bytes_read.sum"], r["dram__bytes_write.sum"]
        fp32 = sum(float(r[m][0]) for m in ("smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
                                           "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
                                           "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum") if m in r)
        rays = RAYS.get(label)
        data[label] = {
            "fp32_lane_inst_per_launch": fp32 or None, "rays_per_launch": rays,
            "fp32_lane_inst_per_ray": (fp32 / rays) if fp32 and rays else None,
            "capture": path.split("/")[-1],
            "kernel": r.get("Kernel Name", ("", ""))[0],
            "dram_bytes_per_launch": float(rd[0]) * UNIT[rd[1]] + float(wr[0]) * UNIT[wr[1]],
            **{m: f"{r[m][0]} {r[m][1]}".strip() for m, _ in METRICS if m in r},
        }
    json.dump(data, open(out_json, "w"), indent=1)
    print("| metric | " + " | ".join(l for l, _ in caps) + " |")
    print("|---|" + "---|" * len(caps))
    for m, desc in METRICS:
        cells = []
        for label, _ in caps:
            v, u = table[label].get(m, ("-", ""))
            try:
                v = f"{float(v):,.4g}"
            except ValueError:
                pass
            cells.append(f"{v} {u}".strip())
        print(f"| {desc} (`{m}`) | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
