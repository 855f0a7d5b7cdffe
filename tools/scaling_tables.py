#!/usr/bin/env python
"""Markdown tables of profiles/README.md §1 from the committed bench lines (profiles/scaling_<tag>/bench_n{1,2,4,8}_*.json).
usage: python tools/scaling_tables.py [tag]"""
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
D = {}
for n in (1, 2, 4, 8):
    f = glob.glob(os.path.join(ROOT, "profiles", f"scaling_{tag}", f"bench_n{n}_*.json"))
    if f:
        D[n] = json.load(open(f[0]))
b = D[1]
MB = {c: v["e2e"]["d2h_bytes_per_step"] / 1e6 for c, v in b["configs"].items()}
print("| N | kernel Grays/s (ms) | speed-up | e2e Grays/s (frames/s) | e2e speed-up | `.txt` 8K frame ms | `.txt` 4K frame ms |\n|---|---|---|---|---|---|---|")
for n, d in D.items():
    print(f"| {n} | {d['value'] / 1e3:.1f} ({d['ms_per_step']:.2f}) | {d['value'] / b['value']:.2f} | {d['e2e']['value'] / 1e3:.1f} "
          f"({d['e2e']['frames_per_s']:.0f}) | {d['e2e']['value'] / b['e2e']['value']:.2f} | {d['e2e_txt']['ms_per_frame']:.0f} | "
          f"{d['e2e_txt_4k']['ms_per_frame']:.0f} |")
print("\n| config | " + " | ".join(f"N={n}" + (": Grays/s (ms)" if n == 1 else "") for n in D) + " | speed-up at 8 |\n|---|" + "---|" * (len(D) + 1))
for c in b["configs"]:
    row = [f"{D[n]['configs'][c]['mrays_s'] / 1e3:.1f} ({D[n]['configs'][c]['ms']:.3f})" for n in D]
    last = max(D)
    print(f"| {c} | " + " | ".join(row) + f" | {D[last]['configs'][c]['mrays_s'] / b['configs'][c]['mrays_s']:.2f} |")
print("\n| config | " + " | ".join(f"N={n}" + (": e2e frames/s" if n == 1 else "") for n in D) + " | frame bytes x frames/s at N = " +
      " / ".join(str(n) for n in D) + " |\n|---|" + "---|" * (len(D) + 1))
for c in b["configs"]:
    fps = [D[n]["configs"][c]["e2e"]["frames_per_s"] for n in D]
    print(f"| {c} | " + " | ".join(f"{v:.0f}" for v in fps) + " | " + " / ".join(f"{v * MB[c] / 1e3:.0f}" for v in fps) + " GB/s |")
print("\nchecks:", {n: (d["parity"]["ok"], d["parity"]["all_configs_ok"], (d.get("multi_device_ctx") or {}).get("ok"),
                       d["e2e_txt_4k"].get("md5_equals_reference_file"), d["clocks"]["sm_mhz"], d["clocks"]["reasons"]) for n, d in D.items()})
print("e2e_multiframe_1080p frames/s:", [round(d["e2e_multiframe_1080p"]["frames_per_s"]) for d in D.values()])
print("e2e_async frames/s:", [round(d["e2e_async"]["frames_per_s"], 1) for d in D.values()])
bal = D[max(D)].get("balance_max_over_mean_per_iteration") or []
if bal and isinstance(bal[0], list):
    bal = bal[0]
print("balance (max/mean per iteration) at the largest N:", [round(x, 3) for x in bal[:6]])
r = glob.glob(os.path.join(ROOT, "profiles", f"scaling_{tag}", "bench_reference_arm_*.json"))
if r:
    r = json.load(open(r[0]))
    print(f"reference arm: {r['value']:.2f} Mrays/s on {r['cpu_baseline']['cores']} cores, single thread {r['cpu_baseline']['single_thread_us_per_pixel']:.2f} us/pixel;"
          f" kernel {b['value'] / r['value']:.0f}x, e2e {b['e2e']['value'] / r['value']:.0f}x")
print("roofline N=1:", {k: b["roofline"][k] for k in ("frac", "frac_algorithmic", "fp32_lane_inst_per_ray", "capture_issue_slot_pct", "capture_threads_per_inst") if k in b["roofline"]})
