#!/usr/bin/env python
"""Developer check of tcrt_balance_columns on ONE GPU: kernel time of every band of the cut, next to
the equal-width cut.  usage: python tools/band_balance_check.py [workload] [n_bands]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS  # noqa: E402
from tilecoderaytracer_b200 import api  # noqa: E402
from tilecoderaytracer_b200.partition import column_bands  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "default_1080p_d5"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
scene_name, w, h, d = WORKLOADS[name]
cam = api.Camera()
scene = api.Scene().build(scene_name, cam)
ctx = api.Context([0])
ctx.upload(scene, cam)
p = api.default_params(w, h, d)
full = min(ctx.render_device(p).render_ms[0] for _ in range(4))
for label, bands in (("balanced", ctx.balance_columns(p, n)), ("equal", column_bands(w, n))):
    ms = []
    for x0, x1 in bands:
        ms.append(min(ctx.render_device(p, x0, x1).render_ms[0] for _ in range(4)))
    print(f"{name} n={n} {label:9s} full {full:.3f} ms  bands {bands}  ms {[round(m, 4) for m in ms]}  "
          f"max {max(ms):.4f}  speed-up {full / max(ms):.2f}x  sum/full {sum(ms) / full:.3f}", flush=True)
# feedback refinement from measured band times (what bench.py does during warm-up at N > 1)
from tilecoderaytracer_b200.partition import rebalance  # noqa: E402

bands = ctx.balance_columns(p, n)
for it in range(4):
    ms = [min(ctx.render_device(p, x0, x1).render_ms[0] for _ in range(3)) for x0, x1 in bands]
    print(f"{name} n={n} feedback{it} ms {[round(m, 4) for m in ms]} max {max(ms):.4f} speed-up {full / max(ms):.2f}x "
          f"sum/full {sum(ms) / full:.3f}", flush=True)
    bands = rebalance(bands, ms, w)
ctx.close()
