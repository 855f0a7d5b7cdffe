#!/usr/bin/env python
"""Developer tool: where an end-to-end frame's time goes (wall clock per call, one GPU).
usage: python tools/e2e_breakdown.py [workload] [bands]   — renders band 0 of `bands` equal bands"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from bench import WORKLOADS  # noqa: E402
from tilecoderaytracer_b200 import api  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "synth256_8k_d10"
bands = int(sys.argv[2]) if len(sys.argv) > 2 else 1
scene_name, w, h, d = WORKLOADS[name]
ctx = api.Context([0])
cam = api.Camera()
scene = api.Scene().build(scene_name, cam)
flat, camx = scene.flatten(), cam.export()
p = api.default_params(w, h, d)
x0, x1 = 0, (w // bands) // 4 * 4
hb = api.HostBuffer((x1 - x0) * h * 12)
out = hb.array(np.float32, (x1 - x0, h, 3))
pageable = np.empty((x1 - x0, h, 3), np.float32)


def timed(label, fn, n=20):
    for _ in range(3):
        fn()
    t0 = time.perf_counter()
    for _ in range(n):
        r = fn()
    ms = (time.perf_counter() - t0) / n * 1e3
    print(f"  {label:46s} {ms:8.3f} ms", flush=True)
    return r


print(f"== {name}, columns [{x0},{x1}) of {w} ({(x1 - x0) * h * 12 / 1e6:.1f} MB)")
ctx.upload_flat(flat, camx)
timed("tcrt_upload_scene (unchanged scene)", lambda: ctx.upload_flat(flat, camx))
st = timed("tcrt_render_device (one launch, no copy)", lambda: ctx.render_device(p, x0, x1))
print(f"      kernel {st.render_ms[0]:.3f} ms")
r = timed("tcrt_render_columns -> pinned host", lambda: ctx.render(p, x0, x1, out))
print(f"      kernels {r[1].render_ms[0]:.3f} ms, copy not hidden {r[1].d2h_ms[0]:.3f} ms, launches {r[1].gpu_launches}")
timed("upload + render_columns -> pinned host", lambda: (ctx.upload_flat(flat, camx), ctx.render(p, x0, x1, out)))
if (x1 - x0) * h <= 9_000_000:
    timed("tcrt_render_columns -> pageable host", lambda: ctx.render(p, x0, x1, pageable), n=5)
