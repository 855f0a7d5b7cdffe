#!/usr/bin/env python
"""Minimal driver for ncu: upload a workload's scene and launch the render kernel a few times
(optionally the .txt formatter).  usage: python tools/profile_run.py <workload> [launches] [--txt]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS  # noqa: E402
from tilecoderaytracer_b200 import api  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "default_1080p_d5"
n = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 4
scene_name, w, h, d = WORKLOADS[name]
cam = api.Camera()
scene = api.Scene().build(scene_name, cam)
ctx = api.Context([0])
ctx.upload(scene, cam)
p = api.default_params(w, h, d)
for i in range(n):
    st = ctx.render_device(p)
    print(name, "launch", i, "kernel ms", st.render_ms[0], "Mrays/s", st.rays / st.render_ms[0] / 1e3, "rays", st.rays, flush=True)
if "--txt" in sys.argv:
    print("txt bytes", len(ctx.format_txt()))
ctx.close()
