#!/usr/bin/env python
"""Where the .txt leg spends its time: GPU formatting + D2H (tcrt_format_txt) vs the file write."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS  # noqa: E402
from tilecoderaytracer_b200 import api  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "default_1080p_d5"
scene_name, w, h, d = WORKLOADS[name]
cam = api.Camera()
scene = api.Scene().build(scene_name, cam)
ctx = api.Context([0])
ctx.upload(scene, cam)
p = api.default_params(w, h, d)
ctx.render_device(p)
for i in range(3):
    t0 = time.perf_counter()
    txt = ctx.format_txt()
    t1 = time.perf_counter()
    print(f"format_txt (GPU format + D2H into a fresh bytes object): {1e3 * (t1 - t0):.2f} ms, {len(txt)} B", flush=True)
for d_ in ("/dev/shm", "/tmp"):
    path = os.path.join(d_, "tcrt_writer_timing.txt")
    for i in range(3):
        t0 = time.perf_counter()
        ctx.write_txt(p, path, 0.0)
        t1 = time.perf_counter()
        with open(path, "wb") as f:
            t2 = time.perf_counter()
            f.write(txt)
        t3 = time.perf_counter()
        print(f"{d_}: tcrt_write_txt {1e3 * (t1 - t0):.2f} ms; python f.write of the same bytes {1e3 * (t3 - t2):.2f} ms", flush=True)
    os.unlink(path)
