#!/usr/bin/env python
"""Developer loop on a GPU box: bit-exact parity on small frames + kernel time of the BASELINE
configs.  TCRT_LIB selects a library variant.  usage: python tools/quick_perf.py [--big] [--noparity]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from oracle import oracle_py as O  # noqa: E402
from tilecoderaytracer_b200 import _ffi, api  # noqa: E402

print("lib:", os.path.basename(_ffi.LIB_PATH), flush=True)
ctx = api.Context([0])
ok = True
if "--noparity" not in sys.argv:
    for name, w, h, d in [("default", 250, 252, 50), ("synth1024", 120, 96, 50), ("synth256", 120, 96, 10),
                          ("random:4:80", 120, 96, 9), ("random:7:120", 120, 96, 9), ("random:24:200", 112, 80, 9),
                          ("random:27:1500", 64, 48, 6), ("two_mirrors", 64, 48, 50),
                          ("boxes:1:6", 120, 96, 9), ("boxes:2:10", 120, 96, 12), ("boxes:3:14", 96, 80, 12),
                          ("boxes:6:9", 120, 96, 9)]:
        cam = api.Camera()
        sc = api.Scene().build(name, cam)
        ctx.upload(sc, cam)
        p = api.default_params(w, h, d)
        img, st = ctx.render(p)
        ref, cnt = O.render(sc.flatten(), cam.export(), p)
        bad = int((img.view(np.uint32) != ref.view(np.uint32)).any(-1).sum())
        rays_ok = st.rays == cnt["rays_primary"] + cnt["rays_shadow"] + cnt["rays_reflect"]
        ok &= bad == 0 and rays_ok
        print(f"parity {name:16s} mismatch {bad:6d} rays_ok {rays_ok}", flush=True)
cfgs = [("default", 500, 504, 50), ("default", 1920, 1080, 5), ("default", 3840, 2160, 50), ("synth1024", 3840, 2160, 50)]
if "--bands" in sys.argv:      # one eighth of the 4K frame: what a GPU renders at N = 8 (latency of the deepest paths)
    cfgs = [("default", 3840, 2160, 50)]
if "--big" in sys.argv:
    cfgs += [("synth256", 7680, 4320, 10), ("two_mirrors", 1920, 1080, 50)]
for name, w, h, d in cfgs:
    cam = api.Camera()
    sc = api.Scene().build(name, cam)
    ctx.upload(sc, cam)
    p = api.default_params(w, h, d)
    ms = []
    for i in range(4):
        st = ctx.render_device(p, 0, w // 8) if "--bands" in sys.argv else ctx.render_device(p)
        ms.append(st.render_ms[0])
    best = min(ms[1:])
    print(f"perf {name:12s} {w}x{h} d{d}: {best:9.3f} ms  {st.rays / best / 1e3:9.1f} Mrays/s", flush=True)
print("PARITY_OK" if ok else "PARITY_FAIL")
