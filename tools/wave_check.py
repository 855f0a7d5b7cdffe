#!/usr/bin/env python
"""Developer check of the wavefront path (tcrt_render_wave.cu): a frame above its size threshold rendered whole
(wavefront) must equal the same frame rendered in bands below the threshold (megakernel), bit for bit, with the same
ray counters; then timings.  usage: TCRT_WAVE=1 TCRT_LIB=<library built with -DTCRT_DEV_KNOBS> python tools/wave_check.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from tilecoderaytracer_b200 import api  # noqa: E402

ctx = api.Context([0])
ok = True
for name, w, h, d, kw in [("synth256", 1280, 832, 10, {}), ("synth256", 1288, 836, 10, {}), ("two_mirrors", 1280, 832, 20, {}),
                          ("synth1024", 1280, 832, 20, {}), ("synth256", 1280, 832, 0, {}), ("synth256", 1280, 832, 10, {"shadows": False}),
                          ("synth256", 1280, 832, 10, {"reflections": False})]:
    cam = api.Camera()
    sc = api.Scene().build(name, cam)
    ctx.upload(sc, cam)
    p = api.default_params(w, h, d, **kw)
    st = ctx.render_device(p)                 # one launch sequence over the whole frame: the wavefront path
    whole = np.empty((w, h, 3), np.float32)
    ctx.download(whole)
    parts, rays = [], [0, 0, 0]
    step = 256
    for x0 in range(0, w, step):
        img, s = ctx.render(p, x0, min(w, x0 + step))
        parts.append(img)
        rays = [rays[0] + s.rays_primary, rays[1] + s.rays_shadow, rays[2] + s.rays_reflect]
    mega = np.concatenate(parts, 0)
    bad = int((whole.view(np.uint32) != mega.view(np.uint32)).any(-1).sum())
    rays_ok = [st.rays_primary, st.rays_shadow, st.rays_reflect] == rays
    ok &= bad == 0 and rays_ok
    print(f"{name:12s} {w}x{h} d{d} {kw}: mismatching pixels {bad}, rays_ok {rays_ok}, launches {st.gpu_launches}, whole {st.render_ms[0]:.3f} ms", flush=True)
for name, w, h, d in [("synth256", 7680, 4320, 10), ("synth256", 3840, 2160, 10), ("two_mirrors", 3840, 2160, 20), ("synth256", 960, 4320, 10)]:
    cam = api.Camera()
    sc = api.Scene().build(name, cam)
    ctx.upload(sc, cam)
    p = api.default_params(w, h, d)
    ms = [ctx.render_device(p).render_ms[0] for _ in range(4)]
    st = ctx.render_device(p)
    print(f"perf {name:12s} {w}x{h} d{d}: {min(ms[1:]):9.3f} ms  {st.rays / min(ms[1:]) / 1e3:9.1f} Mrays/s  launches {st.gpu_launches}", flush=True)
print("WAVE_OK" if ok else "WAVE_FAIL")
