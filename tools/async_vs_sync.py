import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from bench import WORKLOADS
from tilecoderaytracer_b200 import api
scene_name, w, h, d = WORKLOADS['synth256_8k_d10']
ctx = api.Context([0]); cam = api.Camera(); scene = api.Scene().build(scene_name, cam)
flat, camx = scene.flatten(), cam.export(); p = api.default_params(w, h, d)
for bands in (1, 2, 8):
    x0, x1 = 0, (w // bands) // 4 * 4
    hb = [api.HostBuffer((x1 - x0) * h * 12) for _ in range(2)]
    outs = [b.array(np.float32, (x1 - x0, h, 3)) for b in hb]
    for _ in range(2):
        ctx.upload_flat(flat, camx); ctx.render(p, x0, x1, outs[0])
    n = 20
    t0 = time.perf_counter()
    for _ in range(n):
        ctx.upload_flat(flat, camx); ctx.render(p, x0, x1, outs[0])
    ts = (time.perf_counter() - t0) / n
    t0 = time.perf_counter(); tickets = []
    for i in range(n):
        ctx.upload_flat(flat, camx)
        tickets.append(ctx.render_async(p, x0, x1, out=outs[i % 2]))
        if len(tickets) == 2: ctx.wait(tickets.pop(0))
    for t in tickets: ctx.wait(t)
    ta = (time.perf_counter() - t0) / n
    t0 = time.perf_counter()
    for i in range(n):
        ctx.upload_flat(flat, camx); ctx.render(p, x0, x1, outs[i % 2])
    tb = (time.perf_counter() - t0) / n
    print(f"bands {bands}: sync {ts*1e3:.3f} ms  async(2 in flight) {ta*1e3:.3f} ms  sync alternating buffers {tb*1e3:.3f} ms", flush=True)
    del outs, hb
