"""CPU, world_size 2 over gloo: the N>1 host path (rank -> band, reductions, stitch).  The
per-rank render is done by the oracle here (no GPU in this container); the property checked is
the one the multi-GPU run relies on: bands rendered independently concatenate to exactly the
single-process frame, and ray counters add up."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    import numpy as np
    from tilecoderaytracer_b200 import api, distributed as D
    from oracle import oracle_py as O
    rank, world, local = D.init("gloo")
    assert world == 2
    cam = api.Camera(); scene = api.Scene().initialize()
    p = api.default_params(61, 40, 6)
    x0, x1 = D.rank_band(p.width)
    band, cnt = O.render(scene.flatten(), cam.export(), p, x0, x1)
    D.barrier()
    rays = cnt["rays_primary"] + cnt["rays_shadow"] + cnt["rays_reflect"]
    total = D.reduce_sum(rays)
    slowest = D.reduce_max(float(rank + 1))
    full = D.stitch_bands(band, p.width)
    if rank == 0:
        want, c = O.render(scene.flatten(), cam.export(), p)
        assert full.shape == want.shape
        assert np.array_equal(full.view(np.uint32), want.view(np.uint32))
        assert int(total) == c["rays_primary"] + c["rays_shadow"] + c["rays_reflect"]
        assert slowest == 2.0
        print("GLOO_OK", x0, x1)
    else:
        assert full is None and (x0, x1) == (30, 61)
""") % ROOT


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_band_render_over_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "GLOO_OK 0 30" in outs[0]
