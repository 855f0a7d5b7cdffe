#!/usr/bin/env python
"""Regenerates tests/golden/bench_columns.json from THE REFERENCE ITSELF (oracle/_ref/ref_render, built by
oracle/build_ref.sh from /root/reference/src): for every bench workload, the md5 of the raw float32 bytes
(H*3 floats, z fastest) of every `stride`-th column of the full-size frame, plus the reference's ray
counters of that column sample (oracle/_ref/ref_count).  bench.py checks its own bands against these
OUTSIDE the timed region at every N ("parity" in the JSON line); tests/ use them too.

usage: python tests/golden/make_bench_columns.py        (about two minutes of CPU)
"""
from __future__ import annotations

import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from bench import WORKLOADS  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
N_COLUMNS = 32   # per workload: at least 4 per band at N = 8


def main() -> None:
    out = {"generator": "tests/golden/make_bench_columns.py", "what": "md5 of float32 column bytes from ref_render",
           "workloads": {}}
    for name, (scene, w, h, depth) in sorted(WORKLOADS.items()):
        stride = max(1, w // N_COLUMNS)
        x0 = stride // 2          # not column 0: samples sit inside the bands
        with tempfile.TemporaryDirectory() as td:
            f32 = os.path.join(td, "o.f32")
            r = subprocess.run([os.path.join(REF, "ref_count"), scene, str(w), str(h), str(depth), str(x0), str(w), f32,
                                "--stride", str(stride)], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, cwd=td)
            info = json.loads(r.stderr.decode().strip().splitlines()[-1])
            arr = np.fromfile(f32, dtype=np.float32).reshape(-1, h, 3)
        cols = list(range(x0, w, stride))
        assert len(cols) == arr.shape[0]
        out["workloads"][name] = {
            "scene": scene, "W": w, "H": h, "depth": depth, "x0": x0, "stride": stride,
            "columns": cols, "md5": [hashlib.md5(arr[i].tobytes()).hexdigest() for i in range(len(cols))],
            "rays_primary": info["rays_primary"], "rays_shadow": info["rays_shadow"], "rays_reflect": info["rays_reflect"],
        }
        print(name, len(cols), "columns", info["render_s"], "s", flush=True)
    with open(os.path.join(HERE, "bench_columns.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
