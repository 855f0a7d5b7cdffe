#!/usr/bin/env python
"""Sanitizer substitute (compute-sanitizer is closed on the GPU pool): renders the parity scenes with a -DTCRT_CHECKED
library — every index range-checked, the level stack poisoned and checked for reads of unwritten records — compares
every frame with the oracle bit for bit, and reports the check flags (0 = nothing fired).
usage: TCRT_LIB=.../libtcrt_checked.so python tools/checked_run.py"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from oracle import oracle_py as O  # noqa: E402
from tilecoderaytracer_b200 import _ffi, api  # noqa: E402

NAMES = ["pixel store", "level stack index", "BVH stack pointer", "primitive key", "object index", "leaf range", "node index",
         "read of an unwritten level record", "cluster face slot", "queue position"]
lib = _ffi.load()
print("lib:", os.path.basename(_ffi.LIB_PATH), flush=True)
ctx = api.Context([0])
flags = C.c_uint(0)
lib.tcrt_dev_check_flags(C.byref(flags), 1)
bad_total = 0
CASES = [("default", 160, 128, 50), ("boxes:3:14", 128, 96, 12), ("synth256", 160, 128, 10), ("two_mirrors", 96, 80, 50),
         ("synth1024", 160, 128, 50), ("random:24:200", 112, 80, 9), ("random:27:1500", 64, 48, 6), ("random:7:120", 120, 96, 9),
         ("boxes:7:9", 100, 60, 50), ("default", 97, 61, 0), ("default", 64, 48, 200), ("default", 1, 37, 50)]
for name, w, h, d in CASES:
    cam = api.Camera()
    sc = api.Scene().build(name, cam)
    ctx.upload(sc, cam)
    p = api.default_params(w, h, d)
    img, st = ctx.render(p)
    ref, cnt = O.render(sc.flatten(), cam.export(), p)
    bad = int((img.view(np.uint32) != ref.view(np.uint32)).any(-1).sum())
    lib.tcrt_dev_check_flags(C.byref(flags), 1)
    fired = [NAMES[b] for b in range(len(NAMES)) if flags.value >> b & 1]
    bad_total += bad + len(fired)
    print(f"{name:16s} {w}x{h} d{d}: mismatching pixels {bad}, finite {bool(np.isfinite(img).all())}, checks fired: {fired or 'none'}", flush=True)
print("CHECKED_OK" if bad_total == 0 else "CHECKED_FAIL")
