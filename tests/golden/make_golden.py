#!/usr/bin/env python
"""Regenerates tests/golden/* from THE REFERENCE ITSELF (oracle/_ref, built by
oracle/build_ref.sh from /root/reference/src).  Run in the build container, where
/root/reference exists; the outputs are committed so that the GPU box (which has no
/root/reference) can check against them.

  golden.json        md5 of the pixel lines of the reference's .txt for each configuration
                     (the reference program as checked in — rt_asis — for the shipped 500x504
                     depth-50 case, the ref_render harness for the others), plus file sizes and
                     the header lines of the as-is program's file.
  small_frames.npz   full float32 frames at small resolutions for every scene, straight from
                     the reference's calculatePixel.

usage: python tests/golden/make_golden.py [--big]     (--big adds the 4K md5, ~25 s)
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
from oracle import oracle_py as O  # noqa: E402

# (key, scene, W, H, depth) — md5 of pixel lines
MD5_CONFIGS = [
    ("default_500x504_d50", "default", 500, 504, 50),          # BASELINE configs[0]
    ("default_1920x1080_d5", "default", 1920, 1080, 5),        # configs[1]
    ("synth1024_500x504_d50", "synth1024", 500, 504, 50),      # configs[3] at the reference's resolution
    ("synth256_500x504_d10", "synth256", 500, 504, 10),        # configs[4] at the reference's resolution
    ("two_mirrors_500x504_d50", "two_mirrors", 500, 504, 50),  # the reference's SCENE 2
    ("default_200x120_d0", "default", 200, 120, 0),
    ("random_4_80_200x160_d8", "random:4:80", 200, 160, 8),
]
BIG_CONFIGS = [("default_3840x2160_d50", "default", 3840, 2160, 50)]  # configs[2]

# (key, scene, W, H, depth) — full float frames
FRAME_CONFIGS = [
    ("default_125x126_d50", "default", 125, 126, 50),
    ("default_96x54_d5", "default", 96, 54, 5),
    ("synth1024_96x80_d50", "synth1024", 96, 80, 50),
    ("synth256_96x80_d10", "synth256", 96, 80, 10),
    ("two_mirrors_64x48_d50", "two_mirrors", 64, 48, 50),
    ("random_2_40_96x80_d12", "random:2:40", 96, 80, 12),
    ("random_4_80_96x80_d12", "random:4:80", 96, 80, 12),
    ("random_5_25_96x80_d3", "random:5:25", 96, 80, 3),
    ("random_7_120_96x80_d12", "random:7:120", 96, 80, 12),
    ("random_3_300_64x48_d20", "random:3:300", 64, 48, 20),
    ("boxes_1_6_96x80_d8", "boxes:1:6", 96, 80, 8),          # axis-aligned boxes: the box-cluster path
    ("boxes_3_14_96x80_d12", "boxes:3:14", 96, 80, 12),
]


def pixel_md5(txt: bytes) -> str:
    lines = [ln for ln in txt.split(b"\n") if ln.startswith(b"(")]
    return hashlib.md5(b"\n".join(lines) + b"\n").hexdigest()


def main() -> None:
    if not O.have_reference():
        raise SystemExit("oracle/_ref missing: run oracle/build_ref.sh where /root/reference exists")
    out = {"generator": "tests/golden/make_golden.py", "md5": {}, "asis": {}}
    # the reference program exactly as checked in
    for prog in ("rt_asis", "rt_fixed"):
        with tempfile.TemporaryDirectory() as td:
            subprocess.run([os.path.join(O.REF_DIR, prog)], cwd=td, check=True, stdout=subprocess.DEVNULL)
            with open(os.path.join(td, "raytracer_screen.txt"), "rb") as f:
                txt = f.read()
        head = txt.split(b"(", 1)[0].decode().splitlines()
        out["asis"][prog] = {"pixel_md5": pixel_md5(txt), "bytes": len(txt), "header_lines": head}
    configs = MD5_CONFIGS + (BIG_CONFIGS if "--big" in sys.argv else [])
    for key, scene, w, h, d in configs:
        arr, info, txt = O.ref_render(scene, w, h, d, want_txt=True)
        out["md5"][key] = {"scene": scene, "W": w, "H": h, "depth": d, "pixel_md5": pixel_md5(txt),
                           "pixel_bytes": len(txt), "max": float(arr.max()), "objects": info["objects"]}
        print(key, out["md5"][key]["pixel_md5"], flush=True)
    # keep a previously generated big entry when --big is not given
    path = os.path.join(HERE, "golden.json")
    if os.path.exists(path) and "--big" not in sys.argv:
        old = json.load(open(path))
        for k, v in old.get("md5", {}).items():
            out["md5"].setdefault(k, v)
    frames = {}
    meta = {}
    for key, scene, w, h, d in FRAME_CONFIGS:
        arr, info = O.ref_render(scene, w, h, d)
        frames[key] = arr
        meta[key] = {"scene": scene, "W": w, "H": h, "depth": d}
        print(key, arr.shape, float(arr.mean()), flush=True)
    out["frames"] = meta
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "small_frames.npz"), **frames)


if __name__ == "__main__":
    main()
