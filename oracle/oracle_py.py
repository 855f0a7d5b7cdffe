"""oracle/oracle_py.py — TEST INFRASTRUCTURE.  ctypes access to the two checkers:

  * liboracle.so  (oracle/tcrt_oracle.c, the CPU restatement; "port")
  * oracle/_ref/ref_render, rt_asis, rt_fixed (the reference itself, built by
    oracle/build_ref.sh from /root/reference/src; "reference")

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
import this module.  The product (tilecoderaytracer_b200, libtcrt.so) never does.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_LIB = os.path.join(_HERE, "_build", "liboracle.so")
REF_DIR = os.path.join(_HERE, "_ref")
REF_RENDER = os.path.join(REF_DIR, "ref_render")


class OracleCounters(C.Structure):
    _fields_ = [(n, C.c_ulonglong) for n in
                ("rays_primary", "rays_shadow", "rays_reflect", "tests_sphere", "tests_fin", "tests_inf",
                 "sph_exit_v", "sph_exit_d2", "sph_hit", "fin_exit_t", "fin_bounds",
                 "hits_sphere", "hits_plane", "hits_reflective", "hits_textured", "hits_light", "misses",
                 "lit_pairs", "lit_pairs_spec", "combines", "flops")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_LIB):
            raise ImportError(f"{ORACLE_LIB} missing: run __graft_entry__.build() / tilecoderaytracer_b200.build.build_oracle()")
        _lib = C.CDLL(ORACLE_LIB)
        _lib.tcrt_oracle_render.restype = C.c_int
        _lib.tcrt_oracle_render.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                            C.c_void_p, C.POINTER(OracleCounters)]
        _lib.tcrt_oracle_primary_ray.restype = C.c_int
        _lib.tcrt_oracle_primary_ray.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _lib.tcrt_oracle_format_txt.restype = C.c_size_t
        _lib.tcrt_oracle_format_txt.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    return _lib


def render(flat, cam, params, x0=0, x1=None, stride=1):
    """CPU restatement render of columns x0, x0+stride, ... < x1 -> (array[ncols,H,3], counters dict)."""
    x1 = params.width if x1 is None else x1
    ncols = (x1 - x0 + stride - 1) // stride
    out = np.empty((ncols, params.height, 3), dtype=np.float32)
    cnt = OracleCounters()
    rc = lib().tcrt_oracle_render(C.byref(flat), C.byref(cam), C.byref(params), x0, x1, stride,
                                  out.ctypes.data_as(C.c_void_p), C.byref(cnt))
    if rc != 0:
        raise RuntimeError("tcrt_oracle_render: bad arguments")
    return out, cnt.as_dict()


def primary_ray(cam, params, x, z):
    o = np.empty(3, np.float32)
    d = np.empty(3, np.float32)
    lib().tcrt_oracle_primary_ray(C.byref(cam), C.byref(params), x, z, o.ctypes.data_as(C.c_void_p),
                                  d.ctypes.data_as(C.c_void_p))
    return o, d


def format_txt(rgb: np.ndarray) -> bytes:
    """glibc "(%f, %f, %f)\\n" per pixel (RayTracer.cpp:1601)."""
    rgb = np.ascontiguousarray(rgb, dtype=np.float32).reshape(-1, 3)
    n = rgb.shape[0]
    cap = lib().tcrt_oracle_format_txt(rgb.ctypes.data_as(C.c_void_p), n, None, 0)
    buf = C.create_string_buffer(cap)
    got = lib().tcrt_oracle_format_txt(rgb.ctypes.data_as(C.c_void_p), n, buf, cap)
    return buf.raw[:got]


def have_reference() -> bool:
    return os.access(REF_RENDER, os.X_OK)


def ref_render(scene: str, width: int, height: int, depth: int, x0: int = 0, x1: int | None = None,
               stride: int = 1, want_txt: bool = False):
    """Runs the reference itself (oracle/_ref/ref_render).  Returns (array[ncols,H,3], info dict[, txt bytes])."""
    if not have_reference():
        raise FileNotFoundError(f"{REF_RENDER} missing (built by oracle/build_ref.sh where /root/reference exists)")
    x1 = width if x1 is None else x1
    with tempfile.TemporaryDirectory(prefix="tcrt_ref_") as td:
        outp = os.path.join(td, "o.f32")
        cmd = [REF_RENDER, scene, str(width), str(height), str(depth), str(x0), str(x1), outp, "--stride", str(stride)]
        txtp = os.path.join(td, "o.txt")
        if want_txt:
            cmd += ["--txt", txtp]
        r = subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, check=True, cwd=td)
        info = json.loads(r.stderr.decode().strip().splitlines()[-1])
        ncols = (x1 - x0 + stride - 1) // stride
        arr = np.fromfile(outp, dtype=np.float32).reshape(ncols, height, 3)
        if want_txt:
            with open(txtp, "rb") as f:
                return arr, info, f.read()
        return arr, info
