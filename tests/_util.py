"""Shared helpers for the parity tests."""
from __future__ import annotations

import hashlib

import numpy as np

from tilecoderaytracer_b200 import api


def make_scene(name: str):
    cam = api.Camera()
    scene = api.Scene().build(name, cam)
    return scene, cam


def bits(a: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def n_mismatch(a: np.ndarray, b: np.ndarray) -> int:
    """Pixels whose float bit patterns differ in any channel."""
    return int((bits(a) != bits(b)).any(-1).sum())


def assert_bit_identical(a, b, what=""):
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    bad = n_mismatch(a, b)
    assert bad == 0, f"{what}: {bad} of {a.shape[0] * a.shape[1]} pixels differ bitwise"


def quant8(c: np.ndarray) -> np.ndarray:
    """q(c) = floor(clamp(c,0,1)*255 + 0.5): the 8-bit quantisation the north_star tolerance is
    stated in (SURVEY §8a last row)."""
    return np.floor(np.clip(c.astype(np.float64), 0.0, 1.0) * 255.0 + 0.5).astype(np.int32)


def pixel_md5(txt: bytes) -> str:
    lines = [ln for ln in txt.split(b"\n") if ln.startswith(b"(")]
    return hashlib.md5(b"\n".join(lines) + b"\n").hexdigest()


def parse_txt(txt: bytes) -> np.ndarray:
    body = txt[txt.index(b"("):] if b"(" in txt else b""
    vals = body.replace(b"(", b"").replace(b")", b"").replace(b",", b" ").split()
    return np.array(vals, dtype=np.float64).reshape(-1, 3)
