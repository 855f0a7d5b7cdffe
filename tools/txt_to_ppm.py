#!/usr/bin/env python
"""Offline viewer replacement: turns a raytracer_screen.txt — the reference's or ours —
into a binary PPM.  The reference's Java client is not in its repository; its tag parser is described
at RayTracer.cpp:2063-2069: header lines "Tag:value." then one "(r, g, b)" line per pixel, x-major
(all z of column 0, then column 1, ...), z running bottom to top.

usage: python tools/txt_to_ppm.py raytracer_screen.txt out.ppm
"""
import sys

import numpy as np


def read_txt(path):
    tags, first_pixel = {}, None
    with open(path, "rb") as f:
        data = f.read()
    pos = 0
    while pos < len(data) and data[pos:pos + 1] != b"(":
        end = data.index(b"\n", pos)
        line = data[pos:end].decode()
        if ":" in line:
            k, v = line.split(":", 1)
            tags[k] = v.rstrip(".")
        pos = end + 1
    w, h = int(tags["Horizontal_Resolution"]), int(tags["Vertical_Resolution"])
    body = data[pos:].replace(b"(", b"").replace(b")", b"").replace(b",", b" ")
    px = np.array(body.split(), dtype=np.float64).reshape(w, h, 3)
    return w, h, px, tags


def quant8(c):
    """q(c) = floor(clamp(c,0,1)*255 + 0.5): the quantisation of the parity tolerance."""
    return np.floor(np.clip(np.nan_to_num(c, nan=0.0), 0.0, 1.0) * 255.0 + 0.5).astype(np.uint8)


def to_image(px):
    """[W, H, 3] x-major, z up  ->  [H, W, 3] rows, top row first."""
    return np.ascontiguousarray(np.transpose(px, (1, 0, 2))[::-1])


def main():
    if len(sys.argv) != 3:
        raise SystemExit(__doc__)
    w, h, px, _ = read_txt(sys.argv[1])
    img = to_image(quant8(px))
    with open(sys.argv[2], "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (w, h))
        f.write(img.tobytes())
    print(f"{sys.argv[2]}: {w}x{h}")


if __name__ == "__main__":
    main()
