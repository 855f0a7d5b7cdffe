"""Column-band partition of an image over ranks / devices.

The reference's static scheme (strategy 1, RayTracer.cpp:904-906) gives rank r the rows
[r*dz, r*dz+dz) with dz = H / CORE_NUM.  Here the band axis is x, because the pixel array and
the .txt are x-major (RayTracer.h:44, RayTracer.cpp:1589-1590): a band of columns is one
contiguous slice of both, so bands are copied back and concatenated without a transpose and
without any inter-GPU exchange.  libtcrt.so uses the same formula for a multi-device ctx.
"""
from __future__ import annotations


def column_band(width: int, rank: int, world: int) -> tuple[int, int]:
    """Columns [x0, x1) of rank `rank` out of `world`; bands tile [0, width) exactly."""
    if world < 1 or not (0 <= rank < world) or width < 0:
        raise ValueError("bad partition arguments")
    return (width * rank) // world, (width * (rank + 1)) // world


def column_bands(width: int, world: int) -> list[tuple[int, int]]:
    return [column_band(width, r, world) for r in range(world)]


def bands_from_costs(costs, width: int, world: int) -> list[tuple[int, int]]:
    """Cost-balanced cut of [0, width) into `world` bands from per-column-group cost estimates
    (libtcrt.so: tcrt_bands_from_costs, which tcrt_balance_columns feeds with bounce counts of a
    low-resolution pre-pass).  The reference answered load imbalance with the TILE64 work queues
    (PixelQueue.cpp, strategies 2-7); with one band per GPU the band widths do the balancing."""
    import ctypes as C

    from . import _ffi

    lib = _ffi.load()
    n = len(costs)
    arr = (C.c_double * n)(*[float(c) for c in costs])
    b = (C.c_int * (world + 1))()
    rc = lib.tcrt_bands_from_costs(arr, n, width, world, b)
    if rc != 0:
        raise ValueError("bad cost partition arguments")
    return [(int(b[i]), int(b[i + 1])) for i in range(world)]


def rebalance(bands: list[tuple[int, int]], ms: list[float], width: int) -> list[tuple[int, int]]:
    """Feedback step of the band balancer (libtcrt.so: tcrt_rebalance_columns): the cut that would
    equalise the measured band times `ms` if cost were uniform inside each band."""
    import ctypes as C

    from . import _ffi

    lib = _ffi.load()
    n = len(bands)
    b = (C.c_int * (n + 1))(*([x0 for x0, _ in bands] + [bands[-1][1]]))
    t = (C.c_double * n)(*[float(m) for m in ms])
    out = (C.c_int * (n + 1))()
    if lib.tcrt_rebalance_columns(b, t, n, width, out) != 0:
        raise ValueError("bad rebalance arguments")
    return [(int(out[i]), int(out[i + 1])) for i in range(n)]
