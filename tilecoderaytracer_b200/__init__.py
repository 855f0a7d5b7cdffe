"""tilecoderaytracer_b200 — B200-native (sm_100a CUDA) implementation of the per-pixel render
path of ccelio/TileCodeRayTracer behind a C ABI (include/tcrt.h).

Python is only the harness for tests and bench.py: `api` forwards to libtcrt.so through
ctypes.  Importing `api` without the built library raises ImportError — there is no CPU path.
"""
from . import build  # noqa: F401  (build recipes need no compiled code)

__all__ = ["build", "api"]


def __getattr__(name):
    if name == "api":
        import importlib
        return importlib.import_module(".api", __name__)
    raise AttributeError(name)
