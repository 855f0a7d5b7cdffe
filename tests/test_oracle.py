"""CPU tests: the oracle (oracle/tcrt_oracle.c) is pinned to the reference.

* against the committed goldens (tests/golden, generated from the reference's own binaries)
* against the reference itself, when oracle/_ref is present (built where /root/reference is)
"""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from oracle import oracle_py as O
from tilecoderaytracer_b200 import api

from _util import assert_bit_identical, make_scene, pixel_md5

needs_ref = pytest.mark.skipif(not O.have_reference(), reason="oracle/_ref not built (no /root/reference here)")


def oracle_frame(scene_name, w, h, depth, **kw):
    scene, cam = make_scene(scene_name)
    params = api.default_params(w, h, depth, **kw)
    return O.render(scene.flatten(), cam.export(), params)


def test_oracle_matches_golden_frames(golden, golden_frames):
    """Bit-exact on every committed small frame (10 scenes incl. 5 random ones)."""
    assert len(golden["frames"]) >= 12
    for key, m in golden["frames"].items():
        img, _ = oracle_frame(m["scene"], m["W"], m["H"], m["depth"])
        assert_bit_identical(img, golden_frames[key], key)


def test_oracle_txt_md5_shipped_configuration(golden):
    """BASELINE configs[0]: 500x504, depth 50 — the md5 of the reference program's own
    raytracer_screen.txt pixel lines."""
    g = golden["md5"]["default_500x504_d50"]
    assert g["pixel_md5"] == "b4020733f838b88d2783b27aaedbe758"
    assert golden["asis"]["rt_asis"]["pixel_md5"] == g["pixel_md5"]
    assert golden["asis"]["rt_fixed"]["pixel_md5"] == g["pixel_md5"]   # the UB init fix is output-neutral
    img, cnt = oracle_frame("default", 500, 504, 50)
    txt = O.format_txt(img)
    assert len(txt) == g["pixel_bytes"] == 500 * 504 * 31
    assert pixel_md5(txt) == g["pixel_md5"]
    # work counters of SURVEY §6.2
    assert cnt["rays_primary"] == 252000
    assert cnt["rays_reflect"] == 230689
    assert cnt["rays_shadow"] == 963640


@pytest.mark.parametrize("key", ["synth256_500x504_d10", "default_200x120_d0", "random_4_80_200x160_d8"])
def test_oracle_txt_md5_other_configs(golden, key):
    g = golden["md5"][key]
    img, _ = oracle_frame(g["scene"], g["W"], g["H"], g["depth"])
    assert pixel_md5(O.format_txt(img)) == g["pixel_md5"]


@needs_ref
@pytest.mark.parametrize("scene,w,h,d", [
    ("default", 160, 128, 50), ("default", 64, 36, 5), ("default", 50, 40, 0),
    ("synth1024", 80, 64, 50), ("synth256", 80, 64, 10), ("two_mirrors", 48, 40, 50),
    ("random:11:60", 96, 80, 10), ("random:12:200", 64, 48, 6), ("random:13:15", 96, 80, 30),
    ("random:14:500", 48, 40, 4), ("boxes:1:6", 96, 80, 8), ("boxes:2:10", 80, 64, 20), ("boxes:5:0", 64, 48, 5),
    ("boxes:11:9", 72, 56, 15), ("boxes:12:4", 72, 56, 50), ("random:15:90", 72, 56, 25), ("random:16:7", 72, 56, 50),
    # jittered lattices with touching / overlapping spheres (the scenes the uniform sphere grid takes)
    ("lattice:1:4", 96, 80, 12), ("lattice:2:5", 80, 64, 8), ("lattice:3:3", 72, 56, 30), ("lattice:4:6", 64, 48, 6),
])
def test_oracle_matches_reference_binary(scene, w, h, d):
    """The unmodified reference (calculatePixel & co. compiled from /root/reference/src)."""
    ref, info = O.ref_render(scene, w, h, d)
    img, _ = oracle_frame(scene, w, h, d)
    assert info["objects"] == make_scene(scene)[0].getObjectCount()
    assert_bit_identical(img, ref, f"{scene} {w}x{h} d{d}")


@needs_ref
def test_reference_band_and_stride_are_consistent():
    """ref_render's column band / stride sampling (used by bench.py's CPU baselines) selects the
    same pixels a full render has."""
    full, _ = O.ref_render("default", 60, 40, 5)
    band, _ = O.ref_render("default", 60, 40, 5, x0=13, x1=41)
    assert_bit_identical(band, full[13:41], "band")
    samp, _ = O.ref_render("default", 60, 40, 5, x0=3, x1=60, stride=7)
    assert_bit_identical(samp, full[3:60:7], "stride")
    img, _ = oracle_frame("default", 60, 40, 5)
    scene, cam = make_scene("default")
    o_samp, _ = O.render(scene.flatten(), cam.export(), api.default_params(60, 40, 5), 3, 60, 7)
    assert_bit_identical(o_samp, img[3:60:7], "oracle stride")


@needs_ref
def test_reference_as_checked_in_program_matches_golden(golden):
    """rt_asis = the reference's own main(): pixel loop RayTracer.cpp:911-923 and writer
    :1574-1626, parameters as shipped."""
    with tempfile.TemporaryDirectory() as td:
        subprocess.run([os.path.join(O.REF_DIR, "rt_asis")], cwd=td, check=True, stdout=subprocess.DEVNULL)
        txt = open(os.path.join(td, "raytracer_screen.txt"), "rb").read()
    assert pixel_md5(txt) == golden["md5"]["default_500x504_d50"]["pixel_md5"]
    head = txt.split(b"(", 1)[0].decode().splitlines()
    want = golden["asis"]["rt_asis"]["header_lines"]
    assert [l for l in head if not l.startswith(("Run_Time", "us/pixel"))] == \
           [l for l in want if not l.startswith(("Run_Time", "us/pixel"))]


def test_oracle_switches_and_edge_sizes():
    """SHADOWS_ON / REFLECTIONS_ON (rt_project_parameters.h:24-25) and degenerate sizes."""
    base, c0 = oracle_frame("default", 40, 30, 5)
    nosh, c1 = oracle_frame("default", 40, 30, 5, shadows=False)
    nore, c2 = oracle_frame("default", 40, 30, 5, reflections=False)
    assert c1["rays_shadow"] == 0 and c2["rays_reflect"] == 0
    assert (nosh >= base - 1e-6).mean() > 0.99      # removing shadows only adds light
    assert not np.array_equal(nore, base)
    one, c = oracle_frame("default", 1, 1, 50)
    assert one.shape == (1, 1, 3) and c["rays_primary"] == 1
    # depth 0: a reflective first hit adds k * NULL_COLOR * obj and stops
    d0, c = oracle_frame("default", 40, 30, 0)
    assert c["rays_reflect"] == 0


def test_oracle_empty_scene_is_null_color():
    scene, cam = api.Scene(), api.Camera()
    img, cnt = O.render(scene.flatten(), cam.export(), api.default_params(8, 6, 3))
    assert np.all(img == np.float32(0.75))
    assert cnt["rays_shadow"] == 0 and cnt["rays_reflect"] == 0


@needs_ref
@pytest.mark.parametrize("scene,w,h,d,stride", [("default", 120, 96, 50, 1), ("default", 64, 36, 5, 3),
                                                ("synth256", 80, 64, 10, 2), ("synth1024", 60, 48, 50, 1),
                                                ("two_mirrors", 40, 32, 50, 1), ("random:12:200", 64, 48, 6, 5)])
def test_reference_ray_counters_equal_the_oracles(scene, w, h, d, stride):
    """oracle/_ref/ref_count = the reference + two counter increments (build_ref.sh (c)): its primary /
    shadow / reflection ray counts — what bench.py --impl reference divides by its time — equal the
    oracle's counters (which the GPU kernel's counters are tested against), and its pixels equal ref_render's."""
    import json

    with tempfile.TemporaryDirectory() as td:
        f32 = os.path.join(td, "o.f32")
        r = subprocess.run([os.path.join(O.REF_DIR, "ref_count"), scene, str(w), str(h), str(d), "0", str(w), f32,
                            "--stride", str(stride)], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, cwd=td)
        info = json.loads(r.stderr.decode().strip().splitlines()[-1])
        got = np.fromfile(f32, dtype=np.float32).reshape(-1, h, 3)
    s, cam = make_scene(scene)
    want, cnt = O.render(s.flatten(), cam.export(), api.default_params(w, h, d), 0, w, stride)
    assert_bit_identical(got, want, "ref_count pixels")
    for k in ("rays_primary", "rays_shadow", "rays_reflect"):
        assert info[k] == cnt[k], k


def test_bench_golden_columns_equal_the_oracle():
    """tests/golden/bench_columns.json (made by the reference, checked by bench.py at every N): a few columns
    of every workload re-rendered by the oracle give the same md5 and the same ray counts."""
    import hashlib
    import json

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    g = json.load(open(os.path.join(root, "tests", "golden", "bench_columns.json")))["workloads"]
    assert {"default_500x504_d50", "default_1080p_d5", "default_4k_d50", "synth1024_4k_d50", "synth256_8k_d10"} <= set(g)
    for name, e in g.items():
        s, cam = make_scene(e["scene"])
        p = api.default_params(e["W"], e["H"], e["depth"])
        for i in (0, len(e["columns"]) // 2):
            c = e["columns"][i]
            col, _ = O.render(s.flatten(), cam.export(), p, c, c + 1, 1)
            assert hashlib.md5(col[0].tobytes()).hexdigest() == e["md5"][i], (name, c)
        if e["H"] <= 1080:     # the whole sample: ray counters too
            _, cnt = O.render(s.flatten(), cam.export(), p, e["x0"], e["W"], e["stride"])
            assert (cnt["rays_primary"], cnt["rays_shadow"], cnt["rays_reflect"]) == \
                   (e["rays_primary"], e["rays_shadow"], e["rays_reflect"]), name
