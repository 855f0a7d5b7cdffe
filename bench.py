#!/usr/bin/env python
"""bench.py — Mrays/s (primary + shadow + reflection) and frames/s of the render path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One "step" = one frame of the workload.  With N ranks (torchrun, one process per GPU) the frame
is split into N column bands, one per rank, no collective on the data path; each rank's band is
copied back to its own pinned host buffer in the e2e leg.  Rank 0 prints ONE JSON line.

  value      rays / kernel time, scene resident in HBM, output left in HBM; CUDA events on the
             launching stream around every launch (tcrt_stats.render_ms), L2 flushed between
             steps, max over ranks.
  e2e        the same metric through the C-ABI call a user makes with HOST buffers: every step
             re-uploads the flattened scene (H2D) and tcrt_render_columns() copies the finished
             band into pinned host memory (D2H); wall clock around the calls, max over ranks.
  roofline   FP32-pipe bound: algorithmic flops (oracle counters x §8d constants, on
             a column sample, scaled by the exact GPU ray count) / kernel time, against the FP32
             lane-instruction rate measured live by tcrt_fp32_peak (unfused FMUL+FADD mix).
  cpu_baseline  the reference itself (oracle/_ref/ref_render: calculatePixel & co. compiled from
             the reference's sources) on the host cores, bounded sample; rank 0, N=1 only.
  also       e2e_multiframe (scene resident, new camera per frame, frame to the host), e2e_txt
             (render + GPU "%f" formatting + D2H + file write), bands / kernel_ms_per_rank /
             balance_max_over_mean_per_iteration (the cost-balanced cut and how its feedback converged).

At N > 1 the frame is split into cost-balanced column bands (tcrt_balance_columns on rank 0, shared
once, refined from the gathered band times before the timed region).  Without torchrun, --gpus N
drives N GPUs from one process through a multi-device tcrt_ctx instead.

--impl reference times the reference on all host cores (one process per core, column bands —
the reference's own strategy-1 partition), each step a bounded column sample of the workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# name -> (scene, W, H, depth): BASELINE.json configs[1..4] (+ the reference's shipped case)
WORKLOADS = {
    "default_1080p_d5": ("default", 1920, 1080, 5),        # configs[1]  <- default at N=1 (and every N)
    "default_4k_d50": ("default", 3840, 2160, 50),         # configs[2]
    "synth1024_4k_d50": ("synth1024", 3840, 2160, 50),     # configs[3]
    "synth256_8k_d10": ("synth256", 7680, 4320, 10),       # configs[4]
    "default_500x504_d50": ("default", 500, 504, 50),      # configs[0], the reference's own case
    "two_mirrors_1080p_d50": ("two_mirrors", 1920, 1080, 50),
    # reduced-resolution copies of configs[3], [4] for quick ncu captures (tools/profile_run.py)
    "synth1024_1080p_d50": ("synth1024", 1920, 1080, 50),
    "synth256_1080p_d10": ("synth256", 1920, 1080, 10),
}
DEFAULT_WORKLOAD = "default_1080p_d5"
METRIC = "Mrays/s (primary+shadow+reflection)"


# ---- clocks during the timed region (NVML) ---------------------------------------------------------
class ClockSampler:
    REASONS = {
        0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
        0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
        0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting",
    }

    def __init__(self, index: int, period_s: float = 0.002):
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:   # NVML missing: report that, do not invent clocks
            self.nv, self.err = None, repr(e)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self.nv:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        if not self.nv:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "error": self.err}
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---- the reference on the host cores ------------------------------------------------------------------
def ref_processes(scene, w, h, depth, bands, stride):
    """Runs one ref_render per band concurrently; returns (max in-loop render seconds, wall seconds)."""
    from oracle import oracle_py as O

    t0 = time.perf_counter()
    procs = [subprocess.Popen([O.REF_RENDER, scene, str(w), str(h), str(depth), str(x0), str(x1), "-",
                               "--stride", str(stride)], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE,
                              cwd=tempfile.gettempdir()) for x0, x1 in bands if x1 > x0]
    infos = []
    for p in procs:
        err = p.communicate()[1].decode()
        if p.returncode != 0:
            raise RuntimeError("ref_render failed: " + err[-300:])
        infos.append(json.loads(err.strip().splitlines()[-1]))
    wall = time.perf_counter() - t0
    return max(i["render_s"] for i in infos), wall, sum(i["pixels"] for i in infos)


def oracle_sample(scene_name, w, h, depth, stride):
    """Counters (rays, flops) of the column sample x = 0, stride, 2*stride, ... from the oracle."""
    from oracle import oracle_py as O
    from tilecoderaytracer_b200 import api

    cam = api.Camera()
    scene = api.Scene().build(scene_name, cam)
    _, cnt = O.render(scene.flatten(), cam.export(), api.default_params(w, h, depth), 0, w, stride)
    cnt["rays"] = cnt["rays_primary"] + cnt["rays_shadow"] + cnt["rays_reflect"]
    return cnt


def pick_stride(w, h, us_per_pixel, cores, budget_s):
    """Column stride so that (w/stride)*h pixels on `cores` processes take about budget_s."""
    full_s = w * h * us_per_pixel * 1e-6 / max(1, cores)
    stride = max(1, int(full_s / budget_s + 0.999))
    return min(stride, max(1, w // max(1, cores)))


def calibrate_us_per_pixel(scene, w, h, depth):
    """One thin column sample of the reference (about a second) -> single-thread us/pixel."""
    stride = max(1, w // 12)
    render_s, _, pixels = ref_processes(scene, w, h, depth, [(0, w)], stride)
    return render_s * 1e6 / pixels


def run_reference(args):
    """--impl reference: the reference's CPU implementation on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle_py as O
    from tilecoderaytracer_b200.partition import column_bands

    scene, w, h, depth = WORKLOADS[args.workload]
    if not O.have_reference():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_render not built"}))
        return 0
    cores = os.cpu_count() or 1
    procs = min(cores, w)
    us_pp = calibrate_us_per_pixel(scene, w, h, depth)
    total_budget = float(os.environ.get("TCRT_REF_BUDGET_S", "150"))
    stride = pick_stride(w, h, us_pp, procs, total_budget / max(1, args.steps + args.warmup))
    bands = column_bands(w, procs)
    cnt = oracle_sample_bands(scene, w, h, depth, bands, stride)
    for _ in range(args.warmup):
        ref_processes(scene, w, h, depth, bands, stride)
    times = []
    pixels = 0
    for _ in range(args.steps):
        render_s, wall, pixels = ref_processes(scene, w, h, depth, bands, stride)
        times.append(render_s)
    t = sum(times)
    mrays = cnt["rays"] * args.steps / t / 1e6
    frac = pixels / float(w * h)
    sample = (f"{procs} processes x column bands, every {stride}th column of each band "
              f"({pixels} of {w * h} pixels per step); time = slowest process's pixel loop")
    line = {
        "impl": "reference", "metric": METRIC, "value": mrays, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, 1), "frames_per_s_equiv": args.steps * frac / t,
        "cpu_baseline": {"value": mrays, "unit": "Mrays/s", "cores": procs, "kind": "reference", "sample": sample,
                         "single_thread_us_per_pixel": us_pp},
        "e2e": {"value": mrays, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def oracle_sample_bands(scene_name, w, h, depth, bands, stride):
    """Ray count of exactly the pixels ref_processes renders (per-band strided columns)."""
    from oracle import oracle_py as O
    from tilecoderaytracer_b200 import api

    cam = api.Camera()
    scene = api.Scene().build(scene_name, cam)
    flat, c, p = scene.flatten(), cam.export(), api.default_params(w, h, depth)
    rays = flops = 0
    for x0, x1 in bands:
        if x1 <= x0:
            continue
        _, cnt = O.render(flat, c, p, x0, x1, stride)
        rays += cnt["rays_primary"] + cnt["rays_shadow"] + cnt["rays_reflect"]
        flops += cnt["flops"]
    return {"rays": rays, "flops": flops}


def workload_config(name, n_gpus):
    scene, w, h, depth = WORKLOADS[name]
    return {"workload": name, "scene": scene, "width": w, "height": h, "max_depth": depth, "shadows": True,
            "reflections": True, "partition": f"{n_gpus} cost-balanced column band(s), one per GPU, no collective"
                         + ("" if int(os.environ.get("WORLD_SIZE", "1")) > 1 or n_gpus == 1 else
                            " (one process, multi-device tcrt_ctx)"),
            "l2": "flushed between timed steps (the kernel reads no large input: scene lives in shared memory)"}


# ---- the B200 arm -----------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np

    from tilecoderaytracer_b200 import api, distributed as D

    rank, world, local = D.env_world()
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    tdev = None
    if world > 1:
        import torch

        torch.cuda.set_device(local)
        tdev = torch.device("cuda", local)
        D.init("nccl")
    scene_name, w, h, depth = WORKLOADS[args.workload]
    cam = api.Camera()
    scene = api.Scene().build(scene_name, cam)
    params = api.default_params(w, h, depth)
    # --gpus N without torchrun: ONE process drives N GPUs through a multi-device tcrt_ctx (the
    # library cuts, balances and stitches the bands itself); under torchrun: one rank per GPU.
    inproc = world == 1 and args.gpus > 1
    ctx = api.Context(list(range(args.gpus)) if inproc else [local])
    ctx.upload(scene, cam)
    # one column band per rank; the cut is cost-balanced (a low-resolution pre-pass on rank 0 that
    # counts bounces per column, shared once at setup) and outside the timed region
    bands = ctx.balance_columns(params, world) if rank == 0 else None
    bands = D.broadcast_object(bands)
    x0, x1 = bands[rank]
    flat, camx = scene.flatten(), cam.export()
    if inproc:
        for _ in range(12):                   # the ctx re-cuts after every whole-frame render
            ctx.render_device(params)

    # ---- at N > 1, feedback on the cut: every rank re-cuts from the same gathered band times (so all
    # agree) until the slowest band is within 1 % of the mean; the cut is then frozen -------------------
    from tilecoderaytracer_b200.partition import rebalance

    balance_log = []
    if world > 1:
        for it in range(16):
            ms = min(ctx.render_device(params, x0, x1).render_ms[0] for _ in range(2))
            all_ms = D.gather_floats(ms, tdev)
            balance_log.append(max(all_ms) / (sum(all_ms) / world))
            if balance_log[-1] < 1.01:
                break
            bands = rebalance(bands, all_ms, w)
            x0, x1 = bands[rank]
            if x1 <= x0:
                raise SystemExit("empty band after rebalancing")
    # ---- warm-up ------------------------------------------------------------------------------------
    for i in range(max(3, args.warmup)):
        st = ctx.render_device(params, x0, x1)
    # ---- timed: K launches, CUDA events around each, L2 flushed in between ------------------------------
    if world > 1:
        import torch

        D.barrier()
        torch.cuda.synchronize()
    kernel_ms = []
    rays_rank = 0
    launches = 0
    wall0 = time.perf_counter()
    with ClockSampler(local) as clocks:
        for _ in range(args.steps):
            ctx.flush_l2()
            st = ctx.render_device(params, x0, x1)
            kernel_ms.append(max(st.render_ms))       # in-process multi-device: the slowest device
            rays_rank = st.rays
            launches += st.gpu_launches
    wall = time.perf_counter() - wall0
    if world > 1:
        torch.cuda.synchronize()
        D.barrier()
    t_rank_ms = sum(kernel_ms)
    t_ms = D.reduce_max(t_rank_ms, tdev)              # slowest rank, device time
    rank_ms = [t / args.steps for t in D.gather_floats(t_rank_ms, tdev)]
    if inproc:
        rank_ms, bands = list(st.render_ms), st.bands
    rays_total = D.reduce_sum(rays_rank, tdev)        # per frame
    launches_total = int(D.reduce_sum(launches, tdev))
    value = rays_total * args.steps / (t_ms * 1e-3) / 1e6
    frames_per_s = args.steps / (t_ms * 1e-3)

    # ---- e2e: host buffers, H2D scene + D2H band inside the timed region ------------------------------
    band_floats = (x1 - x0) * h * 3
    host = api.HostBuffer(band_floats * 4)
    out = host.array(np.float32, (x1 - x0, h, 3))
    e2e_steps = max(3, min(args.steps, 30))
    for _ in range(2):
        ctx.upload_flat(flat, camx)
        ctx.render(params, x0, x1, out)
    if world > 1:
        D.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.upload_flat(flat, camx)
        _, st_e = ctx.render(params, x0, x1, out)
    e2e_s = time.perf_counter() - t0
    e2e_s = D.reduce_max(e2e_s, tdev)
    e2e_value = rays_total * e2e_steps / e2e_s / 1e6
    h2d = scene_bytes(flat) + 64
    d2h = band_floats * 4 + 32

    line = {
        "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": t_ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, args.gpus),
        "frames_per_s": frames_per_s, "rays_per_frame": int(rays_total),
        "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "frames_per_s": e2e_steps / e2e_s, "steps": e2e_steps,
                "what": "tcrt_upload_scene + tcrt_render_columns into pinned host memory, wall clock"},
        "gpu_launches": launches_total,
        "clocks": clocks.summary(),
        "wall_ms_per_step_incl_flush": 1e3 * wall / args.steps,
        "bands": [list(b) for b in bands], "kernel_ms_per_rank": rank_ms,
        "balance_max_over_mean_per_iteration": balance_log,
        "note": ("strong scaling: the frame is fixed and split into N column bands; this workload is "
                 f"{1e3 * t_ms / args.steps * world:.0f} us of GPU work in total, so at N = 8 a band is a few pixel "
                 "lifetimes long and latency-bound (profiles/README.md §2); BASELINE configs[4] "
                 "(--workload synth256_8k_d10) scales 7.1x on 8 B200 (profiles/scaling_r1/)"),
    }

    if world == 1:
        # ---- multi-frame mode: scene resident, a new camera per frame, frame to the host ----
        import copy

        n_mf = e2e_steps
        cams = []
        for i in range(n_mf):
            c = copy.copy(camx)
            c.eye[0] = camx.eye[0] + 0.002 * i       # a slow dolly: every frame is a different image
            cams.append(c)
        ctx.set_camera(cams[0])
        ctx.render(params, x0, x1, out)
        t0 = time.perf_counter()
        for c in cams:
            ctx.set_camera(c)
            ctx.render(params, x0, x1, out)
        mf_s = time.perf_counter() - t0
        ctx.set_camera(camx)
        line["e2e_multiframe"] = {"frames_per_s": n_mf / mf_s, "frames": n_mf,
                                  "what": "tcrt_set_camera + tcrt_render_columns into pinned host memory per frame; "
                                          "scene stays on the device"}
    if rank == 0:
        # ---- .txt writer leg (the reference's output format), N=1 only --------------------------------------
        if world == 1:
            with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as td:
                path = os.path.join(td, "raytracer_screen.txt")
                ctx.render_device(params)
                ctx.write_txt(params, path, 0.0)
                n_txt = 3
                t0 = time.perf_counter()
                for _ in range(n_txt):
                    ctx.render_device(params)
                    ctx.write_txt(params, path, 0.0)
                txt_s = (time.perf_counter() - t0) / n_txt
                line["e2e_txt"] = {"frames_per_s": 1.0 / txt_s, "ms_per_frame": 1e3 * txt_s,
                                   "file_bytes": os.path.getsize(path),
                                   "what": "render + GPU %f formatting + D2H of text + fwrite to tmpfs"}
        # ---- roofline -----------------------------------------------------------------------------------------
        peak = ctx.fp32_peak()
        stride = max(1, w // 24)
        cnt = oracle_sample(scene_name, w, h, depth, stride)
        flops_per_ray = cnt["flops"] / cnt["rays"]
        flops_per_launch = flops_per_ray * rays_rank
        kernel_s = statistics.mean(kernel_ms) * 1e-3
        achieved = flops_per_launch / kernel_s / 1e12
        line["roofline"] = {
            "bound": "fp32_pipe", "achieved": achieved, "peak": peak["unfused_tera_inst"], "unit": "TFLOP/s",
            "frac": achieved / peak["unfused_tera_inst"], "traffic": ncu_traffic(args.workload),
            "traffic_what": "dram__bytes_read.sum + dram__bytes_write.sum of one launch, bytes, from the committed "
                            "ncu --set full capture of this workload (profiles/ncu_metrics_r1.json); the frame "
                            "(12 B/pixel) mostly stays in the 126 MB L2 during the launch",

            "kernel": "render_kernel", "kernel_ms": 1e3 * kernel_s,
            "algorithmic_flops_per_ray": flops_per_ray, "algorithmic_flops_per_launch": flops_per_launch,
            "peak_source": "measured live: tcrt_fp32_peak, dependent FMUL+FADD chains (1 flop per lane-instruction; "
                           "parity-exact code cannot use FFMA); MEASURED_PEAKS.json has no FP32-pipe figure",
            "peak_fma_tera_inst": peak["fma_tera_inst"],
            "note": "achieved = ALGORITHMIC flops (what the reference's linear sweep over every object computes for "
                    "these rays, SURVEY 8d) / kernel time; box clusters and BVHs skip most of that work, so on "
                    "many-object scenes frac exceeds 1; the executed FP32-pipe and issue-slot utilisation are in "
                    "profiles/ncu_metrics_r1.json",
            "hbm_algorithmic_bytes_per_launch": band_floats * 4,
            "hbm_frac_of_measured": (band_floats * 4 / kernel_s / 1e9) / measured_hbm_gbs(),
        }
        # ---- CPU baseline: the reference itself, bounded sample, N=1 only ---------------------------------------
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(scene_name, w, h, depth)
        print(json.dumps(line))
    ctx.close()
    D.shutdown()
    return 0


def ncu_traffic(workload):
    try:
        return float(json.load(open(os.path.join(ROOT, "profiles", "ncu_metrics_r1.json")))[workload]["dram_bytes_per_launch"])
    except Exception:
        return None


def measured_hbm_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0   # fallback of B200_PROFILING.md


def scene_bytes(flat):
    n = flat.n_objects
    return 4 * (4 * flat.n_spheres + 16 * flat.n_fin_planes + 16 * flat.n_inf_planes + (4 + 4 + 4 + 8 + 4) * n
                + flat.n_spheres + flat.n_fin_planes + flat.n_inf_planes + flat.n_lights + 8 * flat.n_textures)


def cpu_baseline(scene, w, h, depth):
    from oracle import oracle_py as O
    from tilecoderaytracer_b200.partition import column_bands

    if not O.have_reference():
        # the oracle port instead (kind "port"), single thread
        stride = max(1, w // 16)
        t0 = time.perf_counter()
        cnt = oracle_sample(scene, w, h, depth, stride)
        dt = time.perf_counter() - t0
        return {"value": cnt["rays"] / dt / 1e6, "unit": "Mrays/s", "cores": 1, "kind": "port",
                "sample": f"every {stride}th column, oracle/tcrt_oracle.c single thread"}
    cores = os.cpu_count() or 1
    us_pp = calibrate_us_per_pixel(scene, w, h, depth)
    # single thread: ~6 s of work
    s1 = pick_stride(w, h, us_pp, 1, 6.0)
    r1, _, px1 = ref_processes(scene, w, h, depth, [(0, w)], s1)
    c1 = oracle_sample_bands(scene, w, h, depth, [(0, w)], s1)
    # all cores: ~6 s per process
    procs = min(cores, w)
    bands = column_bands(w, procs)
    sa = pick_stride(w, h, us_pp, procs, 6.0)
    ra, _, pxa = ref_processes(scene, w, h, depth, bands, sa)
    ca = oracle_sample_bands(scene, w, h, depth, bands, sa)
    return {
        "value": ca["rays"] / ra / 1e6, "unit": "Mrays/s", "cores": procs, "kind": "reference",
        "sample": f"{procs} processes (one per host core) x column bands, every {sa}th column ({pxa} of {w * h} pixels); "
                  f"time = slowest process's pixel loop",
        "single_thread": {"value": c1["rays"] / r1 / 1e6, "unit": "Mrays/s", "cores": 1,
                          "sample": f"every {s1}th column ({px1} pixels)", "us_per_pixel": r1 * 1e6 / px1},
        "frames_per_s_equiv": (pxa / float(w * h)) / ra,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    from tilecoderaytracer_b200 import build as B

    # the prebuilt in-tree libraries travel with the repo; only a missing one is built here
    # (rank 0 only — under torchrun the other ranks would race the compiler)
    if int(os.environ.get("LOCAL_RANK", "0")) == 0:
        if not os.path.exists(B.LIB_PATH):
            B.build_lib()
        if not os.path.exists(B.ORACLE_LIB_PATH):
            B.build_oracle()
        B.build_ref()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
