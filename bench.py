#!/usr/bin/env python3
"""Benchmark of the hot path on BASELINE.json's headline workload: Mray/s on a sponza-class ~257k-triangle scene at
1920x1080 with PathTracerMIS (max depth 8, both lights sampled per vertex), 1/2/4/8 B200.

    python bench.py --gpus N --steps K --warmup W            # N > 1: launched by torchrun, one rank per GPU
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's own CPU implementation

A "step" is one pass (one sample per pixel) over the whole frame.  Multi-GPU: the frame's 32x32 tiles are
interleaved over the ranks (replicated scene), no data-path collective inside a pass; the per-rank films are
all-gathered once per frame over NCCL.  Prints ONE JSON line on rank 0 (see DESIGN.md "Measurement").
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WIDTH, HEIGHT = 1920, 1080
MAX_DEPTH, DIMS = 8, 128
KERNELS_PER_PASS = 2 + 4 * (MAX_DEPTH + 1)        # raygen + film + (traverse, shade, shadow, resolve) per depth
SCENE_DIR = os.path.join(ROOT, "bench_data", "scenes")
WORKLOAD = ("sponza-class synthetic atrium, 256768 triangles + 3 analytic spheres + 1 box, background + 1-degree directional light, "
            "1920x1080, Path Tracer MIS, maxRayDepth 8, lightSampling all, 128 sampler dims")


def ensure_scene():
    """Synthetic sponza-class OBJ + JSON (deterministic generator, tests/scenes/make_scenes.py)."""
    path = os.path.join(SCENE_DIR, "sponza_class.json")
    if not os.path.exists(path) or not os.path.exists(os.path.join(SCENE_DIR, "MODELS", "sponza_class.obj")):
        sys.path.insert(0, os.path.join(ROOT, "tests", "scenes"))
        import make_scenes
        make_scenes.write_all(SCENE_DIR, sponza=True)
    return path


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     stdout=subprocess.PIPE, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "power_w_max": max(float(s[2]) for s in self.samples),
                "samples": len(self.samples), "reasons": reasons}


def config_for(world):
    """`config` of the JSON line: the SAME dict for the product arm and for --impl reference (same workload, same step)."""
    return {"workload": WORKLOAD,
            "step": "one pass (1 sample per pixel) of the 1920x1080 frame",
            "parallelism": "tiles32x32-interleaved x%d" % world,
            "l2": "per-pass path/ray/shadow queues (>400 MB) exceed L2; the 20 MB BVH+triangles stay L2-resident by design"}


def cpu_reference_run(steps, warmup, width, height, threads=0):
    """The reference's own CPU implementation (oracle/_ref/avx: as-shipped -O2 -mavx2 -mfma build, asserts off)."""
    from tests import util
    if not util.oracle_available("avx"):
        return None
    scene = ensure_scene()
    t0 = time.time()
    meta = util.run_rtref(["render", "--scene", scene, "--data", SCENE_DIR + "/", "--width", width, "--height", height,
                           "--passes", steps, "--warmup", warmup, "--dims", DIMS, "--depth", MAX_DEPTH, "--strategy", "all",
                           "--threads", threads], variant="avx", timeout=3000)
    meta["wall_s"] = time.time() - t0
    return meta


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    w, h = WIDTH, HEIGHT          # the product arm's configuration: full 1920x1080 passes of the same scene and camera
    meta = cpu_reference_run(args.steps, args.warmup, w, h)
    if meta is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/avx (the reference built from /root/reference) is not present on this box"}))
        return
    value = meta["mrays_per_s"]
    sample = "%d passes of %dx%d (full frames of the same scene/camera), %d threads" % (args.steps, w, h, meta["threads"])
    print(json.dumps({
        "impl": "reference", "metric": "Mray/s", "value": value, "unit": "Mray/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * meta["seconds"] / max(1, args.steps), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config_for(max(1, args.gpus)),
        "msamples_per_s": meta["msamples_per_s"],
        "cpu_baseline": {"value": value, "unit": "Mray/s", "cores": meta["threads"], "kind": "reference", "sample": sample, "host_cores": cores},
        "e2e": {"value": value, "unit": "Mray/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


class CudaArray:
    """Zero-copy view of a raw device pointer for torch (used to hand the film to NCCL)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None}


def run_gpu(args):
    import torch
    from raytracer_b200 import capi, host_api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("WORLD_SIZE %d != --gpus %d" % (world, args.gpus))
    if args.gpus > 1 and world == 1:
        raise SystemExit("--gpus %d needs torchrun (one rank per GPU)" % args.gpus)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product has no CPU fallback (use --impl reference for the CPU reference arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # main stream of the context: the wavefront streams fork from it and rtb_join makes it wait for them again, so
    # torch.cuda.Event records on it bracket every kernel of the timed region
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)

    lib = capi.load_library()
    scene_path = ensure_scene() if rank == 0 else None
    if dist is not None:
        dist.barrier()
        scene_path = ensure_scene()
    t_load = time.time()
    scene = host_api.Scene(scene_path, SCENE_DIR + "/")          # JSON + OBJ load, SAH BVH build (host, C++)
    load_s = time.time() - t_load

    def make_viewport(flags=0):
        vp = host_api.Viewport(scene, device=local_rank, num_shards=world, shard_index=rank, flags=flags, stream=stream.cuda_stream)
        vp.set_renderer("Path Tracer MIS")
        vp.set_params(dimensions=DIMS, max_ray_depth=MAX_DEPTH, strategy="all", seed=12345)
        vp.resize(WIDTH, HEIGHT)
        return vp

    vp = make_viewport()
    ctx = vp.context()

    def sync_all():
        vp.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput: K passes enqueued back to back, film gathered once per frame ----------------------
    film_ptr, film_n = C.c_void_p(), C.c_uint64()
    lib.rtb_film_device_buffer(ctx, 0, C.byref(film_ptr), C.byref(film_n))
    film = torch.as_tensor(CudaArray(film_ptr.value, film_n.value), device="cuda")
    gathered = torch.empty(world * film_n.value, dtype=torch.float32, device="cuda") if world > 1 else None

    vp.render(args.warmup)
    sync_all()
    vp.reset()
    c0 = vp.counters()
    sampler = ClockSampler(local_rank)
    sampler.start()
    sync_all()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    vp.render(args.steps)
    lib.rtb_join(ctx)                                    # the torch stream waits (on the device) for the wavefront streams
    if world > 1:
        dist.all_gather_into_tensor(gathered, film)      # ONE collective per frame
    ev1.record()
    sync_all()
    wall_s = time.perf_counter() - t0
    sampler.stop_flag = True
    c1 = vp.counters()
    rays = (c1["numRays"] - c0["numRays"]) + (c1["numShadowRays"] - c0["numShadowRays"])
    samples = c1["numPrimaryRays"] - c0["numPrimaryRays"]
    gpu_launches = c1["numKernelLaunches"] - c0["numKernelLaunches"]
    dev_s = ev0.elapsed_time(ev1) * 1e-3   # device time of the K passes (+ the film gather), CUDA events on the launching stream
    t = torch.tensor([dev_s, float(rays), float(samples)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        seconds, rays, samples = float(tmax[0]), float(tsum[1]), float(tsum[2])
    else:
        seconds = dev_s
    if world > 1 and rank == 0:
        img = np.empty((HEIGHT, WIDTH, 3), dtype=np.float32)
        lib.rtb_unshard_film(ctx, C.c_void_p(gathered.data_ptr()), img.ctypes.data)
        assert np.isfinite(img).all() and img.mean() > 0

    # ---- end to end through the public API with host buffers: EVERY step's frame is read back to the host ---------------
    # A step is still one pass, and the frame as of every pass lands in page-locked host memory; the passes are
    # submitted E2E_GROUP at a time (Viewport.render(n): n Halton seed vectors go up, the passes share wavefronts) and
    # rtb_film_snapshots leaves the film as of each of them on the device, from where each frame is (all-gathered when
    # sharded,) un-permuted to raster order and copied out asynchronously.
    vp.reset()
    pinned = [capi.PinnedImage(WIDTH, HEIGHT), capi.PinnedImage(WIDTH, HEIGHT)]   # page-locked host frames (rtb_host_alloc)
    E2E_GROUP = 8
    e2e_steps = max(E2E_GROUP, (min(args.steps, 32) // E2E_GROUP) * E2E_GROUP)
    snap_ptr, snap_stride = C.c_void_p(), C.c_uint64()
    rc = lib.rtb_film_snapshots(ctx, E2E_GROUP, C.byref(snap_ptr), C.byref(snap_stride))
    assert rc == 0, lib.rtb_last_error()
    snaps = [torch.as_tensor(CudaArray(snap_ptr.value + 4 * snap_stride.value * i, snap_stride.value), device="cuda") for i in range(E2E_GROUP)]
    e2e_count = [0]

    def e2e_group():
        vp.render(E2E_GROUP)              # H2D: the seed vectors + pass descriptors of these steps
        lib.rtb_join(ctx)
        for i in range(E2E_GROUP):
            dst = pinned[e2e_count[0] & 1]
            e2e_count[0] += 1
            if world > 1:
                dist.all_gather_into_tensor(gathered, snaps[i])
                if rank == 0:
                    rc = lib.rtb_unshard_film_async(ctx, C.c_void_p(gathered.data_ptr()), C.c_void_p(dst.ptr))
                    assert rc == 0, lib.rtb_last_error()
            else:
                rc = lib.rtb_unshard_film_async(ctx, C.c_void_p(snaps[i].data_ptr()), C.c_void_p(dst.ptr))
                assert rc == 0, lib.rtb_last_error()

    e2e_group()
    sync_all()
    c0 = vp.counters()
    t0 = time.perf_counter()
    for _ in range(e2e_steps // E2E_GROUP):
        e2e_group()
    sync_all()
    e2e_s = time.perf_counter() - t0
    c1 = vp.counters()
    lib.rtb_film_snapshots(ctx, 0, None, None)
    e2e_rays = (c1["numRays"] - c0["numRays"]) + (c1["numShadowRays"] - c0["numShadowRays"])
    te = torch.tensor([e2e_s, float(e2e_rays)], dtype=torch.float64, device="cuda")
    if world > 1:
        a = te.clone(); dist.all_reduce(a, op=dist.ReduceOp.MAX)
        b = te.clone(); dist.all_reduce(b, op=dist.ReduceOp.SUM)
        e2e_s, e2e_rays = float(a[0]), float(b[1])

    # ---- roofline of the dominant kernel (closest-hit traversal), measured live with CUDA events on its stream ---------
    roofline = None
    cpu_baseline = None
    if rank == 0:
        lib.rtb_set_profiling(ctx, 1)
        tm = {"traverseMs": 0.0, "shadowMs": 0.0, "shadeMs": 0.0, "totalMs": 0.0, "traverseRays": 0, "shadowRays": 0, "traverseLaunches": 0}
        for _ in range(2):
            vp.render(min(args.steps, 10))       # same batched wavefronts as the timed region
            pt = capi.RtbPassTimings()
            lib.rtb_get_pass_timings(ctx, C.byref(pt))
            for k in tm:
                tm[k] += getattr(pt, k)
        lib.rtb_set_profiling(ctx, 0)
        # per-ray traversal work from an instrumented (counting) context on the same inputs
        vpc = make_viewport(flags=capi.RTB_FLAG_COUNT_TESTS)
        vpc.render(1)
        cc = vpc.counters()
        vpc.close()
        nb = cc["numRayBoxTests"] / max(1, cc["numRays"])
        nt = cc["numRayTriangleTests"] / max(1, cc["numRays"])
        nbs = cc["numShadowBoxTests"] / max(1, cc["numShadowRays"])
        nts = cc["numShadowTriangleTests"] / max(1, cc["numShadowRays"])
        bytes_per_ray = 32 + 24 + 32 * nb + 48 * nt              # ray in + hit out + 32 B per box test + 48 B per (padded) triangle
        bytes_per_shadow_ray = 32 + 4 + 32 * nbs + 48 * nts
        peak, peak_src = measured_peak()
        traffic = {}
        prof = os.path.join(ROOT, "profiles", "traverse_dram_bytes.json")
        if os.path.exists(prof):
            try:
                traffic = json.load(open(prof))
            except Exception:
                traffic = {}

        def kernel_roofline(name, what, rays_n, ms, n_launches, bpr, box, tri):
            ach = rays_n * bpr / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
            n_launches = max(1, n_launches)
            return {"bound": "hbm", "kernel": "%s (%s)" % (name, what), "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic.get(name + "_dram_bytes_per_launch"), "peak_source": peak_src,
                    "algorithmic_bytes_per_ray": bpr, "box_tests_per_ray": box, "tri_tests_per_ray": tri,
                    "rays_per_launch": rays_n / n_launches, "algorithmic_bytes_per_launch": rays_n * bpr / n_launches,
                    "ms_per_launch": ms / n_launches, "gray_per_s": rays_n / (ms * 1e-3) / 1e9 if ms > 0 else 0.0,
                    "share_of_pass": ms / tm["totalMs"] if tm["totalMs"] > 0 else 0.0}

        # one k_shadow launch per k_traverse launch (one per depth and wavefront)
        r_trav = kernel_roofline("k_traverse", "closest-hit two-level BVH traversal", tm["traverseRays"], tm["traverseMs"], tm["traverseLaunches"], bytes_per_ray, nb, nt)
        r_shad = kernel_roofline("k_shadow", "any-hit two-level BVH traversal of the light-sample rays", tm["shadowRays"], tm["shadowMs"], tm["traverseLaunches"], bytes_per_shadow_ray, nbs, nts)
        # the dominant kernel is the one with the larger share of the pass; the other one rides along under its name
        roofline, other = (r_shad, r_trav) if r_shad["share_of_pass"] >= r_trav["share_of_pass"] else (r_trav, r_shad)
        roofline[other["kernel"].split(" ")[0]] = other
        roofline["k_shade_share_of_pass"] = tm["shadeMs"] / tm["totalMs"] if tm["totalMs"] > 0 else 0.0
        if world == 1 and not args.no_cpu_baseline:
            meta = cpu_reference_run(4, 1, WIDTH, HEIGHT)
            if meta is not None:
                cpu_baseline = {"value": meta["mrays_per_s"], "unit": "Mray/s", "cores": meta["threads"], "kind": "reference",
                                "sample": "4 timed + 1 warm-up passes of %dx%d (full frames) of the same scene with the reference's AVX2 build on all host threads" % (WIDTH, HEIGHT),
                                "msamples_per_s": meta["msamples_per_s"], "host_cores": os.cpu_count()}

    if rank == 0:      # the frames are gathered on rank 0 only
        assert np.isfinite(pinned[0].array).all() and pinned[0].array.mean() > 0
    vp.close()
    for im in pinned:
        im.close()
    if rank == 0:
        value = rays / seconds * 1e-6
        line = {
            "metric": "Mray/s", "value": value, "unit": "Mray/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * seconds / max(1, args.steps), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": config_for(world),
            "scene_load_s": load_s, "batching": "consecutive passes share wavefronts (rtb_render_passes)",
            "msamples_per_s": samples / seconds * 1e-6, "rays_per_sample": rays / max(1.0, samples),
            "gpu_launches": gpu_launches,
            "e2e": {"value": e2e_rays / e2e_s * 1e-6, "unit": "Mray/s", "steps": e2e_steps,
                    "h2d_bytes_per_step": DIMS * 4 + C.sizeof(capi.RtbPassDesc), "d2h_bytes_per_step": WIDTH * HEIGHT * 12,
                    "group": E2E_GROUP,
                    "api": ("raytracer_b200.host_api.Viewport.render(8) (ViewportB200::Render x 8, one call) with rtb_film_snapshots; the frame as of EVERY pass is copied into a page-locked host frame (rtb_unshard_film_async), synchronize at the end" if world == 1 else
                            "Viewport.render(8) on every rank with rtb_film_snapshots; for EVERY pass: NCCL all-gather of the shard snapshots, rtb_unshard_film_async into rank 0's page-locked host frame; synchronize at the end")},
            "wall_s_timed_region": wall_s,
            "clocks": sampler.summary(),
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
