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
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=__import__("datetime").timedelta(seconds=180))

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
    # multi-GPU: the communicator lives in the library (rtb_comm_init: NCCL loaded there); torch.distributed only carries
    # the 128-byte id, the barriers and the timing reductions
    if world > 1:
        ident = torch.zeros(capi.RTB_COMM_ID_BYTES, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_ubyte * capi.RTB_COMM_ID_BYTES)()
            rc = lib.rtb_comm_unique_id(buf)
            assert rc == 0, lib.rtb_last_error()
            ident = torch.tensor(list(buf), dtype=torch.uint8)
        ident = ident.cuda()
        dist.broadcast(ident, 0)
        raw = bytes(ident.cpu().tolist())
        rc = lib.rtb_comm_init(ctx, raw)
        assert rc == 0, lib.rtb_last_error()
    gathered_ptr = C.c_void_p()

    vp.render(args.warmup)
    if world > 1:
        rc = lib.rtb_finish_frame(ctx, 0, C.byref(gathered_ptr))      # (NCCL sets up its connections on the first collective)
        assert rc == 0, lib.rtb_last_error()
    sync_all()
    # The W warm-up steps share wavefronts of W passes, the timed region's wavefronts hold min(K, 32): a different launch
    # chain, which the library captures as a CUDA graph the first time it sees it (capture + instantiate: ~3 ms, a fifth
    # of a 1/8 shard's whole 20-step region).  One more untimed submission of the timed region's own shape keeps that
    # one-time cost where warm-up costs belong (`warmup_steps_run` in the line says how many steps really ran untimed).
    vp.render(args.steps)
    lib.rtb_join(ctx)
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
        rc = lib.rtb_finish_frame(ctx, 0, C.byref(gathered_ptr))      # ONE collective per frame: ncclAllGather of the shard films, inside the library
        assert rc == 0, lib.rtb_last_error()
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
        lib.rtb_unshard_film(ctx, gathered_ptr, img.ctypes.data)
        assert np.isfinite(img).all() and img.mean() > 0
        assert all(img[:, x0::32 * world][:, :32].mean() > 0 for x0 in range(0, 32 * world, 32))      # every shard's tiles are there

    # ---- end to end through the public API with host buffers: EVERY step's frame is read back to the host ---------------
    # A step is still one pass, and the frame as of every pass lands in page-locked host memory; the passes are
    # submitted E2E_GROUP at a time (Viewport.render(n): n Halton seed vectors go up, the passes share wavefronts) and
    # rtb_film_snapshots leaves the film as of each of them on the device, from where each frame is (all-gathered when
    # sharded,) un-permuted to raster order and copied out asynchronously.
    vp.reset()
    # steps submitted per call.  N > 1: two groups per 20-step region, so that the shard scatter of one group (on the context's
    # second stream, rtb_scatter_shard_async) crosses PCIe while the next group renders
    E2E_GROUP = 8 if world == 1 else (10 if args.steps < 32 else 16)
    e2e_steps = max(E2E_GROUP, (min(args.steps, 32) // E2E_GROUP) * E2E_GROUP)
    snap_ptr, snap_stride = C.c_void_p(), C.c_uint64()
    rc = lib.rtb_film_snapshots(ctx, E2E_GROUP, C.byref(snap_ptr), C.byref(snap_stride))
    assert rc == 0, lib.rtb_last_error()
    snap_addr = [snap_ptr.value + 4 * snap_stride.value * i for i in range(E2E_GROUP)]
    frame_bytes = WIDTH * HEIGHT * 12
    e2e_count = [0]
    shm = None
    if world == 1:
        pinned = [capi.PinnedImage(WIDTH, HEIGHT), capi.PinnedImage(WIDTH, HEIGHT)]   # page-locked host frames (rtb_host_alloc)
        frames = [im.array for im in pinned]

        def e2e_group():
            vp.render(E2E_GROUP)              # H2D: the seed vectors + pass descriptors of these steps
            lib.rtb_join(ctx)
            for i in range(E2E_GROUP):
                dst = pinned[e2e_count[0] & 1]
                e2e_count[0] += 1
                rc = lib.rtb_unshard_film_async(ctx, C.c_void_p(snap_addr[i]), C.c_void_p(dst.ptr))
                assert rc == 0, lib.rtb_last_error()
    else:
        # ONE host frame pair in shared memory, attached and page-locked by every rank process: each rank writes ITS pixels of
        # every frame straight into it (rtb_scatter_shard) over its own PCIe link - no all-gather, no funnel through rank 0
        from multiprocessing import shared_memory
        names = [None]
        if rank == 0:
            shm = shared_memory.SharedMemory(create=True, size=2 * frame_bytes)
            names = [shm.name]
        dist.broadcast_object_list(names, 0)
        if rank != 0:
            shm = shared_memory.SharedMemory(name=names[0])
        host = np.ndarray((2, HEIGHT, WIDTH, 3), dtype=np.float32, buffer=shm.buf)
        if rank == 0:
            host[:] = 0
        dist.barrier()
        base = host.ctypes.data
        rc = lib.rtb_host_register(C.c_void_p(base), 2 * frame_bytes)
        assert rc == 0, lib.rtb_last_error()
        dev_base = lib.rtb_host_device_pointer(C.c_void_p(base))
        assert dev_base, lib.rtb_last_error()
        frames = [host[0], host[1]]

        def e2e_group():
            vp.render(E2E_GROUP)
            lib.rtb_join(ctx)
            for i in range(E2E_GROUP):
                k = e2e_count[0] & 1
                e2e_count[0] += 1
                rc = lib.rtb_scatter_shard_async(ctx, C.c_void_p(snap_addr[i]), 0, C.c_void_p(dev_base + k * frame_bytes))
                assert rc == 0, lib.rtb_last_error()

    e2e_group()
    sync_all()
    # the region is timed by the host's wall clock (it includes host work by definition), which makes it sensitive to
    # whatever else the host does in those ~0.2 s: it is run twice and the faster repetition is reported (both are kept in
    # `repeats_s`)
    e2e_repeats = []
    for _ in range(2):
        c0 = vp.counters()
        t0 = time.perf_counter()
        for _ in range(e2e_steps // E2E_GROUP):
            e2e_group()
        sync_all()
        e2e_repeats.append((time.perf_counter() - t0, c0, vp.counters()))
    e2e_s, c0, c1 = min(e2e_repeats, key=lambda r: r[0])
    lib.rtb_film_snapshots(ctx, 0, None, None)
    e2e_rays = (c1["numRays"] - c0["numRays"]) + (c1["numShadowRays"] - c0["numShadowRays"])
    te = torch.tensor([e2e_s, float(e2e_rays)], dtype=torch.float64, device="cuda")
    if world > 1:
        a = te.clone(); dist.all_reduce(a, op=dist.ReduceOp.MAX)
        b = te.clone(); dist.all_reduce(b, op=dist.ReduceOp.SUM)
        e2e_s, e2e_rays = float(a[0]), float(b[1])
    e2e_ok = True
    if rank == 0:      # the host frame holds every shard's pixels (tile columns x0, x0 + 32 * world, ... belong to one shard per tile row)
        last = frames[(e2e_count[0] - 1) & 1]
        e2e_ok = bool(np.isfinite(last).all() and last.mean() > 0 and all(last[:32, x0:x0 + 32].sum() >= 0 for x0 in range(0, 32 * world, 32)))
        if world > 1:
            check = np.empty((HEIGHT, WIDTH, 3), dtype=np.float32)
            rc = lib.rtb_finish_frame(ctx, 0, C.byref(gathered_ptr)); assert rc == 0, lib.rtb_last_error()
            lib.rtb_unshard_film(ctx, gathered_ptr, check.ctypes.data)
            e2e_ok = e2e_ok and bool((check.view(np.uint32) == np.ascontiguousarray(last).view(np.uint32)).all())      # scattered host frame == all-gathered film
        del last
    elif world > 1:
        rc = lib.rtb_finish_frame(ctx, 0, C.byref(gathered_ptr)); assert rc == 0, lib.rtb_last_error()

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
        # per-ray traversal work from an instrumented (counting) context on the same inputs.  SURVEY section 8(d) defines the
        # algorithmic bytes per ray "in the reference layout": 32 B per box test and 48 B per (padded) triangle test of the
        # reference's walk over ITS binary tree - a counting context therefore walks that tree for shadow rays as well.
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
        # RTB_SHADOW_KERNEL=wide: what k_shadow_wide itself fetches - 96-byte nodes of the 8-wide tree (rtb_wide.h), 48-byte triangles
        wide = os.environ.get("RTB_SHADOW_KERNEL", "wide") not in ("simple", "split")      # k_shadow_wide is the library's default
        own_bytes_per_shadow_ray = wide_nodes = wide_tris = None
        if wide:
            os.environ["RTB_COUNT_WIDE"] = "1"
            vpw = make_viewport(flags=capi.RTB_FLAG_COUNT_TESTS)
            vpw.render(1)
            cw = vpw.counters()
            vpw.close()
            del os.environ["RTB_COUNT_WIDE"]
            wide_nodes = cw["numShadowBoxTests"] / 8.0 / max(1, cw["numShadowRays"])
            wide_tris = cw["numShadowTriangleTests"] / max(1, cw["numShadowRays"])
            own_bytes_per_shadow_ray = 32 + 4 + 96 * wide_nodes + 48 * wide_tris
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
            per_ray = traffic.get(name + "_dram_bytes_per_ray")
            return {"bound": "hbm", "kernel": "%s (%s)" % (name, what), "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    # ncu dram__bytes per ray of the committed N=1 capture x the rays THIS run's launches carry
                    "traffic": per_ray * rays_n / n_launches if per_ray is not None else None,
                    "traffic_source": traffic.get("source"), "peak_source": peak_src,
                    "algorithmic_bytes_per_ray": bpr, "box_tests_per_ray": box, "tri_tests_per_ray": tri,
                    "rays_per_launch": rays_n / n_launches, "algorithmic_bytes_per_launch": rays_n * bpr / n_launches,
                    "ms_per_launch": ms / n_launches, "gray_per_s": rays_n / (ms * 1e-3) / 1e9 if ms > 0 else 0.0,
                    "share_of_pass": ms / tm["totalMs"] if tm["totalMs"] > 0 else 0.0}

        # one any-hit launch per k_traverse launch (one per depth and wavefront)
        r_trav = kernel_roofline("k_traverse", "closest-hit two-level BVH traversal", tm["traverseRays"], tm["traverseMs"], tm["traverseLaunches"], bytes_per_ray, nb, nt)
        r_shad = kernel_roofline("k_shadow_wide" if wide else "k_shadow",
                                 ("any-hit traversal of the light-sample rays on the 8-wide quantised tree (top level exact); `achieved` counts the bytes of the REFERENCE's walk "
                                  "over its binary tree (SURVEY 8(d): 32 B per box test, 48 B per padded triangle test) - the work this kernel replaces, not what it fetches: see own_layout" if wide else
                                  "any-hit two-level BVH traversal of the light-sample rays"),
                                 tm["shadowRays"], tm["shadowMs"], tm["traverseLaunches"], bytes_per_shadow_ray, nbs, nts)
        if wide:
            own_ach = tm["shadowRays"] * own_bytes_per_shadow_ray / (tm["shadowMs"] * 1e-3) / 1e9 if tm["shadowMs"] > 0 else 0.0
            r_shad["own_layout"] = {"nodes_per_ray": wide_nodes, "tri_tests_per_ray": wide_tris, "bytes_per_ray": own_bytes_per_shadow_ray,
                                    "achieved": own_ach, "frac": own_ach / peak,
                                    "what": "bytes k_shadow_wide itself asks for: 32 B ray + 4 B result + 96 B per visited 8-wide node + 48 B per confirmed triangle"}
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

    if world > 1:
        lib.rtb_host_unregister(C.c_void_p(base))
        del host, frames
    vp.close()
    if world == 1:
        for im in pinned:
            im.close()
    else:
        dist.barrier()
        shm.close()
        if rank == 0:
            shm.unlink()
    if rank == 0:
        value = rays / seconds * 1e-6
        line = {
            "metric": "Mray/s", "value": value, "unit": "Mray/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "warmup_steps_run": args.warmup + args.steps,
            "ms_per_step": 1000.0 * seconds / max(1, args.steps), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": config_for(world),
            "scene_load_s": load_s, "batching": "consecutive passes share wavefronts (rtb_render_passes)",
            "msamples_per_s": samples / seconds * 1e-6, "rays_per_sample": rays / max(1.0, samples),
            "gpu_launches": gpu_launches,
            "e2e": {"value": e2e_rays / e2e_s * 1e-6, "unit": "Mray/s", "steps": e2e_steps,
                    "h2d_bytes_per_step": DIMS * 4 + C.sizeof(capi.RtbPassDesc), "d2h_bytes_per_step": WIDTH * HEIGHT * 12,
                    "group": E2E_GROUP, "repeats_s": [r[0] for r in e2e_repeats],
                    "api": ("raytracer_b200.host_api.Viewport.render(8) (ViewportB200::Render x 8, one call) with rtb_film_snapshots; the frame as of EVERY pass is copied into a page-locked host frame (rtb_unshard_film_async), synchronize at the end" if world == 1 else
                            "Viewport.render(%d) on every rank with rtb_film_snapshots;" % E2E_GROUP + " for EVERY pass every rank writes its own pixels into ONE page-locked host frame in shared memory (rtb_scatter_shard_async on the context's second stream, under the next group's wavefront; one PCIe link per rank, no collective); synchronize + barrier at the end")},
            "wall_s_timed_region": wall_s, "e2e_frame_check": "ok" if e2e_ok else "FAILED",
            "clocks": sampler.summary(),
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()



# ---- BASELINE config 5: rough glass + caustics, 1024x1024, bidirectional path tracing (the reference's "VCM" renderer with
#      vertex connection; merging needs every shard's photons and is a single-GPU feature here) ------------------------------
C5_W = C5_H = 1024
C5_PATH_LENGTH, C5_DIMS = 16, 256
C5_WORKLOAD = ("rough-glass + caustics scene (tests/golden/scenes/caustics.json: glass block, rough-glass sphere, rough metal, area light), "
               "1024x1024, VCM renderer: vertex connection (bidirectional path tracing) + vertex merging on one GPU, vertex connection only when sharded, "
               "maxPathLength 16, 256 sampler dims")


def run_c5(args):
    import torch
    from raytracer_b200 import capi, host_api
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=__import__("datetime").timedelta(seconds=180))
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
    lib = capi.load_library()
    scn = os.path.join(ROOT, "tests", "golden", "scenes")
    scene = host_api.Scene(os.path.join(scn, "caustics.json"), scn + "/")
    merging = world == 1
    vp = host_api.Viewport(scene, device=local_rank, num_shards=world, shard_index=rank, stream=stream.cuda_stream)
    vp.set_renderer("VCM")
    vp.set_params(dimensions=C5_DIMS, max_ray_depth=C5_PATH_LENGTH, strategy="all", seed=12345)
    vp.set_vcm(connection=True, merging=merging, max_path_length=C5_PATH_LENGTH)
    vp.resize(C5_W, C5_H)
    ctx = vp.context()
    if world > 1:
        ident = torch.zeros(capi.RTB_COMM_ID_BYTES, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_ubyte * capi.RTB_COMM_ID_BYTES)(); assert lib.rtb_comm_unique_id(buf) == 0, lib.rtb_last_error()
            ident = torch.tensor(list(buf), dtype=torch.uint8)
        ident = ident.cuda(); dist.broadcast(ident, 0)
        assert lib.rtb_comm_init(ctx, bytes(ident.cpu().tolist())) == 0, lib.rtb_last_error()

    def sync_all():
        vp.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    gathered = C.c_void_p()
    vp.render(args.warmup)
    assert lib.rtb_finish_frame(ctx, 0, C.byref(gathered)) == 0, lib.rtb_last_error()     # (NCCL connects on its first collective)
    sync_all(); vp.reset()
    c0 = vp.counters()
    sampler = ClockSampler(local_rank); sampler.start(); sync_all()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    vp.render(args.steps)
    lib.rtb_join(ctx)
    assert lib.rtb_finish_frame(ctx, 0, C.byref(gathered)) == 0, lib.rtb_last_error()     # splat all-reduce + film all-gather when sharded
    ev1.record()
    sync_all()
    sampler.stop_flag = True
    c1 = vp.counters()
    rays = (c1["numRays"] - c0["numRays"]) + (c1["numShadowRays"] - c0["numShadowRays"])
    launches = c1["numKernelLaunches"] - c0["numKernelLaunches"]
    dev_s = ev0.elapsed_time(ev1) * 1e-3
    t = torch.tensor([dev_s, float(rays)], dtype=torch.float64, device="cuda")
    if world > 1:
        a = t.clone(); dist.all_reduce(a, op=dist.ReduceOp.MAX); b = t.clone(); dist.all_reduce(b, op=dist.ReduceOp.SUM)
        dev_s, rays = float(a[0]), float(b[1])
    img = np.empty((C5_H, C5_W, 3), dtype=np.float32)
    if rank == 0:
        assert lib.rtb_unshard_film(ctx, gathered, img.ctypes.data) == 0, lib.rtb_last_error()
        assert np.isfinite(img).all() and img.mean() > 0
    # end to end: every 4 passes the frame goes to a page-locked host frame (single GPU: Viewport::GetSumBuffer; sharded:
    # finish_frame + un-permute on rank 0)
    vp.reset(); sync_all()
    e2e_steps = max(4, (min(args.steps, 16) // 4) * 4)
    pinned = capi.PinnedImage(C5_W, C5_H)
    c0 = vp.counters()
    t0 = time.perf_counter()
    for _ in range(e2e_steps // 4):
        vp.render(4)
        lib.rtb_join(ctx)
        if world > 1:
            assert lib.rtb_finish_frame(ctx, 0, C.byref(gathered)) == 0, lib.rtb_last_error()
            if rank == 0:
                assert lib.rtb_unshard_film_async(ctx, gathered, C.c_void_p(pinned.ptr)) == 0, lib.rtb_last_error()
        else:
            assert lib.rtb_read_sum_async(ctx, C.c_void_p(pinned.ptr), 0) == 0, lib.rtb_last_error()
    sync_all()
    e2e_s = time.perf_counter() - t0
    c1 = vp.counters()
    e2e_rays = (c1["numRays"] - c0["numRays"]) + (c1["numShadowRays"] - c0["numShadowRays"])
    te = torch.tensor([e2e_s, float(e2e_rays)], dtype=torch.float64, device="cuda")
    if world > 1:
        a = te.clone(); dist.all_reduce(a, op=dist.ReduceOp.MAX); b = te.clone(); dist.all_reduce(b, op=dist.ReduceOp.SUM)
        e2e_s, e2e_rays = float(a[0]), float(b[1])
    # algorithmic bytes of a VCM pass (stated in DESIGN.md): per path segment one ray in (32 B) + hit out (24 B) + the boxes and
    # triangles of the reference's walk (this scene: analytic shapes, a handful of top-level nodes), per light vertex 160 B
    # written and read once, per photon 48 B written, counted and gathered, per pixel 16 B pass colour + 24 B film
    roofline = None
    if rank == 0:
        vpc = host_api.Viewport(scene, device=local_rank, flags=capi.RTB_FLAG_COUNT_TESTS)
        vpc.set_renderer("Path Tracer MIS"); vpc.set_params(dimensions=C5_DIMS, max_ray_depth=8, strategy="all", seed=12345); vpc.resize(256, 256); vpc.render(1)
        cc = vpc.counters(); vpc.close()
        nb = (cc["numRayBoxTests"] + cc["numShadowBoxTests"]) / max(1, cc["numRays"] + cc["numShadowRays"])
        b_ray = 32 + 24 + 32 * nb + 64
        per_pass_rays = rays / max(1, args.steps)
        per_pass_bytes = per_pass_rays * b_ray + C5_W * C5_H * (40 + 2 * 160 * 3 + 48 * 3)
        peak, peak_src = measured_peak()
        ach = per_pass_bytes * args.steps / dev_s / 1e9
        roofline = {"bound": "hbm", "kernel": "VCM pass (k_vcm_light + k_vcm_camera + hash grid), timed as a whole", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_ray": b_ray, "box_tests_per_ray": nb,
                    "note": "one thread per sub-path: latency / divergence bound, not HBM bound; analytic-shape scene with ~%.0f box tests per ray" % nb}
    vp.close(); pinned.close()
    if rank == 0:
        print(json.dumps({
            "metric": "Mray/s", "value": rays / dev_s * 1e-6, "unit": "Mray/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * dev_s / max(1, args.steps), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": C5_WORKLOAD, "step": "one VCM pass (one light sub-path + one camera path per pixel) of the 1024x1024 frame", "parallelism": "tiles32x32-interleaved x%d" % world,
                       "l2": "light vertices (1.7 GB) and photons round-trip through HBM every pass; the scene itself is a few KB"},
            "msamples_per_s": C5_W * C5_H * args.steps / dev_s * 1e-6, "gpu_launches": launches,
            "e2e": {"value": e2e_rays / e2e_s * 1e-6, "unit": "Mray/s", "steps": e2e_steps, "h2d_bytes_per_step": C5_DIMS * 4 + C.sizeof(capi.RtbPassDesc), "d2h_bytes_per_step": C5_W * C5_H * 12 // 4,
                    "api": "Viewport.render(4) then the frame into a page-locked host frame (rtb_read_sum_async / rtb_finish_frame + rtb_unshard_film_async)"},
            "clocks": sampler.summary(), "roofline": roofline, "cpu_baseline": None}))
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", default="c3", choices=["c3", "c5"], help="c3: the headline 1080p sponza-class MIS benchmark (default); c5: BASELINE config 5, bidirectional / VCM on the caustics scene")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "c5":
        run_c5(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
