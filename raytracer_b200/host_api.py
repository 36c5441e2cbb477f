"""ctypes binding of librtb200_host.so: the C++ host object model (Scene / Camera / Material / lights, JSON + OBJ
loaders, BVH builder, ViewportB200) that mirrors the reference's plugin surface.  The viewport half needs a CUDA
device; loading / flattening scenes is pure host code."""
import ctypes as C
import os

import numpy as np

from . import capi

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(_HERE, "librtb200_host.so")
BLUE_NOISE_PATH = os.path.join(_HERE, "data", "blue_noise_128.rtbn")

_lib = None


def load_host_library():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(HOST_LIB_PATH):
        raise capi.RtbError("%s not found - run raytracer_b200/host/build.sh (or __graft_entry__.build())" % HOST_LIB_PATH)
    lib = C.CDLL(HOST_LIB_PATH)
    lib.rtbh_scene_load.restype = C.c_void_p
    lib.rtbh_scene_load.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_size_t]
    lib.rtbh_scene_free.argtypes = [C.c_void_p]
    lib.rtbh_scene_free.restype = None
    lib.rtbh_scene_save_flat.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.c_uint32]
    lib.rtbh_scene_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    lib.rtbh_scene_stats.restype = None
    lib.rtbh_viewport_create.restype = C.c_void_p
    lib.rtbh_viewport_create.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_char_p, C.c_void_p, C.c_char_p, C.c_size_t]
    lib.rtbh_viewport_free.argtypes = [C.c_void_p]
    lib.rtbh_viewport_free.restype = None
    lib.rtbh_viewport_last_error.restype = C.c_char_p
    lib.rtbh_viewport_last_error.argtypes = [C.c_void_p]
    lib.rtbh_viewport_context.restype = C.c_void_p
    lib.rtbh_viewport_context.argtypes = [C.c_void_p]
    lib.rtbh_viewport_set_renderer.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32]
    lib.rtbh_viewport_set_params.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64]
    lib.rtbh_viewport_resize.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
    lib.rtbh_viewport_reset.argtypes = [C.c_void_p]
    lib.rtbh_viewport_reset.restype = None
    lib.rtbh_viewport_render.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    lib.rtbh_viewport_synchronize.argtypes = [C.c_void_p]
    lib.rtbh_viewport_read_sum.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    lib.rtbh_viewport_set_adaptive.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, C.c_float]
    lib.rtbh_viewport_set_vcm.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, C.c_float, C.c_float, C.c_uint64]
    lib.rtbh_viewport_progress.argtypes = [C.c_void_p, C.c_void_p]
    lib.rtbh_viewport_blocks.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    lib.rtbh_viewport_blocks.restype = C.c_uint32
    lib.rtbh_visualize_blocks.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
    lib.rtbh_visualize_blocks.restype = None
    lib.rtbh_pass_script.argtypes = [C.c_uint64, C.c_uint32, C.c_float, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
    lib.rtbh_pass_script.restype = None
    lib.rtbh_revise_blocks.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_float, C.c_float]
    lib.rtbh_revise_blocks.restype = C.c_uint32
    lib.rtbh_viewport_set_postprocess.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_uint32]
    lib.rtbh_viewport_read_front.argtypes = [C.c_void_p, C.c_void_p]
    lib.rtbh_viewport_counters.argtypes = [C.c_void_p, C.POINTER(capi.RtbCounters)]
    _lib = lib
    return lib


def pass_script(num_passes, width, height, dimensions=128, seed=12345, aa_spread=0.5, extra_resets=1):
    """The host prologue of `num_passes` Viewport::Render calls (Halton seed vectors + AA offsets) of a viewport seeded
    with `seed`, resized to width x height and reset `extra_resets` more times (the oracle harness resets once after
    Resize): what ViewportB200 hands to rtb_render_passes.  No device involved."""
    lib = load_host_library()
    seeds = np.zeros((num_passes, dimensions), dtype=np.uint32)
    offsets = np.zeros((num_passes, 2), dtype=np.float32)
    lib.rtbh_pass_script(seed, dimensions, aa_spread, width, height, extra_resets, num_passes, seeds.ctypes.data, offsets.ctypes.data)
    return seeds, offsets


def visualize_blocks(front, blocks):
    """Viewport::VisualizeActiveBlocks on a (height, width) uint32 front-buffer image: returns the tinted copy."""
    lib = load_host_library()
    out = np.ascontiguousarray(front, dtype=np.uint32).copy()
    blocks = np.ascontiguousarray(blocks, dtype=np.uint32).reshape(-1, 4)
    lib.rtbh_visualize_blocks(out.ctypes.data, out.shape[1], blocks.ctypes.data, len(blocks))
    return out


def revise_blocks(blocks, errors, min_block_size=4, subdivision_treshold=0.005, convergence_treshold=0.0001):
    """The block-list revision of Viewport::UpdateBlocksList (host logic only, no device): (n, 4) blocks + n errors -> (m, 4)."""
    lib = load_host_library()
    blocks = np.ascontiguousarray(blocks, dtype=np.uint32).reshape(-1, 4)
    errors = np.ascontiguousarray(errors, dtype=np.float32)
    assert len(errors) == len(blocks)
    buf = np.zeros((2 * len(blocks) + 1, 4), dtype=np.uint32)
    buf[:len(blocks)] = blocks
    m = lib.rtbh_revise_blocks(buf.ctypes.data, len(blocks), len(buf), errors.ctypes.data, min_block_size, subdivision_treshold, convergence_treshold)
    return buf[:m].copy()


class Scene:
    """helpers::LoadScene + Scene::BuildBVH (the reference's Demo/Demo.cpp:156-198 SwitchScene sequence)."""

    def __init__(self, json_path, data_path=""):
        self.lib = load_host_library()
        err = C.create_string_buffer(1024)
        self.h = self.lib.rtbh_scene_load(json_path.encode(), data_path.encode(), err, len(err))
        if not self.h:
            raise capi.RtbError("scene load failed: %s" % err.value.decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.rtbh_scene_free(self.h)
            self.h = None

    __del__ = close

    def save_flat(self, path, width, height):
        if self.lib.rtbh_scene_save_flat(self.h, path.encode(), width, height) != 0:
            raise capi.RtbError("cannot write %s" % path)

    def stats(self):
        a, b, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
        self.lib.rtbh_scene_stats(self.h, C.byref(a), C.byref(b), C.byref(c))
        return dict(objects=a.value, lights=b.value, tlas_nodes=c.value)


class Viewport:
    """rt::ViewportB200: Viewport's public methods (Core/Rendering/Viewport.h:37-58) on the CUDA path."""

    def __init__(self, scene, device=0, num_shards=1, shard_index=0, flags=0, blue_noise=True, stream=None):
        self.lib = load_host_library()
        self.scene = scene
        err = C.create_string_buffer(1024)
        bn = BLUE_NOISE_PATH.encode() if blue_noise else None
        self.h = self.lib.rtbh_viewport_create(scene.h, device, num_shards, shard_index, flags, bn, stream, err, len(err))
        if not self.h:
            raise capi.RtbError("viewport creation failed: %s" % err.value.decode())
        self.width = self.height = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.rtbh_viewport_free(self.h)
            self.h = None

    __del__ = close

    def _check(self, rc):
        if rc != 0:
            raise capi.RtbError(self.lib.rtbh_viewport_last_error(self.h).decode())

    def context(self):
        """Raw RtbContext* (for rtb_film_device_buffer / rtb_unshard_film in the multi-GPU gather)."""
        return self.lib.rtbh_viewport_context(self.h)

    def set_renderer(self, name, debug_mode=1):
        self._check(self.lib.rtbh_viewport_set_renderer(self.h, name.encode(), debug_mode))

    def set_params(self, dimensions=128, use_blue_noise=True, aa_spread=0.5, max_ray_depth=8, min_rr_depth=1, strategy="all", seed=12345):
        s = capi.RTB_LIGHT_SAMPLING_ALL if strategy == "all" else capi.RTB_LIGHT_SAMPLING_SINGLE
        self._check(self.lib.rtbh_viewport_set_params(self.h, dimensions, 1 if use_blue_noise else 0, aa_spread, max_ray_depth, min_rr_depth, s, seed))

    def resize(self, w, h):
        self._check(self.lib.rtbh_viewport_resize(self.h, w, h))
        self.width, self.height = w, h

    def reset(self):
        self.lib.rtbh_viewport_reset(self.h)

    def render(self, num_passes=1):
        self._check(self.lib.rtbh_viewport_render(self.h, self.scene.h, num_passes))

    def synchronize(self):
        self._check(self.lib.rtbh_viewport_synchronize(self.h))

    def read_sum(self, which=0, out=None):
        if out is None:
            out = np.empty((self.height, self.width, 3), dtype=np.float32)
        self._check(self.lib.rtbh_viewport_read_sum(self.h, out.ctypes.data, which))
        return out

    def set_adaptive(self, enable=True, num_initial_passes=10, min_block_size=4, max_block_size=256, subdivision_treshold=0.005, convergence_treshold=0.0001):
        """AdaptiveRenderingSettings (Context.h:35-43); applies from the next reset() / resize()."""
        self._check(self.lib.rtbh_viewport_set_adaptive(self.h, 1 if enable else 0, num_initial_passes, min_block_size, max_block_size, subdivision_treshold, convergence_treshold))

    def set_vcm(self, connection=True, merging=True, max_path_length=10, initial_radius=0.02, min_radius=0.02, radius_multiplier=1.0, seed=0x5EED):
        """The public members of the reference's VCM renderer (VertexConnectionAndMerging.h:48-56), its defaults."""
        self._check(self.lib.rtbh_viewport_set_vcm(self.h, 1 if connection else 0, 1 if merging else 0, max_path_length, initial_radius, min_radius, radius_multiplier, seed))

    def progress(self):
        """Viewport::GetProgress (Viewport.h:25-32, 55)."""
        out = np.zeros(5, dtype=np.uint32)
        self._check(self.lib.rtbh_viewport_progress(self.h, out.ctypes.data))
        return {"passesFinished": int(out[0]), "activePixels": int(out[1]), "activeBlocks": int(out[2]),
                "converged": out[3:4].view(np.float32)[0], "averageError": out[4:5].view(np.float32)[0]}

    def blocks(self):
        """The block list of adaptive rendering: (n, 4) uint32 rows minX, maxX, minY, maxY."""
        n = self.lib.rtbh_viewport_blocks(self.h, None, 0)
        out = np.zeros((n, 4), dtype=np.uint32)
        if n:
            self.lib.rtbh_viewport_blocks(self.h, out.ctypes.data, n)
        return out

    def set_postprocess(self, color_filter=(1.0, 1.0, 1.0), exposure=0.0, contrast=0.8, saturation=0.98, dithering=0.005, bloom=0.0,
                        tonemapper=capi.RTB_TONEMAP_ACES):
        """Viewport::SetPostprocessParams; defaults = PostprocessParams::PostprocessParams (PostProcess.cpp:6-14)."""
        f = (C.c_float * 3)(*color_filter)
        self._check(self.lib.rtbh_viewport_set_postprocess(self.h, f, exposure, contrast, saturation, dithering, bloom, tonemapper))

    def read_front(self):
        """Viewport::GetFrontBuffer: (height, width) uint32 0x00RRGGBB."""
        out = np.empty((self.height, self.width), dtype=np.uint32)
        self._check(self.lib.rtbh_viewport_read_front(self.h, out.ctypes.data))
        return out

    def counters(self):
        c = capi.RtbCounters()
        self._check(self.lib.rtbh_viewport_counters(self.h, C.byref(c)))
        return {k: int(getattr(c, k)) for k, _ in capi.RtbCounters._fields_}
