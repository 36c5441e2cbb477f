"""ctypes binding of the C ABI declared in include/rtb200.h (librtb200.so, CUDA sm_100a).

This is plumbing for tests and bench.py: every call goes straight into the extern "C" entry points, the same ones a
C++ host (raytracer_b200/host) or the reference's own Viewport would bind.  There is no CPU fallback: if the shared
library is missing or no CUDA device is present the calls fail loudly.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RTB200_LIB", os.path.join(_HERE, "librtb200.so"))   # override only for A/B builds of the same sources

RTB_INVALID_OBJECT = 0xFFFFFFFF
RTB_LIGHT_OBJECT = 0xFFFFFFFE
RTB_RENDERER_PATH_TRACER, RTB_RENDERER_PATH_TRACER_MIS, RTB_RENDERER_DEBUG, RTB_RENDERER_VCM, RTB_RENDERER_LIGHT_TRACER = 0, 1, 2, 3, 4
RTB_COMM_ID_BYTES = 128
RTB_TRACE_CLOSEST, RTB_TRACE_SHADOW = 0, 1
RTB_LIGHT_SAMPLING_SINGLE, RTB_LIGHT_SAMPLING_ALL = 0, 1
RTB_FLAG_COUNT_TESTS = 1
RENDERER_NAMES = {"Path Tracer": 0, "Path Tracer MIS": 1, "Debug": 2}   # CreateRenderer names, Renderer.cpp:45-69


class RtbBvhNode(C.Structure):
    _fields_ = [("min", C.c_float * 3), ("childIndex", C.c_uint32), ("max", C.c_float * 3), ("leafInfo", C.c_uint32)]


class RtbMat4(C.Structure):
    _fields_ = [("r", (C.c_float * 4) * 4)]


class RtbShape(C.Structure):
    _fields_ = [("type", C.c_uint32), ("meshIndex", C.c_uint32), ("pad0", C.c_uint32), ("pad1", C.c_uint32),
                ("size", C.c_float * 4), ("invSize", C.c_float * 4), ("aux", C.c_float * 4)]


class RtbMaterial(C.Structure):
    _fields_ = [("bsdf", C.c_uint32), ("flags", C.c_uint32), ("roughness", C.c_float), ("metalness", C.c_float),
                ("baseColor", C.c_float * 4), ("emission", C.c_float * 4), ("IoR", C.c_float), ("K", C.c_float),
                ("pad0", C.c_float), ("pad1", C.c_float)]


class RtbTexture(C.Structure):
    _fields_ = [("type", C.c_uint32), ("width", C.c_uint32), ("height", C.c_uint32), ("numOctaves", C.c_uint32),
                ("texA", C.c_uint32), ("texB", C.c_uint32), ("texWeight", C.c_uint32), ("pad0", C.c_uint32),
                ("colorA", C.c_float * 4), ("colorB", C.c_float * 4), ("texels", C.c_void_p)]


class RtbMaterialTextures(C.Structure):
    _fields_ = [("baseColor", C.c_uint32), ("emission", C.c_uint32), ("roughness", C.c_uint32), ("metalness", C.c_uint32),
                ("normalMap", C.c_uint32), ("maskMap", C.c_uint32), ("normalMapStrength", C.c_float), ("pad0", C.c_uint32)]


RTB_TEXTURE_BITMAP, RTB_TEXTURE_CHECKERBOARD, RTB_TEXTURE_NOISE, RTB_TEXTURE_MIX = 0, 1, 2, 3
RTB_NO_TEXTURE = 0xFFFFFFFF


class RtbLight(C.Structure):
    _fields_ = [("type", C.c_uint32), ("flags", C.c_uint32), ("shapeIndex", C.c_uint32), ("isDelta", C.c_uint32),
                ("color", C.c_float * 4), ("cosAngle", C.c_float), ("pad0", C.c_float), ("pad1", C.c_float), ("pad2", C.c_float),
                ("transform", RtbMat4), ("invTransform", RtbMat4)]


class RtbObject(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("shapeIndex", C.c_uint32), ("materialIndex", C.c_uint32), ("lightIndex", C.c_uint32),
                ("transform", RtbMat4), ("invTransform", RtbMat4)]


class RtbMesh(C.Structure):
    _fields_ = [("numNodes", C.c_uint32), ("numTriangles", C.c_uint32), ("numVertices", C.c_uint32), ("numMaterials", C.c_uint32),
                ("nodes", C.c_void_p), ("triangles", C.c_void_p), ("indices", C.c_void_p), ("shading", C.c_void_p),
                ("materialIds", C.c_void_p)]


class RtbSceneDesc(C.Structure):
    _fields_ = [("numTlasNodes", C.c_uint32), ("tlasNodes", C.c_void_p),
                ("numObjects", C.c_uint32), ("objects", C.c_void_p),
                ("numShapes", C.c_uint32), ("shapes", C.c_void_p),
                ("numMeshes", C.c_uint32), ("meshes", C.c_void_p),
                ("numMaterials", C.c_uint32), ("materials", C.c_void_p),
                ("numLights", C.c_uint32), ("lights", C.c_void_p),
                ("numGlobalLights", C.c_uint32), ("globalLights", C.c_void_p),
                ("numTextures", C.c_uint32), ("textures", C.c_void_p),
                ("materialTextures", C.c_void_p), ("lightTextures", C.c_void_p),
                ("bokehWidth", C.c_uint32), ("bokehHeight", C.c_uint32), ("bokehCdf", C.c_void_p)]


class RtbCamera(C.Structure):
    _fields_ = [("localToWorld", RtbMat4), ("aspectRatio", C.c_float), ("tanHalfFoV", C.c_float),
                ("dofEnable", C.c_uint32), ("bokehShape", C.c_uint32), ("aperture", C.c_float),
                ("focalPlaneDistance", C.c_float), ("barrelDistortionConstFactor", C.c_float),
                ("barrelDistortionVariableFactor", C.c_float)]


class RtbPassDesc(C.Structure):
    _fields_ = [("renderer", C.c_uint32), ("debugMode", C.c_uint32), ("camera", RtbCamera),
                ("maxRayDepth", C.c_uint32), ("minRussianRouletteDepth", C.c_uint32),
                ("lightSamplingStrategy", C.c_uint32), ("useBlueNoise", C.c_uint32),
                ("passIndex", C.c_uint32), ("numSeedDims", C.c_uint32), ("seed", C.c_void_p),
                ("sampleOffset", C.c_float * 2)]


RTB_TONEMAP_CLAMPED, RTB_TONEMAP_REINHARD, RTB_TONEMAP_JIM_HEJL, RTB_TONEMAP_ACES = 0, 1, 2, 3


class RtbPostprocessDesc(C.Structure):
    """PostprocessParams (Rendering/PostProcess.h:11-29) + the pass count the film holds."""
    _fields_ = [("colorFilter", C.c_float * 4), ("exposure", C.c_float), ("contrast", C.c_float), ("saturation", C.c_float),
                ("ditheringStrength", C.c_float), ("bloomFactor", C.c_float), ("tonemapper", C.c_uint32),
                ("numPasses", C.c_uint32), ("ditherSeed", C.c_uint32)]


def postprocess_desc(num_passes, color_filter=(1.0, 1.0, 1.0), exposure=0.0, contrast=0.8, saturation=0.98, dithering=0.005,
                     bloom=0.0, tonemapper=RTB_TONEMAP_ACES, dither_seed=0):
    """Defaults = PostprocessParams::PostprocessParams (Rendering/PostProcess.cpp:6-14)."""
    d = RtbPostprocessDesc()
    d.colorFilter[:] = [color_filter[0], color_filter[1], color_filter[2], 1.0]
    d.exposure, d.contrast, d.saturation, d.ditheringStrength, d.bloomFactor = exposure, contrast, saturation, dithering, bloom
    d.tonemapper, d.numPasses, d.ditherSeed = tonemapper, num_passes, dither_seed
    return d


class RtbVcmParams(C.Structure):
    _fields_ = [("useVertexConnection", C.c_uint32), ("useVertexMerging", C.c_uint32), ("maxPathLength", C.c_uint32),
                ("initialMergingRadius", C.c_float), ("minMergingRadius", C.c_float), ("mergingRadiusMultiplier", C.c_float),
                ("lightPathSeed", C.c_uint64)]


class RtbCounters(C.Structure):
    _fields_ = [("numRays", C.c_uint64), ("numShadowRays", C.c_uint64), ("numShadowRaysHit", C.c_uint64),
                ("numPrimaryRays", C.c_uint64), ("numRayBoxTests", C.c_uint64), ("numRayTriangleTests", C.c_uint64),
                ("numShadowBoxTests", C.c_uint64), ("numShadowTriangleTests", C.c_uint64), ("numKernelLaunches", C.c_uint64)]


class RtbCreateDesc(C.Structure):
    _fields_ = [("deviceId", C.c_int32), ("flags", C.c_uint32), ("numShards", C.c_uint32), ("shardIndex", C.c_uint32),
                ("tileSize", C.c_uint32), ("stream", C.c_void_p)]


class RtbPassTimings(C.Structure):
    _fields_ = [("raygenMs", C.c_float), ("traverseMs", C.c_float), ("shadeMs", C.c_float), ("shadowMs", C.c_float),
                ("totalMs", C.c_float), ("traverseLaunches", C.c_uint32), ("shadeLaunches", C.c_uint32),
                ("shadowLaunches", C.c_uint32), ("totalLaunches", C.c_uint32), ("traverseRays", C.c_uint64),
                ("shadowRays", C.c_uint64)]


RAY_DTYPE = np.dtype([("origin", np.float32, 4), ("dir", np.float32, 4), ("invDir", np.float32, 4), ("originDivDir", np.float32, 4)])
HIT_DTYPE = np.dtype([("objectId", np.uint32), ("subObjectId", np.uint32), ("distance", np.float32), ("u", np.float32),
                      ("v", np.float32), ("occluded", np.uint32)])

# every symbol include/rtb200.h declares (checked by tests/test_capi_symbols.py)
EXPORTED_SYMBOLS = [
    "rtb_last_error", "rtb_api_version", "rtb_create", "rtb_destroy", "rtb_upload_scene", "rtb_set_blue_noise",
    "rtb_resize", "rtb_reset", "rtb_render_pass", "rtb_render_passes", "rtb_synchronize", "rtb_join", "rtb_read_sum", "rtb_film_device_buffer", "rtb_film_snapshots",
    "rtb_unshard_film", "rtb_read_sum_async", "rtb_unshard_film_async", "rtb_host_alloc", "rtb_host_free", "rtb_get_counters", "rtb_trace_batch", "rtb_trace_batch_device", "rtb_texture_batch", "rtb_sample_batch",
    "rtb_set_vcm_params", "rtb_vcm_splat_buffers", "rtb_vcm_fold_splats",
    "rtb_scatter_shard", "rtb_scatter_shard_async", "rtb_comm_unique_id", "rtb_comm_init", "rtb_comm_init_all", "rtb_finish_frame", "rtb_finish_frame_all",
    "rtb_host_register", "rtb_host_unregister", "rtb_host_device_pointer",
    "rtb_primary_rays", "rtb_set_profiling", "rtb_get_pass_timings", "rtb_read_front", "rtb_block_errors", "rtb_set_active_tiles",
]

_lib = None


class RtbError(RuntimeError):
    pass


def load_library(path=None):
    """dlopen librtb200.so (no fallback: raises if it has not been built)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RtbError("%s not found - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(raytracer_b200/csrc/build.sh); there is no CPU fallback" % p)
    lib = C.CDLL(p)
    lib.rtb_last_error.restype = C.c_char_p
    lib.rtb_api_version.restype = C.c_int
    lib.rtb_create.argtypes = [C.POINTER(RtbCreateDesc), C.POINTER(C.c_void_p)]
    lib.rtb_destroy.argtypes = [C.c_void_p]
    lib.rtb_destroy.restype = None
    lib.rtb_upload_scene.argtypes = [C.c_void_p, C.POINTER(RtbSceneDesc)]
    lib.rtb_set_blue_noise.argtypes = [C.c_void_p, C.c_void_p]
    lib.rtb_resize.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
    lib.rtb_reset.argtypes = [C.c_void_p]
    lib.rtb_render_pass.argtypes = [C.c_void_p, C.POINTER(RtbPassDesc)]
    lib.rtb_render_passes.argtypes = [C.c_void_p, C.POINTER(RtbPassDesc), C.c_uint32]
    lib.rtb_synchronize.argtypes = [C.c_void_p]
    lib.rtb_join.argtypes = [C.c_void_p]
    lib.rtb_read_sum.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    lib.rtb_film_device_buffer.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    lib.rtb_unshard_film.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.rtb_film_snapshots.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    lib.rtb_read_sum_async.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    lib.rtb_unshard_film_async.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.rtb_host_alloc.argtypes = [C.c_uint64]
    lib.rtb_host_alloc.restype = C.c_void_p
    lib.rtb_host_free.argtypes = [C.c_void_p]
    lib.rtb_host_free.restype = None
    lib.rtb_get_counters.argtypes = [C.c_void_p, C.POINTER(RtbCounters)]
    lib.rtb_trace_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32]
    lib.rtb_trace_batch_device.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.POINTER(C.c_float)]
    lib.rtb_texture_batch.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p]
    lib.rtb_block_errors.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p]
    lib.rtb_set_active_tiles.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    lib.rtb_read_front.argtypes = [C.c_void_p, C.POINTER(RtbPostprocessDesc), C.c_void_p, C.c_void_p]
    lib.rtb_sample_batch.argtypes = [C.c_void_p, C.POINTER(RtbPassDesc), C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p]
    lib.rtb_primary_rays.argtypes = [C.c_void_p, C.POINTER(RtbPassDesc), C.c_void_p]
    lib.rtb_set_profiling.argtypes = [C.c_void_p, C.c_int]
    lib.rtb_set_vcm_params.argtypes = [C.c_void_p, C.POINTER(RtbVcmParams)]
    lib.rtb_vcm_splat_buffers.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    lib.rtb_vcm_fold_splats.argtypes = [C.c_void_p]
    lib.rtb_scatter_shard.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.rtb_scatter_shard_async.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.rtb_comm_unique_id.argtypes = [C.c_void_p]
    lib.rtb_comm_init.argtypes = [C.c_void_p, C.c_char_p]
    lib.rtb_comm_init_all.argtypes = [C.POINTER(C.c_void_p), C.c_uint32]
    lib.rtb_finish_frame.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
    lib.rtb_finish_frame_all.argtypes = [C.POINTER(C.c_void_p), C.c_uint32, C.c_int, C.POINTER(C.c_void_p)]
    lib.rtb_host_register.argtypes = [C.c_void_p, C.c_uint64]
    lib.rtb_host_unregister.argtypes = [C.c_void_p]
    lib.rtb_host_device_pointer.argtypes = [C.c_void_p]
    lib.rtb_host_device_pointer.restype = C.c_void_p
    lib.rtb_get_pass_timings.argtypes = [C.c_void_p, C.POINTER(RtbPassTimings)]
    if path is None:
        _lib = lib
    return lib


def _check(lib, rc):
    if rc != 0:
        raise RtbError("rtb error %d: %s" % (rc, lib.rtb_last_error().decode()))


# ---- "RTBS" flat scene files (raytracer_b200/csrc/rtb_scene_file.h) ------------------------------------------------

class SceneFile:
    """Parsed RTBS file: numpy views of every array + a ctypes RtbSceneDesc pointing into them."""

    def __init__(self, path):
        data = np.fromfile(path, dtype=np.uint8)
        self._buf = data
        pos = 0

        def u32():
            nonlocal pos
            v = int(data[pos:pos + 4].view(np.uint32)[0])
            pos += 4
            return v

        def arr(nbytes):
            nonlocal pos
            a = np.ascontiguousarray(data[pos:pos + nbytes])
            pos += nbytes
            return a

        assert u32() == 0x53425452, "not an RTBS file"
        version = u32()
        assert version in (1, 2), "not an RTBS v1/v2 file"
        n = u32(); self.tlas = arr(n * C.sizeof(RtbBvhNode)); self.num_tlas = n
        n = u32(); self.objects = arr(n * C.sizeof(RtbObject)); self.num_objects = n
        n = u32(); self.shapes = arr(n * C.sizeof(RtbShape)); self.num_shapes = n
        self.meshes = []
        for _ in range(u32()):
            nn, nt, nv, nm = u32(), u32(), u32(), u32()
            self.meshes.append(dict(numNodes=nn, numTriangles=nt, numVertices=nv, numMaterials=nm,
                                    nodes=arr(nn * 32), triangles=arr(nt * 36), indices=arr(nt * 16),
                                    shading=arr(nv * 32), materialIds=arr(nm * 4)))
        n = u32(); self.materials = arr(n * C.sizeof(RtbMaterial)); self.num_materials = n
        n = u32(); self.lights = arr(n * C.sizeof(RtbLight)); self.num_lights = n
        n = u32(); self.global_lights = arr(n * 4); self.num_global_lights = n
        self.camera = RtbCamera.from_buffer_copy(arr(C.sizeof(RtbCamera)).tobytes())
        self.field_of_view = float(arr(4).view(np.float32)[0])
        self.num_triangles = sum(m["numTriangles"] for m in self.meshes)
        # version 2: textures
        self.textures, self.material_textures, self.light_textures = [], None, None
        self.bokeh_size, self.bokeh_cdf = (0, 0), None
        if version >= 2:
            for _ in range(u32()):
                h = arr(32).view(np.uint32)
                ca, cb = arr(16).view(np.float32), arr(16).view(np.float32)
                t = dict(type=int(h[0]), width=int(h[1]), height=int(h[2]), numOctaves=int(h[3]), texA=int(h[4]), texB=int(h[5]),
                         texWeight=int(h[6]), colorA=ca, colorB=cb, texels=None)
                if t["type"] == RTB_TEXTURE_BITMAP:
                    t["texels"] = arr(t["width"] * t["height"] * 16)
                self.textures.append(t)
            n = u32()
            if n:
                self.material_textures = arr(n * C.sizeof(RtbMaterialTextures))
            n = u32()
            if n:
                self.light_textures = arr(n * 4)
            if pos + 8 <= len(data):
                bw, bh = u32(), u32()
                if bw and bh:
                    self.bokeh_size, self.bokeh_cdf = (bw, bh), arr((bw * bh + 1) * 4)

    def desc(self):
        self._mesh_structs = (RtbMesh * max(1, len(self.meshes)))()
        for i, m in enumerate(self.meshes):
            s = self._mesh_structs[i]
            s.numNodes, s.numTriangles, s.numVertices, s.numMaterials = m["numNodes"], m["numTriangles"], m["numVertices"], m["numMaterials"]
            s.nodes, s.triangles, s.indices = m["nodes"].ctypes.data, m["triangles"].ctypes.data, m["indices"].ctypes.data
            s.shading, s.materialIds = m["shading"].ctypes.data, m["materialIds"].ctypes.data
        d = RtbSceneDesc()
        d.numTlasNodes, d.tlasNodes = self.num_tlas, self.tlas.ctypes.data
        d.numObjects, d.objects = self.num_objects, self.objects.ctypes.data
        d.numShapes, d.shapes = self.num_shapes, self.shapes.ctypes.data
        d.numMeshes, d.meshes = len(self.meshes), C.addressof(self._mesh_structs)
        d.numMaterials, d.materials = self.num_materials, self.materials.ctypes.data
        d.numLights, d.lights = self.num_lights, self.lights.ctypes.data
        d.numGlobalLights, d.globalLights = self.num_global_lights, self.global_lights.ctypes.data
        self._texture_structs = (RtbTexture * max(1, len(self.textures)))()
        for i, t in enumerate(self.textures):
            s = self._texture_structs[i]
            s.type, s.width, s.height, s.numOctaves = t["type"], t["width"], t["height"], t["numOctaves"]
            s.texA, s.texB, s.texWeight = t["texA"], t["texB"], t["texWeight"]
            for k in range(4):
                s.colorA[k], s.colorB[k] = float(t["colorA"][k]), float(t["colorB"][k])
            s.texels = t["texels"].ctypes.data if t["texels"] is not None else None
        d.numTextures = len(self.textures)
        d.textures = C.addressof(self._texture_structs) if self.textures else None
        d.materialTextures = self.material_textures.ctypes.data if self.material_textures is not None else None
        d.lightTextures = self.light_textures.ctypes.data if self.light_textures is not None else None
        bokeh = getattr(self, "bokeh_cdf", None)
        if bokeh is not None:
            d.bokehWidth, d.bokehHeight, d.bokehCdf = self.bokeh_size[0], self.bokeh_size[1], bokeh.ctypes.data
        return d


class PassScript:
    """Per-pass Halton seed vectors + AA offsets (written by the oracle's `rtref render --dump-passes`)."""

    def __init__(self, path=None, seeds=None, offsets=None):
        if path is not None:
            raw = np.fromfile(path, dtype=np.uint32)
            self.num_passes, self.num_dims = int(raw[0]), int(raw[1])
            body = raw[2:].reshape(self.num_passes, self.num_dims + 2)
            self.seeds = np.ascontiguousarray(body[:, :self.num_dims])
            self.offsets = np.ascontiguousarray(body[:, self.num_dims:]).view(np.float32)
        else:
            self.seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
            self.offsets = np.ascontiguousarray(offsets, dtype=np.float32)
            self.num_passes, self.num_dims = self.seeds.shape


class PinnedImage:
    """height x width x 3 float32 array in page-locked host memory (rtb_host_alloc), for rtb_read_sum / rtb_unshard_film."""

    def __init__(self, width, height):
        self.lib = load_library()
        self.nbytes = width * height * 12
        self.ptr = self.lib.rtb_host_alloc(self.nbytes)
        if not self.ptr:
            raise RtbError(self.lib.rtb_last_error().decode())
        buf = (C.c_float * (width * height * 3)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=np.float32).reshape(height, width, 3)

    def close(self):
        if self.ptr:
            self.array = None
            self.lib.rtb_host_free(self.ptr)
            self.ptr = None


def load_blue_noise(path=None):
    """128x128x4 uint16 table, returned in the interleaved RGBA16 order GenericSampler reads (GenericSampler.cpp:91-96)."""
    p = path or os.path.join(_HERE, "data", "blue_noise_128.rtbn")
    raw = np.fromfile(p, dtype=np.uint16)
    assert raw.size == 4 + 4 * 128 * 128 and raw[0] == 0x5442 and raw[1] == 0x4E52, "bad blue noise file"
    planar = raw[4:].reshape(4, 128, 128)
    return np.ascontiguousarray(np.transpose(planar, (1, 2, 0)))


class Context:
    """Thin RAII wrapper around RtbContext*."""

    def set_vcm_params(self, connection=True, merging=True, max_path_length=10, initial_radius=0.02, min_radius=0.02, radius_multiplier=1.0, seed=0x5EED):
        """VertexConnectionAndMerging's public members (defaults: its constructor's)."""
        p = RtbVcmParams(1 if connection else 0, 1 if merging else 0, max_path_length, initial_radius, min_radius, radius_multiplier, seed)
        _check(self.lib, self.lib.rtb_set_vcm_params(self.h, C.byref(p)))

    def __init__(self, device=0, flags=0, num_shards=1, shard_index=0, tile_size=0, stream=None, blue_noise=True):
        self.lib = load_library()
        d = RtbCreateDesc(device, flags, num_shards, shard_index, tile_size, stream)
        h = C.c_void_p()
        _check(self.lib, self.lib.rtb_create(C.byref(d), C.byref(h)))
        self.h = h
        self.width = self.height = 0
        if blue_noise:
            bn = load_blue_noise()
            _check(self.lib, self.lib.rtb_set_blue_noise(self.h, bn.ctypes.data))

    def close(self):
        if getattr(self, "h", None):
            self.lib.rtb_destroy(self.h)
            self.h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def upload_scene(self, scene_file):
        d = scene_file.desc()
        _check(self.lib, self.lib.rtb_upload_scene(self.h, C.byref(d)))

    def resize(self, w, h):
        _check(self.lib, self.lib.rtb_resize(self.h, w, h))
        self.width, self.height = w, h

    def reset(self):
        _check(self.lib, self.lib.rtb_reset(self.h))

    def make_pass(self, camera, seeds, offset, pass_index, renderer=RTB_RENDERER_PATH_TRACER_MIS, max_depth=8,
                  rr_depth=1, strategy=RTB_LIGHT_SAMPLING_ALL, blue_noise=True, debug_mode=1):
        p = RtbPassDesc()
        p.renderer, p.debugMode = renderer, debug_mode
        p.camera = camera
        p.maxRayDepth, p.minRussianRouletteDepth = max_depth, rr_depth
        p.lightSamplingStrategy, p.useBlueNoise = strategy, 1 if blue_noise else 0
        p.passIndex, p.numSeedDims = pass_index, len(seeds)
        self._seed_keepalive = np.ascontiguousarray(seeds, dtype=np.uint32)
        p.seed = self._seed_keepalive.ctypes.data
        p.sampleOffset[0], p.sampleOffset[1] = float(offset[0]), float(offset[1])
        return p

    def render_pass(self, p):
        _check(self.lib, self.lib.rtb_render_pass(self.h, C.byref(p)))

    def render_passes(self, camera, seeds, offsets, first_pass=0, **kw):
        """rtb_render_passes: seeds (n, dims) uint32, offsets (n, 2) float32; consecutive passes share wavefronts."""
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        n = len(seeds)
        arr = (RtbPassDesc * n)()
        for i in range(n):
            p = self.make_pass(camera, seeds[i], offsets[i], first_pass + i, **kw)
            p.seed = seeds[i].ctypes.data
            arr[i] = p
        self._batch_keepalive = seeds
        _check(self.lib, self.lib.rtb_render_passes(self.h, arr, n))

    def synchronize(self):
        _check(self.lib, self.lib.rtb_synchronize(self.h))

    def read_sum(self, which=0, out=None):
        if out is None:
            out = np.empty((self.height, self.width, 3), dtype=np.float32)
        _check(self.lib, self.lib.rtb_read_sum(self.h, out.ctypes.data, which))
        return out

    def block_errors(self, num_passes, blocks, gathered_sum=None, gathered_sum2=None):
        """rtb_block_errors: Viewport::ComputeBlockError of each (minX, maxX, minY, maxY) row -> float32 array."""
        blocks = np.ascontiguousarray(blocks, dtype=np.uint32).reshape(-1, 4)
        out = np.zeros(len(blocks), dtype=np.float32)
        _check(self.lib, self.lib.rtb_block_errors(self.h, gathered_sum, gathered_sum2, num_passes, blocks.ctypes.data, len(blocks), out.ctypes.data))
        return out

    def set_active_tiles(self, tiles):
        """rtb_set_active_tiles: None = whole image; else (n, 4) uint32 rows minX, maxX, minY, maxY (n may be 0)."""
        if tiles is None:
            _check(self.lib, self.lib.rtb_set_active_tiles(self.h, None, 0))
            return
        tiles = np.ascontiguousarray(tiles, dtype=np.uint32).reshape(-1, 4)
        keep = tiles if len(tiles) else np.zeros((1, 4), dtype=np.uint32)       # a non-null pointer even for an empty list
        _check(self.lib, self.lib.rtb_set_active_tiles(self.h, keep.ctypes.data, len(tiles)))

    def read_front(self, post, gathered_device_ptr=None):
        """rtb_read_front: Viewport::GetFrontBuffer - (height, width) uint32 0x00RRGGBB."""
        out = np.empty((self.height, self.width), dtype=np.uint32)
        _check(self.lib, self.lib.rtb_read_front(self.h, C.byref(post), gathered_device_ptr, out.ctypes.data))
        return out

    def texture_batch(self, texture_index, uv):
        """rtb_texture_batch: ITexture::Evaluate at uv (n, 2) float32 -> (n, 4) float32."""
        uv = np.ascontiguousarray(uv, dtype=np.float32)
        out = np.empty((len(uv), 4), dtype=np.float32)
        _check(self.lib, self.lib.rtb_texture_batch(self.h, texture_index, uv.ctypes.data, len(uv), out.ctypes.data))
        return out

    def film_snapshots(self, max_passes):
        """rtb_film_snapshots: (device pointer, stride in floats) of the per-pass film snapshots."""
        p, n = C.c_void_p(), C.c_uint64()
        _check(self.lib, self.lib.rtb_film_snapshots(self.h, max_passes, C.byref(p), C.byref(n)))
        return p.value, n.value

    def film_device_buffer(self, which=0):
        p, n = C.c_void_p(), C.c_uint64()
        _check(self.lib, self.lib.rtb_film_device_buffer(self.h, which, C.byref(p), C.byref(n)))
        return p.value, n.value

    def read_sum_async(self, pinned, which=0):
        """Snapshot + asynchronous copy into a PinnedImage; complete after synchronize()."""
        _check(self.lib, self.lib.rtb_read_sum_async(self.h, pinned.ptr, which))

    def unshard_film(self, gathered_device_ptr):
        out = np.empty((self.height, self.width, 3), dtype=np.float32)
        _check(self.lib, self.lib.rtb_unshard_film(self.h, gathered_device_ptr, out.ctypes.data))
        return out

    def vcm_splat_buffers(self):
        """rtb_vcm_splat_buffers as two torch tensors (views of the device buffers) for the caller's all-reduce."""
        import torch
        a, b, n = C.c_void_p(), C.c_void_p(), C.c_uint64()
        _check(self.lib, self.lib.rtb_vcm_splat_buffers(self.h, C.byref(a), C.byref(b), C.byref(n)))

        class _View:
            def __init__(self, ptr, count):
                self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None}
        return torch.as_tensor(_View(a.value, n.value), device="cuda"), torch.as_tensor(_View(b.value, n.value), device="cuda")

    def vcm_fold_splats(self):
        _check(self.lib, self.lib.rtb_vcm_fold_splats(self.h))

    def counters(self):
        c = RtbCounters()
        _check(self.lib, self.lib.rtb_get_counters(self.h, C.byref(c)))
        return {k: int(getattr(c, k)) for k, _ in RtbCounters._fields_}

    def trace_batch(self, rays, mode=RTB_TRACE_CLOSEST, tmax=None):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.zeros(len(rays), dtype=HIT_DTYPE)
        if tmax is not None:
            hits["distance"] = tmax
        _check(self.lib, self.lib.rtb_trace_batch(self.h, rays.ctypes.data, len(rays), hits.ctypes.data, mode))
        return hits

    def trace_batch_device(self, rays_ptr, n, hits_ptr, mode=RTB_TRACE_CLOSEST):
        ms = C.c_float()
        _check(self.lib, self.lib.rtb_trace_batch_device(self.h, rays_ptr, n, hits_ptr, mode, C.byref(ms)))
        return ms.value

    def sample_batch(self, p, xy, num_dims):
        xy = np.ascontiguousarray(xy, dtype=np.uint32)
        out = np.empty((len(xy), num_dims), dtype=np.uint32)
        _check(self.lib, self.lib.rtb_sample_batch(self.h, C.byref(p), xy.ctypes.data, len(xy), num_dims, out.ctypes.data))
        return out

    def primary_rays(self, p):
        out = np.empty(self.width * self.height, dtype=RAY_DTYPE)
        _check(self.lib, self.lib.rtb_primary_rays(self.h, C.byref(p), out.ctypes.data))
        return out

    def set_profiling(self, enable):
        _check(self.lib, self.lib.rtb_set_profiling(self.h, 1 if enable else 0))

    def pass_timings(self):
        t = RtbPassTimings()
        _check(self.lib, self.lib.rtb_get_pass_timings(self.h, C.byref(t)))
        return {k: getattr(t, k) for k, _ in RtbPassTimings._fields_}


def camera_for(scene_file, width, height):
    """Camera of the scene file at a given aspect ratio (Camera::SetPerspective, Camera.cpp:30-49).

    tanHalfFoV is taken from the file (the host computed it with tanf, Camera.cpp:37); only the aspect changes.
    """
    cam = RtbCamera.from_buffer_copy(bytes(scene_file.camera))
    cam.aspectRatio = float(np.float32(width) / np.float32(height))
    return cam
