"""Host-side parity (CPU only): the product's own scene model (JSON + OBJ loaders, transforms, BVH builder, mesh
preprocessing) flattens a scene to EXACTLY the arrays the reference's own loader + Scene::BuildBVH produce (the
golden .rtb files were dumped from the reference's Scene object by oracle/rtref.cpp)."""
import ctypes as C
import os

import numpy as np
import pytest

from raytracer_b200 import capi, host_api
from tests import util

SCENES = ["c1_sphere_area", "cornell", "materials", "mesh_small"]


def nodes_view(raw):
    return np.frombuffer(raw.tobytes(), dtype=np.uint32).reshape(-1, 8).copy()


def normalise_nodes(n):
    """node 1 is never written by the builder, and leaves carry no split axis (uninitialised bits in the reference)."""
    n = n.copy()
    if len(n) > 1:
        n[1] = 0
    leaf = (n[:, 7] & 0x3FFFFFFF) != 0
    n[leaf, 7] &= 0x3FFFFFFF
    return n


@pytest.fixture(scope="module")
def built():
    if not os.path.exists(host_api.HOST_LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return True


@pytest.mark.parametrize("name", SCENES)
def test_flattened_scene_matches_reference(built, tmp_path, name):
    scn = os.path.join(util.GOLDEN, "scenes")
    scene = host_api.Scene(os.path.join(scn, name + ".json"), scn + "/")
    out = str(tmp_path / "mine.rtb")
    scene.save_flat(out, 64, 64)
    mine, ref = capi.SceneFile(out), capi.SceneFile(util.golden(name + ".rtb"))

    assert (mine.num_objects, mine.num_shapes, mine.num_materials, mine.num_lights, mine.num_global_lights, len(mine.meshes)) == \
           (ref.num_objects, ref.num_shapes, ref.num_materials, ref.num_lights, ref.num_global_lights, len(ref.meshes))
    assert (normalise_nodes(nodes_view(mine.tlas)) == normalise_nodes(nodes_view(ref.tlas))).all(), "top-level BVH differs"
    for what in ("objects", "shapes", "materials", "lights", "global_lights"):
        assert getattr(mine, what).tobytes() == getattr(ref, what).tobytes(), "%s differ" % what
    for a, b in zip(mine.meshes, ref.meshes):
        for k in ("numNodes", "numTriangles", "numVertices", "numMaterials"):
            assert a[k] == b[k], k
        assert (normalise_nodes(nodes_view(a["nodes"])) == normalise_nodes(nodes_view(b["nodes"]))).all(), "mesh BVH differs"
        for k in ("triangles", "indices", "shading", "materialIds"):
            assert a[k].tobytes() == b[k].tobytes(), "mesh %s differ" % k
    assert bytes(mine.camera) == bytes(ref.camera), "camera differs"
    assert mine.field_of_view == ref.field_of_view


def test_loader_error_paths(built, tmp_path):
    """Error behaviour of helpers::LoadScene (Demo/SceneLoader.cpp): message + failure, never a crash."""
    def load(text):
        p = tmp_path / "s.json"
        p.write_text(text)
        return host_api.Scene(str(p), str(tmp_path) + "/")

    with pytest.raises(capi.RtbError, match="Failed to open"):
        host_api.Scene(str(tmp_path / "missing.json"))
    with pytest.raises(capi.RtbError, match="parse"):
        load("{ not json")
    with pytest.raises(capi.RtbError, match="must be a float"):      # TryParseFloat rejects integer literals
        load('{"objects": [{"type": "sphere", "radius": 1}]}')
    with pytest.raises(capi.RtbError, match="does not exist"):
        load('{"objects": [{"type": "sphere", "radius": 1.0, "material": "nope"}]}')
    with pytest.raises(capi.RtbError, match="Unknown scene object type"):
        load('{"objects": [{"type": "torus"}]}')
    with pytest.raises(capi.RtbError, match="Unknown light type"):
        load('{"lights": [{"type": "sphere", "color": [1.0, 1.0, 1.0], "radius": 1.0}]}')
    with pytest.raises(capi.RtbError, match="Duplicated material"):
        load('{"materials": [{"name": "a"}, {"name": "a"}]}')
    with pytest.raises(capi.RtbError, match="Unknown BSDF"):
        load('{"materials": [{"name": "a", "bsdf": "velvet"}]}')
    with pytest.raises(capi.RtbError, match="textures are not supported"):
        load('{"materials": [{"name": "a", "baseColorTexture": "x.bmp"}]}')
    s = load('{}')                     # empty scene loads (RenderingTest.EmptyScene)
    assert s.stats() == dict(objects=0, lights=0, tlas_nodes=0)
    s = load('{"objects": [{"type": "sphere", "radius": 1.0}]}')          # default material, identity transform
    assert s.stats()["objects"] == 1
