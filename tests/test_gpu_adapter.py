"""The reference-side adapter of INTEGRATION.md section 2, compiled for real (oracle/rtadapter.cpp, built by
oracle/build_ref.sh against the reference's own classes): in ONE process the reference's helpers::LoadScene ->
Scene::BuildBVH -> rt::Viewport (CPU) and, on the same Scene / Camera objects, rt::ViewportB200 -> SceneFlattener ->
rtb_* (this repo's library) - the two sum buffers and ray counters must agree bit for bit."""
import json
import os
import subprocess

import pytest

from tests import util

pytestmark = pytest.mark.gpu
ADAPTER = os.path.join(util.ORACLE_EXACT, "rtadapter")


@pytest.mark.skipif(not os.path.exists(ADAPTER), reason="oracle/_ref/exact/rtadapter did not travel to this box")
@pytest.mark.parametrize("scene,renderer,w,h,passes,dims", [
    ("cornell", "Path Tracer MIS", 96, 80, 4, 128),
    ("mesh_small", "Path Tracer MIS", 80, 64, 3, 128),
    ("materials", "Path Tracer MIS", 72, 56, 3, 256),      # six lights: 194 sample dimensions at depth 8
    ("textures", "Path Tracer", 64, 48, 2, 128),
])
def test_reference_scene_through_the_compiled_adapter(scene, renderer, w, h, passes, dims):
    scn = os.path.join(util.GOLDEN, "scenes")
    run_dir = os.path.join(util.ORACLE_EXACT, "run")
    os.makedirs(run_dir, exist_ok=True)
    r = subprocess.run([ADAPTER, "--scene", os.path.join(scn, scene + ".json"), "--data", scn + "/", "--width", str(w), "--height", str(h), "--passes", str(passes),
                        "--renderer", renderer, "--dims", str(dims), "--depth", "8"],
                       cwd=run_dir, env=dict(os.environ, RT_FIXED_SEED="777"), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert lines, r.stderr[-2000:]
    meta = json.loads(lines[-1])
    assert meta["differentPixels"] == 0, meta
    assert (meta["rays"], meta["shadowRays"], meta["shadowRaysHit"]) == (meta["refRays"], meta["refShadowRays"], meta["refShadowRaysHit"]), meta
    assert r.returncode == 0, r.stderr[-2000:]
