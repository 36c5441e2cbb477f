"""Multi-GPU frames through the C ABI and the C++ host: needs two devices on the box
(`gpurun --gpus 2`); skipped otherwise.  The tile-interleaved shards of several devices driven from ONE process
(rt::ViewportGroupB200, rtb200_render --gpus N) give the single-device film bit for bit, both assembled in a page-locked
host frame (rtb_scatter_shard) and all-gathered on the devices by the library's own NCCL communicator
(rtb_comm_init_all / rtb_finish_frame_all)."""
import json
import os
import subprocess

import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu
RENDER = os.path.join(util.ROOT, "raytracer_b200", "rtb200_render")


def device_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def cli(tmp_path, name, extra, gpus):
    scn = os.path.join(util.GOLDEN, "scenes")
    out = str(tmp_path / (name + ".f32"))
    cmd = [RENDER, "-s", os.path.join(scn, "mesh_small.json"), "--data", scn + "/", "-w", "200", "-h", "152", "--spp", "6", "--dims", "128", "--out", out,
           "--bluenoise", os.path.join(util.ROOT, "raytracer_b200", "data", "blue_noise_128.rtbn")] + extra
    if gpus > 1:
        cmd += ["--gpus", str(gpus), "--check-gather", "1"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    meta = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    return np.fromfile(out, dtype=np.float32), meta, r.stderr


@pytest.mark.skipif(device_count() < 2, reason="needs two CUDA devices")
@pytest.mark.parametrize("renderer", ["Path Tracer MIS", "Path Tracer"])
def test_two_devices_from_one_process_equal_one_device(tmp_path, renderer):
    one, m1, _ = cli(tmp_path, "one", ["--renderer", renderer], 1)
    two, m2, err = cli(tmp_path, "two", ["--renderer", renderer], 2)
    assert "all-gathered film on 2 devices == scattered host frame" in err
    util.assert_bit_identical(two, one, "2-device frame vs 1 device")
    for k in ("numRays", "numShadowRays", "numShadowRaysHit", "numPrimaryRays"):
        assert m1[k] == m2[k], k
