// Compile-only check (tests/test_hostsim_parity.py::test_bidirectional_functions_compile_for_the_device): the bidir_* building
// blocks of csrc/rtb_bidir.h instantiate inside a kernel for sm_100a with the product's flags.  Never launched.
#include <cuda_runtime.h>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include "../../raytracer_b200/csrc/rtb_integrator.h"
#include "../../raytracer_b200/csrc/rtb_bidir.h"
using namespace rtb;
struct DevOccluded { const SceneView* S; __device__ bool operator()(const Ray& r, float tMax) const { return scene_traverse_shadow<false>(*S, r, tMax, nullptr); } };
__global__ void k_check(const __grid_constant__ SceneView S, CameraConnection camera, BidirPass pass, BidirLightVertex* lvs, BidirPhoton* photons, float4* out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    BidirPath path;
    if (!bidir_generate_light_sample<true>(S, pass, i, 0.1f, 0.2f, 0.3f, 0.4f, 0.5f, path)) return;
    Hit hit; hit.objectId = RTB_INVALID_OBJECT; hit.subObjectId = 0; hit.distance = INFINITY; hit.u = hit.v = 0.0f;
    scene_traverse<false>(S, path.ray, hit, nullptr);
    if (hit.distance == INFINITY) return;
    BidirLightVertex lv;
    evaluate_intersection<true>(S, path.ray, hit, lv.vx, RTB_NO_MATERIAL);
    lv.vx.outgoingDirWorld = neg(path.ray.dir);
    evaluate_shading_data<true>(S, lv.vx);
    bidir_arrive(path, lv.vx, hit.distance, true);
    lv.pathLength = path.length; lv.throughput = path.throughput; lv.dVC = path.dVC; lv.dVM = path.dVM; lv.dVCM = path.dVCM;
    const DevOccluded occ{ &S };
    float fx, fy; V3 value;
    V4 acc = v4s(0.0f);
    if (bidir_connect_to_camera(pass, camera, lv, occ, fx, fy, value)) acc = acc + v4(value.x, value.y, value.z, fx + fy);
    BidirPath cam; bidir_begin_camera_path(cam, path.ray, camera);
    acc = acc + bidir_connect_vertices(pass, cam, lv.vx, lvs[i], occ);
    acc = acc + bidir_sample_light<true>(S, pass, S.lights[0], lv.vx, cam, 0.3f, 0.6f, 0.9f, occ);
    acc = acc + bidir_evaluate_light<true>(S, pass, S.lights[0], &lv.vx, cam);
    bidir_merge_photon(pass, cam, lv.vx, photons[i], acc);
    if (bidir_advance_path(pass, path, lv.vx, 0.2f, 0.4f, 0.6f)) acc = acc + path.throughput;
    out[i] = make_float4(acc.x, acc.y, acc.z, acc.w);
    lvs[i] = lv;
}
