// Sanitizer harness for the host-side file readers (development aid; not part of the product or the test suites).
//   fuzz_harness bitmap <file>            Bitmap::Load: BMP, DDS, OpenEXR
//   fuzz_harness scene <scene.json> <dir> helpers::LoadScene (JSON, OBJ, MTL, textures) + BuildBVH + Flatten
// Built by tools/loader_fuzz/run.py with -fsanitize=address,undefined.
#include "../../raytracer_b200/host/rt_loader.h"

int main(int argc, char** argv)
{
    if (argc >= 3 && std::string(argv[1]) == "bitmap")
    {
        rt::Bitmap b;
        const bool ok = b.Load(argv[2]);
        printf("%d %u %u %s\n", (int)ok, b.GetWidth(), b.GetHeight(), b.GetError().c_str());
        return 0;
    }
    if (argc >= 4 && std::string(argv[1]) == "scene")
    {
        rt::Scene scene; rt::Camera camera; std::string err;
        bool ok = rt::helpers::LoadScene(argv[2], argv[3], scene, camera, err);
        if (ok) ok = scene.BuildBVH();
        if (ok) { rtb::SceneFile flat; scene.Flatten(flat); camera.Describe(flat.camera); camera.DescribeBokeh(flat); }
        printf("%d %s\n", (int)ok, err.c_str());
        return 0;
    }
    fprintf(stderr, "usage: fuzz_harness bitmap <file> | scene <scene.json> <data dir>\n");
    return 2;
}
