// oracle/ref_harness.cpp — TEST INFRASTRUCTURE, not product code.
//
// Drives the UNMODIFIED reference implementation of the render path (its own
// Camera::createEyeRay, calculatePixel, getCollision, SceneObject::collision, cosineShade,
// inShade ... compiled from /root/reference/src by oracle/build_ref.sh) at a run-time
// chosen resolution / recursion cap / scene, and dumps the pixel array as raw float32.
// The reference fixes all of these at compile time (rt_project_parameters.h:27,65-66,73);
// build_ref.sh turns MAX_RECURSION_LEVEL into the global below and this file supplies the
// pixel loop, which is the two statements of RayTracer.cpp:916-920 verbatim in meaning:
//     ray = my_camera.createEyeRay(((float) x) / W, ((float) z) / H);
//     pixels[x][z] = calculatePixel(ray, 0);
//
// Only the public construction API is used to make scenes (scenes/scene_builders.inc);
// the default and two-mirror scenes are the reference's own Scene::initialize /
// Scene::initializeTwoMirrors.
//
// Output: <out.f32> = (x1-x0)*H*3 floats, x-major, z fastest, (r,g,b) per pixel — the
// layout of the reference's pixels[W][H] (RayTracer.h:44).  With --txt the same values
// are also written with the reference's pixel-line format string (RayTracer.cpp:1601).
//
// usage: ref_render <scene> <W> <H> <max_depth> <x0> <x1> <out.f32|-> [--txt file] [--stride s]
//   scene: default | two_mirrors | synth1024 | synth256 | random:<seed>:<n_objects> | boxes:<seed>:<n_boxes>
//   --stride s : render only columns x0, x0+s, x0+2s, ... (bounded CPU samples for bench.py)

int g_ref_max_depth = 50;   // replaces the literal of rt_project_parameters.h:73
// ref_count only (build_ref.sh (c)): [0] primary, [1] shadow, [2] reflection rays
unsigned long long g_ref_rays[3] = {0, 0, 0};

#define main tcrt_reference_main_unused
#include "RayTracer.cpp"
#undef main
#include "Camera.cpp"
#include "SceneObject.cpp"
#include "SceneSphere.cpp"
#include "SceneInfinitePlane.cpp"
#include "SceneFinitePlane.cpp"
#include "Scene.cpp"            // brings in Color_Values.h (definitions, one TU only)

#include "scene_builders.inc"

#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <vector>

static double now_s() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int main(int argc, char** argv) {
    if (argc < 8) {
        fprintf(stderr, "usage: %s <scene> <W> <H> <max_depth> <x0> <x1> <out.f32|-> [--txt file] [--stride s]\n", argv[0]);
        return 2;
    }
    const char* scene_name = argv[1];
    int W = atoi(argv[2]), H = atoi(argv[3]);
    g_ref_max_depth = atoi(argv[4]);
    int x0 = atoi(argv[5]), x1 = atoi(argv[6]);
    const char* out_path = argv[7];
    const char* txt_path = NULL;
    int stride = 1;
    for (int i = 8; i < argc; i++) {
        if (!strcmp(argv[i], "--txt") && i + 1 < argc) txt_path = argv[++i];
        else if (!strcmp(argv[i], "--stride") && i + 1 < argc) stride = atoi(argv[++i]);
    }
    if (W <= 0 || H <= 0 || x0 < 0 || x1 > W || x0 > x1 || stride < 1) {
        fprintf(stderr, "bad geometry\n");
        return 2;
    }

    // the reference's own global scene / camera (RayTracer.h:50-51)
    if (!strcmp(scene_name, "default")) {
        if (my_scene.initialize()) return 1;
    } else if (!strcmp(scene_name, "two_mirrors")) {
        if (my_scene.initializeTwoMirrors(&my_camera)) return 1;
    } else if (!strcmp(scene_name, "synth1024")) {
        tcrt_scenes::build_synth1024(my_scene);
    } else if (!strcmp(scene_name, "synth256")) {
        tcrt_scenes::build_synth256(my_scene);
    } else if (!strncmp(scene_name, "random:", 7)) {
        unsigned int seed = 0; int n = 0;
        if (sscanf(scene_name + 7, "%u:%d", &seed, &n) != 2) { fprintf(stderr, "bad random spec\n"); return 2; }
        tcrt_scenes::build_random(my_scene, seed, n);
    } else if (!strncmp(scene_name, "lattice:", 8)) {
        unsigned int seed = 0; int n = 0;
        if (sscanf(scene_name + 8, "%u:%d", &seed, &n) != 2) { fprintf(stderr, "bad lattice spec\n"); return 2; }
        tcrt_scenes::build_lattice(my_scene, seed, n);
    } else if (!strncmp(scene_name, "boxes:", 6)) {
        unsigned int seed = 0; int n = 0;
        if (sscanf(scene_name + 6, "%u:%d", &seed, &n) != 2) { fprintf(stderr, "bad boxes spec\n"); return 2; }
        tcrt_scenes::build_boxes(my_scene, seed, n);
    } else {
        fprintf(stderr, "unknown scene %s\n", scene_name);
        return 2;
    }
    // The reference's shadow loop runs over [getSceneObjectStartIndex(), getSceneObjectFinalIndex())
    // (RayTracer.cpp:717-719), which only Scene::initialize / initializeTwoMirrors set (Scene.cpp:
    // 204-205,383-384) — addObject does not.  For a scene assembled through addObject alone the public
    // Scene::SetObjectIndices(rank 0, group of 1) (Scene.cpp:486-505) sets the same [0, object_count).
    my_scene.SetObjectIndices(0, 1);

    int ncols = (x1 - x0 + stride - 1) / stride;
    std::vector<float> out((size_t)ncols * (size_t)H * 3u);
    double t0 = now_s();
    size_t k = 0;
    for (int x = x0; x < x1; x += stride) {
        for (int z = 0; z < H; z++) {
            CelioRayTracer::Ray* ray = my_camera.createEyeRay(((float)x) / W, ((float)z) / H);
            CelioRayTracer::vector3d c = calculatePixel(ray, 0);
            free(ray);
            out[k++] = c.x; out[k++] = c.y; out[k++] = c.z;
        }
    }
    double t1 = now_s();

    if (strcmp(out_path, "-")) {
        FILE* f = fopen(out_path, "wb");
        if (!f) { fprintf(stderr, "cannot open %s\n", out_path); return 1; }
        fwrite(out.data(), sizeof(float), out.size(), f);
        fclose(f);
    }
    double t2 = now_s();
    if (txt_path) {
        FILE* f = fopen(txt_path, "w");
        if (!f) { fprintf(stderr, "cannot open %s\n", txt_path); return 1; }
        char line_str[80];
        for (size_t p = 0; p < out.size(); p += 3) {
            sprintf(line_str, "(%f, %f, %f)\n", out[p], out[p + 1], out[p + 2]);   // RayTracer.cpp:1601
            fputs(line_str, f);
        }
        fclose(f);
    }
    double t3 = now_s();
    fprintf(stderr, "{\"scene\": \"%s\", \"objects\": %d, \"W\": %d, \"H\": %d, \"depth\": %d, \"x0\": %d, \"x1\": %d, "
           "\"stride\": %d, \"pixels\": %zu, \"render_s\": %.6f, \"txt_s\": %.6f"
#ifdef TCRT_REF_COUNT
           ", \"rays_primary\": %llu, \"rays_shadow\": %llu, \"rays_reflect\": %llu"
#endif
           "}\n",
           scene_name, my_scene.getObjectCount(), W, H, g_ref_max_depth, x0, x1, stride, out.size() / 3, t1 - t0,
           txt_path ? (t3 - t2) : 0.0
#ifdef TCRT_REF_COUNT
           , g_ref_rays[0], g_ref_rays[1], g_ref_rays[2]
#endif
           );
    return 0;
}
