// examples/raytrace_main.cpp — the reference's program (main -> raytrace_main, RayTracer.cpp:1122,855)
// on B200s: build the scene with the reference's construction API, render, write
// raytracer_screen.txt.  Host code is plain C++; the GPU work is behind libtcrt.so.
//
//   g++ -std=c++17 -O2 -ffp-contract=off -Iinclude examples/raytrace_main.cpp \
//       -Ltilecoderaytracer_b200 -ltcrt -Wl,-rpath,$PWD/tilecoderaytracer_b200 -o raytrace_main
//   ./raytrace_main [scene 1|2] [width height max_depth] [n_gpus]
#include <stdio.h>
#include <stdlib.h>

#include "Camera.h"
#include "RayTracer.h"
#include "Scene.h"

int main(int argc, char** argv) {
    using namespace CelioRayTracer;
    int which = argc > 1 ? atoi(argv[1]) : 1;          // SCENE (rt_project_parameters.h:27)
    RenderSettings rs;
    if (argc > 4) { rs.width = atoi(argv[2]); rs.height = atoi(argv[3]); rs.max_depth = atoi(argv[4]); }
    int n_gpus = argc > 5 ? atoi(argv[5]) : 1;

    Scene my_scene;
    Camera my_camera;
    if (which == 2 ? my_scene.initializeTwoMirrors(&my_camera) : my_scene.initialize()) return 1;

    std::vector<int> devices;
    for (int i = 0; i < n_gpus; i++) devices.push_back(i);
    printf("****** Start Ray Tracing. *******\n");
    int rc = raytrace_main(my_scene, my_camera, rs, LOG_FILE_NAME, devices);
    printf(rc ? "Failed.\n" : "Program Done.\n");
    return rc;
}
