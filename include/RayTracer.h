// include/RayTracer.h — C++ convenience over the C ABI (include/tcrt.h), shaped like the
// reference's driver: raytrace_main() (RayTracer.cpp:855-1114) = build scene, render every
// pixel, printPixelsToLog().  Header-only; link with libtcrt.so.
#ifndef TCRT_RAYTRACER_H_
#define TCRT_RAYTRACER_H_

#include <chrono>
#include <string>
#include <vector>

#include "CelioRayTracer.hpp"
#include "tcrt.h"

#define LOG_FILE_NAME "raytracer_screen.txt"   // RayTracer.h:134

namespace CelioRayTracer {

// The compile-time knobs of rt_project_parameters.h, at run time.
struct RenderSettings {
    int width = 500;          // SCREEN_HORIZONTAL_RESOLUTION
    int height = 504;         // SCREEN_VERTICAL_RESOLUTION
    int max_depth = 50;       // MAX_RECURSION_LEVEL
    bool shadows = true;      // SHADOWS_ON
    bool reflections = true;  // REFLECTIONS_ON
    tcrt_params params() const {
        tcrt_params p;
        tcrt_default_params(&p);
        p.width = width; p.height = height; p.max_depth = max_depth;
        p.shadows_on = shadows; p.reflections_on = reflections;
        return p;
    }
};

// Owns a tcrt_ctx on one or more B200s.
class GpuRayTracer {
public:
    explicit GpuRayTracer(const std::vector<int>& devices = std::vector<int>(1, 0)) : ctx_(NULL), run_time_s_(0) {
        rc_ = tcrt_create(&ctx_, devices.data(), (int)devices.size());
    }
    ~GpuRayTracer() { tcrt_destroy(ctx_); }
    bool ok() const { return rc_ == TCRT_OK; }
    const char* lastError() const { return tcrt_last_error(ctx_); }

    // my_scene / my_camera of the reference (RayTracer.h:50-51) become resident device data
    int setScene(const Scene& scene, const Camera& camera) {
        scene.flatten(flat_);
        tcrt_scene s = flat_.view();
        tcrt_camera c;
        camera.exportTo(&c);
        return rc_ = tcrt_upload_scene(ctx_, &s, &c);
    }
    // the pixel loop (RayTracer.cpp:911-923): pixels[x][z] -> pixels_xmajor[(x*H + z)*3 + c]
    int render(const RenderSettings& rs, float* pixels_xmajor, tcrt_stats* stats = NULL) {
        settings_ = rs;
        tcrt_params p = rs.params();
        std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
        rc_ = pixels_xmajor ? tcrt_render(ctx_, &p, pixels_xmajor, stats)
                            : tcrt_render_device(ctx_, &p, 0, p.width, stats);
        run_time_s_ = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        return rc_;
    }
    // init_log + printPixelsToLog (RayTracer.cpp:2022-2061,1574-1626)
    int printPixelsToLog(const char* path = LOG_FILE_NAME) {
        tcrt_params p = settings_.params();
        return rc_ = tcrt_write_txt(ctx_, &p, path, run_time_s_);
    }
    double runTimeSeconds() const { return run_time_s_; }
    tcrt_ctx* handle() { return ctx_; }

private:
    tcrt_ctx* ctx_;
    int rc_;
    double run_time_s_;
    RenderSettings settings_;
    SceneFlattener flat_;
};

// raytrace_main (RayTracer.cpp:855): 0 ok, 1 error, like the reference.
inline int raytrace_main(const Scene& scene, const Camera& camera, const RenderSettings& rs = RenderSettings(),
                         const char* log_path = LOG_FILE_NAME, const std::vector<int>& devices = std::vector<int>(1, 0)) {
    GpuRayTracer rt(devices);
    if (!rt.ok()) { fprintf(stderr, "%s\n", rt.lastError()); return 1; }
    if (rt.setScene(scene, camera) || rt.render(rs, NULL) || rt.printPixelsToLog(log_path)) {
        fprintf(stderr, "%s\n", rt.lastError());
        return 1;
    }
    return 0;
}

}  // namespace CelioRayTracer
#endif  // TCRT_RAYTRACER_H_
