// tcrt_render_wave.cu — EXPERIMENTAL, compiled only into developer builds (-DTCRT_DEV_KNOBS, selected at run time with
// TCRT_WAVE=1): the render path as a WAVEFRONT for large frames of sphere-BVH scenes (BASELINE configs[4]: the
// mirror-heavy 256-sphere scene at 7680x4320) — rays live in queues in device memory, not in the registers of the lane
// that owns the pixel.  Bit-identical to tcrt_render.cu (frames and ray counters, tools/wave_check.py), and measured
// TWICE AS SLOW: 46.6 ms against 23.9 ms on the 8K frame (trace 23.3 ms + shade 23.6 ms).  Per-lane task fetch raises
// the lanes busy per traversal step only from 14.7 to 22.1 of 32 (5.6 lanes hold two leaves and wait for the leaf
// phase, 2.5 have no task), while every ray now costs a scattered 32-byte descriptor read and an answer write, and the
// shade kernel moves ~280 B per path and wave through uncoalesced slots (profiles/README.md, round 2 §4).  Kept as the
// measured record of the experiment; the product path does not use it.
//
// Same contract as tcrt_render.cu (same reference functions replaced, same arithmetic, bit-identical results).
// Why a second shape: in the megakernel a warp walks the BVH for its own 32 rays and runs as long as its longest
// walk; on incoherent mirror rays 14.7 of 32 lanes are busy per traversal step (profiles/README.md, round 2 §4), and
// every warp-sized remedy (refill, task pool, stealing) lost what it gained because a warp's own progress stayed
// tied to the walks it happened to hold.  Here the recursion level is the unit of work:
//   wave w:   trace   a persistent kernel; every LANE takes the next ray task of the wave's queue as soon as its walk
//                     ends (nearest-hit and shadow rays share one instruction stream), answers go to per-path slots;
//             shade   one thread per live path: light loop of the previous hit with the shadow answers
//                     (cosineShade, specular — RayTracer.cpp:537-591), reflection record, then the hit record of the
//                     ray just traced (CollisionObject ctor, SceneObject.h:47-105), its shadow rays and its
//                     reflected ray become the tasks of wave w + 1.
// A path's state between waves is 3 float4 (+ its ray descriptors and its level records, deepest-first fold at the
// end: RayTracer.cpp:601).  Which lane walks a ray cannot change its answer (exact reference tests, ties on the object
// index, boolean any-hit), so the frame is the megakernel's, bit for bit.
//
// The frame is processed in chunks of up to kWaveChunk pixels (queue order: 4x8-pixel tiles, so a warp's primary rays
// are neighbours); per chunk: primary, then (trace, shade) x (max_depth + 2) launches, no host synchronisation —
// every kernel reads its work count from device memory and a wave without work returns at once.
#ifdef TCRT_DEV_KNOBS
#include <vector>

#include "tcrt_render_common.cuh"

namespace {

constexpr int kWaveBlock = 256;
constexpr int kClaim = 128;            // tasks a warp claims from the wave's queue with one global atomic

// per-path flags (state1.z)
enum : unsigned { kCur = 1u, kNxt = 2u, kModeShift = 2, kWantA = 16u, kWantB = 32u, kLevelShift = 8 };

struct WaveMem {
    int cap;                 // paths per chunk
    int n_waves;             // max_depth + 2
    unsigned* task_count;    // [n_waves + 1]
    unsigned* live_count;    // [n_waves + 1]
    unsigned* head;          // [n_waves + 1] claim counter of the trace kernel
    float4* state0;          // [cap] colour rgb, diffuse
    float4* state1;          // [cap] specular, kref, flags (bits), pixel id (bits)
    float4* cd;              // [cap] cosine / specular dot products of the two lights: c_a, d_a, c_b, d_b
    float4* ray;             // [3 kinds][2][cap]: (origin, limit) (direction, -); kind 0 next ray, 1 / 2 shadow rays
    float* near_best;        // [cap]
    int* near_key;           // [cap]
    unsigned char* occl;     // [2][cap]
    unsigned* tasks;         // [2 (wave parity)][3 * cap]: path << 2 | kind
    unsigned* live;          // [2 (wave parity)][cap]
    float* rec;              // [levels][7][cap]: local rgb, k, object rgb per recursion level
};

__device__ __forceinline__ float4* ray_slot(const WaveMem& wm, int kind, int part, unsigned i) {
    return wm.ray + ((size_t)(kind * 2 + part) * wm.cap + i);
}

__device__ __forceinline__ void stage_scene(const DeviceScene& sc, float4* smem4, Sm& sm) {
    const int stage_off = sc.stage_off;
    const int n_stage = sc.blob_f4 - stage_off;
    for (int i = threadIdx.x; i < n_stage; i += kWaveBlock) smem4[i] = __ldg(sc.blob + stage_off + i);
    __syncthreads();
    const float4* base = smem4 - stage_off;
    sm.sph = base;
    sm.fin = base + sc.fin_off;
    sm.inf = base + sc.inf_off;
    sm.light = base + sc.light_off;
    sm.clu = base + sc.clu_off;
    sm.cslot = reinterpret_cast<const int*>(base + sc.cslot_off);
    sm.idx = reinterpret_cast<const int*>(base + sc.idx_off);
}

// ---- wave 0: primary rays (Camera::createEyeRay, Camera.cpp:71-84) -----------------------------------------------------
__global__ void __launch_bounds__(kWaveBlock) wave_primary_kernel(const __grid_constant__ RenderLaunch rl, const WaveMem wm,
                                                                  unsigned q0, unsigned n) {
    const unsigned j = blockIdx.x * kWaveBlock + threadIdx.x;
    if (j == 0) {
        wm.task_count[0] = n;
        wm.live_count[0] = n;
    }
    if (j >= n) return;
    int xc, z;
    queue_to_pixel(rl, (int)(q0 + j), xc, z);
    V3 O, D;
    primary_ray(rl, xc, z, O, D);
    *ray_slot(wm, 0, 0, j) = make_float4(O.x, O.y, O.z, rl.far_dist);
    *ray_slot(wm, 0, 1, j) = make_float4(D.x, D.y, D.z, 0.f);
    wm.state1[j] = make_float4(0.f, 0.f, __uint_as_float(kNxt), __int_as_float(xc * rl.height + z));
    wm.live[j] = j;
    wm.tasks[j] = j << 2;
}

// ---- trace: every lane walks ray tasks until the wave's queue is empty ---------------------------------------------------
template <int FM>
__global__ void __launch_bounds__(kWaveBlock, 4) wave_trace_kernel(const __grid_constant__ RenderLaunch rl, const WaveMem wm, int w) {
    extern __shared__ float4 smem4[];
    const DeviceScene& sc = rl.scene;
    const unsigned n_tasks = wm.task_count[w];
    if (n_tasks == 0) return;
    Sm sm;
    stage_scene(sc, smem4, sm);
    const unsigned* __restrict__ tasks = wm.tasks + (size_t)(w & 1) * 3 * wm.cap;
    const float4* __restrict__ gnodes = sc.bvh_sph;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;

    unsigned cur = 0, end = 0;         // the warp's claimed block of the queue (warp-uniform)
    bool exhausted = false;
    V3 O = mk(0.f, 0.f, 0.f), D = mk(1.f, 0.f, 0.f);
    float best = 0.f;
    int key = -1, kind = -1;
    unsigned path = 0;
    int stack[kBvhStack];
    stack[0] = kDone;
    int sp = 1;
    int node = kDone, leaf = 0;
    bool found = false;
    Trav tv;
    tv.inv = tv.OpI = tv.OmI = mk(0.f, 0.f, 0.f);
    tv.m = 0.f;
    float lim_s = 0.f;
    for (;;) {
        // ---- fetch: free lanes take the next tasks (in batches: setting a task up costs the warp a few traversal steps
        // whatever the number of lanes doing it) ----------------------------------------------------------------------
        const unsigned free_m = __ballot_sync(kFull, kind < 0);
        if (!exhausted && (__popc(free_m) >= 8 || free_m == kFull)) {
            if (cur == end) {
                unsigned b = 0;
                if (lane == 0) b = atomicAdd(wm.head + w, (unsigned)kClaim);
                b = __shfl_sync(kFull, b, 0);
                if (b >= n_tasks) {
                    exhausted = true;
                } else {
                    cur = b;
                    end = min(b + (unsigned)kClaim, n_tasks);
                }
            }
            if (!exhausted) {
                const unsigned k = cur + __popc(free_m & lt_mask);
                cur = min(end, cur + (unsigned)__popc(free_m));
                TCRT_STAT(6, __ballot_sync(kFull, kind < 0 && k < end));
                if (kind < 0 && k < end) {
                    const unsigned t = __ldg(tasks + k);
                    path = t >> 2;
                    kind = (int)(t & 3u);
                    const float4 r0 = *ray_slot(wm, kind, 0, path), r1 = *ray_slot(wm, kind, 1, path);
                    O = xyz(r0);
                    D = xyz(r1);
                    best = r0.w;
                    key = -1;
                    found = false;
                    task_linear<FM>(sm, sc, O, D, kind == 0, best, key, found);
                    tv = make_trav(sc, O, D, true);
                    lim_s = with_slack(best, tv.m);
                    node = found ? kDone : sc.bvh_sph_root;
                    leaf = 0;
                    sp = 1;
                }
            }
        }
        if (!__any_sync(kFull, kind >= 0)) {
            if (exhausted) break;
            continue;
        }
        // ---- inner nodes, until no lane is still looking for a leaf -----------------------------------------
        for (;;) {
            const bool inner = (unsigned)node < (unsigned)kDone;
            if (!__any_sync(kFull, inner && leaf == 0)) break;
            TCRT_STAT(2, __ballot_sync(kFull, inner));
            TCRT_STAT(8, __ballot_sync(kFull, !inner && leaf != 0));
            TCRT_STAT(12, __ballot_sync(kFull, kind < 0));
            if (inner) {
                const float4 a = __ldg(gnodes + 4 * node), b = __ldg(gnodes + 4 * node + 1);
                const float4 c = __ldg(gnodes + 4 * node + 2), ch = __ldg(gnodes + 4 * node + 3);
                float tn0, tn1;
                const bool h0 = box_hit_fma(a.x, a.y, a.z, a.w, b.x, b.y, tv, lim_s, tn0);
                const bool h1 = box_hit_fma(b.z, b.w, c.x, c.y, c.z, c.w, tv, lim_s, tn1);
                const int c0 = __float_as_int(ch.x), c1 = __float_as_int(ch.y);
                if (h0 && h1) {
                    const bool swap = kind == 0 && (tn1 < tn0);     // nearer child first: tightens `best` early
                    stack[sp++] = swap ? c0 : c1;
                    node = swap ? c1 : c0;
                } else if (h0 || h1) {
                    node = h0 ? c0 : c1;
                } else {
                    node = stack[--sp];
                }
                if (node < 0 && leaf == 0) {   // postpone the leaf, keep walking
                    leaf = node;
                    node = stack[--sp];
                }
            }
        }
        // ---- leaves ---------------------------------------------------------------------------------------
        TCRT_STAT(4, __ballot_sync(kFull, leaf != 0));
        if (leaf != 0) {
            const int v = ~leaf;
            const int first = v & 0xffffff, last = first + (v >> 24);
            TCRT_UNROLL_LOOP
            for (int i = first; i < last; ++i) task_sphere(sm, __ldg(sc.blob + i), i, O, D, kind == 0, best, key, found);
            lim_s = with_slack(best, tv.m);      // a nearest task's limit shrinks; a shadow task's is its light's distance
            if (found) {
                node = kDone;
                sp = 1;
            }
            if (node < 0) {      // the walk stopped on a second leaf: it is next
                leaf = node;
                node = stack[--sp];
            } else {
                leaf = 0;
            }
        }
        // ---- finished walks answer into their path's slots and free their lane ------------------------------------
        if (kind >= 0 && node == kDone && leaf == 0) {
            if (kind == 0) {
                wm.near_best[path] = best;
                wm.near_key[path] = key;
            } else if (found) {
                wm.occl[(size_t)(kind - 1) * wm.cap + path] = 1;
            }
            kind = -1;
        }
    }
}

// ---- shade: one thread per live path ---------------------------------------------------------------------------------
template <int FM>
__global__ void __launch_bounds__(kWaveBlock, 3) wave_shade_kernel(const __grid_constant__ RenderLaunch rl, const WaveMem wm, int w) {
    extern __shared__ float4 smem4[];
    const DeviceScene& sc = rl.scene;
    const unsigned n_live = wm.live_count[w];
    if (n_live == 0) return;
    Sm sm;
    stage_scene(sc, smem4, sm);
    const int stage_off = sc.stage_off;
    const unsigned* __restrict__ live = wm.live + (size_t)(w & 1) * wm.cap;
    unsigned* __restrict__ live_next = wm.live + (size_t)((w + 1) & 1) * wm.cap;
    unsigned* __restrict__ tasks_next = wm.tasks + (size_t)((w + 1) & 1) * 3 * wm.cap;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const V3 null_color = mk(rl.null_r, rl.null_g, rl.null_b);
    const size_t cap = (size_t)wm.cap;
    unsigned n_shadow = 0, n_reflect = 0;

    const unsigned warps_total = gridDim.x * (kWaveBlock / 32);
    const unsigned warp_id = blockIdx.x * (kWaveBlock / 32) + (threadIdx.x >> 5);
    for (unsigned base = warp_id * 32u; base < n_live; base += warps_total * 32u) {
        const bool valid = base + lane < n_live;
        const unsigned i = valid ? __ldg(live + base + lane) : 0u;
        float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0, cd = s0;
        if (valid) {
            s0 = wm.state0[i];
            s1 = wm.state1[i];
            cd = wm.cd[i];
        }
        unsigned flags = __float_as_uint(s1.z);
        const int pix = __float_as_int(s1.w);
        int level = (int)(flags >> kLevelShift);
        bool cur = valid && (flags & kCur), nxt = valid && (flags & kNxt);
        const int mode = (int)((flags >> kModeShift) & 3u);
        V3 color = xyz(s0);
        float diffuse = s0.w, specular = s1.x, kref = s1.y;

        // ---- light loop of `cur` (RayTracer.cpp:537-591) with the answers of its shadow rays ----------------------
        V3 local = mk(0.f, 0.f, 0.f);
        if (cur) {
            const bool occl_a = rl.shadows_on && wm.occl[i] != 0, occl_b = rl.shadows_on && wm.occl[cap + i] != 0;
#pragma unroll 1
            for (int s = 0; s < 2; ++s) {
                const bool want = flags & (s ? kWantB : kWantA);
                if (want && !(s ? occl_b : occl_a)) {
                    const float4 lp = sm.light[2 * s];
                    const float4 lc = sm.light[2 * s + 1];
                    if (diffuse > 0.0f) {        // cosineShade (:654-701)
                        const float c = s ? cd.z : cd.x;
                        if (c > 0.0f) {
                            float f = c * diffuse * lp.w;
                            local.x += f * color.x * lc.x;
                            local.y += f * color.y * lc.y;
                            local.z += f * color.z * lc.z;
                        }
                        local.x = (local.x > 1.0f) ? 1.0f : local.x;
                        local.y = (local.y > 1.0f) ? 1.0f : local.y;
                        local.z = (local.z > 1.0f) ? 1.0f : local.z;
                    }
                    const float d = s ? cd.w : cd.y;   // specular (:561-588): (V.R)^20 by 19 multiplies
                    if (d > 0.0f) {
                        float pw = d;
#pragma unroll
                        for (int k = 0; k < 19; ++k) pw *= d;
                        float sp = pw * specular;
                        local.x += lc.x * sp;
                        local.y += lc.y * sp;
                        local.z += lc.z * sp;
                    }
                }
            }
        }
        // ---- `cur` is shaded: push its record and go on, or finish the path ---------------------------------------
        bool done = false;
        V3 tail = null_color;
        int n_stacked = level;
        if (cur) {
            if (mode == 0) {
                tail = local;
                done = true;
            } else {
                float* rec = wm.rec + (size_t)level * 7 * cap + i;
                rec[0 * cap] = local.x; rec[1 * cap] = local.y; rec[2 * cap] = local.z;
                rec[3 * cap] = kref;
                rec[4 * cap] = color.x; rec[5 * cap] = color.y; rec[6 * cap] = color.z;
                ++n_stacked;
                if (mode == 2) done = true;        // recursion cap: the child is NULL_COLOR (:454-455)
                else ++level;                       // the reflected ray's level
            }
            cur = false;
        }
        // ---- hit record of the ray just traced (CollisionObject ctor, SceneObject.h:47-105) --------------------------
        unsigned new_flags = 0u;
        bool reflected = false;
        V3 Q = mk(0.f, 0.f, 0.f), n2 = mk(0.f, 0.f, 1.f), N = n2, Din = mk(1.f, 0.f, 0.f), nD = Din;
        if (nxt) {
            const float best = wm.near_best[i];
            const int bkey = wm.near_key[i];
            n_stacked = level;
            if (bkey < 0) {
                tail = null_color;                        // :507-509
                done = true;
            } else {
                const int obj = sm.idx[bkey];
                const float4 surf = __ldg(sc.obj_surface + obj);
                const float4 mat = __ldg(sc.obj_material + obj);
                const int mflags = __float_as_int(mat.w);   // bit31 light, low bits texture id + 1
                color = xyz(surf);
                if (mflags < 0) {
                    tail = scale(color, mat.z);           // :520-527
                    done = true;
                } else {
                    const V3 O = xyz(*ray_slot(wm, 0, 0, i)), D = xyz(*ray_slot(wm, 0, 1, i));
                    diffuse = surf.w;
                    specular = mat.x;
                    kref = mat.y;
                    const int tex = (mflags & 0x7fffffff) - 1;
                    V3 n1;
                    const V3 Pp = scale(D, best) + O;   // t*D + O  (SceneSphere.cpp:122, SceneFinitePlane.cpp:108)
                    if (bkey < sc.n_sph) {
                        Q = Pp;
                        const float4 sg = (bkey < stage_off) ? __ldg(sc.blob + bkey) : sm.sph[bkey];
                        // n1 = normalize(P - C) (SceneSphere.cpp:129-130); n2 = normalize(n1): Ray(point, normal)
                        // re-normalises (Ray.h:21-25); N = normalize(n2): the specular term's N (:565-566)
                        V3 v = Pp - xyz(sg);
#pragma unroll 1
                        for (int k = 0; k < 3; ++k) {
                            v = normalize(v);
                            if (k == 0) n1 = v;
                            if (k == 1) n2 = v;
                        }
                        N = v;
                    } else {
                        float den;
                        float px = 0.f, py = 0.f;
                        if (bkey < sc.n_sph + sc.n_fin) {
                            const float4* g = sm.fin + 4 * (bkey - sc.n_sph);
                            den = dot(D, xyz(g[0]));
                            if (tex >= 0) {   // x, y of SceneFinitePlane.cpp:116-120, recomputed for the winner
                                V3 PO = Pp - xyz(g[3]);
                                px = dot(PO, xyz(g[1]));
                                py = dot(PO, xyz(g[2]));
                            }
                        } else {
                            const int slot = bkey - sc.n_sph - sc.n_fin;
                            den = dot(D, xyz(sm.inf[slot]));
                            if (tex >= 0) {   // SceneInfinitePlane.cpp:59-74
                                V3 PO = Pp - xyz(__ldg(sc.inf_frame + 3 * slot + 2));
                                px = dot(PO, xyz(__ldg(sc.inf_frame + 3 * slot + 0)));
                                py = dot(PO, xyz(__ldg(sc.inf_frame + 3 * slot + 1)));
                            }
                        }
                        if (tex >= 0) color = checker(__ldg(sc.textures + 2 * tex), __ldg(sc.textures + 2 * tex + 1), px, py);
                        // computeNormal + the two re-normalisations of the plane side, tabulated at upload
                        const float4* nt = sc.obj_normals + 6 * obj + (den < 0.0f ? 0 : 3);
                        n1 = xyz(__ldg(nt));
                        n2 = xyz(__ldg(nt + 1));
                        N = xyz(__ldg(nt + 2));
                        Q = Pp + scale(n1, 0.0010000000475f);   // + temp_normal * INTERSECTION_OFFSET_DIST
                    }
                    Din = D;
                    cur = true;
                    new_flags = kCur;
                    if (rl.reflections_on && kref > 0.0f) {   // :595-604
                        if (level + 1 > rl.max_depth) {
                            new_flags |= 2u << kModeShift;
                        } else {
                            new_flags |= (1u << kModeShift) | kNxt;
                            float ndi = dot(n1, D);     // SceneObject.h:63
                            // -2*normal.k * n_dot_incoming + incoming.k  (SceneObject.h:81-83), then Ray() normalises
                            nD = normalize(mk(-2.0f * n1.x * ndi + D.x, -2.0f * n1.y * ndi + D.y, -2.0f * n1.z * ndi + D.z));
                            reflected = true;
                        }
                    }
                }
            }
        }
        // ---- the new hit's shadow rays: which answers matter (see tcrt_render.cu), descriptors for the next wave -----
        float4 cd_new = make_float4(0.f, 0.f, 0.f, 0.f);
        if (__any_sync(kFull, cur)) {
            float bound = 0.f;    // upper bound of any channel of `local`: can the clamp of cosineShade act yet?
#pragma unroll 1
            for (int s = 0; s < 2; ++s) {
                if (s >= sc.n_lights) break;
                n_shadow += rl.shadows_on ? (unsigned)__popc(__ballot_sync(kFull, cur)) : 0u;
                if (cur) {
                    const float4 lp = sm.light[2 * s];
                    const float4 lc = sm.light[2 * s + 1];
                    // inShade (:743-752): dir = L - P, |dir|, Ray(P, dir) normalises with the same length
                    const V3 dir = xyz(lp) - Q;
                    const float dist = __fsqrt_rn(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
                    const V3 lr = div3(dir, dist);
                    const float c = dot(n2, lr);                          // cosineShade (:654-701)
                    const float two_ln = 2.0f * dot(lr, N);               // specular (:561-588)
                    const float d = dot(Din, lr - scale(N, two_ln));
                    const bool basic = (d > 0.0f) || (diffuse > 0.0f && c > 0.0f);
                    const bool want = basic || (diffuse > 0.0f && bound > 0.999f);
                    if (basic) {
                        const float lmax = fmaxf(fmaxf(lc.x, lc.y), lc.z);
                        const float cmax = fmaxf(fmaxf(color.x, color.y), color.z);
                        float u = 0.f;
                        if (diffuse > 0.0f && c > 0.0f) u = c * diffuse * lp.w * cmax;
                        if (d > 0.0f) {
                            const float dd = fminf(d, 1.0f) * fminf(d, 1.0f);      // d^20 <= d^4 for d <= 1
                            u += specular * (d > 1.0f ? 1.001f : dd * dd);
                        }
                        bound = (u >= 0.0f && u < 4.0f && lmax >= 0.0f && lmax < 4.0f) ? bound + 1.001f * u * lmax : 8.0f;
                    }
                    if (s == 0) { cd_new.x = c; cd_new.y = d; } else { cd_new.z = c; cd_new.w = d; }
                    if (want) {
                        new_flags |= s ? kWantB : kWantA;
                        if (rl.shadows_on) {
                            *ray_slot(wm, s + 1, 0, i) = make_float4(Q.x, Q.y, Q.z, dist);
                            *ray_slot(wm, s + 1, 1, i) = make_float4(lr.x, lr.y, lr.z, 0.f);
                            wm.occl[(size_t)s * cap + i] = 0;
                        }
                    }
                }
            }
        }
        n_reflect += (unsigned)__popc(__ballot_sync(kFull, reflected));
        // ---- finish the path, or hand it to the next wave -----------------------------------------------------------------
        if (valid && done) {
            // final += (k * child) * obj, deepest level first (:601)
            TCRT_UNROLL_LOOP
            for (int l = n_stacked - 1; l >= 0; --l) {
                const float* rec = wm.rec + (size_t)l * 7 * cap + i;
                V3 kc = scale(tail, rec[3 * cap]);
                tail.x = rec[0 * cap] + kc.x * rec[4 * cap];
                tail.y = rec[1 * cap] + kc.y * rec[5 * cap];
                tail.z = rec[2 * cap] + kc.z * rec[6 * cap];
            }
            float* o = rl.out + 3 * (size_t)pix;
            o[0] = tail.x;
            o[1] = tail.y;
            o[2] = tail.z;
        }
        const bool goes_on = valid && !done && cur;
        if (goes_on) {
            if (new_flags & kNxt) {
                *ray_slot(wm, 0, 0, i) = make_float4(Q.x, Q.y, Q.z, rl.far_dist);
                *ray_slot(wm, 0, 1, i) = make_float4(nD.x, nD.y, nD.z, 0.f);
            }
            wm.state0[i] = make_float4(color.x, color.y, color.z, diffuse);
            wm.state1[i] = make_float4(specular, kref, __uint_as_float(new_flags | ((unsigned)level << kLevelShift)), __int_as_float(pix));
            wm.cd[i] = cd_new;
        }
        // warp-aggregated appends: next wave's live paths and ray tasks (next ray first, then the shadow rays)
        const bool t_n = goes_on && (new_flags & kNxt);
        const bool t_a = goes_on && (new_flags & kWantA) && rl.shadows_on, t_b = goes_on && (new_flags & kWantB) && rl.shadows_on;
        const unsigned m_l = __ballot_sync(kFull, goes_on), m_n = __ballot_sync(kFull, t_n);
        const unsigned m_a = __ballot_sync(kFull, t_a), m_b = __ballot_sync(kFull, t_b);
        if (m_l != 0u) {
            unsigned b_l = 0, b_t = 0;
            if (lane == 0) {
                b_l = atomicAdd(wm.live_count + w + 1, (unsigned)__popc(m_l));
                b_t = atomicAdd(wm.task_count + w + 1, (unsigned)(__popc(m_n) + __popc(m_a) + __popc(m_b)));
            }
            b_l = __shfl_sync(kFull, b_l, 0);
            b_t = __shfl_sync(kFull, b_t, 0);
            if (goes_on) live_next[b_l + __popc(m_l & lt_mask)] = i;
            if (t_n) tasks_next[b_t + __popc(m_n & lt_mask)] = i << 2;
            if (t_a) tasks_next[b_t + __popc(m_n) + __popc(m_a & lt_mask)] = (i << 2) | 1u;
            if (t_b) tasks_next[b_t + __popc(m_n) + __popc(m_a) + __popc(m_b & lt_mask)] = (i << 2) | 2u;
        }
    }
    if (lane == 0) {
        if (n_shadow) atomicAdd(rl.counters + 1, (unsigned long long)n_shadow);
        if (n_reflect) atomicAdd(rl.counters + 2, (unsigned long long)n_reflect);
    }
}

__global__ void wave_count_primary_kernel(unsigned long long* counters, unsigned long long n) { atomicAdd(counters, n); }

}  // namespace

#ifdef TCRT_LANE_STATS
extern "C" int tcrt_dev_lane_stats_wave(unsigned long long* out16, int reset) {
    cudaDeviceSynchronize();
    if (out16 && cudaMemcpyFromSymbol(out16, g_lane_stats, sizeof(unsigned long long) * 16) != cudaSuccess) return -1;
    if (reset) {
        unsigned long long z[16] = {};
        if (cudaMemcpyToSymbol(g_lane_stats, z, sizeof z) != cudaSuccess) return -1;
    }
    return 0;
}
#endif

// paths per chunk: enough to fill the machine many times over, small enough that a chunk's queues stay modest
static constexpr size_t kWaveChunk = (size_t)4 << 20;

size_t tcrt_wave_mem_bytes(size_t band_pixels, int max_depth) {
    const size_t cap = std::min((band_pixels + 31) / 32 * 32, kWaveChunk);
    const size_t levels = (size_t)max_depth + 1;
    const size_t n_waves = (size_t)max_depth + 3;
    size_t b = 3 * n_waves * sizeof(unsigned) + 256;
    b += cap * (3 * sizeof(float4) + 6 * sizeof(float4) + sizeof(float) + sizeof(int) + 2 + 2 * 3 * sizeof(unsigned) + 2 * sizeof(unsigned));
    b += cap * levels * 7 * sizeof(float);
    return b + 4096;
}

bool tcrt_wave_applies(const RenderLaunch& rl, int fm) {
    const size_t px = (size_t)(rl.x1 - rl.x0) * rl.height;
    return rl.scene.bvh_sph != nullptr && (fm == 0 || fm == 3) && rl.scene.n_lights <= 2 && rl.max_depth <= 24 && px >= ((size_t)1 << 20) &&
           rl.col_cost == nullptr;
}

cudaError_t tcrt_launch_render_wave(const RenderLaunch& rl_in, int fm, int sm_count, size_t smem_scene, void* mem, size_t mem_bytes,
                                    cudaStream_t stream, int* launches) {
    RenderLaunch rl = rl_in;
    rl.tiled = ((rl.x1 - rl.x0) % TCRT_TILE_W == 0 && rl.height % TCRT_TILE_H == 0) ? 1 : 0;
    rl.row_order = nullptr;
    const size_t total = (size_t)(rl.x1 - rl.x0) * rl.height;
    if (mem == nullptr || mem_bytes < tcrt_wave_mem_bytes(total, rl.max_depth)) return cudaErrorInvalidValue;
    WaveMem wm;
    wm.cap = (int)std::min((total + 31) / 32 * 32, kWaveChunk);
    wm.n_waves = rl.max_depth + 2;
    const size_t cap = (size_t)wm.cap, levels = (size_t)rl.max_depth + 1, nw = (size_t)wm.n_waves + 1;
    unsigned char* p = static_cast<unsigned char*>(mem);
    auto take = [&](size_t bytes) {
        unsigned char* r = p;
        p += (bytes + 255) / 256 * 256;
        return r;
    };
    wm.task_count = reinterpret_cast<unsigned*>(take(3 * nw * sizeof(unsigned)));
    wm.live_count = wm.task_count + nw;
    wm.head = wm.live_count + nw;
    wm.state0 = reinterpret_cast<float4*>(take(cap * sizeof(float4)));
    wm.state1 = reinterpret_cast<float4*>(take(cap * sizeof(float4)));
    wm.cd = reinterpret_cast<float4*>(take(cap * sizeof(float4)));
    wm.ray = reinterpret_cast<float4*>(take(6 * cap * sizeof(float4)));
    wm.near_best = reinterpret_cast<float*>(take(cap * sizeof(float)));
    wm.near_key = reinterpret_cast<int*>(take(cap * sizeof(int)));
    wm.occl = take(2 * cap);
    wm.tasks = reinterpret_cast<unsigned*>(take(2 * 3 * cap * sizeof(unsigned)));
    wm.live = reinterpret_cast<unsigned*>(take(2 * cap * sizeof(unsigned)));
    wm.rec = reinterpret_cast<float*>(take(cap * levels * 7 * sizeof(float)));
    if ((size_t)(p - static_cast<unsigned char*>(mem)) > mem_bytes) return cudaErrorInvalidValue;

    const size_t smem = smem_scene;
    cudaError_t e;
    auto attr = [&](const void* f) { return cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); };
    const void* k_trace = fm == 0 ? (const void*)wave_trace_kernel<0> : (const void*)wave_trace_kernel<3>;
    const void* k_shade = fm == 0 ? (const void*)wave_shade_kernel<0> : (const void*)wave_shade_kernel<3>;
    if ((e = attr(k_trace)) != cudaSuccess || (e = attr(k_shade)) != cudaSuccess) return e;
    const int grid_trace = sm_count * 4, grid_shade = sm_count * 6;
#ifdef TCRT_DEV_KNOBS   // per-kernel-kind timing of one call (developer builds, TCRT_WAVE_TIMING=1)
    const bool timing = getenv("TCRT_WAVE_TIMING") != nullptr;
    std::vector<cudaEvent_t> evs;
    auto mark = [&]() {
        if (!timing) return;
        cudaEvent_t ev;
        cudaEventCreate(&ev);
        cudaEventRecord(ev, stream);
        evs.push_back(ev);
    };
#else
    auto mark = []() {};
#endif
    for (size_t q0 = 0; q0 < total; q0 += cap) {
        const unsigned n = (unsigned)std::min(cap, total - q0);
        if ((e = cudaMemsetAsync(wm.task_count, 0, 3 * nw * sizeof(unsigned), stream)) != cudaSuccess) return e;
        mark();
        wave_primary_kernel<<<(n + kWaveBlock - 1) / kWaveBlock, kWaveBlock, 0, stream>>>(rl, wm, (unsigned)q0, n);
        mark();
        for (int w = 0; w < wm.n_waves; ++w) {
            if (fm == 0) {
                wave_trace_kernel<0><<<grid_trace, kWaveBlock, smem, stream>>>(rl, wm, w);
                mark();
                wave_shade_kernel<0><<<grid_shade, kWaveBlock, smem, stream>>>(rl, wm, w);
                mark();
            } else {
                wave_trace_kernel<3><<<grid_trace, kWaveBlock, smem, stream>>>(rl, wm, w);
                mark();
                wave_shade_kernel<3><<<grid_shade, kWaveBlock, smem, stream>>>(rl, wm, w);
                mark();
            }
        }
        if (launches) *launches += 1 + 2 * wm.n_waves;
    }
    wave_count_primary_kernel<<<1, 1, 0, stream>>>(rl.counters, (unsigned long long)total);
    if (launches) *launches += 1;
#ifdef TCRT_DEV_KNOBS
    if (timing && !evs.empty()) {
        cudaStreamSynchronize(stream);
        const int per_chunk = 2 + 2 * wm.n_waves;      // marks per chunk
        double t_prim = 0, t_trace = 0, t_shade = 0;
        std::vector<double> tw(wm.n_waves, 0.0), sw(wm.n_waves, 0.0);
        for (size_t c = 0; c * per_chunk < evs.size(); ++c) {
            const size_t b = c * per_chunk;
            float ms;
            cudaEventElapsedTime(&ms, evs[b], evs[b + 1]);
            t_prim += ms;
            for (int w = 0; w < wm.n_waves; ++w) {
                cudaEventElapsedTime(&ms, evs[b + 1 + 2 * w], evs[b + 2 + 2 * w]);
                t_trace += ms;
                tw[w] += ms;
                cudaEventElapsedTime(&ms, evs[b + 2 + 2 * w], evs[b + 3 + 2 * w]);
                t_shade += ms;
                sw[w] += ms;
            }
        }
        fprintf(stderr, "wave timing: primary %.3f ms, trace %.3f ms, shade %.3f ms\n", t_prim, t_trace, t_shade);
        for (int w = 0; w < wm.n_waves; ++w) fprintf(stderr, "  wave %2d: trace %.3f shade %.3f\n", w, tw[w], sw[w]);
        for (auto ev : evs) cudaEventDestroy(ev);
    }
#endif
    return cudaGetLastError();
}
#endif  // TCRT_DEV_KNOBS
