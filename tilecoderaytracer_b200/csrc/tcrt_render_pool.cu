// tcrt_render_pool.cu — EXPERIMENTAL, compiled only into developer builds (-DTCRT_DEV_KNOBS, selected at run
// time with TCRT_POOL=1): an alternative render kernel for scenes whose spheres sit in a BVH (BASELINE configs
// 3 and 4, the reference's SCENE 2), in which a warp's BVH walks are TASKS of a warp-wide pool.  Bit-identical
// to tcrt_render.cu on every parity scene, and it raises the lanes busy per instruction from 17.9 to 20.8 of
// 32 on the mirror scene — but it measured 1-13 % SLOWER than tcrt_render.cu (profiles/README.md, round 2:
// what the pool saves in traversal steps it pays in task set-up, one extra loop iteration per path and a
// lower issue rate), so the product path does not use it.  Kept as the measured record of that experiment.
//
// Same contract as tcrt_render.cu (same reference functions replaced, same arithmetic, bit-identical
// results); what differs is how the traversal work of a warp is scheduled.  In tcrt_render.cu every lane
// walks the BVH for its own ray: nearest hit, then one shadow ray per light, one after the other.  With
// incoherent mirror rays that left 18 of 32 lanes busy in the nearest-hit walk, 12 in the shadow walks and
// 9 in their leaf tests (profiles/r2/lane_stats_r1_kernel.txt): lanes whose path has ended, lanes whose
// hit cannot be lit by this light, and lanes whose walk is shorter than the warp's longest all idle.
//
// Here the bounce loop is software-pipelined and the walks are pooled:
//   * after a hit is found, BOTH things that depend on it are known: the reflected ray of the next
//     recursion level (SceneObject.h:81-83) and the shadow rays to the lights (RayTracer.cpp:743-752).
//     They are independent of each other, so one pool holds, per lane, up to three tasks: "nearest hit of
//     my next ray", "is light A visible", "is light B visible";
//   * the 32 lanes execute the pool's tasks, not their own rays: task k goes to the k-th free lane, the
//     ray is handed over with warp shuffles, and a lane whose walk ends takes the next unclaimed task
//     (dynamic fetch inside the traversal loop).  Results return through a few words of shared memory
//     per warp.  Idle lanes (dead paths, unlit hits) therefore work for the others;
//   * shading (cosineShade, specular, the reflection record) runs afterwards, per lane, in the
//     reference's light order with the reference's arithmetic.
// Which lane runs a walk cannot change its answer: every primitive test is the reference's exact
// operation sequence, nearest-hit ties go to the lower object index, any-hit answers are booleans.
//
// Citations: calculatePixel RayTracer.cpp:448-638, getCollision :50-89, inShade :709-771, cosineShade
// :654-701, CollisionObject ctor SceneObject.h:47-105, SceneSphere::collision SceneSphere.cpp:50-168,
// SceneInfinitePlane.cpp:29-108, SceneFinitePlane.cpp:86-164.
#ifdef TCRT_DEV_KNOBS
#include "tcrt_render_common.cuh"

namespace {

#ifndef TCRT_POOL_MIN_BLOCKS
#define TCRT_POOL_MIN_BLOCKS 3
#endif
constexpr int kPoolMinBlocks = TCRT_POOL_MIN_BLOCKS;
constexpr int kWarpsPerCta = kBlock / 32;
#ifndef TCRT_POOL_FETCH_MIN
#define TCRT_POOL_FETCH_MIN 8
#endif
constexpr int kPoolFetchMin = TCRT_POOL_FETCH_MIN;

// Per-warp mailbox of the task pool (shared memory).  The lane that owns a ray writes its task descriptor;
// whichever lane ends up walking it writes the answer back.
struct WarpPool {
    float4 ray[96][2];          // task k: (origin xyz, limit) (direction xyz, bits: owner lane | kind << 5)
                                //   kind 0 nearest hit (limit = far distance), 1 / 2 shadow ray to the round's first /
                                //   second light (limit = distance to the light)
    float best[32];             // nearest-hit result of the lane's next ray: distance ...
    int key[32];                // ... and primitive key (-1: miss)
    unsigned char occl[2][32];  // any-hit results of the lane's two shadow rays
};
// Scenes with more than two lights run the light loop in rounds of two; what the next round's preparation
// needs (hit point, normals, incoming direction: 12 floats) waits in shared memory meanwhile, so that no
// scene pays registers for it during the walks.
constexpr int kStashFloats = 12;

// What a lane executes: one ray through the sphere BVH + the linearly swept rest of the scene.
struct Task {
    V3 O, D;
    float best;     // nearest: best distance so far; shadow: distance to the light
    int key;        // nearest: best primitive key
    int src;        // lane that owns the ray
    int kind;       // 0 nearest, 1/2 shadow ray, -1 none
};

// Runs the warp's pool of task descriptors wp.ray[first .. n_tasks) (warp-uniform bounds).  Task k goes to the k-th
// free lane; a lane whose walk has ended takes the next unclaimed task.  Answers land in wp.best / wp.key
// (nearest) and wp.occl (shadow), indexed by the owner lane.
template <int FM>
__device__ __forceinline__ void trace_pool(const Sm& sm, const DeviceScene& sc, WarpPool& wp, unsigned lt_mask, int first_task,
                                           int n_tasks) {
    const float4* __restrict__ gnodes = sc.bvh_sph;
    int next = first_task;        // first unclaimed task (warp-uniform)
    Task t;
    t.kind = -1;
    t.O = mk(0.f, 0.f, 0.f);
    t.D = mk(1.f, 0.f, 0.f);
    t.best = 0.f;
    t.key = -1;
    t.src = 0;
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
        // ---- fetch: free lanes take the next unclaimed tasks ---------------------------------------------
        // (setting a task up costs about as much as a few traversal steps, and costs the WARP that whatever
        // the number of lanes doing it: free lanes are refilled in batches, not one by one)
        const unsigned free_m = __ballot_sync(kFull, t.kind < 0);
        if (next < n_tasks && __popc(free_m) >= min(kPoolFetchMin, n_tasks - next)) {
            TCRT_STAT(6, free_m);
            const int k = next + __popc(free_m & lt_mask);
            next = min(n_tasks, next + __popc(free_m));
            if (t.kind < 0 && k < n_tasks) {
                const float4 r0 = wp.ray[k][0], r1 = wp.ray[k][1];
                const int code = __float_as_int(r1.w);
                t.src = code & 31;
                t.kind = code >> 5;
                t.O = xyz(r0);
                t.D = xyz(r1);
                t.best = r0.w;
                t.key = -1;
                found = false;
                task_linear<FM>(sm, sc, t.O, t.D, t.kind == 0, t.best, t.key, found);
                tv = make_trav(sc, t.O, t.D, true);
                lim_s = with_slack(t.best, tv.m);
                node = found ? kDone : sc.bvh_sph_root;
                leaf = 0;
                sp = 1;
            }
        }
        if (!__any_sync(kFull, t.kind >= 0)) break;
        // ---- inner nodes, until no lane is still looking for a leaf -----------------------------------------
        for (;;) {
            const bool inner = (unsigned)node < (unsigned)kDone;
            if (!__any_sync(kFull, inner && leaf == 0)) break;
            TCRT_STAT(2, __ballot_sync(kFull, inner));
            TCRT_STAT(8, __ballot_sync(kFull, !inner && leaf != 0));     // lanes waiting for the leaf phase
            TCRT_STAT(12, __ballot_sync(kFull, t.kind < 0));              // lanes without a task
            if (inner) {
                const float4 a = __ldg(gnodes + 4 * node), b = __ldg(gnodes + 4 * node + 1);
                const float4 c = __ldg(gnodes + 4 * node + 2), ch = __ldg(gnodes + 4 * node + 3);
                float tn0, tn1;
                const bool h0 = box_hit_fma(a.x, a.y, a.z, a.w, b.x, b.y, tv, lim_s, tn0);
                const bool h1 = box_hit_fma(b.z, b.w, c.x, c.y, c.z, c.w, tv, lim_s, tn1);
                const int c0 = __float_as_int(ch.x), c1 = __float_as_int(ch.y);
                if (h0 && h1) {
                    const bool swap = t.kind == 0 && (tn1 < tn0);     // nearer child first: tightens `best` early
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
            for (int i = first; i < last; ++i) task_sphere(sm, __ldg(sc.blob + i), i, t.O, t.D, t.kind == 0, t.best, t.key, found);
            lim_s = with_slack(t.best, tv.m);      // a nearest task's limit shrinks; a shadow task's is its light's distance
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
        // ---- finished walks report and free their lane ----------------------------------------------------
        if (t.kind >= 0 && node == kDone && leaf == 0) {
            if (t.kind == 0) {
                wp.best[t.src] = t.best;
                wp.key[t.src] = t.key;
            } else if (found) {
                wp.occl[t.kind - 1][t.src] = 1;
            }
            t.kind = -1;
        }
    }
    __syncwarp();
}

template <int CAP, int FM>
__global__ void __launch_bounds__(kBlock, kPoolMinBlocks) render_pool_kernel(const __grid_constant__ RenderLaunch rl) {
    extern __shared__ float4 smem4[];
    const DeviceScene& sc = rl.scene;
    // ---- stage the sweep blob (without the BVH-covered spheres, which leaves read through L1) ----------
    const int stage_off = sc.stage_off;
    const int n_stage = sc.blob_f4 - stage_off;
    for (int i = threadIdx.x; i < n_stage; i += kBlock) smem4[i] = __ldg(sc.blob + stage_off + i);
    __syncthreads();
    Sm sm;
    const float4* base = smem4 - stage_off;
    sm.sph = base;
    sm.fin = base + sc.fin_off;
    sm.inf = base + sc.inf_off;
    sm.light = base + sc.light_off;
    sm.clu = base + sc.clu_off;
    sm.cslot = reinterpret_cast<const int*>(base + sc.cslot_off);
    sm.idx = reinterpret_cast<const int*>(base + sc.idx_off);
    WarpPool& wp = reinterpret_cast<WarpPool*>(smem4 + n_stage)[threadIdx.x >> 5];
    // [kStashFloats][kBlock] floats behind the mailboxes, present only when the scene has more than two lights
    float* const stash = reinterpret_cast<float*>(reinterpret_cast<WarpPool*>(smem4 + n_stage) + kWarpsPerCta) + threadIdx.x;

    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const unsigned total = (unsigned)(rl.x1 - rl.x0) * (unsigned)rl.height;
    const V3 null_color = mk(rl.null_r, rl.null_g, rl.null_b);
    const int n_rounds = max(1, (sc.n_lights + 1) >> 1);

    float stack[CAP * 7];   // per level: local rgb, k, object rgb (local memory)
    // ---- per-lane path state ------------------------------------------------------------------------------
    int pix = -1;           // pixel id within the band, -1 = idle
    int level = 0;          // recursion level of the hit awaiting shading (`cur`), or of the ray in flight when there is none
    bool cur = false;       // a hit (not a light) waits for its light loop
    bool nxt = false;       // a ray (Q, nD) waits for its nearest hit: a fresh primary ray, or the reflected ray of `cur`
    int mode = 0;           // after shading `cur`: 0 finish with its local colour, 1 push record and continue with the
                            // reflected ray, 2 push record, recursion cap reached: the child is NULL_COLOR (:454-455)
    // hit point of `cur` == origin of its shadow rays and of its reflected ray; the light loop's view of the hit.
    // None of these is alive while the pool runs: rays travel through the mailbox, and what the shading after
    // the walks needs of a light is its cosine and specular dot products (c_*, d_*).
    V3 Q = mk(0.f, 0.f, 0.f), nD = mk(1.f, 0.f, 0.f);
    V3 n2 = mk(0.f, 0.f, 1.f), N = n2, Din = nD;
    V3 color = null_color;
    float diffuse = 0.f, specular = 0.f, kref = 0.f;
    unsigned n_primary = 0, n_shadow = 0, n_reflect = 0;
    unsigned wcur = 0, wend = 0;
    bool exhausted = false;

    for (;;) {
        // ---- refill: idle lanes take the next pixels of the warp's tile -------------------------------------
        unsigned idle = __ballot_sync(kFull, pix < 0);
        while (__popc(idle) >= rl.refill_min && !exhausted) {
            if (wcur == wend) {
                unsigned qb = 0;
                if (lane == 0) qb = atomicAdd(rl.queue, 32u);
                qb = __shfl_sync(kFull, qb, 0);
                if (qb >= total) {
                    exhausted = true;
                    break;
                }
                wcur = qb;
                wend = min(qb + 32u, total);
            }
            const unsigned avail = wend - wcur;
            const unsigned rank = __popc(idle & lt_mask);
            if (pix < 0 && rank < avail) {
                int xc, z;
                queue_to_pixel(rl, (int)(wcur + rank), xc, z);
                pix = xc * rl.height + z;
                level = 0;
                cur = false;
                nxt = true;
                primary_ray(rl, xc, z, Q, nD);
            }
            n_primary += min((unsigned)__popc(idle), avail);
            wcur += min((unsigned)__popc(idle), avail);
            idle = __ballot_sync(kFull, pix < 0);
        }
        if (idle == kFull) break;
        TCRT_STAT(0, ~idle);

        // ---- the lane's next ray becomes a task ---------------------------------------------------------------
        const unsigned wn = __ballot_sync(kFull, nxt);
        const int near_slot = __popc(wn & lt_mask);
        if (nxt) {
            wp.ray[near_slot][0] = make_float4(Q.x, Q.y, Q.z, rl.far_dist);
            wp.ray[near_slot][1] = make_float4(nD.x, nD.y, nD.z, __int_as_float((int)lane));
        }
        const int n_near = __popc(wn);     // slots [0, n_near) keep the next rays until their hit records are built
        int n_tasks = n_near;

        // ---- light loop of `cur` (RayTracer.cpp:537-591) around the pool: two lights per round ------------------
        V3 local = mk(0.f, 0.f, 0.f);
        float bound = 0.f;        // upper bound of any channel of `local` (decides whether the clamp of cosineShade can act)
        for (int r = 0; r < n_rounds; ++r) {
            const int la = 2 * r;
            // Does the answer of a shadow ray matter?  Only if the light can change `local`: a positive
            // cosine term, a positive specular dot product, or the clamp of cosineShade (which runs for every
            // unshadowed light once a channel exceeds 1 — impossible while `bound` <= 1).  Otherwise the ray
            // is counted (the reference casts it) but not traced.
            // (the two lights of a round share one copy of this code: the bounce loop has to fit the instruction cache)
            float c_a = 0.f, c_b = 0.f, d_a = 0.f, d_b = 0.f;
            bool want_a = false, want_b = false;
            wp.occl[0][lane] = 0;
            wp.occl[1][lane] = 0;
#pragma unroll 1
            for (int s = 0; s < 2; ++s) {
                const int l = la + s;
                if (l >= sc.n_lights) break;
                n_shadow += rl.shadows_on ? (unsigned)__popc(__ballot_sync(kFull, cur)) : 0u;
                float c = 0.f, d = 0.f, dist = 0.f;
                V3 lr = mk(0.f, 0.f, 0.f);
                bool w = false;
                if (cur) {
                    const float4 lp = sm.light[2 * l];
                    const float4 lc = sm.light[2 * l + 1];
                    // inShade (:743-752): dir = L - P, |dir|, Ray(P, dir) normalises with the same length
                    const V3 dir = xyz(lp) - Q;
                    dist = __fsqrt_rn(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
                    lr = div3(dir, dist);
                    c = dot(n2, lr);                                   // cosineShade (:654-701)
                    const float two_ln = 2.0f * dot(lr, N);            // specular (:561-588)
                    d = dot(Din, lr - scale(N, two_ln));
                    const bool basic = (d > 0.0f) || (diffuse > 0.0f && c > 0.0f);
                    w = basic || (diffuse > 0.0f && bound > 0.999f);
                    if (basic) {
                        const float lmax = fmaxf(fmaxf(lc.x, lc.y), lc.z);
                        const float cmax = fmaxf(fmaxf(color.x, color.y), color.z);
                        float u = 0.f;
                        if (diffuse > 0.0f && c > 0.0f) u = c * diffuse * lp.w * cmax;
                        if (d > 0.0f) {
                            const float dd = fminf(d, 1.0f) * fminf(d, 1.0f);      // d^20 <= d^4 for d <= 1
                            u += specular * (d > 1.0f ? 1.001f : dd * dd);
                        }
                        // anything not plainly finite and small counts as "may exceed 1"
                        bound = (u >= 0.0f && u < 4.0f && lmax >= 0.0f && lmax < 4.0f) ? bound + 1.001f * u * lmax : 8.0f;
                    }
                }
                // the shadow ray becomes a task
                const bool trace = w && rl.shadows_on;
                const unsigned wm = __ballot_sync(kFull, trace);
                if (trace) {
                    const int slot = n_tasks + __popc(wm & lt_mask);
                    wp.ray[slot][0] = make_float4(Q.x, Q.y, Q.z, dist);
                    wp.ray[slot][1] = make_float4(lr.x, lr.y, lr.z, __int_as_float((int)(lane | ((unsigned)(s + 1) << 5))));
                }
                n_tasks += __popc(wm);
                if (s == 0) { c_a = c; d_a = d; want_a = w; } else { c_b = c; d_b = d; want_b = w; }
            }
            const bool more = r + 1 < n_rounds;     // (only with more than two lights)
            if (more) {
                stash[0 * kBlock] = Q.x;   stash[1 * kBlock] = Q.y;   stash[2 * kBlock] = Q.z;
                stash[3 * kBlock] = n2.x;  stash[4 * kBlock] = n2.y;  stash[5 * kBlock] = n2.z;
                stash[6 * kBlock] = N.x;   stash[7 * kBlock] = N.y;   stash[8 * kBlock] = N.z;
                stash[9 * kBlock] = Din.x; stash[10 * kBlock] = Din.y; stash[11 * kBlock] = Din.z;
            }
            __syncwarp();
            TCRT_STAT(10, __ballot_sync(kFull, want_a && rl.shadows_on));
            trace_pool<FM>(sm, sc, wp, lt_mask, r == 0 ? 0 : n_near, n_tasks);
            n_tasks = n_near;
            if (more) {
                Q = mk(stash[0 * kBlock], stash[1 * kBlock], stash[2 * kBlock]);
                n2 = mk(stash[3 * kBlock], stash[4 * kBlock], stash[5 * kBlock]);
                N = mk(stash[6 * kBlock], stash[7 * kBlock], stash[8 * kBlock]);
                Din = mk(stash[9 * kBlock], stash[10 * kBlock], stash[11 * kBlock]);
            } else {
                Q = n2 = N = Din = mk(0.f, 0.f, 0.f);      // dead from here on: the next hit record defines them
            }
            const bool occl_a = wp.occl[0][lane] != 0, occl_b = wp.occl[1][lane] != 0;
#pragma unroll 1
            for (int s = 0; s < 2; ++s) {
                if (s ? (want_b && !occl_b) : (want_a && !occl_a)) {
                    const int l = la + s;
                    const float4 lp = sm.light[2 * l];
                    const float4 lc = sm.light[2 * l + 1];
                    // cosineShade (:654-701)
                    if (diffuse > 0.0f) {
                        const float c = s ? c_b : c_a;
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
                    // specular (:561-588): (V.R)^20 by 19 multiplies
                    const float d = s ? d_b : d_a;
                    if (d > 0.0f) {
                        float pw = d;
#pragma unroll
                        for (int i = 0; i < 19; ++i) pw *= d;
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
                float* rec = stack + 7 * level;
                rec[0] = local.x; rec[1] = local.y; rec[2] = local.z;
                rec[3] = kref;
                rec[4] = color.x; rec[5] = color.y; rec[6] = color.z;
                ++n_stacked;
                if (mode == 2) done = true;        // tail = NULL_COLOR
                else ++level;                       // the reflected ray's level
            }
            cur = false;
        }
        // ---- the next ray's hit record (CollisionObject ctor, SceneObject.h:47-105) ---------------------------------
        bool reflected = false;
        if (nxt) {
            nxt = false;
            const float best = wp.best[lane];
            const int bkey = wp.key[lane];
            n_stacked = level;
            if (bkey < 0) {
                tail = null_color;                        // :507-509
                done = true;
            } else {
                const int obj = sm.idx[bkey];
                const float4 surf = __ldg(sc.obj_surface + obj);
                const float4 mat = __ldg(sc.obj_material + obj);
                const int flags = __float_as_int(mat.w);   // bit31 light, low bits texture id + 1
                color = xyz(surf);
                if (flags < 0) {
                    tail = scale(color, mat.z);           // :520-527
                    done = true;
                } else {
                    const V3 O = xyz(wp.ray[near_slot][0]), D = xyz(wp.ray[near_slot][1]);   // the ray, back from the mailbox
                    diffuse = surf.w;
                    specular = mat.x;
                    kref = mat.y;
                    const int tex = (flags & 0x7fffffff) - 1;
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
                    mode = 0;
                    if (rl.reflections_on && kref > 0.0f) {   // :595-604
                        if (level + 1 > rl.max_depth) {
                            mode = 2;
                        } else {
                            mode = 1;
                            float ndi = dot(n1, D);     // SceneObject.h:63
                            // -2*normal.k * n_dot_incoming + incoming.k  (SceneObject.h:81-83), then Ray() normalises
                            nD = normalize(mk(-2.0f * n1.x * ndi + D.x, -2.0f * n1.y * ndi + D.y, -2.0f * n1.z * ndi + D.z));
                            nxt = true;
                            reflected = true;
                        }
                    }
                }
            }
        }
        __syncwarp();      // the mailbox is rewritten at the top of the loop
        n_reflect += (unsigned)__popc(__ballot_sync(kFull, reflected));
        if (done) {
            // final += (k * child) * obj, deepest level first (:601)
            TCRT_UNROLL_LOOP
            for (int i = n_stacked - 1; i >= 0; --i) {
                const float* rec = stack + 7 * i;
                V3 kc = scale(tail, rec[3]);
                tail.x = rec[0] + kc.x * rec[4];
                tail.y = rec[1] + kc.y * rec[5];
                tail.z = rec[2] + kc.z * rec[6];
            }
            float* o = rl.out + 3 * (size_t)pix;
            o[0] = tail.x;
            o[1] = tail.y;
            o[2] = tail.z;
            if (rl.col_cost != nullptr) {      // balancer pre-pass: bounces of this pixel, per column and per row
                const int xc = pix / rl.height;
                atomicAdd(rl.col_cost + xc, (unsigned)(level + 1));
                atomicAdd(rl.col_cost + (rl.x1 - rl.x0) + (pix - xc * rl.height), (unsigned)(level + 1));
            }
            pix = -1;
        }
    }

    if (lane == 0) {
        atomicAdd(rl.counters + 0, (unsigned long long)n_primary);
        atomicAdd(rl.counters + 1, (unsigned long long)n_shadow);
        atomicAdd(rl.counters + 2, (unsigned long long)n_reflect);
    }
}

template <int CAP, int FM>
cudaError_t launch_pool_one(const RenderLaunch& rl, int grid, size_t smem, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(render_pool_kernel<CAP, FM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    render_pool_kernel<CAP, FM><<<grid, kBlock, smem, stream>>>(rl);
    return cudaGetLastError();
}

template <int CAP>
cudaError_t launch_pool_cap(const RenderLaunch& rl, int fm, int grid, size_t smem, cudaStream_t stream) {
    return fm == 0 ? launch_pool_one<CAP, 0>(rl, grid, smem, stream) : launch_pool_one<CAP, 3>(rl, grid, smem, stream);
}

}  // namespace

#ifdef TCRT_LANE_STATS
extern "C" int tcrt_dev_lane_stats_pool(unsigned long long* out16, int reset) {
    cudaDeviceSynchronize();
    if (out16 && cudaMemcpyFromSymbol(out16, g_lane_stats, sizeof(unsigned long long) * 16) != cudaSuccess) return -1;
    if (reset) {
        unsigned long long z[16] = {};
        if (cudaMemcpyToSymbol(g_lane_stats, z, sizeof z) != cudaSuccess) return -1;
    }
    return 0;
}
#endif

static size_t pool_smem_extra(int n_lights) {
    return sizeof(WarpPool) * kWarpsPerCta + (n_lights > 2 ? sizeof(float) * kStashFloats * kBlock : 0);
}

// fm: 0 no finite planes, 3 a handful swept linearly.  The caller has checked that the scene has a sphere BVH.
cudaError_t tcrt_launch_render_pool(const RenderLaunch& rl, int fm, int sm_count, size_t smem_scene, cudaStream_t stream) {
    const size_t smem = smem_scene + pool_smem_extra(rl.scene.n_lights);
    int ctas_per_sm = kPoolMinBlocks;
    while (ctas_per_sm > 1 && (smem + 1024) * ctas_per_sm > 220 * 1024) --ctas_per_sm;
    const int grid = sm_count * ctas_per_sm;
    const int levels = rl.max_depth + 1;
    if (levels <= 8) return launch_pool_cap<8>(rl, fm, grid, smem, stream);
    if (levels <= 16) return launch_pool_cap<16>(rl, fm, grid, smem, stream);
    if (levels <= 64) return launch_pool_cap<64>(rl, fm, grid, smem, stream);
    return launch_pool_cap<256>(rl, fm, grid, smem, stream);
}
#endif  // TCRT_DEV_KNOBS
