/* oracle/tcrt_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Scalar CPU restatement ("port") of the reference's per-pixel render path over the same
 * flattened scene (include/tcrt.h) the CUDA kernels consume.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this; libtcrt.so never does.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file bit-for-bit against the
 * reference itself (oracle/_ref/ref_render, built from /root/reference/src by
 * oracle/build_ref.sh) on every scene, and against the committed goldens in tests/golden/
 * (made by tests/golden/make_golden.py from the same reference binaries).
 *
 * Arithmetic: IEEE binary32, round-to-nearest-even, NO contraction (build with
 * -ffp-contract=off), sqrtf and '/' correctly rounded — vector3d.h:40-44,57-74 with
 * USING_FIXED_POINT false (rt_project_parameters.h:35-42).  Every expression below keeps
 * the reference's operand order and association; comments give file:line.
 *
 * Structure differs from the reference on purpose (it is the shape the GPU kernel has):
 *   - per-type sweeps (spheres, finite planes, infinite planes) instead of one loop of
 *     virtual collision() calls; ties on distance go to the lower object index, which is
 *     what "first strictly smaller in index order" (RayTracer.cpp:73-86) means;
 *   - distance-only tests; the hit record (CollisionObject ctor, SceneObject.h:47-105) is
 *     built for the winner only;
 *   - calculatePixel's recursion (RayTracer.cpp:448-638) is a loop that stores
 *     (local colour, k, object colour) per level and folds from the deepest level up, which
 *     reproduces  final += k * child * obj  (RayTracer.cpp:601) bit for bit.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "../include/tcrt.h"

typedef struct { float x, y, z; } v3;

static inline v3 v3_make(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 v3_ld(const float* p) { v3 r = {p[0], p[1], p[2]}; return r; }
static inline v3 v3_sub(v3 a, v3 b) { return v3_make(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v3_add(v3 a, v3 b) { return v3_make(a.x + b.x, a.y + b.y, a.z + b.z); }
/* vector3d.h:93-99 */
static inline float v3_dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
/* operator*(v, f): v.k * f  (vector3d.h:119-122) */
static inline v3 v3_scale(v3 a, float f) { return v3_make(a.x * f, a.y * f, a.z * f); }
/* vector3d.h:57-74: one sqrt, three divides */
static inline v3 v3_normalize(v3 a) {
    float len = sqrtf(a.x * a.x + a.y * a.y + a.z * a.z);
    return v3_make(a.x / len, a.y / len, a.z / len);
}

/* Work counters.  The primitive tests are counted as the REFERENCE executes them: every object
 * for a nearest-hit query (RayTracer.cpp:73-86), objects in index order up to the first blocker
 * for a shadow query (:724-736) — see in_shade().  `flops` applies SURVEY.md §8(d)'s per-unit
 * constants (add/sub/mul/div/sqrt/fmod = 1, negate/compare/select/convert = 0). */
typedef struct {
    unsigned long long rays_primary, rays_shadow, rays_reflect;
    unsigned long long tests_sphere, tests_fin, tests_inf; /* executed primitive tests */
    unsigned long long sph_exit_v, sph_exit_d2, sph_hit;   /* sphere test outcome classes */
    unsigned long long fin_exit_t, fin_bounds;              /* finite plane: left at t / bounds evaluated */
    unsigned long long hits_sphere, hits_plane, hits_reflective, hits_textured, hits_light, misses;
    unsigned long long lit_pairs, lit_pairs_spec;           /* unshadowed (hit, light) pairs; with V.R > 0 */
    unsigned long long combines;                            /* reflection combine steps */
    unsigned long long flops;                               /* algorithmic flops, SURVEY §8(d) */
} oracle_counters;

static oracle_counters* g_cnt; /* current counters (single-threaded test code) */

/* ---- primitive tests: distance only -------------------------------------------------- */

/* SceneSphere::collision, SceneSphere.cpp:50-130.  Returns 1 and *dist = v - sqrt(d2) (may
 * be negative: the reference always reports root1, :139). */
static inline int sphere_dist(const float* g, v3 O, v3 D, float* dist) {
    v3 OE = v3_sub(v3_ld(g), O);                        /* :54 */
    float v = v3_dot(OE, D);                            /* :56 */
    if (v < 0.0f) { g_cnt->sph_exit_v++; return 0; }    /* :58 */
    float d2 = g[3] - (v3_dot(OE, OE) - v * v);         /* :66 */
    if (d2 < (float)1E-9) { g_cnt->sph_exit_d2++; return 0; } /* :68 */
    *dist = v - sqrtf(d2);                              /* :85,:139 */
    g_cnt->sph_hit++;
    return 1;
}

/* Shared start of both plane tests: t = (-dto - O.n) / (D.n)
 * SceneInfinitePlane.cpp:39-46, SceneFinitePlane.cpp:92-99. */
static inline int plane_t(const float* g, v3 O, v3 D, float* t, float* den_out) {
    v3 n = v3_ld(g);
    float num = g[3] - v3_dot(O, n);   /* g[3] holds -distance_to_origin */
    float den = v3_dot(D, n);
    if (den == 0.0f) return 0;
    *t = num / den;
    *den_out = den;
    return 1;
}

/* SceneFinitePlane::collision, SceneFinitePlane.cpp:86-127.  x,y are the in-plane
 * coordinates (needed by the texture lookup of the winner). */
static inline int fin_dist(const float* g, v3 O, v3 D, float* dist, float* px, float* py, float* den) {
    float t;
    if (!plane_t(g, O, D, &t, den)) { g_cnt->fin_exit_t++; return 0; }
    if ((double)t < 1E-5) { g_cnt->fin_exit_t++; return 0; } /* :102 — a double compare */
    g_cnt->fin_bounds++;
    v3 P = v3_add(v3_scale(D, t), O);                   /* :108 */
    v3 PO = v3_sub(P, v3_ld(g + 12));                   /* :116 */
    float x = v3_dot(PO, v3_ld(g + 4));                 /* :119 */
    float y = v3_dot(PO, v3_ld(g + 8));                 /* :120 */
    if (x < 0 || x > g[7] || y < 0 || y > g[11]) return 0; /* :122 */
    *dist = t; *px = x; *py = y;
    return 1;
}

/* SceneInfinitePlane::collision, SceneInfinitePlane.cpp:29-54. */
static inline int inf_dist(const float* g, v3 O, v3 D, float* dist, float* den) {
    float t;
    if (!plane_t(g, O, D, &t, den)) return 0;
    if (t < (float)1E-10) return 0;                     /* :50 */
    *dist = t;
    return 1;
}

/* Texture_CheckerBoard::getTexturePixel, Texture_CheckerBoard.h:31-65 (fmod on floats is fmodf). */
static inline v3 checker(const float* tex, float x, float y) {
    float w = tex[3], h = tex[7];
    if (x >= 0) x = fmodf(x, w);
    else x = fmodf(fmodf(-x, w) + w / 2.0f, w);
    if (y >= 0) y = fmodf(y, h);
    else y = fmodf(fmodf(-y, h) + h / 2.0f, h);
    int light = (x < w / 2) ? (y < h / 2) : !(y < h / 2);
    return light ? v3_ld(tex) : v3_ld(tex + 4);
}

/* ---- the winner's hit record ----------------------------------------------------------- */
typedef struct {
    v3 P;          /* intersection_point handed to shading (offset for planes) */
    v3 n2;         /* normal_ray direction: the collision normal re-normalised by Ray(point, normal), SceneObject.h:60 */
    v3 refl;       /* reflected ray direction (normalised), valid if reflective */
    v3 color;      /* object or texture colour */
    float diffuse, specular, reflective, intensity;
    int is_light, reflective_material;
} hit_record;

static void build_hit(const tcrt_scene* s, int obj, float dist, float px, float py, float den, v3 O, v3 D,
                      hit_record* h) {
    const int* info = s->obj_info + 4 * obj;
    const float* surf = s->obj_surface + 4 * obj;
    const float* mat = s->obj_material + 4 * obj;
    v3 n1; /* the normal the primitive passes to the CollisionObject ctor */
    h->color = v3_ld(surf);
    if (info[0] == TCRT_SPHERE) {
        const float* g = s->sphere_geom + 4 * info[1];
        h->P = v3_add(v3_scale(D, dist), O);            /* SceneSphere.cpp:122-123 */
        n1 = v3_normalize(v3_sub(h->P, v3_ld(g)));      /* :129-130 */
    } else {
        const float* g = (info[0] == TCRT_FINITE_PLANE) ? s->fin_geom + 16 * info[1] : s->inf_geom + 16 * info[1];
        v3 P = v3_add(v3_scale(D, dist), O);            /* SceneFinitePlane.cpp:108, SceneInfinitePlane.cpp:55 */
        if (info[3] >= 0) {
            if (info[0] == TCRT_INFINITE_PLANE) {       /* SceneInfinitePlane.cpp:59-74 */
                v3 PO = v3_sub(P, v3_ld(g + 12));
                px = v3_dot(PO, v3_ld(g + 4));
                py = v3_dot(PO, v3_ld(g + 8));
            }
            h->color = checker(s->textures + 8 * info[3], px, py); /* SceneFinitePlane.cpp:130-132 */
        }
        /* computeNormal: normal.dot(eyeDir) < 0 ? normal : reverseNormal
         * (SceneFinitePlane.cpp:167-174, SceneInfinitePlane.cpp:110-117); normal.dot(D) has the
         * same products and order as D.dot(normal) = den. */
        const float* nn = s->obj_normals + 8 * obj;
        n1 = (den < 0) ? v3_ld(nn) : v3_ld(nn + 4);
        /* intersection_point + temp_normal*INTERSECTION_OFFSET_DIST, the 1E-3 narrowed to
         * float by operator*(vector3d, sdecimal32)  (SceneFinitePlane.cpp:135-136) */
        h->P = v3_add(P, v3_scale(n1, (float)1E-3));
    }
    /* CollisionObject ctor, SceneObject.h:47-105 */
    h->n2 = v3_normalize(n1);                           /* Ray(point, normal) normalises, Ray.h:21-25 */
    float ndi = v3_dot(n1, D);                          /* :63 */
    h->diffuse = surf[3];
    h->specular = mat[0];
    h->reflective = mat[1];
    h->intensity = mat[2];
    h->is_light = info[2];
    h->reflective_material = (h->reflective > 0.0f);
    if (h->reflective_material) {
        v3 r = v3_make(-2 * n1.x * ndi + D.x, -2 * n1.y * ndi + D.y, -2 * n1.z * ndi + D.z); /* :81-83 */
        h->refl = v3_normalize(r);                      /* :85 via Ray ctor */
    } else {
        h->refl = v3_make(0, 0, 0);
    }
}

/* ---- getCollision, RayTracer.cpp:50-89 --------------------------------------------------- */
static int nearest_hit(const tcrt_scene* s, const tcrt_params* p, v3 O, v3 D, float* dist_o, float* px_o, float* py_o,
                       float* den_o, oracle_counters* c) {
    float best = p->far_dist;       /* FLOAT_MAX_VALUE */
    int best_obj = -1;
    float bx = 0, by = 0, bden = 0;
    float d, x, y, den;
    for (int i = 0; i < s->n_spheres; i++) {
        c->tests_sphere++;
        if (sphere_dist(s->sphere_geom + 4 * i, O, D, &d)) {
            int obj = s->sphere_obj[i];
            if (d < best || (d == best && best_obj >= 0 && obj < best_obj)) { best = d; best_obj = obj; }
        }
    }
    for (int i = 0; i < s->n_fin_planes; i++) {
        c->tests_fin++;
        if (fin_dist(s->fin_geom + 16 * i, O, D, &d, &x, &y, &den)) {
            int obj = s->fin_obj[i];
            if (d < best || (d == best && best_obj >= 0 && obj < best_obj)) {
                best = d; best_obj = obj; bx = x; by = y; bden = den;
            }
        }
    }
    for (int i = 0; i < s->n_inf_planes; i++) {
        c->tests_inf++;
        if (inf_dist(s->inf_geom + 16 * i, O, D, &d, &den)) {
            int obj = s->inf_obj[i];
            if (d < best || (d == best && best_obj >= 0 && obj < best_obj)) { best = d; best_obj = obj; bden = den; }
        }
    }
    *dist_o = best; *px_o = bx; *py_o = by; *den_o = bden;
    return best_obj;
}

/* ---- inShadeCollisionDetection, RayTracer.cpp:709-739: any non-light object closer than
 * the light.  The answer does not depend on the order; the loop runs in object-index order
 * with the reference's early exit so that the executed-test counters are the reference's. ---- */
static int in_shade(const tcrt_scene* s, v3 O, v3 D, float dist_to_light, oracle_counters* c) {
    float d, x, y, den;
    for (int obj = 0; obj < s->n_objects; obj++) {
        const int* info = s->obj_info + 4 * obj;
        if (info[2]) continue;                          /* :727 lights are skipped */
        int hit;
        if (info[0] == TCRT_SPHERE) {
            c->tests_sphere++;
            hit = sphere_dist(s->sphere_geom + 4 * info[1], O, D, &d);
        } else if (info[0] == TCRT_FINITE_PLANE) {
            c->tests_fin++;
            hit = fin_dist(s->fin_geom + 16 * info[1], O, D, &d, &x, &y, &den);
        } else {
            c->tests_inf++;
            hit = inf_dist(s->inf_geom + 16 * info[1], O, D, &d, &den);
        }
        if (hit && d < dist_to_light) return 1;         /* :729-731 */
    }
    return 0;
}

/* ---- the light loop of calculatePixel, RayTracer.cpp:537-591 ------------------------------ */
static v3 shade_lights(const tcrt_scene* s, const tcrt_params* p, const hit_record* h, v3 V, oracle_counters* c) {
    v3 final = v3_make(0, 0, 0);
    for (int l = 0; l < s->n_lights; l++) {
        int lo = s->light_obj[l];
        v3 Lpos = v3_ld(s->obj_origin + 4 * lo);
        v3 Lcol = v3_ld(s->obj_surface + 4 * lo);       /* the light's material colour, never its texture (:563) */
        float Lint = s->obj_material[4 * lo + 2];
        /* inShade, :743-771: dir = L - P; dist = |dir|; Ray(P, dir) normalises */
        v3 dir = v3_sub(Lpos, h->P);
        float dist = sqrtf(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
        v3 lr = v3_make(dir.x / dist, dir.y / dist, dir.z / dist); /* = dir.normalize(): same sqrt value */
        int shaded = 0;
        if (p->shadows_on) {
            c->rays_shadow++;
            shaded = in_shade(s, h->P, lr, dist, c);
        }
        if (shaded) continue;
        c->lit_pairs++;
        /* cosineShade, :654-701 (light_ray there equals lr) */
        if (h->diffuse > 0.0f) {
            float cdf = v3_dot(h->n2, lr);              /* :680 */
            if (cdf > 0.0f) {
                float factor = cdf * h->diffuse * Lint; /* :684 */
                final.x += factor * h->color.x * Lcol.x;
                final.y += factor * h->color.y * Lcol.y;
                final.z += factor * h->color.z * Lcol.z;
            }
            final.x = (final.x > 1.0f) ? 1.0f : final.x; /* :695-697, inside the diffuse>0 block */
            final.y = (final.y > 1.0f) ? 1.0f : final.y;
            final.z = (final.z > 1.0f) ? 1.0f : final.z;
        }
        /* specular, :561-588.  L equals lr; N is the normal_ray direction normalised once more. */
        v3 N = v3_normalize(h->n2);
        float two_ln = 2.0f * v3_dot(lr, N);
        v3 R = v3_sub(lr, v3_scale(N, two_ln));         /* L - 2.0f*L.dot(N)*N */
        float dot = v3_dot(V, R);
        if (dot > 0.0f) {
            c->lit_pairs_spec++;
            float pw = dot;
            for (int i = 0; i < 19; i++) pw *= dot;     /* :581-584 */
            float spec = pw * h->specular;
            final.x += Lcol.x * spec;                   /* spec_factor * light_color */
            final.y += Lcol.y * spec;
            final.z += Lcol.z * spec;
        }
    }
    return final;
}

/* ---- calculatePixel, RayTracer.cpp:448-638, as a loop + bottom-up fold --------------------- */
typedef struct { v3 local; float k; v3 obj; } level_rec;

static v3 trace_pixel(const tcrt_scene* s, const tcrt_params* p, v3 O, v3 D, level_rec* stack, oracle_counters* c) {
    v3 null_color = v3_ld(p->null_color);
    v3 tail;           /* colour returned by the deepest call */
    int depth = 0;     /* number of stacked (reflective) levels */
    int level = 0;
    for (;;) {
        /* level > MAX_RECURSION_LEVEL -> NULL_COLOR (:454-455) is handled where the child
         * would be spawned, below */
        float dist, px, py, den;
        int obj = nearest_hit(s, p, O, D, &dist, &px, &py, &den, c);
        if (obj < 0) { c->misses++; tail = null_color; break; } /* :507-509 */
        hit_record h;
        build_hit(s, obj, dist, px, py, den, O, D, &h);
        if (s->obj_info[4 * obj] == TCRT_SPHERE) c->hits_sphere++; else c->hits_plane++;
        if (h.reflective_material) c->hits_reflective++;
        if (s->obj_info[4 * obj + 3] >= 0) c->hits_textured++;
        if (h.is_light) c->hits_light++;
        if (h.is_light) {                               /* :520-527 */
            tail = v3_scale(h.color, h.intensity);
            break;
        }
        v3 local = shade_lights(s, p, &h, D, c);
        if (p->reflections_on && h.reflective_material) { /* :595-604 */
            stack[depth].local = local;
            stack[depth].k = h.reflective;
            stack[depth].obj = h.color;
            depth++;
            if (level + 1 > p->max_depth) { tail = null_color; break; }
            c->rays_reflect++;
            O = h.P;
            D = h.refl;
            level++;
            continue;
        }
        tail = local;
        break;
    }
    /* final_color += getReflectiveFactor() * reflective_color * object_color  (:601):
     * ((k * child) * obj), added to the level's local colour, deepest level first */
    c->combines += (unsigned long long)depth;
    for (int i = depth - 1; i >= 0; i--) {
        v3 kc = v3_scale(tail, stack[i].k);
        tail.x = stack[i].local.x + kc.x * stack[i].obj.x;
        tail.y = stack[i].local.y + kc.y * stack[i].obj.y;
        tail.z = stack[i].local.z + kc.z * stack[i].obj.z;
    }
    return tail;
}

/* Camera::createEyeRay + Ray(o, pf, pi): Camera.cpp:71-84, Ray.h:26-30; call site
 * RayTracer.cpp:916-918 (((float) x) / W with W an int constant). */
static void primary_ray(const tcrt_camera* cam, const tcrt_params* p, int x, int z, v3* O, v3* D) {
    float dx = ((float)x) / p->width;
    float dy = ((float)z) / p->height;
    float sx = dx * cam->screen_width - cam->screen_halfwidth;
    float sy = dy * cam->screen_height - cam->screen_halfheight;
    v3 pixel = v3_add(v3_ld(cam->screen_origin), v3_scale(v3_ld(cam->horizontal), sx));
    pixel = v3_add(pixel, v3_scale(v3_ld(cam->vertical), sy));
    *O = v3_ld(cam->eye);
    *D = v3_normalize(v3_sub(pixel, *O));
}

/* Renders columns x0, x0+stride, ... < x1 into out (x-major, z fastest, rgb).  counters may
 * be NULL.  Returns 0, or -1 on bad arguments. */
int tcrt_oracle_render(const tcrt_scene* s, const tcrt_camera* cam, const tcrt_params* p, int x0, int x1, int stride,
                       float* out, oracle_counters* counters) {
    if (!s || !cam || !p || !out || p->width <= 0 || p->height <= 0 || x0 < 0 || x1 > p->width || x0 > x1 ||
        stride < 1 || p->max_depth < 0)
        return -1;
    oracle_counters c;
    memset(&c, 0, sizeof(c));
    g_cnt = &c;
    level_rec* stack = (level_rec*)malloc(sizeof(level_rec) * (size_t)(p->max_depth + 2));
    if (!stack) return -1;
    size_t k = 0;
    for (int x = x0; x < x1; x += stride)
        for (int z = 0; z < p->height; z++) {
            v3 O, D;
            primary_ray(cam, p, x, z, &O, &D);
            c.rays_primary++;
            v3 col = trace_pixel(s, p, O, D, stack, &c);
            out[k++] = col.x; out[k++] = col.y; out[k++] = col.z;
        }
    free(stack);
    /* SURVEY §8(d) per-unit constants */
    c.flops = 30ull * c.rays_primary
            + 8ull * c.sph_exit_v + 16ull * c.sph_exit_d2 + 18ull * c.sph_hit
            + 12ull * c.tests_inf
            + 12ull * c.fin_exit_t + 31ull * c.fin_bounds
            + 32ull * c.hits_sphere + 31ull * c.hits_plane + 18ull * c.hits_reflective + 17ull * c.hits_textured
            + 18ull * (s->n_lights > 0 ? (c.rays_shadow ? c.rays_shadow
                                                        : (c.hits_sphere + c.hits_plane - c.hits_light) * (unsigned long long)s->n_lights)
                                       : 0ull)
            + 66ull * c.lit_pairs + 26ull * c.lit_pairs_spec
            + 9ull * c.combines;
    if (counters) *counters = c;
    return 0;
}

/* Single primary ray, for unit tests of ray generation. */
int tcrt_oracle_primary_ray(const tcrt_camera* cam, const tcrt_params* p, int x, int z, float* o3, float* d3) {
    v3 O, D;
    primary_ray(cam, p, x, z, &O, &D);
    o3[0] = O.x; o3[1] = O.y; o3[2] = O.z;
    d3[0] = D.x; d3[1] = D.y; d3[2] = D.z;
    return 0;
}

/* The writer's pixel line, RayTracer.cpp:1601 — glibc's own "%f". */
#include <stdio.h>
size_t tcrt_oracle_format_txt(const float* rgb, size_t n_pixels, char* out, size_t cap) {
    size_t len = 0;
    char line[160];
    for (size_t i = 0; i < n_pixels; i++) {
        int n = snprintf(line, sizeof line, "(%f, %f, %f)\n", rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]);
        if (out && len + (size_t)n <= cap) memcpy(out + len, line, (size_t)n);
        len += (size_t)n;
    }
    return len;
}
