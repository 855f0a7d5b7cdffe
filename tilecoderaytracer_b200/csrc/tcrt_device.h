// tcrt_device.h — internal: structures shared by the CUDA kernels and the C-ABI layer.
#ifndef TCRT_DEVICE_H_
#define TCRT_DEVICE_H_

#include <cuda_runtime.h>
#include <stdint.h>

#include "tcrt.h"

// ---- device-resident scene ---------------------------------------------------------------
// One allocation per device.  The "sweep blob" is what every CTA stages into shared memory
// with coalesced float4 loads; everything else is winner-only data read through the
// read-only path.
//
// Sweep blob layout (float4 units):
//   [0, n_sph)                      sphere   (cx, cy, cz, r^2)
//   [fin_off, fin_off + 4*n_fin)    finite   (n, -dto) (h, h_dist) (v, v_dist) (plane_origin, 0)
//   [inf_off, inf_off + n_inf)      infinite (n, -dto)
//   [light_off, light_off + 2*n_lights)  (light position, intensity) (light colour, 0)
//   [clu_off, clu_off + 4*n_clu)    box clusters of axis-aligned finite planes (tcrt_cluster.cpp):
//                                   (lo.x, hi.x, lo.y, hi.y) (lo.z, hi.z, present mask, -) (c of faces x0 x1 y0 y1) (c of z0 z1, -, -)
//                                   (pairs: one packed FADD2/FMUL2 handles both bounds / both faces of an axis)
//   [cslot_off, ...)                int32 finite-plane slot per cluster face (6 per cluster, -1 = absent)
//   [idx_off, ...)                  int32 object index per primitive: spheres, finite, infinite
// Within each type the non-light primitives come first (counts *_nl): the shadow sweep
// (inShadeCollisionDetection skips lights, RayTracer.cpp:727) just stops there.  The
// nearest-hit order inside a type is irrelevant because ties are resolved on the object
// index explicitly.
struct DeviceScene {
    const float4* blob;        // sweep blob (global)
    int blob_f4;               // size in float4 units (including the index tail, rounded up)
    int stage_off;             // CTAs stage blob[stage_off, blob_f4) into shared memory: the spheres a BVH covers
                               // (the first n_sph_bvh) are read from global memory through L1
    int n_sph, n_fin, n_inf;
    int n_sph_nl, n_fin_nl, n_inf_nl;   // non-light prefix lengths
    int n_sph_bvh, n_fin_bvh;           // BVH-covered prefix of the non-light prefix (0 = none)
    int n_lights;
    int fin_off, inf_off, light_off, idx_off;   // float4 offsets into the blob
    // finite-plane array order: [generic non-light | axis-aligned ("arect", never lights) | lights];
    // arects are reached through the box clusters only.  n_fin_gen == n_fin_nl - n_arect.
    int n_fin_gen, n_arect;
    int n_clu, clu_off, cslot_off;
    float clu_cx, clu_cy, clu_cz, clu_rbig;   // hull centre of the clusters; L1 radius + largest |coordinate|
    // winner-only, indexed by object index
    const float4* obj_surface;   // colour rgb, diffuse
    const float4* obj_material;  // specular, reflective, intensity, 0
    const float4* obj_normals;   // 6 per object: (n1, normalize(n1), normalize^2(n1)) facing, then reverse
    const int4* obj_info;        // type, slot (position in the permuted type array), is_light, texture id
    const float4* inf_frame;     // 3 per infinite plane slot: horizontal, vertical, origin
    const float4* textures;      // 2 per texture: (light rgb, width) (dark rgb, height)
    // optional BVHs over the non-light spheres / finite planes (nullptr = linear sweep); the
    // type's primitive array is then in leaf order.  Node layout: see tcrt_render.cu.
    const float4* bvh_sph;
    const float4* bvh_fin;
    int bvh_sph_root, bvh_fin_root;   // encoded like a child reference
    // optional uniform grid over the BVH-covered spheres (nullptr = none): a 3D-DDA finds the cells a ray crosses,
    // the spheres registered in a cell are tested exactly.  Rays fattened by more than grid_margin use the BVH.
    const float4* grid_cells;         // 2 per cell: (geometry of the cell's first sphere; r^2 = -1e30 in an empty cell: never
                                      // hit) (number of FURTHER spheres, slot of the first sphere, position of the second one
                                      // in grid_items, -) — the usual cell holds one sphere and costs ONE round of loads
    const int* grid_items;            // sphere slots, cell after cell
    float grid_lo[3], grid_hi[3], grid_cell[3], grid_inv_cell[3];   // hi = lo + cell * dims
    int grid_dims[3];
    float grid_margin;
    float grid_k2_max;                // fatten().m <= grid_margin  <=>  its k2 <= grid_k2_max (m grows with k2)
    // per-ray fattening of the boxes (tcrt_render.cu `fatten`): bounding sphere (centre, radius^2)
    // of everything inside the BVHs, smallest BVH sphere radius, largest |coordinate|
    float bvh_cx, bvh_cy, bvh_cz, bvh_r2, bvh_rmin, bvh_cmax;
};

// host BVH builder (tcrt_bvh.cpp).  boxes: n x 6 floats (lo xyz, hi xyz), already inflated.
// On return `order` is the leaf order (a permutation of 0..n-1), `nodes` 4 float4 per node,
// and the function value is the encoded root reference.
#ifdef __cplusplus
#include <vector>
int tcrt_build_bvh(const std::vector<float>& boxes, int n, std::vector<int>& order, std::vector<float4>& nodes,
                   int* depth);   // *depth: deepest node; the kernel's traversal stack holds TCRT_BVH_STACK entries
#define TCRT_BVH_STACK 48
// E = TCRT_SPH_E * 2(|O - Cs|^2 + Rs^2) bounds the rounding of the reference's sphere discriminant (`fatten`,
// tcrt_render_common.cuh); the host derives the grid's far-origin bound from the same constant
#define TCRT_SPH_E 4e-6f

// host builder of the uniform grid over the BVH-covered spheres (tcrt_bvh.cpp)
struct TcrtSphereGrid {
    float lo[3] = {0, 0, 0};
    float cell[3] = {1, 1, 1};
    int dims[3] = {0, 0, 0};
    float reg_margin = 0.f;             // a ray whose fattening is below this may use the grid
    std::vector<int> cell_start;        // dims[0]*dims[1]*dims[2] + 1
    std::vector<int> items;             // sphere slots (leaf order), cell by cell
};
bool tcrt_build_sphere_grid(const std::vector<float>& boxes, int n, TcrtSphereGrid& g);
#endif

// Pixel tile of one queue claim: TCRT_TILE_W columns x TCRT_TILE_H rows = 32 pixels, one per lane.
#ifndef TCRT_TILE_W
#define TCRT_TILE_W 4
#endif
#define TCRT_TILE_H (32 / TCRT_TILE_W)

struct RenderLaunch {
    DeviceScene scene;
    tcrt_camera cam;
    int width, height;
    int x0, x1;              // columns rendered by this launch
    int max_depth;
    int shadows_on, reflections_on;
    float null_r, null_g, null_b;
    float far_dist;
    float* out;              // (x1-x0)*height*3 floats, x-major
    unsigned int* queue;     // pixel queue head (zeroed before launch)
    unsigned long long* counters;   // [0] primary [1] shadow [2] reflect
    const int* row_order;    // optional: permutation of the rows [0, height) — of the 8-row tile rows [0, height/8) when
                             // `tiled` — in which the queue runs
    unsigned int* col_cost;  // optional (balancer pre-pass): [x1-x0] per column then [height] per row, += bounces of every finished pixel
    int n_objects;           // objects of the scene (checked builds verify table indices against it)
    int refill_min;          // idle lanes a warp waits for before it takes new pixels (1..32)
    int tiled;               // queue positions map to 4x8-pixel tiles (band width % 4 == 0, height % 8 == 0)
};

// host builder of the box clusters (tcrt_cluster.cpp).  Face f = 2*axis + k; an absent face has
// c = NaN and plane = -1; `plane` indexes the caller's list.
struct TcrtBoxCluster {
    float lo[3], hi[3];
    float c[6];
    int plane[6];
};
#ifdef __cplusplus
// planes: indices into fin_geom (16 floats each) of the candidate (non-light) finite planes.
// arect_of_plane[i] != 0 when planes[i] is axis-aligned and therefore owned by a cluster.
void tcrt_build_box_clusters(const float* fin_geom, const std::vector<int>& planes, std::vector<int>& arect_of_plane,
                             std::vector<TcrtBoxCluster>& out);
#endif

// render kernels (tcrt_render.cu).  wave_mem / wave_bytes: scratch for the wavefront path (tcrt_render_wave.cu), which
// tcrt_launch_render takes for large frames of sphere-BVH scenes when the scratch is big enough (tcrt_wave_mem_needed).
cudaError_t tcrt_launch_render(const RenderLaunch& rl, int sm_count, cudaStream_t stream, int* launches, void* wave_mem,
                               size_t wave_bytes);
// bytes of scratch the wavefront path wants for this launch, 0 when the launch does not take that path
size_t tcrt_wave_mem_needed(const RenderLaunch& rl);
bool tcrt_wave_applies(const RenderLaunch& rl, int fm);
size_t tcrt_wave_mem_bytes(size_t band_pixels, int max_depth);
cudaError_t tcrt_launch_render_wave(const RenderLaunch& rl, int fm, int sm_count, size_t smem_scene, void* mem, size_t mem_bytes,
                                    cudaStream_t stream, int* launches);
size_t tcrt_render_max_smem();
// kernel for scenes whose spheres sit in a uniform grid (tcrt_render_grid.cu); smem = bytes of the staged blob
cudaError_t tcrt_launch_render_grid(const RenderLaunch& rl, int fm, int sm_count, size_t smem, cudaStream_t stream);
// experimental task-pool kernel for sphere-BVH scenes (tcrt_render_pool.cu, developer builds only); smem_scene = bytes of the staged blob
cudaError_t tcrt_launch_render_pool(const RenderLaunch& rl, int fm, int sm_count, size_t smem_scene, cudaStream_t stream);
// n random (a.xyz, b) cases: div3 (shared-reciprocal division of the render kernel) vs __fdiv_rn; *bad += mismatches
cudaError_t tcrt_launch_div3_check(unsigned long long n, unsigned seed, unsigned long long* bad, cudaStream_t stream);   // dynamic shared memory the kernel may opt in to

// txt formatter (tcrt_format.cu)
// fixed path: 31-byte lines; general path: per-pixel lengths -> two-level exclusive scan
// (offs: n_pixels uint64 tile-local offsets; block_sums: ceil(n/256)+1 uint64, last = total).
cudaError_t tcrt_launch_txt_fixed_check(const float* rgb, size_t n_pixels, unsigned int* not_fixed_flag,
                                        cudaStream_t stream);
cudaError_t tcrt_launch_txt_fixed(const float* rgb, size_t n_pixels, char* text, cudaStream_t stream);
cudaError_t tcrt_launch_txt_lengths(const float* rgb, size_t n_pixels, unsigned long long* offs,
                                    unsigned long long* block_sums, cudaStream_t stream, int* launches);
cudaError_t tcrt_launch_txt_general(const float* rgb, size_t n_pixels, const unsigned long long* offs,
                                    const unsigned long long* block_offs, char* text, cudaStream_t stream);
// 8-bit quantisation + transposition of an x-major band (cols x height) into image rows (top row first, cols pixels each)
cudaError_t tcrt_launch_quantize8_image(const float* rgb, int cols, int height, unsigned char* out, cudaStream_t stream);
cudaError_t tcrt_launch_l2_flush(void* scratch, size_t bytes, cudaStream_t stream);
// 8 independent chains per thread, `iters` rounds: 8*iters FFMA, or 8*iters FMUL + 8*iters FADD
cudaError_t tcrt_launch_fp32_peak(bool fma, float* scratch, int grid, int iters, cudaStream_t stream);

#endif  // TCRT_DEVICE_H_
