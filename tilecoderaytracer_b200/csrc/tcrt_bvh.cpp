// tcrt_bvh.cpp — host-side BVH builder for the render kernel's culling structure.
//
// The reference tests every object for every ray (getCollision, RayTracer.cpp:73-86); its author
// lists "octree parsing" as not supported (README.md:25-29).  The BVH here is only a pruning
// device: the kernel still evaluates the reference's exact arithmetic for every primitive whose
// (inflated) box the ray may touch, so results are unchanged (see tcrt_render.cu).
//
// Binary tree, surface-area heuristic with a full sweep on each axis, leaves of <= 4 primitives
// (<= 8 when splitting does not pay), depth capped at 40 by falling back to median splits.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "tcrt_device.h"

namespace {

struct Box {
    float lo[3], hi[3];
    void reset() {
        for (int k = 0; k < 3; k++) { lo[k] = INFINITY; hi[k] = -INFINITY; }
    }
    void grow(const float* b) {
        for (int k = 0; k < 3; k++) {
            lo[k] = std::min(lo[k], b[k]);
            hi[k] = std::max(hi[k], b[3 + k]);
        }
    }
    double area() const {
        double dx = (double)hi[0] - lo[0], dy = (double)hi[1] - lo[1], dz = (double)hi[2] - lo[2];
        if (!(dx >= 0) || !(dy >= 0) || !(dz >= 0)) return 0.0;
        if (!std::isfinite(dx) || !std::isfinite(dy) || !std::isfinite(dz)) return 1e300;
        return 2.0 * (dx * dy + dy * dz + dz * dx);
    }
};

struct Builder {
    const float* boxes;            // n x 6
    std::vector<int>& order;       // permuted in place
    std::vector<float4>& nodes;
    std::vector<double> suffix;
    int max_depth = 0;
    int max_leaf = 4;   // 4 and 6 measure the same, 1-3 slower (profiles/README.md)

    Builder(const float* b, std::vector<int>& o, std::vector<float4>& n) : boxes(b), order(o), nodes(n) {}

    float centroid(int id, int axis) const {
        float lo = boxes[6 * id + axis], hi = boxes[6 * id + 3 + axis];
        if (!std::isfinite(lo) || !std::isfinite(hi)) return 0.f;
        return 0.5f * (lo + hi);
    }
    Box bounds(int first, int count) const {
        Box b;
        b.reset();
        for (int i = first; i < first + count; i++) b.grow(boxes + 6 * order[i]);
        return b;
    }
    static int leaf_ref(int first, int count) { return ~(first | (count << 24)); }

    // returns the encoded reference of the subtree over order[first, first+count)
    int build(int first, int count, int depth) {
        max_depth = std::max(max_depth, depth);
        if (count <= max_leaf) return leaf_ref(first, count);
        const Box parent = bounds(first, count);
        int best_axis = -1, best_split = -1;
        double best_cost = 1e308;
        if (depth < 32) {
            for (int axis = 0; axis < 3; axis++) {
                std::sort(order.begin() + first, order.begin() + first + count,
                          [&](int a, int b) { return centroid(a, axis) < centroid(b, axis); });
                suffix.assign(count + 1, 0.0);
                Box r;
                r.reset();
                for (int i = count - 1; i >= 1; i--) {
                    r.grow(boxes + 6 * order[first + i]);
                    suffix[i] = r.area();
                }
                Box l;
                l.reset();
                for (int i = 1; i < count; i++) {
                    l.grow(boxes + 6 * order[first + i - 1]);
                    double cost = l.area() * i + suffix[i] * (count - i);
                    if (cost < best_cost) { best_cost = cost; best_axis = axis; best_split = i; }
                }
            }
            // no split beats a leaf: keep small groups together
            if (count <= 2 * max_leaf && best_cost >= parent.area() * count) return leaf_ref(first, count);
        }
        if (best_axis < 0) {   // depth guard: balanced median split on the widest centroid axis
            float clo[3] = {INFINITY, INFINITY, INFINITY}, chi[3] = {-INFINITY, -INFINITY, -INFINITY};
            for (int i = first; i < first + count; i++)
                for (int k = 0; k < 3; k++) {
                    float c = centroid(order[i], k);
                    clo[k] = std::min(clo[k], c);
                    chi[k] = std::max(chi[k], c);
                }
            best_axis = 0;
            for (int k = 1; k < 3; k++)
                if (chi[k] - clo[k] > chi[best_axis] - clo[best_axis]) best_axis = k;
            best_split = count / 2;
        }
        const int axis = best_axis;
        std::sort(order.begin() + first, order.begin() + first + count,
                  [&](int a, int b) { return centroid(a, axis) < centroid(b, axis); });
        const int me = (int)(nodes.size() / 4);
        nodes.resize(nodes.size() + 4);
        const int c0 = build(first, best_split, depth + 1);
        const int c1 = build(first + best_split, count - best_split, depth + 1);
        const Box b0 = bounds(first, best_split), b1 = bounds(first + best_split, count - best_split);
        float4 q3;
        memcpy(&q3.x, &c0, 4);
        memcpy(&q3.y, &c1, 4);
        q3.z = q3.w = 0.f;
        nodes[4 * me + 0] = make_float4(b0.lo[0], b0.lo[1], b0.lo[2], b0.hi[0]);
        nodes[4 * me + 1] = make_float4(b0.hi[1], b0.hi[2], b1.lo[0], b1.lo[1]);
        nodes[4 * me + 2] = make_float4(b1.lo[2], b1.hi[0], b1.hi[1], b1.hi[2]);
        nodes[4 * me + 3] = q3;
        return me;
    }
};

}  // namespace

int tcrt_build_bvh(const std::vector<float>& boxes, int n, std::vector<int>& order, std::vector<float4>& nodes, int* depth) {
    order.resize(n);
    for (int i = 0; i < n; i++) order[i] = i;
    nodes.clear();
    Builder b(boxes.data(), order, nodes);
    const int root = b.build(0, n, 0);
    if (depth) *depth = b.max_depth;
    return root;
}
