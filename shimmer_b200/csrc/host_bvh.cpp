// libshimmer_host.so -- host-side helpers that stand in for the parts of shimmer's Rust host
// this image cannot compile (no cargo/rustc): the BVH build that feeds sg_scene_create.
//
// In the real integration `BvhAggregate::new` (src/aggregate.rs:207-290) runs unchanged in
// Rust and the shim only narrows `LinearBvhNode` to SgBvhNode.  Here the same topology is
// produced natively so the synthetic scenes traverse exactly the tree shimmer would build:
//   * top-down, split axis = largest extent of the centroid bounds (aggregate.rs:340-343)
//   * split position = midpoint of the centroid bounds, primitives with
//     centroid[dim] < pmid go left (:359-364); equal-count median fallback when one side
//     is empty (:366-374)
//   * leaf iff one primitive, zero surface area, or degenerate centroid bounds (:326,:345);
//     `maxnodeprims` is ignored by the reference (:41,:60,:285), so it is ignored here
//   * depth-first linear layout, first child at index+1 (:425-467)
// The work list is explicit (no recursion) and nodes are emitted in pre-order directly.
#include "../../include/shimmer_gpu.h"
#include "sg_host_tables.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

namespace {

struct Item { uint32_t index; float lo[3], hi[3]; };

inline float centroid(const Item& it, int a) { return 0.5f * it.lo[a] + it.hi[a] * 0.5f; }   // aggregate.rs:490-492

struct Frame {
    size_t begin, end;     // primitive range
    uint32_t node;         // node index already reserved for this range (or UINT32_MAX)
    uint32_t parent;       // parent to patch with second_child_offset (UINT32_MAX = none)
    bool is_second;
};

}  // namespace

extern "C" {

// prim_bounds: min.xyz max.xyz per primitive.  out_nodes: capacity 2n-1.  out_order[i] = input
// index of the i-th primitive in leaf order.  Returns the node count, or -1 on bad input.
int64_t sh_bvh_build(int64_t n, const float* prim_bounds, SgBvhNode* out_nodes, uint32_t* out_order) {
    if (n <= 0 || !prim_bounds || !out_nodes || !out_order) return -1;
    std::vector<Item> items((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        items[i].index = (uint32_t)i;
        for (int a = 0; a < 3; ++a) { items[i].lo[a] = prim_bounds[6 * i + a]; items[i].hi[a] = prim_bounds[6 * i + 3 + a]; }
    }
    uint32_t n_nodes = 0, n_ordered = 0;
    std::vector<Frame> work;
    work.push_back({0, (size_t)n, UINT32_MAX, UINT32_MAX, false});
    const float fmax_v = std::numeric_limits<float>::max(), fmin_v = std::numeric_limits<float>::lowest();
    while (!work.empty()) {
        Frame f = work.back(); work.pop_back();
        const uint32_t me = n_nodes++;
        if (f.parent != UINT32_MAX && f.is_second) out_nodes[f.parent].offset = me;
        Item* it = items.data() + f.begin;
        const size_t cnt = f.end - f.begin;
        float lo[3] = {fmax_v, fmax_v, fmax_v}, hi[3] = {fmin_v, fmin_v, fmin_v};
        float clo[3] = {fmax_v, fmax_v, fmax_v}, chi[3] = {fmin_v, fmin_v, fmin_v};
        for (size_t i = 0; i < cnt; ++i) for (int a = 0; a < 3; ++a) {
            lo[a] = std::fmin(lo[a], it[i].lo[a]); hi[a] = std::fmax(hi[a], it[i].hi[a]);
            const float c = centroid(it[i], a);
            clo[a] = std::fmin(clo[a], c); chi[a] = std::fmax(chi[a], c);
        }
        SgBvhNode nd; std::memset(&nd, 0, sizeof nd);
        for (int a = 0; a < 3; ++a) { nd.bmin[a] = lo[a]; nd.bmax[a] = hi[a]; }
        const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        const float area = 2.0f * (dx * dy + dx * dz + dy * dz);                // bounding_box.rs:394-397
        const float ex = chi[0] - clo[0], ey = chi[1] - clo[1], ez = chi[2] - clo[2];
        const int dim = (ex > ey && ex > ez) ? 0 : (ey > ez ? 1 : 2);          // bounding_box.rs:404-415
        if (area == 0.0f || cnt == 1 || chi[dim] == clo[dim]) {
            nd.offset = n_ordered; nd.n_prims = (uint16_t)cnt;
            for (size_t i = 0; i < cnt; ++i) out_order[n_ordered++] = it[i].index;
            out_nodes[me] = nd;
            continue;
        }
        const float pmid = (clo[dim] + chi[dim]) / 2.0f;
        // in-place front/back swap partition -- same element moves as itertools::partition
        size_t i = 0, j = cnt;
        size_t split = 0;
        while (i < j) {
            if (centroid(it[i], dim) < pmid) { ++i; ++split; continue; }
            bool swapped = false;
            while (j > i + 1) { --j; if (centroid(it[j], dim) < pmid) { std::swap(it[i], it[j]); swapped = true; break; } }
            if (!swapped) break;
            ++i; ++split;
        }
        if (split == 0 || split == cnt) {
            split = cnt / 2;
            std::nth_element(it, it + split, it + cnt, [dim](const Item& a, const Item& b) { return centroid(a, dim) < centroid(b, dim); });
        }
        nd.axis = (uint8_t)dim; nd.n_prims = 0; nd.offset = 0;
        out_nodes[me] = nd;
        // interior bounds are the union of the children (init_interior :540-545) == bounds of the range
        work.push_back({f.begin + split, f.end, UINT32_MAX, me, true});    // second child: processed after the whole first subtree
        work.push_back({f.begin, f.begin + split, UINT32_MAX, me, false});
    }
    return (int64_t)n_nodes;
}

// Triangle::bounds (triangle.rs:508-511) for every (mesh-global) triangle: Bounds3f::new(p0,p1).union_point(p2)
void sh_triangle_bounds(int64_t n_tris, const uint32_t* indices, const float* p, float* out_bounds) {
    for (int64_t t = 0; t < n_tris; ++t) {
        const float* a = p + 3 * (size_t)indices[3 * t];
        const float* b = p + 3 * (size_t)indices[3 * t + 1];
        const float* c = p + 3 * (size_t)indices[3 * t + 2];
        for (int k = 0; k < 3; ++k) {
            out_bounds[6 * t + k] = std::fmin(std::fmin(a[k], b[k]), c[k]);
            out_bounds[6 * t + 3 + k] = std::fmax(std::fmax(a[k], b[k]), c[k]);
        }
    }
}

// Interval table of one piecewise-linear spectrum (sg_host_tables.h): out holds kSpecLutBins = 471 entries.  Returns 1 when the
// table was built, 0 when the device keeps the binary search (n < 2, n > 65535, unsorted knots).
int sh_spectrum_lut(const float* lambdas, int32_t n, uint16_t* out) {
    if (!lambdas || !out) return 0;
    return sg::build_spectrum_lut(lambdas, n, out) ? 1 : 0;
}

}  // extern "C"
