// Results-invariant re-indexing of the reference's BVH subtrees, plain C++.
//
// The ABI's item stream encodes `Bvh` objects exactly as the reference builds them (median split
// on the longest axis, src/bvh.rs:22-81) and visits them (left first, src/bvh.rs:85-120).  For a
// subtree whose leaves hold only primitives the *result* of hit() does not depend on the tree
// above the leaves (see DESIGN.md for the argument), so the device library is free to index the
// same leaves — same leaf boxes, same primitive records, same order for ties — with a tree that
// is cheaper to walk: surface-area-heuristic splits, two child boxes per 64-byte node so a visit
// tests both children and descends into the nearer one first.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace rtiow {

constexpr uint32_t kLinkLeaf = 0x80000000u;   // bit 31: leaf; bits 24..30 item count; bits 0..23 first item
constexpr uint32_t kLinkEmpty = 0x7fffffffu;  // a child slot that holds nothing
constexpr uint32_t kMaxLeafItems = 127u;
constexpr int kAccelMaxDepth = 30;            // traversal stack is 32 entries

struct AccelLeaf {
    float mn[3], mx[3];     // the reference's leaf box, bit for bit
    uint32_t first, count;  // item range in the compacted stream
};

struct AccelNode {  // 64 bytes = 4 x float4
    float c0_min[3]; uint32_t link0;
    float c0_max[3]; uint32_t link1;
    float c1_min[3]; uint32_t pad0;
    float c1_max[3]; uint32_t pad1;
};
static_assert(sizeof(AccelNode) == 64, "AccelNode must be 64 bytes");

// The same node for the conservative test (path_logic.cuh trav_node_step_fast): per axis the four
// planes {child0 min, child0 max, child1 min, child1 max} and the same quad with min/max swapped, so
// that a ray reads its near/far-ordered planes with one 16-byte load per axis (which of the two
// quads depends only on the sign of its direction); then the links and `m`, the largest
// |coordinate| in the node (+inf when it is too large for the error bound to be trusted).
struct FastNode {  // 112 bytes = 7 x float4
    float ax[3][2][4];
    uint32_t link0, link1;
    float m;
    uint32_t pad;
};
static_assert(sizeof(FastNode) == 112, "FastNode must be 112 bytes");

inline FastNode to_fast_node(const AccelNode& n) {
    FastNode f{};
    const bool has1 = n.link1 != kLinkEmpty;
    float m = 1e-30f;  // never 0: the "test nothing" slack is m * inf
    for (int a = 0; a < 3; ++a) {
        const float q[4] = {n.c0_min[a], n.c0_max[a], n.c1_min[a], n.c1_max[a]};
        const float w[4] = {n.c0_max[a], n.c0_min[a], n.c1_max[a], n.c1_min[a]};
        std::memcpy(f.ax[a][0], q, 16);
        std::memcpy(f.ax[a][1], w, 16);
        for (int k = 0; k < (has1 ? 4 : 2); ++k) m = std::fmax(m, std::fabs(q[k]));
    }
    if (!(m <= 67108864.f)) m = std::numeric_limits<float>::infinity();  // 2^26; also catches NaN
    f.link0 = n.link0;
    f.link1 = n.link1;
    f.m = m;
    return f;
}

namespace accel_detail {

struct Box {
    float mn[3], mx[3];
    void reset() {
        for (int a = 0; a < 3; ++a) { mn[a] = std::numeric_limits<float>::infinity(); mx[a] = -std::numeric_limits<float>::infinity(); }
    }
    void grow(const float* omn, const float* omx) {  // exact min/max: the union is representable
        // a NaN coordinate is ignored, like fmin/fmax (which do not inline)
        for (int a = 0; a < 3; ++a) { mn[a] = omn[a] < mn[a] ? omn[a] : mn[a]; mx[a] = omx[a] > mx[a] ? omx[a] : mx[a]; }
    }
    double half_area() const {
        const double dx = static_cast<double>(mx[0]) - mn[0], dy = static_cast<double>(mx[1]) - mn[1], dz = static_cast<double>(mx[2]) - mn[2];
        const double a = dx * dy + dy * dz + dz * dx;
        return (a == a && a >= 0.0) ? a : std::numeric_limits<double>::max();
    }
};

inline int ceil_log2(size_t n) {
    int k = 0;
    while ((static_cast<size_t>(1) << k) < n) ++k;
    return k;
}

// Full-sweep SAH over per-axis presorted leaf lists: every range [b, e) of idx[0], idx[1], idx[2]
// holds the same leaves, each sorted by centroid on its axis (ties by leaf number), so a node costs
// three linear sweeps and three stable partitions — O(n log n) for the whole tree.
struct Builder {
    const std::vector<AccelLeaf>& leaves;
    std::vector<AccelNode>& nodes;
    std::vector<uint32_t> idx[3];
    std::vector<unsigned char> left_side;
    std::vector<uint32_t> scratch;
    std::vector<double> right_area;
    int max_depth_seen = 0;

    Builder(const std::vector<AccelLeaf>& l, std::vector<AccelNode>& n) : leaves(l), nodes(n) {
        const size_t count = leaves.size();
        for (int axis = 0; axis < 3; ++axis) {
            idx[axis].resize(count);
            for (size_t i = 0; i < count; ++i) idx[axis][i] = static_cast<uint32_t>(i);
            std::stable_sort(idx[axis].begin(), idx[axis].end(), [&](uint32_t a, uint32_t b) {
                return leaves[a].mn[axis] + leaves[a].mx[axis] < leaves[b].mn[axis] + leaves[b].mx[axis];
            });
        }
        left_side.assign(count, 0);
        scratch.resize(count);
        right_area.resize(count + 1);
    }

    static uint32_t leaf_link(const AccelLeaf& l) { return kLinkLeaf | (l.count << 24) | l.first; }

    // Moves the leaves flagged in left_side to the front of every axis list, keeping their order.
    void partition(size_t b, size_t e) {
        for (int axis = 0; axis < 3; ++axis) {
            std::vector<uint32_t>& v = idx[axis];
            size_t nl = b, nr = 0;
            for (size_t i = b; i < e; ++i) {
                if (left_side[v[i]]) v[nl++] = v[i];
                else scratch[nr++] = v[i];
            }
            std::copy(scratch.begin(), scratch.begin() + static_cast<std::ptrdiff_t>(nr), v.begin() + static_cast<std::ptrdiff_t>(nl));
        }
    }

    // Returns the link of the subtree over the leaves in [b, e) and its box.
    uint32_t build(size_t b, size_t e, int depth, Box* out_box) {
        max_depth_seen = std::max(max_depth_seen, depth);
        const size_t n = e - b;
        if (n == 1) {
            const AccelLeaf& l = leaves[idx[0][b]];
            std::memcpy(out_box->mn, l.mn, 12);
            std::memcpy(out_box->mx, l.mx, 12);
            return leaf_link(l);
        }
        int split_axis = -1;
        size_t best_i = n / 2;
        const bool force_median = (kAccelMaxDepth - depth) <= ceil_log2(n) + 1;
        if (!force_median) {
            double best_cost = std::numeric_limits<double>::max();
            for (int axis = 0; axis < 3; ++axis) {
                const uint32_t* v = idx[axis].data() + b;
                Box acc; acc.reset();
                for (size_t i = n; i-- > 1;) {
                    acc.grow(leaves[v[i]].mn, leaves[v[i]].mx);
                    right_area[i] = acc.half_area();
                }
                acc.reset();
                for (size_t i = 1; i < n; ++i) {
                    acc.grow(leaves[v[i - 1]].mn, leaves[v[i - 1]].mx);
                    const double cost = acc.half_area() * static_cast<double>(i) + right_area[i] * static_cast<double>(n - i);
                    if (cost < best_cost) {
                        best_cost = cost;
                        best_i = i;
                        split_axis = axis;
                    }
                }
            }
        }
        if (split_axis < 0) {  // median on the widest axis (also the NaN / depth-limit fallback)
            Box bx; bx.reset();
            for (size_t i = b; i < e; ++i) bx.grow(leaves[idx[0][i]].mn, leaves[idx[0][i]].mx);
            float ext = -1.f;
            split_axis = 0;
            for (int a = 0; a < 3; ++a) {
                const float x = bx.mx[a] - bx.mn[a];
                if (x > ext) { ext = x; split_axis = a; }
            }
            best_i = n / 2;
        }
        const uint32_t* v = idx[split_axis].data() + b;
        for (size_t i = 0; i < n; ++i) left_side[v[i]] = i < best_i ? 1 : 0;
        partition(b, e);
        const size_t split = b + best_i;
        const uint32_t me = static_cast<uint32_t>(nodes.size());
        nodes.push_back(AccelNode{});
        Box b0, b1;
        const uint32_t l0 = build(b, split, depth + 1, &b0);
        const uint32_t l1 = build(split, e, depth + 1, &b1);
        AccelNode& nd = nodes[me];
        std::memcpy(nd.c0_min, b0.mn, 12); std::memcpy(nd.c0_max, b0.mx, 12);
        std::memcpy(nd.c1_min, b1.mn, 12); std::memcpy(nd.c1_max, b1.mx, 12);
        nd.link0 = l0; nd.link1 = l1;
        out_box->reset();
        out_box->grow(b0.mn, b0.mx);
        out_box->grow(b1.mn, b1.mx);
        return me;
    }
};

}  // namespace accel_detail

// Builds the tree over `leaves`, appends its nodes to `nodes` (indices are absolute into `nodes`)
// and returns the index of the root node.  A single leaf gets a root node with one empty slot so
// that its box test still happens.
inline uint32_t accel_build(const std::vector<AccelLeaf>& leaves, std::vector<AccelNode>* nodes, int* depth_out) {
    if (leaves.size() == 1) {
        AccelNode nd{};
        std::memcpy(nd.c0_min, leaves[0].mn, 12); std::memcpy(nd.c0_max, leaves[0].mx, 12);
        for (int a = 0; a < 3; ++a) { nd.c1_min[a] = std::numeric_limits<float>::infinity(); nd.c1_max[a] = -std::numeric_limits<float>::infinity(); }
        nd.link0 = accel_detail::Builder::leaf_link(leaves[0]);
        nd.link1 = kLinkEmpty;
        nodes->push_back(nd);
        if (depth_out) *depth_out = 1;
        return static_cast<uint32_t>(nodes->size() - 1);
    }
    accel_detail::Builder bd(leaves, *nodes);
    accel_detail::Box root_box;
    const uint32_t root = bd.build(0, leaves.size(), 0, &root_box);
    if (depth_out) *depth_out = bd.max_depth_seen;
    return root;
}

}  // namespace rtiow
