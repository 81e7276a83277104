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

namespace accel_detail {

struct Box {
    float mn[3], mx[3];
    void reset() {
        for (int a = 0; a < 3; ++a) { mn[a] = std::numeric_limits<float>::infinity(); mx[a] = -std::numeric_limits<float>::infinity(); }
    }
    void grow(const float* omn, const float* omx) {  // exact min/max: the union is representable
        for (int a = 0; a < 3; ++a) { mn[a] = std::fmin(mn[a], omn[a]); mx[a] = std::fmax(mx[a], omx[a]); }
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

struct Builder {
    const std::vector<AccelLeaf>& leaves;
    std::vector<AccelNode>& nodes;
    std::vector<uint32_t> order;  // permutation of leaf indices, partitioned in place
    int max_depth_seen = 0;

    static uint32_t leaf_link(const AccelLeaf& l) { return kLinkLeaf | (l.count << 24) | l.first; }

    Box bounds(size_t b, size_t e) const {
        Box bx; bx.reset();
        for (size_t i = b; i < e; ++i) bx.grow(leaves[order[i]].mn, leaves[order[i]].mx);
        return bx;
    }

    // Returns the link of the subtree over order[b, e) and its box.
    uint32_t build(size_t b, size_t e, int depth, Box* out_box) {
        max_depth_seen = std::max(max_depth_seen, depth);
        const size_t n = e - b;
        if (n == 1) {
            const AccelLeaf& l = leaves[order[b]];
            std::memcpy(out_box->mn, l.mn, 12);
            std::memcpy(out_box->mx, l.mx, 12);
            return leaf_link(l);
        }
        size_t split = b + n / 2;
        int split_axis = -1;
        const bool force_median = (kAccelMaxDepth - depth) <= ceil_log2(n) + 1;
        if (!force_median) {
            double best_cost = std::numeric_limits<double>::max();
            size_t best_i = 0;
            std::vector<uint32_t> tmp(order.begin() + static_cast<std::ptrdiff_t>(b), order.begin() + static_cast<std::ptrdiff_t>(e));
            std::vector<double> right_area(n);
            auto sort_axis = [&](int axis) {
                std::copy(order.begin() + static_cast<std::ptrdiff_t>(b), order.begin() + static_cast<std::ptrdiff_t>(e), tmp.begin());
                std::stable_sort(tmp.begin(), tmp.end(), [&](uint32_t l, uint32_t r) {
                    return leaves[l].mn[axis] + leaves[l].mx[axis] < leaves[r].mn[axis] + leaves[r].mx[axis];
                });
            };
            for (int axis = 0; axis < 3; ++axis) {
                sort_axis(axis);
                Box acc; acc.reset();
                for (size_t i = n; i-- > 1;) {
                    acc.grow(leaves[tmp[i]].mn, leaves[tmp[i]].mx);
                    right_area[i] = acc.half_area();
                }
                acc.reset();
                for (size_t i = 1; i < n; ++i) {
                    acc.grow(leaves[tmp[i - 1]].mn, leaves[tmp[i - 1]].mx);
                    const double cost = acc.half_area() * static_cast<double>(i) + right_area[i] * static_cast<double>(n - i);
                    if (cost < best_cost) {
                        best_cost = cost;
                        best_i = i;
                        split_axis = axis;
                    }
                }
            }
            if (split_axis >= 0) {
                sort_axis(split_axis);
                std::copy(tmp.begin(), tmp.end(), order.begin() + static_cast<std::ptrdiff_t>(b));
                split = b + best_i;
            }
        }
        if (split_axis < 0) {  // median on the widest axis (also the NaN / depth-limit fallback)
            const Box bx = bounds(b, e);
            int axis = 0;
            float ext = -1.f;
            for (int a = 0; a < 3; ++a) {
                const float x = bx.mx[a] - bx.mn[a];
                if (x > ext) { ext = x; axis = a; }
            }
            std::stable_sort(order.begin() + static_cast<std::ptrdiff_t>(b), order.begin() + static_cast<std::ptrdiff_t>(e), [&](uint32_t l, uint32_t r) {
                return leaves[l].mn[axis] + leaves[l].mx[axis] < leaves[r].mn[axis] + leaves[r].mx[axis];
            });
            split = b + n / 2;
        }
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
    accel_detail::Builder bd{leaves, *nodes, {}, 0};
    bd.order.resize(leaves.size());
    for (size_t i = 0; i < leaves.size(); ++i) bd.order[i] = static_cast<uint32_t>(i);
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
    accel_detail::Box root_box;
    const uint32_t root = bd.build(0, leaves.size(), 0, &root_box);
    if (depth_out) *depth_out = bd.max_depth_seen;
    return root;
}

}  // namespace rtiow
