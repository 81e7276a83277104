// Scene descriptor validation and the device blob layout, as plain C++ (no CUDA): shared by the
// C ABI implementation (rtiow_b200.cu) and by tests/kernel_host_harness.cpp.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../../include/rtiow_b200.h"
#include "accel_build.hpp"

namespace rtiow {

// Internal item kind (never part of the ABI): a re-indexed Bvh subtree (accel_build.hpp).
//   a_w = 7 | (skip << 4)   skip = index of the item after the subtree's primitives
//   a[0] = bits(root node index);  the subtree's primitive items follow, in stream order.
constexpr uint32_t kItemAccel = 7u;
// A leaf of a re-indexed subtree whose box is more than this many times the subtree's median leaf box becomes an
// "ordered leaf" (see Compactor::run).
constexpr float kFragileExtentRatio = 32.f;

inline uint32_t bits_of(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float float_of(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

inline bool ops_prefix(const rtiow_scene_desc_t* d, uint32_t outer, uint32_t inner) {
    // true if frame `outer`'s op list is a prefix of frame `inner`'s
    const rtiow_frame_t& fo = d->frames[outer];
    const rtiow_frame_t& fi = d->frames[inner];
    if (fo.n_ops > fi.n_ops) return false;
    for (uint32_t k = 0; k < fo.n_ops; ++k)
        if (std::memcmp(&d->ops[fo.first_op + k], &d->ops[fi.first_op + k], sizeof(rtiow_xform_op_t)) != 0) return false;
    return true;
}

inline int validate_desc(const rtiow_scene_desc_t* d, bool* has_frames, bool* uses_perlin, std::string* msg) {
    auto bad = [&](const std::string& m) { *msg = m; return static_cast<int>(RTIOW_ERR_INVALID_SCENE); };
    auto badarg = [&](const std::string& m) { *msg = m; return static_cast<int>(RTIOW_ERR_INVALID_ARG); };
    if (d->abi_version != RTIOW_B200_ABI_VERSION) return badarg("abi_version mismatch");
    if (!d->items || d->n_items == 0) return bad("Can't render a scene with zero items");
    if (d->n_items >= (1u << 28)) return bad("too many items");
    if (!d->frames || d->n_frames == 0 || d->frames[0].n_ops != 0) return bad("frames[0] must be the world frame");
    if (d->n_materials >= (1u << 24)) return bad("too many materials");
    if ((d->n_ops && !d->ops) || (d->n_materials && !d->materials) || (d->n_textures && !d->textures))
        return badarg("null array with non-zero count");
    for (uint32_t f = 0; f < d->n_frames; ++f) {
        const rtiow_frame_t& fr = d->frames[f];
        if (static_cast<uint64_t>(fr.first_op) + fr.n_ops > d->n_ops) return bad("frame op range out of bounds");
    }
    for (uint32_t k = 0; k < d->n_ops; ++k)
        if (d->ops[k].kind > RTIOW_OP_FLIP) return bad("unknown transform op");
    *uses_perlin = false;
    for (uint32_t t = 0; t < d->n_textures; ++t) {
        const rtiow_texture_t& tx = d->textures[t];
        if (tx.kind > RTIOW_TEX_PERLIN) return bad("unknown texture kind");
        if (tx.kind == RTIOW_TEX_CHECKER && (tx.child0 >= t || tx.child1 >= t))
            return bad("checker children must precede their parent (keeps the texture graph acyclic)");
        if (tx.kind == RTIOW_TEX_PERLIN) *uses_perlin = true;
    }
    if (*uses_perlin && (!d->perlin_vecs || !d->perlin_perm)) return bad("Perlin texture without Perlin tables");
    for (uint32_t m = 0; m < d->n_materials; ++m) {
        const rtiow_material_t& mt = d->materials[m];
        if (mt.kind > RTIOW_MAT_ISOTROPIC) return bad("unknown material kind");
        const bool textured = mt.kind == RTIOW_MAT_LAMBERTIAN || mt.kind == RTIOW_MAT_DIFFUSE_LIGHT || mt.kind == RTIOW_MAT_ISOTROPIC;
        if (textured && mt.tex >= d->n_textures) return bad("material texture index out of range");
    }
    if ((d->items[d->n_items - 1].a_w & 15u) != RTIOW_ITEM_END) return bad("item stream must end with RTIOW_ITEM_END");

    *has_frames = false;
    std::vector<uint32_t> frame_at(d->n_items);
    std::vector<char> is_boundary(d->n_items, 0);
    uint32_t cur = 0;
    auto prim_ok = [&](uint32_t i, uint32_t outer_frame) -> const char* {
        const rtiow_item_t& it = d->items[i];
        const uint32_t kind = it.a_w & 15u, frame = it.a_w >> 4;
        if (kind != RTIOW_ITEM_SPHERE && kind != RTIOW_ITEM_RECT && kind != RTIOW_ITEM_PRISM) return "expected a primitive item";
        if (frame >= d->n_frames) return "primitive frame out of range";
        if ((it.b_w & 0x00ffffffu) >= d->n_materials) return "primitive material out of range";
        if (!ops_prefix(d, outer_frame, frame)) return "primitive frame does not extend the enclosing frame";
        if (kind == RTIOW_ITEM_RECT && (((it.b_w >> 24) >> 2) & 3u) > 2u) return "rect axis out of range";
        return nullptr;
    };
    for (uint32_t i = 0; i < d->n_items; ++i) {
        frame_at[i] = cur;
        const rtiow_item_t& it = d->items[i];
        const uint32_t kind = it.a_w & 15u, payload = it.a_w >> 4;
        switch (kind) {
            case RTIOW_ITEM_END:
                if (i + 1 != d->n_items) return bad("RTIOW_ITEM_END before the end of the stream");
                break;
            case RTIOW_ITEM_BBOX:
                if (payload <= i || payload >= d->n_items) return bad("BBOX skip link must point forward, inside the stream");
                break;
            case RTIOW_ITEM_SPHERE:
            case RTIOW_ITEM_RECT:
            case RTIOW_ITEM_PRISM:
                if (const char* m = prim_ok(i, cur)) return bad(m);
                break;
            case RTIOW_ITEM_MEDIUM: {
                if (payload >= d->n_frames || !ops_prefix(d, cur, payload)) return bad("medium frame invalid");
                if ((it.b_w & 0x00ffffffu) >= d->n_materials) return bad("medium material out of range");
                const uint32_t run_end = bits_of(it.a[2]);
                if (run_end <= i + 1 || run_end >= d->n_items) return bad("medium without boundary items (a[2] must hold the index after its boundary run)");
                if (bits_of(it.a[1]) >= 65536u - 16u) return bad("medium id too large");
                // the boundary object, flattened in the medium's frame: primitives, and the boxes of a Bvh boundary
                for (uint32_t j = i + 1; j < run_end; ++j) {
                    const uint32_t jk = d->items[j].a_w & 15u, jp = d->items[j].a_w >> 4;
                    if (jk == RTIOW_ITEM_BBOX) {
                        if (jp <= j || jp > run_end) return bad("medium boundary: BBOX skip link must stay inside the boundary run");
                    } else if (const char* m = prim_ok(j, payload)) {
                        return bad(std::string("medium boundary: ") + m);
                    }
                    frame_at[j] = cur;
                    is_boundary[j] = 1;
                }
                i = run_end - 1;  // the boundary run is consumed with the medium
                break;
            }
            case RTIOW_ITEM_SET_FRAME:
                if (payload >= d->n_frames) return bad("SET_FRAME frame out of range");
                cur = payload;
                *has_frames = true;
                break;
            default:
                return bad("unknown item kind");
        }
    }
    for (uint32_t i = 0; i < d->n_items; ++i) {
        const rtiow_item_t& it = d->items[i];
        if ((it.a_w & 15u) == RTIOW_ITEM_BBOX && !is_boundary[i]) {
            const uint32_t tgt = it.a_w >> 4;
            if (frame_at[tgt] != frame_at[i]) return bad("BBOX skip link crosses a frame change");
            if (is_boundary[tgt]) return bad("BBOX skip link lands on a medium boundary item");
        }
    }
    return RTIOW_OK;
}


// How Bvh subtrees are laid out for the device (rtiow_b200.h RTIOW_TRAVERSAL_*).
enum BlobMode { kBlobFast = 0, kBlobReferenceOrder = 1, kBlobExact = 2 };

// The features of a scene (path_logic.cuh SF_*), without SF_ACCEL and SF_ORDERED (they depend on how the blob is built).
inline uint32_t scene_features(const rtiow_scene_desc_t* d) {
    uint32_t f = 0;
    auto wrapped = [&](uint32_t frame) { return frame < d->n_frames && d->frames[frame].n_ops != 0; };
    for (uint32_t i = 0; i < d->n_items; ++i) {
        const uint32_t kind = d->items[i].a_w & 15u, payload = d->items[i].a_w >> 4;
        if (kind == RTIOW_ITEM_SPHERE) f |= 1u | (wrapped(payload) ? 4u : 0u);
        else if (kind == RTIOW_ITEM_RECT || kind == RTIOW_ITEM_PRISM) f |= 2u | (wrapped(payload) ? 4u : 0u);
        else if (kind == RTIOW_ITEM_MEDIUM) {
            f |= 8u | (wrapped(payload) ? 4u : 0u);
            const uint32_t run_end = bits_of(d->items[i].a[2]);
            if (run_end != i + 2u || (i + 1u < d->n_items && (d->items[i + 1u].a_w & 15u) == RTIOW_ITEM_BBOX)) f |= 512u;  // SF_MEDIUM_RUN
        }
        else if (kind == RTIOW_ITEM_SET_FRAME) f |= 4u;
    }
    for (uint32_t m = 0; m < d->n_materials; ++m) {
        const rtiow_material_t& mt = d->materials[m];
        const bool textured = mt.kind == RTIOW_MAT_LAMBERTIAN || mt.kind == RTIOW_MAT_DIFFUSE_LIGHT || mt.kind == RTIOW_MAT_ISOTROPIC;
        if (textured && d->textures[mt.tex].kind != RTIOW_TEX_CONSTANT) f |= 16u;
        if (mt.kind == RTIOW_MAT_DIFFUSE_LIGHT) f |= 32u;
        if (mt.kind == RTIOW_MAT_ISOTROPIC) f |= 64u;
        if (mt.kind == RTIOW_MAT_METAL || mt.kind == RTIOW_MAT_DIELECTRIC) f |= 128u;
    }
    for (uint32_t t = 0; t < d->n_textures; ++t)
        if (d->textures[t].kind == RTIOW_TEX_CHECKER) f |= 2048u;  // SF_CHECKER
    for (uint32_t k = 0; k < d->n_ops; ++k)
        if (d->ops[k].kind == RTIOW_OP_SCALE) f |= 4096u;          // SF_SCALE
    return f;
}

struct BlobLayout {
    uint32_t off_nodes, off_frames, off_ops, off_mats, off_tex, off_pvecs, off_pperm;
    uint32_t off_fnodes;
    uint32_t n_items, n_nodes, n_accel, accel_depth;
    uint32_t n_derived; // leaves whose box is derived from their primitive's record (no BBOX item)
    uint32_t n_ordered; // leaves kept out of their re-indexed subtree, tested after it in the reference's order
    uint32_t n_prisms;  // rect_prism records, fused from six Rect items each or supplied as RTIOW_ITEM_PRISM
};

// ---------------------------------------------------------------------------------------------
// Stream compaction + re-indexing.  Every BBOX subtree that consists of boxes and primitives only
// (no medium, no frame switch) becomes ONE kItemAccel item followed by the subtree's primitive
// items in their original relative order (so "earlier in the stream" still means "earlier in the
// reference's visiting order" for tie-breaking), plus a tree of AccelNodes over the reference's
// own leaf boxes.  Everything else is copied, with skip links remapped.
// ---------------------------------------------------------------------------------------------
namespace blob_detail {

// ---------------------------------------------------------------------------------------------
// rect_prism (src/object.rs:420-473) arrives as the six Rect items of its And tree.  Six consecutive
// items that are exactly that — same frame and material, axes z, y, x, z, y, x, the first three at
// p1 and the last three (with the opposite FlipNormals state) at p0, ranges equal bit for bit — are
// replaced by ONE prism record {p0, p1}, which the device evaluates as the same six Rect::hit calls
// in the same order (path_logic.cuh prism_hit_t): same image, a sixth of the item loads and of the
// shared-memory footprint (the final scene's 400 boxes: 77 KB -> 13 KB).
// ---------------------------------------------------------------------------------------------
inline bool same_bits(float a, float b) { return bits_of(a) == bits_of(b); }

// Internal primitive flag (never part of the ABI): the primitive is alone in a leaf of a re-indexed subtree and the
// leaf's box — the reference's `bounding_box()` of that object — is what the record itself yields:
//   Translate{Sphere}: (-r + offset, r + offset)      object.rs:113-118, 285-291
//   rect_prism:        (p0 - 0.0001, p1 + 0.0001)     object.rs:220-233 merged by And (object.rs:412-416)
// checked bit for bit against the BBOX item the host sent, which is then dropped: a leaf visit loads one record instead
// of two (path_logic.cuh trav_leaf_visit_fast) and the scene needs 32 bytes less shared memory per leaf.
constexpr uint32_t kFlagBoxDerived = 16u;

inline bool leaf_box_is_derivable(const rtiow_item_t& box, const rtiow_item_t& prim) {
    const uint32_t kind = prim.a_w & 15u, flags = prim.b_w >> 24;
    float mn[3], mx[3];
    if (kind == RTIOW_ITEM_SPHERE) {
        const float r = prim.a[0];
        for (int a = 0; a < 3; ++a) {
            const float off = (flags & RTIOW_FLAG_HAS_OFFSET) ? prim.b[a] : 0.f;
            mn[a] = (flags & RTIOW_FLAG_HAS_OFFSET) ? -r + off : -r;
            mx[a] = (flags & RTIOW_FLAG_HAS_OFFSET) ? r + off : r;
        }
    } else if (kind == RTIOW_ITEM_PRISM) {
        for (int a = 0; a < 3; ++a) {
            mn[a] = prim.a[a] - 0.0001f;
            mx[a] = prim.b[a] + 0.0001f;
        }
    } else {
        return false;
    }
    for (int a = 0; a < 3; ++a)
        if (!same_bits(mn[a], box.a[a]) || !same_bits(mx[a], box.b[a])) return false;
    return true;
}

inline bool is_prism_run(const rtiow_item_t* it, rtiow_item_t* fused) {
    static const uint32_t axes[6] = {2u, 1u, 0u, 2u, 1u, 0u};
    const uint32_t frame = it[0].a_w >> 4, mat = it[0].b_w & 0x00ffffffu;
    const bool flip0 = ((it[0].b_w >> 24) & RTIOW_FLAG_FLIP) != 0;
    for (int f = 0; f < 6; ++f) {
        const uint32_t flags = it[f].b_w >> 24;
        if ((it[f].a_w & 15u) != RTIOW_ITEM_RECT || (it[f].a_w >> 4) != frame || (it[f].b_w & 0x00ffffffu) != mat) return false;
        if (((flags >> 2) & 3u) != axes[f] || (flags & ~(RTIOW_FLAG_FLIP | (3u << 2))) != 0u) return false;
        if (((flags & RTIOW_FLAG_FLIP) != 0) != (f < 3 ? flip0 : !flip0)) return false;
    }
    // item = {k, r0.start, r0.end} {r1.start, r1.end}; z faces: r0 = x, r1 = y; y faces: r0 = x, r1 = z; x faces: r0 = y, r1 = z
    const float p0x = it[0].a[1], p1x = it[0].a[2], p0y = it[0].b[0], p1y = it[0].b[1], p1z = it[0].a[0], p0z = it[3].a[0];
    const float want[6][5] = {{p1z, p0x, p1x, p0y, p1y}, {p1y, p0x, p1x, p0z, p1z}, {p1x, p0y, p1y, p0z, p1z},
                              {p0z, p0x, p1x, p0y, p1y}, {p0y, p0x, p1x, p0z, p1z}, {p0x, p0y, p1y, p0z, p1z}};
    for (int f = 0; f < 6; ++f) {
        const float have[5] = {it[f].a[0], it[f].a[1], it[f].a[2], it[f].b[0], it[f].b[1]};
        for (int k = 0; k < 5; ++k)
            if (!same_bits(have[k], want[f][k])) return false;
    }
    *fused = rtiow_item_t{};
    fused->a[0] = p0x; fused->a[1] = p0y; fused->a[2] = p0z;
    fused->b[0] = p1x; fused->b[1] = p1y; fused->b[2] = p1z;
    fused->a_w = RTIOW_ITEM_PRISM | (frame << 4);
    fused->b_w = mat | ((flip0 ? static_cast<uint32_t>(RTIOW_FLAG_FLIP) : 0u) << 24);
    return true;
}

// Returns the stream with every such run fused, links (BBOX skip, medium run end) remapped.
inline std::vector<rtiow_item_t> fuse_prisms(const rtiow_item_t* items, uint32_t n) {
    std::vector<char> is_target(static_cast<size_t>(n) + 1, 0);  // an item some link jumps to
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t kind = items[i].a_w & 15u;
        if (kind == RTIOW_ITEM_BBOX) is_target[std::min(items[i].a_w >> 4, n)] = 1;
        else if (kind == RTIOW_ITEM_MEDIUM) is_target[std::min(bits_of(items[i].a[2]), n)] = 1;
    }
    std::vector<rtiow_item_t> out;
    out.reserve(n);
    std::vector<uint32_t> new_index(static_cast<size_t>(n) + 1, 0);
    for (uint32_t i = 0; i < n;) {
        rtiow_item_t fused;
        bool ok = i + 6 <= n && (items[i].a_w & 15u) == RTIOW_ITEM_RECT;
        for (uint32_t k = 1; ok && k < 6; ++k) ok = !is_target[i + k];  // nothing may jump into the middle of the six
        if (ok && is_prism_run(items + i, &fused)) {
            for (uint32_t k = 0; k < 6; ++k) new_index[i + k] = static_cast<uint32_t>(out.size());
            out.push_back(fused);
            i += 6;
        } else {
            new_index[i] = static_cast<uint32_t>(out.size());
            out.push_back(items[i]);
            i += 1;
        }
    }
    new_index[n] = static_cast<uint32_t>(out.size());
    for (rtiow_item_t& it : out) {
        const uint32_t kind = it.a_w & 15u;
        if (kind == RTIOW_ITEM_BBOX) it.a_w = RTIOW_ITEM_BBOX | (new_index[std::min(it.a_w >> 4, n)] << 4);
        else if (kind == RTIOW_ITEM_MEDIUM) it.a[2] = float_of(new_index[std::min(bits_of(it.a[2]), n)]);
    }
    return out;
}

struct Compactor {
    const rtiow_scene_desc_t* d;
    bool enable_accel;
    bool keep_leaf_boxes;  // kBlobFast: every leaf's own BBOX item stays in front of its primitives
    std::vector<rtiow_item_t> out;
    std::vector<AccelNode> nodes;
    uint32_t n_accel = 0, n_ordered = 0, n_derived = 0;
    int max_depth = 0;
    bool derive_leaf_boxes = true;
    uint32_t frame_of_box = 0;  // the frame the BBOX items of the subtree being compacted live in
    bool odd_nesting = false;  // a skip link that leaves its enclosing box: the stream is shipped as it is

    uint32_t kind(uint32_t i) const { return d->items[i].a_w & 15u; }
    uint32_t payload(uint32_t i) const { return d->items[i].a_w >> 4; }

    // Is the subtree of BBOX item i "simple"?  Collects its leaves (box item index, prim range).
    struct RawLeaf { uint32_t box, first, end; };
    bool classify(uint32_t i, std::vector<RawLeaf>* leaves) const {
        const uint32_t s = payload(i);
        if (i + 1 >= s) return false;
        if (kind(i + 1) == RTIOW_ITEM_BBOX) {  // inner node: children tile (i, s)
            uint32_t j = i + 1;
            while (j < s) {
                if (kind(j) != RTIOW_ITEM_BBOX || payload(j) > s) return false;
                if (!classify(j, leaves)) return false;
                j = payload(j);
            }
            return j == s;
        }
        for (uint32_t j = i + 1; j < s; ++j)
            if (kind(j) != RTIOW_ITEM_SPHERE && kind(j) != RTIOW_ITEM_RECT && kind(j) != RTIOW_ITEM_PRISM) return false;
        if (s - (i + 1) > kMaxLeafItems) return false;
        leaves->push_back(RawLeaf{i, i + 1, s});
        return true;
    }

    void run(uint32_t begin, uint32_t end) {
        uint32_t i = begin;
        while (i < end && !odd_nesting) {
            const uint32_t k = kind(i);
            if (k == RTIOW_ITEM_BBOX) {
                const uint32_t s = payload(i);
                std::vector<RawLeaf> raw;
                if (enable_accel && s <= end && d->n_items < (1u << 24) && classify(i, &raw)) {
                    // Leaves whose box dwarfs the others' (book-1's radius-1000 ground sphere among radius-0.2 ones) stay OUT
                    // of the tree: f32 cancellation makes such a primitive's t fall outside its own box's computed interval
                    // often enough (1 sample in 4 M on book-1) that the result depends on whether the reference asked it
                    // before or after a competing hit.  They follow the subtree as "ordered leaves" (BBOX item with
                    // b_w = 1 + index of the first subtree item the reference visits after them), tested with exactly the
                    // range the reference's order gives them (path_logic.cuh trav_stream).
                    std::vector<char> fragile(raw.size(), 0);
                    if (raw.size() >= 4) {
                        std::vector<float> ext(raw.size());
                        for (size_t k = 0; k < raw.size(); ++k) {
                            const rtiow_item_t& bx = d->items[raw[k].box];
                            ext[k] = std::fmax(std::fmax(bx.b[0] - bx.a[0], bx.b[1] - bx.a[1]), bx.b[2] - bx.a[2]);
                        }
                        std::vector<float> sorted = ext;
                        std::nth_element(sorted.begin(), sorted.begin() + static_cast<std::ptrdiff_t>(sorted.size() / 2), sorted.end());
                        const float median = sorted[sorted.size() / 2];
                        size_t n_fragile = 0;
                        for (size_t k = 0; k < raw.size(); ++k)
                            if (!(ext[k] <= kFragileExtentRatio * median)) { fragile[k] = 1; ++n_fragile; }   // also NaN / inf
                        if (raw.size() - n_fragile < 2) std::fill(fragile.begin(), fragile.end(), 0);
                    }
                    const size_t at = out.size();
                    out.push_back(rtiow_item_t{});
                    std::vector<AccelLeaf> leaves;
                    leaves.reserve(raw.size());
                    std::vector<uint32_t> next_regular(raw.size() + 1, 0);  // first tree item the reference visits after leaf k
                    for (size_t k = 0; k < raw.size(); ++k) {  // stream order
                        const RawLeaf& rl = raw[k];
                        if (fragile[k]) continue;
                        AccelLeaf l{};
                        std::memcpy(l.mn, d->items[rl.box].a, 12);
                        std::memcpy(l.mx, d->items[rl.box].b, 12);
                        // the frame of the prim must be the subtree's (no wrapper between the Bvh and the leaf object)
                        const bool derived = keep_leaf_boxes && derive_leaf_boxes && rl.end - rl.first == 1 &&
                                             (d->items[rl.first].a_w >> 4) == frame_of_box && leaf_box_is_derivable(d->items[rl.box], d->items[rl.first]);
                        if (keep_leaf_boxes && !derived) {
                            rtiow_item_t box = d->items[rl.box];
                            box.a_w = RTIOW_ITEM_BBOX | (static_cast<uint32_t>(out.size() + 1 + (rl.end - rl.first)) << 4);
                            box.b_w = 0;
                            out.push_back(box);
                        }
                        next_regular[k] = static_cast<uint32_t>(out.size()) - ((keep_leaf_boxes && !derived) ? 1u : 0u);
                        l.first = static_cast<uint32_t>(out.size());
                        l.count = rl.end - rl.first;
                        for (uint32_t j = rl.first; j < rl.end; ++j) out.push_back(d->items[j]);
                        if (derived) {
                            out.back().b_w |= kFlagBoxDerived << 24;
                            ++n_derived;
                        }
                        leaves.push_back(l);
                    }
                    const uint32_t tree_end = static_cast<uint32_t>(out.size());
                    next_regular[raw.size()] = tree_end;
                    for (size_t k = raw.size(); k-- > 0;)
                        if (fragile[k]) next_regular[k] = next_regular[k + 1];
                    int depth = 0;
                    const uint32_t root = accel_build(leaves, &nodes, &depth);
                    max_depth = depth > max_depth ? depth : max_depth;
                    rtiow_item_t& acc = out[at];
                    acc.a_w = kItemAccel | (tree_end << 4);
                    std::memcpy(&acc.a[0], &root, 4);
                    ++n_accel;
                    for (size_t k = 0; k < raw.size(); ++k) {  // the ordered leaves, in the reference's order
                        if (!fragile[k]) continue;
                        const RawLeaf& rl = raw[k];
                        rtiow_item_t box = d->items[rl.box];
                        box.a_w = RTIOW_ITEM_BBOX | (static_cast<uint32_t>(out.size() + 1 + (rl.end - rl.first)) << 4);
                        box.b_w = next_regular[k] + 1u;
                        out.push_back(box);
                        for (uint32_t j = rl.first; j < rl.end; ++j) out.push_back(d->items[j]);
                        ++n_ordered;
                    }
                    i = s;
                } else if (s <= end) {  // keep the box on the reference-order stream, recurse inside
                    const size_t at = out.size();
                    out.push_back(d->items[i]);
                    out[at].b_w = 0;  // b_w of a BBOX item is the library's own (ordered-leaf marker)
                    run(i + 1, s);
                    out[at].a_w = RTIOW_ITEM_BBOX | (static_cast<uint32_t>(out.size()) << 4);
                    i = s;
                } else {  // improperly nested skip link (validated to be forward): cannot happen for streams
                    odd_nesting = true;  // produced by the flatteners; keep semantics by disabling compaction
                }
            } else if (k == RTIOW_ITEM_MEDIUM) {  // the medium and its boundary run, as they are (inner skip links shifted)
                const uint32_t run_end = bits_of(d->items[i].a[2]);
                const size_t at = out.size();
                out.push_back(d->items[i]);
                const uint32_t delta = static_cast<uint32_t>(out.size()) - (i + 1u);  // modulo 2^32: may be "negative"
                for (uint32_t j = i + 1; j < run_end; ++j) {
                    rtiow_item_t it = d->items[j];
                    if ((it.a_w & 15u) == RTIOW_ITEM_BBOX) it.a_w = RTIOW_ITEM_BBOX | (((it.a_w >> 4) + delta) << 4);
                    out.push_back(it);
                }
                out[at].a[2] = float_of(static_cast<uint32_t>(out.size()));
                i = run_end;
            } else {
                if (k == RTIOW_ITEM_SET_FRAME) frame_of_box = payload(i);
                out.push_back(d->items[i]);
                i += 1;
            }
        }
    }
};

}  // namespace blob_detail

// items | accel nodes | frames | ops | materials | textures | perlin vecs (float4) | perlin perms;
// every section starts on a 128-byte boundary so the whole blob can be moved with 128 B-granular
// TMA bulk copies.
inline std::vector<unsigned char> build_blob(const rtiow_scene_desc_t* d_in, bool uses_perlin, BlobLayout* lay,
                                             BlobMode mode = kBlobFast, bool fuse_rect_prisms = true) {
    const bool enable_accel = mode != kBlobReferenceOrder;
    rtiow_scene_desc_t fused_desc = *d_in;
    std::vector<rtiow_item_t> fused_items;
    if (fuse_rect_prisms) {
        fused_items = blob_detail::fuse_prisms(d_in->items, d_in->n_items);
        fused_desc.items = fused_items.data();
        fused_desc.n_items = static_cast<uint32_t>(fused_items.size());
    }
    const rtiow_scene_desc_t* d = &fused_desc;
    std::vector<unsigned char> blob;
    auto align_up = [](size_t v) { return (v + 127u) / 128u * 128u; };
    auto append = [&](const void* src, size_t bytes) -> uint32_t {
        const size_t off = align_up(blob.size());
        blob.resize(off + bytes, 0);
        if (bytes) std::memcpy(blob.data() + off, src, bytes);
        return static_cast<uint32_t>(off);
    };
    blob_detail::Compactor cp{d, enable_accel, mode == kBlobFast, {}, {}};
    cp.run(0, d->n_items);
    if (cp.odd_nesting) {  // ship the stream as it is
        cp = blob_detail::Compactor{d, false, false, {}, {}};
        cp.out.assign(d->items, d->items + d->n_items);
        for (rtiow_item_t& it : cp.out)
            if ((it.a_w & 15u) == RTIOW_ITEM_BBOX) it.b_w = 0;
    }
    append(cp.out.data(), sizeof(rtiow_item_t) * cp.out.size());
    lay->n_items = static_cast<uint32_t>(cp.out.size());
    lay->n_prisms = 0;
    for (const rtiow_item_t& it : cp.out) lay->n_prisms += (it.a_w & 15u) == RTIOW_ITEM_PRISM ? 1u : 0u;
    if (mode == kBlobFast) {
        std::vector<FastNode> fast;
        fast.reserve(cp.nodes.size());
        for (const AccelNode& n : cp.nodes) fast.push_back(to_fast_node(n));
        lay->off_fnodes = append(fast.data(), sizeof(FastNode) * fast.size());
        lay->off_nodes = lay->off_fnodes;
    } else {
        lay->off_nodes = append(cp.nodes.data(), sizeof(AccelNode) * cp.nodes.size());
        lay->off_fnodes = 0;
    }
    lay->n_nodes = static_cast<uint32_t>(cp.nodes.size());
    lay->n_accel = cp.n_accel;
    lay->n_ordered = cp.n_ordered;
    lay->n_derived = cp.n_derived;
    lay->accel_depth = static_cast<uint32_t>(cp.max_depth);
    lay->off_frames = append(d->frames, sizeof(rtiow_frame_t) * d->n_frames);
    lay->off_ops = append(d->ops, sizeof(rtiow_xform_op_t) * d->n_ops);
    {   // materials, with constant textures baked in: {kind | texkind<<8, tex, param, 0} {color/albedo, 0}
        std::vector<float> mats(8 * static_cast<size_t>(d->n_materials), 0.f);
        for (uint32_t m = 0; m < d->n_materials; ++m) {
            const rtiow_material_t& mt = d->materials[m];
            uint32_t texkind = 0xffu;
            float col[3] = {mt.albedo[0], mt.albedo[1], mt.albedo[2]};
            const bool textured = mt.kind == RTIOW_MAT_LAMBERTIAN || mt.kind == RTIOW_MAT_DIFFUSE_LIGHT || mt.kind == RTIOW_MAT_ISOTROPIC;
            if (textured) {
                texkind = d->textures[mt.tex].kind;
                if (texkind == RTIOW_TEX_CONSTANT) std::memcpy(col, d->textures[mt.tex].color, 12);
            }
            const uint32_t w0 = mt.kind | (texkind << 8);
            std::memcpy(&mats[8 * m + 0], &w0, 4);
            std::memcpy(&mats[8 * m + 1], &mt.tex, 4);
            mats[8 * m + 2] = mt.param;
            mats[8 * m + 4] = col[0]; mats[8 * m + 5] = col[1]; mats[8 * m + 6] = col[2];
        }
        lay->off_mats = append(mats.data(), mats.size() * 4);
    }
    lay->off_tex = append(d->textures, sizeof(rtiow_texture_t) * d->n_textures);
    if (uses_perlin) {
        std::vector<float> v4(4 * 256, 0.f);
        for (int i = 0; i < 256; ++i) std::memcpy(&v4[4 * i], d->perlin_vecs + 3 * i, 12);
        lay->off_pvecs = append(v4.data(), v4.size() * 4);
        lay->off_pperm = append(d->perlin_perm, 768);
    } else {
        lay->off_pvecs = lay->off_pperm = 0;
    }
    blob.resize(align_up(blob.size()), 0);
    return blob;
}

}  // namespace rtiow
