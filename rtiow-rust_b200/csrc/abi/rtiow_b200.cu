// C ABI of the B200 render path (include/rtiow_b200.h): scene validation + upload, the
// persistent megakernel launch per sample pass, the sample fold, statistics.
// No CPU fallback: every entry point fails with RTIOW_ERR_NO_DEVICE / RTIOW_ERR_CUDA if the
// device path is unavailable.
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../../include/rtiow_b200.h"
#include <cub/device/device_radix_sort.cuh>

#include "../device/aux_kernels.cuh"
#include "kernel_table.hpp"
#include "unit_plan.hpp"
#include "scene_blob.hpp"

namespace {

thread_local std::string g_err;

int set_err(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e__ = (call);                                                                         \
        if (e__ != cudaSuccess)                                                                           \
            return set_err(RTIOW_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));          \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// Device buffers, stream and events of one render in flight.  A scene owns one for its lifetime;
// destroyed scenes hand theirs to a small per-device cache so that create/render/destroy cycles
// (one per frame in a host application) do not pay cudaMalloc/cudaFree of the sample staging buffer.
struct Workspace {
    int device = 0;
    // Two pipeline slots (staging, pass accumulator, unit counter, a stream for the megakernel): consecutive renders
    // alternate between them, so that render k+1's CTAs move onto the SMs that render k's CTAs leave while its last
    // long paths are still running (enqueue_render).  Frames whose staging is large use slot 0 only.
    struct Slot {
        DevBuf staging, accum;
        cudaStream_t stream = nullptr;
        cudaEvent_t render_done = nullptr, fold_done = nullptr;
        bool fold_recorded = false;
        // Tile order of the slot's next render: after a render's last fold its tiles are sorted by the longest path found
        // in each (DevBufs: keys in reverse tile order, sorted keys, the values n-1 .. 0, the sorted values = the order,
        // CUB's scratch).  `order_shape` says which launch shape the order belongs to.  tiles_in and tile_order each start
        // with a 128-byte header that holds the launch's unit counter: the kernel finds the order behind its counter
        // (KParams), whichever of the two a launch uses.
        DevBuf longest, longest_sorted, tiles_in, tile_order, sort_tmp;
        uint64_t order_shape[4] = {0, 0, 0, 0};
        bool order_valid = false;
        uint32_t order_age = 0;   // renders since the order was learnt
        uint32_t default_order_n = 0;  // tiles_in holds the default order of this many tiles (x 2 + direction)
    } slot[4];
    uint32_t next_slot = 0, n_slots = 2;
    DevBuf out, samples;
    DevBuf scene_blob[3];  // device image of the owning scene, by blob mode
    unsigned long long* d_segs = nullptr;
    cudaStream_t stream = nullptr;
    std::vector<cudaEvent_t> events;  // [start, trace_end, fold_end] per pass
    // two pinned bounce buffers: device -> pinned -> caller's pageable memory, pipelined
    static constexpr size_t kBounce = 4u << 20;
    void* pinned[2] = {nullptr, nullptr};
    cudaEvent_t pin_ev[2] = {nullptr, nullptr};
    // Renders may be enqueued on the caller's streams: `busy` marks the end of the last one, so that whoever
    // touches the buffers next (a render on another stream, a scene that inherits this workspace from the
    // cache, destroy) orders itself after it.
    cudaEvent_t busy = nullptr;
    cudaStream_t busy_stream = nullptr;
    bool busy_recorded = false;

    // Orders `s` after the last render that used these buffers (no-op on the stream that ran it).
    cudaError_t order_after_last_render(cudaStream_t s) {
        if (!busy_recorded || s == busy_stream) return cudaSuccess;
        return cudaStreamWaitEvent(s, busy, 0);
    }
    cudaError_t mark_render_end(cudaStream_t s) {
        cudaError_t e = cudaEventRecord(busy, s);
        if (e == cudaSuccess) { busy_stream = s; busy_recorded = true; }
        return e;
    }
    // Blocks the host until the last render that used these buffers has finished.
    void wait_idle() {
        if (busy_recorded) cudaEventSynchronize(busy);
        if (stream) cudaStreamSynchronize(stream);
        for (Slot& sl : slot)
            if (sl.stream) cudaStreamSynchronize(sl.stream);
    }

    cudaError_t init(int dev) {
        device = dev;
        cudaError_t e;
        if ((e = cudaEventCreateWithFlags(&busy, cudaEventDisableTiming)) != cudaSuccess) return e;
        for (Slot& sl : slot) {
            if ((e = cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking)) != cudaSuccess) return e;
            if ((e = cudaEventCreateWithFlags(&sl.render_done, cudaEventDisableTiming)) != cudaSuccess) return e;
            if ((e = cudaEventCreateWithFlags(&sl.fold_done, cudaEventDisableTiming)) != cudaSuccess) return e;
        }
        if ((e = cudaMalloc(reinterpret_cast<void**>(&d_segs), sizeof(unsigned long long))) != cudaSuccess) return e;
        if ((e = cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking)) != cudaSuccess) return e;
        return cudaSuccess;
    }
    void destroy() {
        cudaSetDevice(device);
        wait_idle();
        if (busy) cudaEventDestroy(busy);
        for (Slot& sl : slot) {
            if (sl.stream) cudaStreamSynchronize(sl.stream);
            sl.staging.release(); sl.accum.release();
            sl.longest.release(); sl.longest_sorted.release(); sl.tiles_in.release(); sl.tile_order.release(); sl.sort_tmp.release();
            if (sl.render_done) cudaEventDestroy(sl.render_done);
            if (sl.fold_done) cudaEventDestroy(sl.fold_done);
            if (sl.stream) cudaStreamDestroy(sl.stream);
        }
        out.release(); samples.release();
        for (DevBuf& b : scene_blob) b.release();
        if (d_segs) cudaFree(d_segs);
        for (cudaEvent_t ev : events) cudaEventDestroy(ev);
        for (int i = 0; i < 2; ++i) {
            if (pinned[i]) cudaFreeHost(pinned[i]);
            if (pin_ev[i]) cudaEventDestroy(pin_ev[i]);
        }
        if (stream) cudaStreamDestroy(stream);
    }
    // Device -> caller's (pageable) host memory, ordered after everything enqueued on `stream`;
    // returns when the data is in `dst`.
    cudaError_t copy_to_host(void* dst, const void* src, size_t bytes) {
        cudaError_t e;
        if (bytes <= (256u << 10)) {
            if ((e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
            return cudaStreamSynchronize(stream);
        }
        for (int i = 0; i < 2; ++i) {
            if (!pinned[i] && (e = cudaMallocHost(&pinned[i], kBounce)) != cudaSuccess) return e;
            if (!pin_ev[i] && (e = cudaEventCreateWithFlags(&pin_ev[i], cudaEventDisableTiming)) != cudaSuccess) return e;
        }
        const size_t n = (bytes + kBounce - 1) / kBounce;
        auto size_of = [&](size_t i) { return std::min(kBounce, bytes - i * kBounce); };
        auto drain = [&](size_t i) -> cudaError_t {
            cudaError_t de = cudaEventSynchronize(pin_ev[i & 1]);
            if (de == cudaSuccess) std::memcpy(static_cast<char*>(dst) + i * kBounce, pinned[i & 1], size_of(i));
            return de;
        };
        for (size_t i = 0; i < n; ++i) {
            if (i >= 2 && (e = drain(i - 2)) != cudaSuccess) return e;
            if ((e = cudaMemcpyAsync(pinned[i & 1], static_cast<const char*>(src) + i * kBounce, size_of(i), cudaMemcpyDeviceToHost,
                                     stream)) != cudaSuccess) return e;
            if ((e = cudaEventRecord(pin_ev[i & 1], stream)) != cudaSuccess) return e;
        }
        for (size_t i = n >= 2 ? n - 2 : 0; i < n; ++i)
            if ((e = drain(i)) != cudaSuccess) return e;
        return cudaSuccess;
    }
};

// The few device attributes the launch logic needs, queried once per device (cudaGetDeviceProperties
// costs milliseconds; a host application creates a scene per frame).
struct DeviceInfo {
    int major = 0, minor = 0, sm_count = 0, max_smem_optin = 0;
};
std::mutex g_dev_mutex;
std::vector<std::pair<int, DeviceInfo>> g_dev_info;

cudaError_t device_info(int device, DeviceInfo* out) {
    std::lock_guard<std::mutex> lock(g_dev_mutex);
    for (const auto& e : g_dev_info)
        if (e.first == device) { *out = e.second; return cudaSuccess; }
    DeviceInfo d;
    cudaError_t e;
    if ((e = cudaDeviceGetAttribute(&d.major, cudaDevAttrComputeCapabilityMajor, device)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&d.minor, cudaDevAttrComputeCapabilityMinor, device)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&d.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device)) != cudaSuccess) return e;
    g_dev_info.emplace_back(device, d);
    *out = d;
    return cudaSuccess;
}

// Per-(kernel, device) launch state, set once: the dynamic shared memory limit is raised to the device's opt-in
// maximum (the attribute is process-wide per function and device: setting it to each scene's own size before
// every launch would let two host threads rendering different scenes undercut each other), registers are read once.
struct KernelInfo {
    const void* fn;
    int device;
    int regs;
};
std::mutex g_kernel_mutex;
std::vector<KernelInfo> g_kernel_info;

cudaError_t kernel_info(const void* fn, int device, int max_smem_optin, int* regs) {
    std::lock_guard<std::mutex> lock(g_kernel_mutex);
    for (const KernelInfo& k : g_kernel_info)
        if (k.fn == fn && k.device == device) { *regs = k.regs; return cudaSuccess; }
    cudaError_t e;
    cudaFuncAttributes fa{};
    if ((e = cudaFuncGetAttributes(&fa, fn)) != cudaSuccess) return e;
    // static + dynamic shared memory share the opt-in limit
    const int dyn_max = max_smem_optin - static_cast<int>(fa.sharedSizeBytes);
    if ((e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_max)) != cudaSuccess) return e;
    g_kernel_info.push_back(KernelInfo{fn, device, fa.numRegs});
    *regs = fa.numRegs;
    return cudaSuccess;
}

std::mutex g_ws_mutex;
std::vector<Workspace*> g_ws_cache;

// rtiow_b200_render_multi keeps the devices' frames between calls of the same shape (at most two shapes)
struct MultiFrames {
    uint32_t nx = 0, ny = 0;
    std::vector<int> devices;
    std::vector<rtiow_peer_frame_t*> pf;
    void destroy();
};
std::mutex g_multi_mutex;
std::vector<MultiFrames> g_multi_cache;

constexpr size_t kWsCachePerDevice = 2;

Workspace* ws_acquire(int device, cudaError_t* err) {
    {
        std::lock_guard<std::mutex> lock(g_ws_mutex);
        for (size_t i = 0; i < g_ws_cache.size(); ++i)
            if (g_ws_cache[i]->device == device) {
                Workspace* w = g_ws_cache[i];
                g_ws_cache.erase(g_ws_cache.begin() + static_cast<std::ptrdiff_t>(i));
                // The learnt tile orders stay with the workspace: a host application that creates a scene per frame renders
                // frame k+1 in the order frame k taught.  They are only schedules (any permutation renders the same image);
                // one inherited from a DIFFERENT scene is merely a worse schedule, and lasts one render: the first render of
                // the new owner learns its own.
                for (Workspace::Slot& sl : w->slot) sl.order_age = 15u;
                return w;
            }
    }
    auto w = new Workspace();
    *err = w->init(device);
    if (*err != cudaSuccess) {
        w->destroy();
        delete w;
        return nullptr;
    }
    return w;
}

void ws_release(Workspace* w) {
    if (!w) return;
    cudaSetDevice(w->device);
    w->wait_idle();  // also renders enqueued on the caller's streams (rtiow_b200_render_rows_device)
    {
        std::lock_guard<std::mutex> lock(g_ws_mutex);
        size_t same = 0;
        for (Workspace* c : g_ws_cache) same += c->device == w->device;
        if (same < kWsCachePerDevice) {
            g_ws_cache.push_back(w);
            return;
        }
    }
    w->destroy();
    delete w;
}

}  // namespace

// A private copy of the caller's descriptor: blobs for the other traversal modes are built on demand.
struct OwnedDesc {
    std::vector<rtiow_item_t> items;
    std::vector<rtiow_frame_t> frames;
    std::vector<rtiow_xform_op_t> ops;
    std::vector<rtiow_material_t> materials;
    std::vector<rtiow_texture_t> textures;
    std::vector<float> perlin_vecs;
    std::vector<uint8_t> perlin_perm;
    rtiow_scene_desc_t d{};
    void assign(const rtiow_scene_desc_t* src) {
        items.assign(src->items, src->items + src->n_items);
        frames.assign(src->frames, src->frames + src->n_frames);
        if (src->n_ops) ops.assign(src->ops, src->ops + src->n_ops);
        if (src->n_materials) materials.assign(src->materials, src->materials + src->n_materials);
        if (src->n_textures) textures.assign(src->textures, src->textures + src->n_textures);
        if (src->perlin_vecs) perlin_vecs.assign(src->perlin_vecs, src->perlin_vecs + 768);
        if (src->perlin_perm) perlin_perm.assign(src->perlin_perm, src->perlin_perm + 768);
        d = *src;
        d.items = items.data();
        d.frames = frames.data();
        d.ops = ops.empty() ? nullptr : ops.data();
        d.materials = materials.empty() ? nullptr : materials.data();
        d.textures = textures.empty() ? nullptr : textures.data();
        d.perlin_vecs = perlin_vecs.empty() ? nullptr : perlin_vecs.data();
        d.perlin_perm = perlin_perm.empty() ? nullptr : perlin_perm.data();
    }
};

struct rtiow_scene {
    int device = 0;
    int sm_count = 0;
    int max_smem_optin = 0;
    OwnedDesc desc;
    bool uses_perlin = false;
    // device blobs by rtiow::BlobMode, built and uploaded on first use
    struct Blob {
        bool built = false;
        std::vector<unsigned char> host;
        unsigned char* d = nullptr;
        uint32_t bytes = 0;
        rtiow::BlobLayout lay{};
    } blobs[3];
    int traversal = RTIOW_TRAVERSAL_REINDEXED;
    bool has_frames = false;
    uint32_t bg_kind = 0;
    float bg0[3] = {0, 0, 0}, bg1[3] = {0, 0, 0};

    Workspace* ws = nullptr;
    uint32_t events_used = 0;
#ifdef RT_CTA_TIMELINE
    std::vector<std::pair<unsigned long long*, uint32_t>> cta_timelines;
#endif
    bool timeline_on = false;                 // RTIOW_B200_TIMELINE=1: keep every launch's events, print them at scene_destroy
    std::vector<cudaEvent_t> timeline;        // per pass: render start / end (slot stream), fold start / end (caller's stream)

    // tuning
    uint32_t cta_threads = 0, ctas_per_sm = 0, staging_mib = 0, sample_chunk = 0;
    bool force_global = false;
    uint32_t features = 0;         // scene_blob.hpp scene_features()
    bool specialise = true;        // use a kernel compiled for a subset of features when the scene allows (same image)
    uint32_t refill_lanes = 0;     // 0 = automatic
    bool costly_segments = false;  // the scene has wrapper frames on subtrees or constant media
    bool pipeline = true;          // consecutive renders alternate between two pipeline slots (RTIOW_B200_PIPELINE=0: one)
    bool fuse_prisms = true;       // six-Rect rect_prism runs become one prism record (RTIOW_B200_FUSE_PRISMS=0: keep the rects)
    bool bottom_first = true;      // unit order (RTIOW_B200_UNIT_ORDER=0: top rows first)
    bool order_hint = true;        // RTIOW_B200_ORDER_HINT=0: never reorder the tiles by the previous render's path lengths
    int phase_sync = -1;           // -1 = automatic (RTIOW_B200_PHASE_SYNC, read once at scene_create)
    uint32_t phase_group_env = 0;  // RTIOW_B200_PHASE_GROUP, 0 = automatic

    // last render
    rtiow_stats_t stats{};
};

namespace {

using rtiow::KParams;

using rtiow::KernelVariant;

rtiow_scene::Blob& blob_of(rtiow_scene* s, rtiow::BlobMode mode) {
    rtiow_scene::Blob& B = s->blobs[mode];
    if (!B.built) {
        B.host = rtiow::build_blob(&s->desc.d, s->uses_perlin, &B.lay, mode, s->fuse_prisms);
        B.bytes = static_cast<uint32_t>(B.host.size());
        B.built = true;
    }
    return B;
}

int ensure_events(rtiow_scene* s, uint32_t n) {
    while (s->ws->events.size() < n) {
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        s->ws->events.push_back(e);
    }
    return RTIOW_OK;
}

int ensure_uploaded(rtiow_scene* s, rtiow_scene::Blob& B) {
    if (B.d) return RTIOW_OK;
    DevBuf& buf = s->ws->scene_blob[&B - s->blobs];  // lives in the (cached) workspace: no cudaMalloc per scene
    CK(buf.reserve(B.bytes));
    CK(cudaMemcpyAsync(buf.p, B.host.data(), B.bytes, cudaMemcpyHostToDevice, s->ws->stream));
    CK(cudaStreamSynchronize(s->ws->stream));  // B.host is pageable; renders may run on other streams
    B.d = static_cast<unsigned char*>(buf.p);
    return RTIOW_OK;
}

int check_render_args(rtiow_scene* s, const rtiow_camera_t* cam, uint32_t nx, uint32_t ny, uint32_t ns, uint32_t r0,
                      uint32_t r1, const void* out, uint32_t step = 1, uint32_t band = 1) {
    if (!s || !cam || !out) return set_err(RTIOW_ERR_INVALID_ARG, "null argument");
    if (nx == 0 || ny == 0 || ns == 0) return set_err(RTIOW_ERR_INVALID_ARG, "nx, ny and ns must be non-zero");
    if (r0 >= r1 || r1 > ny) return set_err(RTIOW_ERR_INVALID_ARG, "row range must satisfy row_begin < row_end <= ny");
    if (step == 0 || band == 0 || band > step) return set_err(RTIOW_ERR_INVALID_ARG, "need 1 <= band_rows <= row_step");
    if (nx > 65535u || ny > 65535u) return set_err(RTIOW_ERR_INVALID_ARG, "image too large (65535 x 65535 at most)");
    if (!(cam->time0 < cam->time1))  // rand's gen_range asserts low < high (camera.rs:55)
        return set_err(RTIOW_ERR_INVALID_ARG, "Uniform::sample_single called with low >= high (camera exposure)");
    return RTIOW_OK;
}

// Where a multi-GPU render puts its pixels and which tiles of the frame are its own (rtiow_b200_render_peers).
struct PeerTarget {
    rtiow::FoldDst dst{};          // every rank's frame (the buffer of this epoch)
    uint32_t tile_first = 0, tile_step = 1;  // KParams: rank, number of ranks
};
constexpr unsigned long long kPeerTimeoutNs = 60ull * 1000ull * 1000ull * 1000ull;

// Enqueue a full render of rows [r0, r1) into device buffer d_out (rgb floats, packed) — or, with `peers`, of this
// rank's tiles of those rows into every rank's frame — and/or d_samples.
int enqueue_render(rtiow_scene* s, const rtiow_camera_t* cam, uint32_t nx, uint32_t ny, uint32_t ns, uint64_t seed,
                   uint32_t r0, uint32_t r1, float* d_out, float4* d_samples, cudaStream_t stream, uint32_t step = 1,
                   uint32_t band = 1, const PeerTarget* peers = nullptr) {
    CK(cudaSetDevice(s->device));
    // bands of `band` rows starting at r0, r0 + step, ..., clipped to r1
    const uint32_t n_full = (r1 - r0) / step, rest = (r1 - r0) - n_full * step;
    const uint32_t n_rows = n_full * band + std::min(rest, band);
    // 8x4-pixel tiles of the row block; this launch's share of them (all, or every tile_step-th for a rank of a peer render)
    const uint32_t tiles_x = (nx + rtiow::kTileW - 1u) / rtiow::kTileW;
    const uint32_t tiles_all = tiles_x * ((n_rows + rtiow::kTileH - 1u) / rtiow::kTileH);
    const uint32_t tile_first = peers ? peers->tile_first : 0u, tile_step = peers ? peers->tile_step : 1u;
    if (tile_first >= tiles_all) return RTIOW_OK;  // more ranks than tiles: nothing to render
    const uint32_t n_groups = (tiles_all - tile_first + tile_step - 1u) / tile_step;
    const uint64_t npix64 = static_cast<uint64_t>(n_groups) * 32u;  // staging slots (tile-major; edge tiles are padded)
    if (npix64 >= (1ull << 32)) return set_err(RTIOW_ERR_INVALID_ARG, "image too large");
    const uint32_t npix = static_cast<uint32_t>(npix64);
    uint64_t real_pix = static_cast<uint64_t>(n_rows) * nx;  // (statistics)
    if (peers) {
        real_pix = npix64;
        const uint32_t edge_w = nx % rtiow::kTileW, edge_h = n_rows % rtiow::kTileH;
        if (edge_w || edge_h)  // the tiles on the right / bottom edge are partly outside the frame
            for (uint32_t t = tile_first; t < tiles_all; t += tile_step) {
                const uint32_t w = (t % tiles_x == tiles_x - 1u && edge_w) ? edge_w : rtiow::kTileW;
                const uint32_t h = (t / tiles_x == tiles_all / tiles_x - 1u && edge_h) ? edge_h : rtiow::kTileH;
                real_pix -= 32u - w * h;
            }
    }
    Workspace& W = *s->ws;
    CK(W.order_after_last_render(stream));
    // Per-sample staging budget.  Automatic: up to 56 GiB, at most 70 % of what is free — a B200 has 180 GB and a
    // pass boundary costs a kernel tail, so C3 (10 GB), C4 (51 GB) and one rank's share of C5 (15 GB) are single passes.
    // Which pipeline slot: alternate for frames whose staging is small (a kernel of a few milliseconds, where the ~0.15 ms
    // in which the last long paths finish is worth hiding behind the next render); big frames stay on slot 0.
    const bool pipelined = s->pipeline && npix64 * ns * 16 <= (4ull << 30);
    const uint32_t slot_id = pipelined ? (W.next_slot++ % W.n_slots) : 0u;
    Workspace::Slot& SL = W.slot[slot_id];
    uint64_t budget = static_cast<uint64_t>(s->staging_mib) << 20;
    if (s->staging_mib == 0 && npix64 * ns * 16 <= SL.staging.cap) {
        budget = SL.staging.cap;  // the whole frame fits what is already there: no need to ask the driver (cudaMemGetInfo is slow)
    } else if (s->staging_mib == 0) {
        size_t free_b = 0, total_b = 0;
        CK(cudaMemGetInfo(&free_b, &total_b));
        budget = std::min<uint64_t>(56ull << 30, (static_cast<uint64_t>(free_b) + SL.staging.cap) / 10 * 7);
        budget = std::max<uint64_t>(budget, 64ull << 20);
    }
    // passes of equal size (the last one is not a sliver)
    const uint32_t s_max = static_cast<uint32_t>(std::max<uint64_t>(1, std::min<uint64_t>(ns, budget / (npix64 * 16))));
    const uint32_t n_pass = (ns + s_max - 1) / s_max;
    const uint32_t s_pass = (ns + n_pass - 1) / n_pass;
    if (npix64 * s_pass * 16 > SL.staging.cap || npix64 * 16 > SL.accum.cap) {
        CK(cudaStreamSynchronize(SL.stream));  // the buffers are about to be replaced
        if (SL.fold_recorded) CK(cudaEventSynchronize(SL.fold_done));
    }
    CK(SL.staging.reserve(npix64 * s_pass * 16));
    CK(SL.accum.reserve(npix64 * 16));

    // ---- which blob: REINDEXED means the conservative-test tree when it fits in shared memory, else the
    // exact-test tree (smaller) when that fits, else the conservative-test tree from global memory
    const uint32_t smem_cap = static_cast<uint32_t>(s->max_smem_optin) - 1024u;
    rtiow::BlobMode mode = s->traversal == RTIOW_TRAVERSAL_REFERENCE_ORDER ? rtiow::kBlobReferenceOrder
                           : (s->traversal == RTIOW_TRAVERSAL_REINDEXED_EXACT ? rtiow::kBlobExact : rtiow::kBlobFast);
    // The conservative-test tree is the larger image (112-byte nodes).  Shared memory and the L1 that caches the per-lane
    // traversal stacks and register spills come out of the same 228 KB: a scene that fills it (final scene: 208 KB) runs
    // 30 % SLOWER with the cheaper box test than with the exact-test tree (140 KB) — measured, profiles/r02 — so the
    // conservative tree is used only while it leaves a third of the array to the L1.
    const uint32_t fast_cap = smem_cap / 3u * 2u;
    if (s->traversal == RTIOW_TRAVERSAL_REINDEXED && !s->force_global && blob_of(s, rtiow::kBlobFast).bytes > fast_cap &&
        blob_of(s, rtiow::kBlobExact).bytes <= smem_cap)
        mode = rtiow::kBlobExact;
    rtiow_scene::Blob& B = blob_of(s, mode);
    if (int rc = ensure_uploaded(s, B)) return rc;
    const bool fast = mode == rtiow::kBlobFast && B.lay.n_nodes != 0;
    const bool smem = B.bytes <= smem_cap && !s->force_global;
    // 0 = automatic: one CTA per SM; 768 threads (80 registers) for the general kernel, 1024 (64 registers) for the
    // spheres-only one, which needs no more
    // kernel profile: the smallest compiled feature mask that covers the scene (the re-indexed subtrees add SF_ACCEL)
    const uint32_t needs = s->features | (B.lay.n_accel ? static_cast<uint32_t>(rtiow::SF_ACCEL) : 0u) |
                           (B.lay.n_ordered ? static_cast<uint32_t>(rtiow::SF_ORDERED) : 0u);
    uint32_t profile = 0;
    if (s->specialise) {
        if (!s->has_frames && (needs & ~rtiow::kFeatSpheres) == 0u) profile = 1;
        else if (!s->has_frames && (needs & ~rtiow::kFeatRects) == 0u) profile = 2;
        else if (smem && (needs & ~rtiow::kFeatLean) == 0u) profile = 3;  // the general kernel minus what this scene cannot contain
    }
    const uint32_t threads = s->cta_threads ? s->cta_threads : ((profile == 1u || profile == 2u) ? 1024u : 768u);
    KernelVariant var = profile == 3u ? rtiow::pick_lean_smem(s->has_frames, fast, threads)
                                      : (smem ? rtiow::pick_plain_smem(s->has_frames, fast, profile, threads)
                                              : rtiow::pick_plain_global(s->has_frames, fast, profile, threads));
    if (!var.fn && profile == 3u) {  // no lean instantiation for this CTA size: the general kernel renders the same image
        profile = 0;
        var = rtiow::pick_plain_smem(s->has_frames, fast, profile, threads);
    }
    if (!var.fn) return set_err(RTIOW_ERR_INVALID_ARG, "no kernel instantiation for this cta_threads");
    // strips of tiles (KParams): as few tiles per strip as the table's room in shared memory allows
    const size_t order_room = static_cast<size_t>(smem_cap) - (smem ? (B.bytes + 127u) / 128u * 128u : 0u);
    uint32_t order_shift = 0, n_strips = 0;
    rtiow::plan_strips(n_groups, static_cast<uint32_t>(std::min<size_t>(rtiow::kMaxStrips, std::max<size_t>(order_room / sizeof(uint32_t), 1))),
                       &order_shift, &n_strips);
    const uint32_t n_tile_slots = n_strips << order_shift;
    const size_t dyn_smem = (smem ? (B.bytes + 127u) / 128u * 128u : 0u) + static_cast<size_t>(n_strips) * sizeof(uint32_t);
    int num_regs = 0;
    CK(kernel_info(reinterpret_cast<const void*>(var.fn), s->device, s->max_smem_optin, &num_regs));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, var.fn, var.threads, dyn_smem));
    if (occ < 1) return set_err(RTIOW_ERR_CUDA, "render kernel does not fit on an SM");
    if (s->ctas_per_sm) occ = std::min<int>(occ, static_cast<int>(s->ctas_per_sm));
    uint32_t grid = static_cast<uint32_t>(s->sm_count) * static_cast<uint32_t>(occ);
    const uint32_t warps_per_cta = static_cast<uint32_t>(var.threads) / 32u;
    // no more CTAs than there is work for: a warp takes at least one unit (one sample of one tile)
    const uint64_t min_units = static_cast<uint64_t>(n_groups) * std::min(s_pass, ns);
    grid = static_cast<uint32_t>(std::max<uint64_t>(1, std::min<uint64_t>(grid, (min_units + warps_per_cta - 1) / warps_per_cta)));

    const rtiow::TileMap tile_map{nx, n_rows, tiles_x, tile_first, tile_step};
    KParams P{};
    P.blob = B.d;
    P.blob_bytes = B.bytes;
    P.off_nodes = B.lay.off_nodes; P.off_frames = B.lay.off_frames; P.off_ops = B.lay.off_ops; P.off_mats = B.lay.off_mats;
    P.off_tex = B.lay.off_tex; P.off_pvecs = B.lay.off_pvecs; P.off_pperm = B.lay.off_pperm;
    P.off_fnodes = B.lay.off_fnodes;
    std::memcpy(P.cam, cam, sizeof(float) * 21);
    P.nx = nx; P.ny = ny; P.row_begin = r0; P.n_rows = n_rows; P.row_step = step; P.row_band = band;
    P.npix = npix; P.tiles_x = tiles_x; P.tile_first = tile_first; P.tile_step = tile_step;
    P.key0 = static_cast<uint32_t>(seed); P.key1 = static_cast<uint32_t>(seed >> 32);
    P.bg_kind = s->bg_kind;
    std::memcpy(P.bg0, s->bg0, 12); std::memcpy(P.bg1, s->bg1, 12);
    P.staging = static_cast<float4*>(SL.staging.p);
    // Scenes whose segments run through a lot of different code (media, wrapper frames, nested subtrees) are
    // bound by instruction fetch: ncu shows `no_instruction` as the top stall with the 99 % hot set at 35 KB against
    // a 32 KB L1.5 I-cache.  Two CTA barriers per round (before hit_top, before shading) keep the 24 warps in the
    // same phase, i.e. the same few KB of code: final scene +16 %; book-1 and Cornell, whose hot set fits, lose 8-30 %.
    // Named barriers over half the CTA (12 warps) wait a little less than __syncthreads and keep the locality.
    P.phase_sync = s->phase_sync >= 0 ? static_cast<uint32_t>(s->phase_sync) : (s->costly_segments ? 2u : 0u);
    P.phase_group = static_cast<uint32_t>(var.threads) / 32u;
    if (P.phase_sync != 0u && P.phase_group % 2u == 0u) P.phase_group /= 2u;  // two barrier groups per CTA: measured best
    if (s->phase_group_env) {
        const uint32_t g = s->phase_group_env, warps = static_cast<uint32_t>(var.threads) / 32u;
        if (warps % g == 0 && warps / g <= 15) P.phase_group = g;
    }
    // Idle lanes get new pixel-samples once `refill_thr` lanes of the warp wait: generating camera rays costs the
    // warp the same for 3 lanes as for 30, and rays started together stay coherent.  Waiting costs idle lane
    // iterations, which are expensive in scenes with media / wrapper frames and frequent where paths are short.
    // Measured (profiles/r01/sweep_v9_refill_threshold.log): book-1 best at 12, Cornell at 4, final at 1.
    P.refill_thr = s->refill_lanes ? s->refill_lanes : (s->costly_segments ? 1u : (s->bg_kind == RTIOW_BG_SKY_GRADIENT ? 12u : 4u));

    // Tile order.  The kernel ends when the last path ends, and a path of 51 segments takes ~0.15 ms however small the frame
    // is: per-CTA time stamps (make EXTRA=-DRT_CTA_TIMELINE=1) show the CTAs of an eighth of book-1 living 0.07-0.23 ms
    // beyond the moment the unit counter runs out, with the paths that started in the last sphere-covered tiles.  So the
    // fold notes every tile's longest path and sorts the tiles by it (one 6-bit radix pass, stable: ties stay bottom
    // first), and the slot's NEXT render of the same shape hands the tiles with long paths out first and ends with the
    // ones that had none.  A hint only: the image does not depend on the order, and a render that has no previous one
    // (or follows one of another shape) takes its tiles bottom rows first.
    const uint64_t shape[4] = {(static_cast<uint64_t>(nx) << 32) | ny, (static_cast<uint64_t>(r0) << 32) | r1,
                               (static_cast<uint64_t>(step) << 32) | band,
                               (static_cast<uint64_t>(tile_first) << 40) | (static_cast<uint64_t>(tile_step) << 16) | order_shift};
    const bool want_order = s->order_hint && n_groups >= 1024u && !d_samples && s->bottom_first;
    P.n_strips = n_strips; P.order_shift = order_shift;
    const bool have_order = want_order && SL.order_valid && std::equal(shape, shape + 4, SL.order_shape);
    // learning costs a sort behind the fold, and with the tail gone nothing hides it: an order is kept for 16 renders
    const bool learn_order = want_order && (!have_order || ++SL.order_age >= 16u);
    const size_t table_bytes = (rtiow::kOrderHeaderWords + static_cast<size_t>(n_strips)) * sizeof(uint32_t);
    auto table_of = [](DevBuf& b) { return static_cast<uint32_t*>(b.p) + rtiow::kOrderHeaderWords; };
    size_t sort_tmp_bytes = 0;
    if (learn_order)
        CK(cub::DeviceRadixSort::SortPairsDescending(nullptr, sort_tmp_bytes, static_cast<const uint32_t*>(nullptr),
                                                      static_cast<uint32_t*>(nullptr), static_cast<const uint32_t*>(nullptr),
                                                      static_cast<uint32_t*>(nullptr), static_cast<int>(n_strips), 0, 6, stream));
    bool order_usable = have_order;
    if (table_bytes > SL.tiles_in.cap || (learn_order && (table_bytes > SL.tile_order.cap || sort_tmp_bytes > SL.sort_tmp.cap))) {
        CK(cudaStreamSynchronize(SL.stream));  // the buffers are about to be replaced
        if (SL.fold_recorded) CK(cudaEventSynchronize(SL.fold_done));
        order_usable = false;
        SL.order_valid = false;
        SL.default_order_n = 0;
    }
    CK(SL.tiles_in.reserve(table_bytes));
    if (learn_order) {
        CK(SL.longest.reserve(table_bytes)); CK(SL.longest_sorted.reserve(table_bytes)); CK(SL.tile_order.reserve(table_bytes));
        CK(SL.sort_tmp.reserve(std::max<size_t>(sort_tmp_bytes, 16)));
    }
    // the default order, bottom tile first (also the sort's values); rewritten only when the number of tiles changes
    const uint32_t default_tag = n_strips * 2u + (s->bottom_first ? 1u : 0u);
    if (SL.default_order_n != default_tag) {
        if (SL.fold_recorded) CK(cudaStreamWaitEvent(SL.stream, SL.fold_done, 0));  // a sort may still be reading the old one
        rtiow::iota_kernel<<<(n_strips + 255u) / 256u, 256, 0, SL.stream>>>(table_of(SL.tiles_in), n_strips, s->bottom_first ? 0 : 1);
        CK(cudaGetLastError());
        SL.default_order_n = default_tag;
    }
    P.work_counter = static_cast<unsigned int*>(order_usable ? SL.tile_order.p : SL.tiles_in.p);  // counter, then the order
    // Work units (KParams): chunks of samples of a tile, sized from the work per resident warp (unit_plan.hpp: one size for
    // open scenes, big chunks and a tail of smaller ones for closed scenes).  Round 1 used ONE size, 1 sample per unit on 8
    // GPUs, and lost 17 % against T(1)/8 there, most of it per-unit overhead and incoherent refills.
    const uint64_t resident_warps = static_cast<uint64_t>(grid) * warps_per_cta;
    auto plan_units = [&](uint32_t s_count, KParams& K) {
        rtiow::plan_units(n_groups, n_tile_slots, s_count, s->sample_chunk, s->bg_kind == RTIOW_BG_SKY_GRADIENT, resident_warps, K);
    };
    // The megakernel runs on the slot's own stream, everything that consumes its samples (export, fold, the peers'
    // hand-shake) on the caller's.  Dependencies: a render needs the previous fold of ITS slot (the staging buffer and the
    // pass accumulator are reused) and nothing of the caller's stream — its inputs are the scene and this call's arguments;
    // a fold needs its render.  So with two slots render k+1 is released while render k is still running: one persistent CTA
    // fills an SM's register file, so k+1's CTAs start exactly where k's have finished — the time in which k's last long
    // paths (up to 51 segments each, ~0.15 ms on book-1 whatever the frame size) keep a few SMs busy is no longer idle time
    // on the others.  Same image: nothing about a render depends on when it runs.
    if (int rc = ensure_events(s, 4 * n_pass)) return rc;
    s->events_used = 0;
    CK(cudaMemsetAsync(W.d_segs, 0, sizeof(unsigned long long), stream));
    uint32_t launches = 0;
    for (uint32_t pass = 0; pass < n_pass; ++pass) {
        P.s_begin = pass * s_pass;
        P.s_count = std::min(s_pass, ns - P.s_begin);
        plan_units(P.s_count, P);
        if (SL.fold_recorded) CK(cudaStreamWaitEvent(SL.stream, SL.fold_done, 0));
        CK(cudaMemsetAsync(P.work_counter, 0, sizeof(unsigned int), SL.stream));
        cudaEvent_t tl[4] = {nullptr, nullptr, nullptr, nullptr};
        if (s->timeline_on && s->timeline.size() < 4096)
            for (cudaEvent_t& e : tl) {
                CK(cudaEventCreate(&e));
                s->timeline.push_back(e);
            }
        CK(cudaEventRecord(W.events[s->events_used++], SL.stream));
        if (tl[0]) CK(cudaEventRecord(tl[0], SL.stream));
#ifdef RT_CTA_TIMELINE
        {   // diagnostic build: six time stamps per CTA, kept until scene_destroy prints them
            std::vector<unsigned long long> init(6u * grid);
            for (uint32_t b = 0; b < grid; ++b) { init[6 * b + 2] = init[6 * b + 4] = ~0ull; }
            unsigned long long* d = nullptr;
            CK(cudaMalloc(reinterpret_cast<void**>(&d), init.size() * 8));
            CK(cudaMemcpy(d, init.data(), init.size() * 8, cudaMemcpyHostToDevice));
            s->cta_timelines.push_back({d, grid});
            P.cta_times = d;
        }
#endif
        var.fn<<<grid, var.threads, dyn_smem, SL.stream>>>(P);
        CK(cudaGetLastError());
        ++launches;
        CK(cudaEventRecord(W.events[s->events_used++], SL.stream));
        if (tl[1]) CK(cudaEventRecord(tl[1], SL.stream));
        CK(cudaEventRecord(SL.render_done, SL.stream));
        CK(cudaStreamWaitEvent(stream, SL.render_done, 0));
        CK(cudaEventRecord(W.events[s->events_used++], stream));
        if (tl[2]) CK(cudaEventRecord(tl[2], stream));
        if (d_samples) {
            const uint64_t n = npix64 * P.s_count;
            rtiow::export_samples_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(
                P.staging, d_samples, tile_map, npix, P.s_begin, P.s_count, ns);
            CK(cudaGetLastError());
            ++launches;
        }
        if (d_out || peers) {
            rtiow::FoldDst dst{};
            if (peers) {
                dst = peers->dst;
            } else {
                dst.p[0] = d_out;
                dst.n = 1u;
            }
            dst.map = tile_map;
            if (learn_order && pass == 0) CK(cudaMemsetAsync(SL.longest.p, 0, static_cast<size_t>(n_strips) * sizeof(uint32_t), stream));
            rtiow::fold_kernel<<<(npix + 255u) / 256u, 256, 0, stream>>>(
                P.staging, static_cast<float4*>(SL.accum.p), dst, npix, P.s_count, pass == 0, pass + 1 == n_pass,
                static_cast<float>(ns), W.d_segs, learn_order ? static_cast<uint32_t*>(SL.longest.p) : nullptr, n_strips, order_shift);
            CK(cudaGetLastError());
            ++launches;
            if (learn_order && pass + 1 == n_pass) {  // the order for this slot's next render (it waits for fold_done)
                CK(cub::DeviceRadixSort::SortPairsDescending(SL.sort_tmp.p, sort_tmp_bytes, static_cast<const uint32_t*>(SL.longest.p),
                                                              static_cast<uint32_t*>(SL.longest_sorted.p),
                                                              static_cast<const uint32_t*>(table_of(SL.tiles_in)),
                                                              table_of(SL.tile_order), static_cast<int>(n_strips), 0, 6, stream));
                ++launches;
                std::copy(shape, shape + 4, SL.order_shape);
                SL.order_valid = true;
                SL.order_age = 0;
            }
        }
        CK(cudaEventRecord(W.events[s->events_used++], stream));
        if (tl[3]) CK(cudaEventRecord(tl[3], stream));
        CK(cudaEventRecord(SL.fold_done, stream));
        SL.fold_recorded = true;
    }
    CK(W.mark_render_end(stream));
    s->stats = rtiow_stats_t{};
    s->stats.samples = real_pix * ns;
    s->stats.kernel_launches = launches;
    s->stats.passes = n_pass;
    s->stats.scene_in_smem = smem ? 1u : 0u;
    s->stats.scene_bytes = B.bytes;
    s->stats.accel_nodes = B.lay.n_nodes;
    s->stats.accel_subtrees = B.lay.n_accel;
    s->stats.grid = grid;
    s->stats.block = static_cast<uint32_t>(var.threads);
    s->stats.dyn_smem_bytes = static_cast<uint32_t>(dyn_smem);
    s->stats.regs_per_thread = static_cast<uint32_t>(num_regs);
    s->stats.kernel_profile = profile;
    s->stats.traversal = static_cast<uint32_t>(mode == rtiow::kBlobFast ? RTIOW_TRAVERSAL_REINDEXED
                                               : (mode == rtiow::kBlobExact ? RTIOW_TRAVERSAL_REINDEXED_EXACT : RTIOW_TRAVERSAL_REFERENCE_ORDER));
    return RTIOW_OK;
}

}  // namespace

extern "C" {

int rtiow_b200_abi_version(void) { return static_cast<int>(RTIOW_B200_ABI_VERSION); }
const char* rtiow_b200_build_flavour(void) {
#if RT_FAST_MATH
    return "fast: FMA contraction, approximate division / square root / transcendentals; within tolerance of the reference arithmetic, not bit-exact";
#else
    return "parity: no FMA contraction, IEEE division and square root, fixed transcendentals; decision-for-decision the reference's f32 arithmetic";
#endif
}
const char* rtiow_b200_last_error(void) { return g_err.c_str(); }

int rtiow_b200_scene_validate(const rtiow_scene_desc_t* d) {
    if (!d) return set_err(RTIOW_ERR_INVALID_ARG, "null argument");
    bool has_frames = false, uses_perlin = false;
    std::string msg;
    if (int rc = rtiow::validate_desc(d, &has_frames, &uses_perlin, &msg)) return set_err(rc, msg);
    return RTIOW_OK;
}

int rtiow_b200_scene_create(const rtiow_scene_desc_t* d, int device, rtiow_scene_t** out) {
    if (!d || !out) return set_err(RTIOW_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    bool has_frames = false, uses_perlin = false;
    {
        std::string msg;
        if (int rc = rtiow::validate_desc(d, &has_frames, &uses_perlin, &msg)) return set_err(rc, msg);
    }

    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return set_err(RTIOW_ERR_NO_DEVICE, std::string("no CUDA device (this library has no CPU fallback): ") +
                                                cudaGetErrorString(e));
    if (device < 0 || device >= n_dev) return set_err(RTIOW_ERR_INVALID_ARG, "device ordinal out of range");
    CK(cudaSetDevice(device));
    DeviceInfo info{};
    CK(device_info(device, &info));
    if (info.major != 10)
        return set_err(RTIOW_ERR_NO_DEVICE, std::string("device is sm_") + std::to_string(info.major) + std::to_string(info.minor) +
                                                "; this library carries sm_100a code only");

    auto s = new rtiow_scene();
    s->device = device;
    s->sm_count = info.sm_count;
    s->max_smem_optin = info.max_smem_optin;
    s->has_frames = has_frames;
    s->costly_segments = has_frames;
    for (uint32_t i = 0; i < d->n_items; ++i) s->costly_segments |= (d->items[i].a_w & 15u) == RTIOW_ITEM_MEDIUM;
    s->features = rtiow::scene_features(d);
    s->bg_kind = d->background_kind;
    std::memcpy(s->bg0, d->background_c0, 12);
    std::memcpy(s->bg1, d->background_c1, 12);

    auto fail = [&](cudaError_t ce, const char* what) {
        rtiow_b200_scene_destroy(s);
        return set_err(RTIOW_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(ce));
    };
    s->desc.assign(d);
    s->uses_perlin = uses_perlin;
    s->ws = ws_acquire(device, &e);
    if (!s->ws) return fail(e, "workspace");

    if (const char* env = std::getenv("RTIOW_B200_CTA_THREADS")) s->cta_threads = static_cast<uint32_t>(std::atoi(env));
    if (const char* env = std::getenv("RTIOW_B200_CTAS_PER_SM")) s->ctas_per_sm = static_cast<uint32_t>(std::atoi(env));
    if (const char* env = std::getenv("RTIOW_B200_STAGING_MIB")) s->staging_mib = static_cast<uint32_t>(std::atoi(env));
    if (const char* env = std::getenv("RTIOW_B200_FORCE_GLOBAL")) s->force_global = std::atoi(env) != 0;
    if (const char* env = std::getenv("RTIOW_B200_TRAVERSAL")) s->traversal = std::min(2, std::max(0, std::atoi(env)));
    if (const char* env = std::getenv("RTIOW_B200_SPECIALISE")) s->specialise = std::atoi(env) != 0;
    if (const char* env = std::getenv("RTIOW_B200_REFILL_LANES")) s->refill_lanes = static_cast<uint32_t>(std::min(32, std::max(0, std::atoi(env))));
    if (const char* env = std::getenv("RTIOW_B200_PIPELINE")) s->pipeline = std::atoi(env) != 0;
    if (const char* env = std::getenv("RTIOW_B200_FUSE_PRISMS")) s->fuse_prisms = std::atoi(env) != 0;
    if (const char* env = std::getenv("RTIOW_B200_UNIT_ORDER")) s->bottom_first = std::atoi(env) != 0;
    if (const char* env = std::getenv("RTIOW_B200_ORDER_HINT")) s->order_hint = std::atoi(env) != 0;
    if (const char* env = std::getenv("RTIOW_B200_PHASE_SYNC")) s->phase_sync = std::max(0, std::atoi(env));
    if (const char* env = std::getenv("RTIOW_B200_PHASE_GROUP")) s->phase_group_env = static_cast<uint32_t>(std::max(1, std::atoi(env)));
    if (const char* env = std::getenv("RTIOW_B200_TIMELINE")) s->timeline_on = std::atoi(env) != 0;
    if (const char* env = std::getenv("RTIOW_B200_SLOTS")) s->ws->n_slots = static_cast<uint32_t>(std::min(4, std::max(1, std::atoi(env))));
    if (const char* env = std::getenv("RTIOW_B200_SAMPLE_CHUNK")) s->sample_chunk = static_cast<uint32_t>(std::max(0, std::atoi(env)));
    {   // build + upload the blob of the selected traversal now, so that render calls only launch
        const rtiow::BlobMode mode = s->traversal == RTIOW_TRAVERSAL_REFERENCE_ORDER ? rtiow::kBlobReferenceOrder
                                     : (s->traversal == RTIOW_TRAVERSAL_REINDEXED_EXACT ? rtiow::kBlobExact : rtiow::kBlobFast);
        if (int rc = ensure_uploaded(s, blob_of(s, mode))) {
            rtiow_b200_scene_destroy(s);
            return rc;
        }
    }
    *out = s;
    return RTIOW_OK;
}

void rtiow_b200_release_cached_memory(void) {
    {
        std::lock_guard<std::mutex> lock(g_multi_mutex);
        for (MultiFrames& c : g_multi_cache) c.destroy();
        g_multi_cache.clear();
    }
    std::vector<Workspace*> drop;
    {
        std::lock_guard<std::mutex> lock(g_ws_mutex);
        drop.swap(g_ws_cache);
    }
    for (Workspace* w : drop) {
        w->destroy();
        delete w;
    }
}

void rtiow_b200_scene_destroy(rtiow_scene_t* s) {
    if (!s) return;
    cudaSetDevice(s->device);
#ifdef RT_CTA_TIMELINE
    {
        cudaDeviceSynchronize();
        unsigned long long t00 = 0;
        size_t li = 0;
        for (auto& rec : s->cta_timelines) {
            std::vector<unsigned long long> h(6u * rec.second);
            cudaMemcpy(h.data(), rec.first, h.size() * 8, cudaMemcpyDeviceToHost);
            cudaFree(rec.first);
            unsigned long long k0 = ~0ull;
            for (uint32_t b = 0; b < rec.second; ++b) k0 = std::min(k0, h[6 * b]);
            if (!t00) t00 = k0;
            // per stamp: min / mean / max over the CTAs, in microseconds since the first CTA of this launch started
            std::fprintf(stderr, "cta-timeline %3zu: launch at %10.1f us;", li++, (k0 - t00) * 1e-3);
            static const char* names[6] = {"start", "staged", "first-warp-dry", "last-warp-dry", "first-warp-exit", "last-warp-exit"};
            for (int k = 0; k < 6; ++k) {
                double mn = 1e30, mx = 0, sum = 0;
                for (uint32_t b = 0; b < rec.second; ++b) {
                    const double v = (h[6 * b + k] - k0) * 1e-3;
                    mn = std::min(mn, v); mx = std::max(mx, v); sum += v;
                }
                std::fprintf(stderr, " %s %.1f/%.1f/%.1f", names[k], mn, sum / rec.second, mx);
            }
            std::fprintf(stderr, "\n");
        }
    }
#endif
    if (!s->timeline.empty()) {  // RTIOW_B200_TIMELINE: when each launch ran, in ms since the first one
        cudaDeviceSynchronize();
        for (size_t i = 0; i + 4 <= s->timeline.size(); i += 4) {
            float t[4] = {0, 0, 0, 0};
            for (int k = 0; k < 4; ++k) cudaEventElapsedTime(&t[k], s->timeline[0], s->timeline[i + k]);
            std::fprintf(stderr, "timeline %4zu: render %9.4f .. %9.4f  fold %9.4f .. %9.4f\n", i / 4, t[0], t[1], t[2], t[3]);
        }
        for (cudaEvent_t e : s->timeline) cudaEventDestroy(e);
    }
    ws_release(s->ws);  // synchronises the scene's stream first; the blobs' device memory stays with the workspace
    delete s;
}

int rtiow_b200_set_tuning(rtiow_scene_t* s, uint32_t cta_threads, uint32_t ctas_per_sm, uint32_t staging_mib, int force_global) {
    if (!s) return set_err(RTIOW_ERR_INVALID_ARG, "null scene");
    if (cta_threads) {
        if (cta_threads != 256 && cta_threads != 512 && cta_threads != 768 && cta_threads != 1024)
            return set_err(RTIOW_ERR_INVALID_ARG, "cta_threads must be 256, 512, 768 or (specialised kernels) 1024");
    }
    s->cta_threads = cta_threads;
    s->ctas_per_sm = ctas_per_sm;
    s->staging_mib = staging_mib;
    s->force_global = force_global != 0;
    return RTIOW_OK;
}

int rtiow_b200_set_specialisation(rtiow_scene_t* s, int enable) {
    if (!s) return set_err(RTIOW_ERR_INVALID_ARG, "null scene");
    s->specialise = enable != 0;
    return RTIOW_OK;
}

int rtiow_b200_set_traversal(rtiow_scene_t* s, int mode) {
    if (!s) return set_err(RTIOW_ERR_INVALID_ARG, "null scene");
    if (mode != RTIOW_TRAVERSAL_REINDEXED && mode != RTIOW_TRAVERSAL_REFERENCE_ORDER && mode != RTIOW_TRAVERSAL_REINDEXED_EXACT)
        return set_err(RTIOW_ERR_INVALID_ARG, "unknown traversal mode");
    s->traversal = mode;
    return RTIOW_OK;
}

int rtiow_b200_render_rows_device(rtiow_scene_t* s, const rtiow_camera_t* cam, uint32_t nx, uint32_t ny, uint32_t ns,
                                  uint64_t seed, uint32_t r0, uint32_t r1, float* d_out, void* cuda_stream) {
    if (int rc = check_render_args(s, cam, nx, ny, ns, r0, r1, d_out)) return rc;
    return enqueue_render(s, cam, nx, ny, ns, seed, r0, r1, d_out, nullptr, static_cast<cudaStream_t>(cuda_stream));
}

int rtiow_b200_render_rows_strided_device(rtiow_scene_t* s, const rtiow_camera_t* cam, uint32_t nx, uint32_t ny, uint32_t ns,
                                          uint64_t seed, uint32_t r0, uint32_t r1, uint32_t step, uint32_t band, float* d_out,
                                          void* cuda_stream) {
    if (int rc = check_render_args(s, cam, nx, ny, ns, r0, r1, d_out, step, band)) return rc;
    return enqueue_render(s, cam, nx, ny, ns, seed, r0, r1, d_out, nullptr, static_cast<cudaStream_t>(cuda_stream), step, band);
}

int rtiow_b200_render_rows(rtiow_scene_t* s, const rtiow_camera_t* cam, uint32_t nx, uint32_t ny, uint32_t ns, uint64_t seed,
                           uint32_t r0, uint32_t r1, float* out_rows) {
    if (int rc = check_render_args(s, cam, nx, ny, ns, r0, r1, out_rows)) return rc;
    CK(cudaSetDevice(s->device));
    const size_t bytes = static_cast<size_t>(r1 - r0) * nx * 3 * sizeof(float);
    Workspace& W = *s->ws;
    CK(W.out.reserve(bytes));
    if (int rc = enqueue_render(s, cam, nx, ny, ns, seed, r0, r1, static_cast<float*>(W.out.p), nullptr, W.stream)) return rc;
    CK(W.copy_to_host(out_rows, W.out.p, bytes));
    return RTIOW_OK;
}

int rtiow_b200_render(rtiow_scene_t* s, const rtiow_camera_t* cam, uint32_t nx, uint32_t ny, uint32_t ns, uint64_t seed,
                      float* out_rgb) {
    return rtiow_b200_render_rows(s, cam, nx, ny, ns, seed, 0, ny, out_rgb);
}

int rtiow_b200_render_samples(rtiow_scene_t* s, const rtiow_camera_t* cam, uint32_t nx, uint32_t ny, uint32_t ns, uint64_t seed,
                              uint32_t r0, uint32_t r1, float* out_samples) {
    if (int rc = check_render_args(s, cam, nx, ny, ns, r0, r1, out_samples)) return rc;
    CK(cudaSetDevice(s->device));
    const size_t bytes = static_cast<size_t>(r1 - r0) * nx * ns * 4 * sizeof(float);
    Workspace& W = *s->ws;
    CK(W.samples.reserve(bytes));
    if (int rc = enqueue_render(s, cam, nx, ny, ns, seed, r0, r1, nullptr, static_cast<float4*>(W.samples.p), W.stream)) return rc;
    CK(W.copy_to_host(out_samples, W.samples.p, bytes));
    return RTIOW_OK;
}

int rtiow_b200_ppm_quantise(rtiow_scene_t* s, const float* linear, size_t n, uint8_t* out) {
    if (!s || !linear || !out || n == 0) return set_err(RTIOW_ERR_INVALID_ARG, "null argument");
    CK(cudaSetDevice(s->device));
    Workspace& W = *s->ws;
    CK(W.out.reserve(n * sizeof(float)));
    CK(W.samples.reserve(n));
    CK(cudaMemcpyAsync(W.out.p, linear, n * sizeof(float), cudaMemcpyHostToDevice, W.stream));
    rtiow::ppm_quantise_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, W.stream>>>(
        static_cast<const float*>(W.out.p), static_cast<unsigned char*>(W.samples.p), n);
    CK(cudaGetLastError());
    CK(W.copy_to_host(out, W.samples.p, n));
    return RTIOW_OK;
}

int rtiow_b200_ppm_quantise_device(rtiow_scene_t* s, const float* d_linear, size_t n, uint8_t* d_out, void* cuda_stream) {
    if (!s || !d_linear || !d_out || n == 0) return set_err(RTIOW_ERR_INVALID_ARG, "null argument");
    CK(cudaSetDevice(s->device));
    rtiow::ppm_quantise_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
        d_linear, d_out, n);
    CK(cudaGetLastError());
    return RTIOW_OK;
}

int rtiow_b200_render_ppm(rtiow_scene_t* s, const rtiow_camera_t* cam, uint32_t nx, uint32_t ny, uint32_t ns, uint64_t seed,
                          uint8_t* out_rgb8) {
    if (int rc = check_render_args(s, cam, nx, ny, ns, 0, ny, out_rgb8)) return rc;
    CK(cudaSetDevice(s->device));
    const size_t n = static_cast<size_t>(ny) * nx * 3;
    Workspace& W = *s->ws;
    CK(W.out.reserve(n * sizeof(float)));
    CK(W.samples.reserve(n));
    if (int rc = enqueue_render(s, cam, nx, ny, ns, seed, 0, ny, static_cast<float*>(W.out.p), nullptr, W.stream)) return rc;
    rtiow::ppm_quantise_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, W.stream>>>(
        static_cast<const float*>(W.out.p), static_cast<unsigned char*>(W.samples.p), n);
    CK(cudaGetLastError());
    CK(W.mark_render_end(W.stream));
    CK(W.copy_to_host(out_rgb8, W.samples.p, n));  // a quarter of the float frame's bytes over PCIe
    return RTIOW_OK;
}

}  // extern "C"

// One rank's two copies of the frame (epoch e is assembled in buffer e & 1, aux_kernels.cuh) plus its hand-shake flags, in
// ONE cudaMalloc allocation that the other ranks map (CUDA IPC between processes, peer access inside one):
// [frame 0: ny*nx*3 f32 | frame 1 | arrived[16] u32].
struct rtiow_peer_frame {
    int device = 0;
    uint32_t nx = 0, ny = 0, rank = 0, n_ranks = 1;
    size_t frame_bytes = 0, bytes = 0;
    unsigned char* base = nullptr;             // my allocation
    unsigned char* peer_base[rtiow::kMaxFoldDst] = {};  // everybody's (peer_base[rank] == base)
    bool opened_ipc[rtiow::kMaxFoldDst] = {};
    bool connected = false;
    unsigned int epoch = 0;
    unsigned int* timed_out = nullptr;         // pinned, mapped: the barrier kernel raises it, the host reads it without a sync
    cudaEvent_t done_ev = nullptr;             // single-process multi-GPU: end of this device's part

    float* frame(uint32_t q, unsigned int e) const { return reinterpret_cast<float*>(peer_base[q] + (e & 1u) * frame_bytes); }
    unsigned int* arrived(uint32_t q) const { return reinterpret_cast<unsigned int*>(peer_base[q] + 2u * frame_bytes); }
};

namespace {

struct PeerHandleWire {  // RTIOW_PEER_HANDLE_BYTES on the wire
    cudaIpcMemHandle_t ipc;   // 64 bytes
    uint64_t pid;
    uint64_t ptr;             // valid inside process `pid`
    uint64_t bytes;
    int32_t device;
    uint32_t nx, ny, rank, n_ranks;
    uint32_t magic;
};
static_assert(sizeof(PeerHandleWire) <= RTIOW_PEER_HANDLE_BYTES, "peer handle does not fit");
constexpr uint32_t kPeerMagic = 0x52543230u;

uint64_t my_pid();

// Renders rank `pf->rank`'s tiles of the frame into the frames listed in `dst_ranks` (a bit mask) and, if `handshake`,
// ends with the barrier that makes the frame whole on every rank.
int render_share(rtiow_scene* s, const rtiow_camera_t* cam, uint32_t nx, uint32_t ny, uint32_t ns, uint64_t seed,
                 rtiow_peer_frame* pf, uint32_t dst_mask, bool handshake, cudaStream_t stream) {
    if (!pf->connected && pf->n_ranks > 1) return set_err(RTIOW_ERR_INVALID_ARG, "peer frame is not connected");
    if (pf->nx != nx || pf->ny != ny) return set_err(RTIOW_ERR_INVALID_ARG, "peer frame has another size");
    if (s->device != pf->device) return set_err(RTIOW_ERR_INVALID_ARG, "scene and peer frame live on different devices");
    if (*pf->timed_out) return set_err(RTIOW_ERR_CUDA, "a peer did not arrive within the time-out in an earlier multi-GPU render");
    CK(cudaSetDevice(s->device));
    // The frame's 8x4-pixel tiles are dealt round-robin: rank r renders tiles r, r + G, ... of the whole frame (KParams).
    const uint32_t G = pf->n_ranks;
    PeerTarget T{};
    T.tile_first = pf->rank;
    T.tile_step = G;
    const unsigned int epoch = ++pf->epoch;
    rtiow::PeerFlags arrived{};
    for (uint32_t q = 0; q < G; ++q) {
        if (dst_mask & (1u << q)) T.dst.p[T.dst.n++] = pf->frame(q, epoch);
        arrived.p[q] = pf->arrived(q);
    }
    arrived.n = G;
    // No wait before the fold: it overwrites the peers' copies of frame epoch - 2, and this stream is already past the
    // barrier of epoch - 1, which every peer entered after its reads of that frame (aux_kernels.cuh).
    // (more ranks than tiles: nothing to render, but the barrier still runs)
    if (int rc = enqueue_render(s, cam, nx, ny, ns, seed, 0, ny, nullptr, nullptr, stream, 1, 1, &T)) return rc;
    if (handshake && G > 1) {  // my rows are in every frame -> tell everybody; the frame is whole once everybody has told me
        rtiow::peer_barrier_kernel<<<1, 32, 0, stream>>>(arrived, pf->rank, epoch, kPeerTimeoutNs, pf->timed_out);
        CK(cudaGetLastError());
        CK(s->ws->mark_render_end(stream));
    }
    return RTIOW_OK;
}

uint64_t my_pid() { return static_cast<uint64_t>(getpid()); }


}  // namespace

extern "C" {

int rtiow_b200_peer_frame_create(int device, uint32_t nx, uint32_t ny, uint32_t rank, uint32_t n_ranks, rtiow_peer_frame_t** out) {
    if (!out) return set_err(RTIOW_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    if (nx == 0 || ny == 0 || n_ranks == 0 || rank >= n_ranks || n_ranks > static_cast<uint32_t>(rtiow::kMaxFoldDst))
        return set_err(RTIOW_ERR_INVALID_ARG, "need nx, ny > 0 and rank < n_ranks <= 16");
    CK(cudaSetDevice(device));
    auto pf = new rtiow_peer_frame();
    pf->device = device; pf->nx = nx; pf->ny = ny; pf->rank = rank; pf->n_ranks = n_ranks;
    pf->frame_bytes = (static_cast<size_t>(nx) * ny * 3 * sizeof(float) + 255u) / 256u * 256u;
    pf->bytes = 2u * pf->frame_bytes + rtiow::kMaxFoldDst * sizeof(unsigned int);
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&pf->base), pf->bytes);  // plain cudaMalloc: exportable with CUDA IPC
    if (e == cudaSuccess) e = cudaMemset(pf->base, 0, pf->bytes);
    if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void**>(&pf->timed_out), sizeof(unsigned int), cudaHostAllocMapped);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&pf->done_ev, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        rtiow_b200_peer_frame_destroy(pf);
        return set_err(RTIOW_ERR_CUDA, std::string("peer frame: ") + cudaGetErrorString(e));
    }
    *pf->timed_out = 0u;
    pf->peer_base[rank] = pf->base;
    pf->connected = n_ranks == 1;
    *out = pf;
    return RTIOW_OK;
}

int rtiow_b200_peer_frame_export(rtiow_peer_frame_t* pf, uint8_t* handle) {
    if (!pf || !handle) return set_err(RTIOW_ERR_INVALID_ARG, "null argument");
    CK(cudaSetDevice(pf->device));
    PeerHandleWire w{};
    CK(cudaIpcGetMemHandle(&w.ipc, pf->base));
    w.pid = my_pid();
    w.ptr = reinterpret_cast<uint64_t>(pf->base);
    w.bytes = pf->bytes;
    w.device = pf->device;
    w.nx = pf->nx; w.ny = pf->ny; w.rank = pf->rank; w.n_ranks = pf->n_ranks;
    w.magic = kPeerMagic;
    std::memset(handle, 0, RTIOW_PEER_HANDLE_BYTES);
    std::memcpy(handle, &w, sizeof(w));
    return RTIOW_OK;
}

int rtiow_b200_peer_frame_connect(rtiow_peer_frame_t* pf, const uint8_t* handles) {
    if (!pf || !handles) return set_err(RTIOW_ERR_INVALID_ARG, "null argument");
    CK(cudaSetDevice(pf->device));
    for (uint32_t q = 0; q < pf->n_ranks; ++q) {
        PeerHandleWire w{};
        std::memcpy(&w, handles + static_cast<size_t>(q) * RTIOW_PEER_HANDLE_BYTES, sizeof(w));
        if (w.magic != kPeerMagic || w.rank != q || w.n_ranks != pf->n_ranks || w.nx != pf->nx || w.ny != pf->ny || w.bytes != pf->bytes)
            return set_err(RTIOW_ERR_INVALID_ARG, "peer handle " + std::to_string(q) + " does not describe rank " + std::to_string(q) +
                                                      " of this frame");
        if (q == pf->rank) continue;
        if (w.pid == my_pid()) {  // same process: the pointer itself, with peer access switched on
            if (w.device != pf->device) {
                int can = 0;
                CK(cudaDeviceCanAccessPeer(&can, pf->device, w.device));
                if (!can) return set_err(RTIOW_ERR_CUDA, "no peer access between devices " + std::to_string(pf->device) + " and " +
                                                             std::to_string(w.device));
                cudaError_t e = cudaDeviceEnablePeerAccess(w.device, 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                else if (e != cudaSuccess) return set_err(RTIOW_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
            }
            pf->peer_base[q] = reinterpret_cast<unsigned char*>(w.ptr);
        } else {
            void* p = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&p, w.ipc, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) return set_err(RTIOW_ERR_CUDA, std::string("cudaIpcOpenMemHandle (rank ") + std::to_string(q) + "): " +
                                                                     cudaGetErrorString(e));
            pf->peer_base[q] = static_cast<unsigned char*>(p);
            pf->opened_ipc[q] = true;
        }
    }
    pf->connected = true;
    return RTIOW_OK;
}

int rtiow_b200_peer_frame_ptr(rtiow_peer_frame_t* pf, float** d_frame) {
    if (!pf || !d_frame) return set_err(RTIOW_ERR_INVALID_ARG, "null argument");
    *d_frame = pf->frame(pf->rank, pf->epoch);
    return RTIOW_OK;
}

void rtiow_b200_peer_frame_destroy(rtiow_peer_frame_t* pf) {
    if (!pf) return;
    cudaSetDevice(pf->device);
    cudaDeviceSynchronize();
    for (uint32_t q = 0; q < pf->n_ranks; ++q)
        if (pf->opened_ipc[q]) cudaIpcCloseMemHandle(pf->peer_base[q]);
    if (pf->base) cudaFree(pf->base);
    if (pf->timed_out) cudaFreeHost(pf->timed_out);
    if (pf->done_ev) cudaEventDestroy(pf->done_ev);
    delete pf;
}

}  // extern "C"
namespace {
void MultiFrames::destroy() {
    for (rtiow_peer_frame_t* p : pf) rtiow_b200_peer_frame_destroy(p);
    pf.clear();
}
}  // namespace
extern "C" {

int rtiow_b200_render_peers(rtiow_scene_t* s, const rtiow_camera_t* cam, uint32_t nx, uint32_t ny, uint32_t ns, uint64_t seed,
                                 rtiow_peer_frame_t* pf, void* cuda_stream) {
    if (!pf) return set_err(RTIOW_ERR_INVALID_ARG, "null argument");
    if (int rc = check_render_args(s, cam, nx, ny, ns, 0, ny, pf)) return rc;
    return render_share(s, cam, nx, ny, ns, seed, pf, (1u << pf->n_ranks) - 1u, true, static_cast<cudaStream_t>(cuda_stream));
}

int rtiow_b200_render_multi(rtiow_scene_t* const* scenes, int ngpus, const rtiow_camera_t* cam, uint32_t nx, uint32_t ny, uint32_t ns,
                            uint64_t seed, float* out_rgb) {
    if (!scenes || ngpus < 1 || ngpus > rtiow::kMaxFoldDst) return set_err(RTIOW_ERR_INVALID_ARG, "need 1 <= ngpus <= 16 scene handles");
    for (int g = 0; g < ngpus; ++g)
        if (!scenes[g]) return set_err(RTIOW_ERR_INVALID_ARG, "null scene handle");
    if (int rc = check_render_args(scenes[0], cam, nx, ny, ns, 0, ny, out_rgb)) return rc;
    if (ngpus == 1) return rtiow_b200_render(scenes[0], cam, nx, ny, ns, seed, out_rgb);
    for (int g = 0; g < ngpus; ++g)
        for (int h = 0; h < g; ++h)
            if (scenes[g]->device == scenes[h]->device) return set_err(RTIOW_ERR_INVALID_ARG, "two scene handles on the same device");
    // the frame's 8x4-pixel tiles are dealt round-robin (render_share): every GPU gets the same mix of cheap and expensive pixels
    const uint32_t G = static_cast<uint32_t>(ngpus);
    // the devices' frames (GPU 0's is the one that gets assembled) are kept between calls of the same shape
    std::vector<int> devices;
    for (uint32_t g = 0; g < G; ++g) devices.push_back(scenes[g]->device);
    std::lock_guard<std::mutex> lock(g_multi_mutex);
    MultiFrames* mf = nullptr;
    for (MultiFrames& c : g_multi_cache)
        if (c.nx == nx && c.ny == ny && c.devices == devices) mf = &c;
    int rc = RTIOW_OK;
    if (!mf) {
        while (g_multi_cache.size() >= 2) {
            g_multi_cache.front().destroy();
            g_multi_cache.erase(g_multi_cache.begin());
        }
        MultiFrames fresh;
        fresh.nx = nx; fresh.ny = ny; fresh.devices = devices;
        fresh.pf.assign(G, nullptr);
        std::vector<uint8_t> handles(static_cast<size_t>(G) * RTIOW_PEER_HANDLE_BYTES);
        for (uint32_t g = 0; g < G && rc == RTIOW_OK; ++g) {
            rc = rtiow_b200_peer_frame_create(devices[g], nx, ny, g, G, &fresh.pf[g]);
            if (rc == RTIOW_OK) rc = rtiow_b200_peer_frame_export(fresh.pf[g], handles.data() + static_cast<size_t>(g) * RTIOW_PEER_HANDLE_BYTES);
        }
        for (uint32_t g = 0; g < G && rc == RTIOW_OK; ++g) rc = rtiow_b200_peer_frame_connect(fresh.pf[g], handles.data());
        if (rc != RTIOW_OK) {
            const std::string keep = g_err;
            fresh.destroy();
            g_err = keep;
            return rc;
        }
        g_multi_cache.push_back(fresh);
        mf = &g_multi_cache.back();
    }
    std::vector<rtiow_peer_frame_t*>& pf = mf->pf;
    // every GPU renders its tiles and its fold stores them straight into GPU 0's frame (the only consumer here); one host
    // thread drives all of them, the devices run concurrently
    for (uint32_t g = 0; g < G && rc == RTIOW_OK; ++g) {
        rc = render_share(scenes[g], cam, nx, ny, ns, seed, pf[g], 1u, false, scenes[g]->ws->stream);
        if (rc == RTIOW_OK && cudaEventRecord(pf[g]->done_ev, scenes[g]->ws->stream) != cudaSuccess) rc = set_err(RTIOW_ERR_CUDA, "cudaEventRecord");
        if (rc == RTIOW_OK && scenes[g]->ws->mark_render_end(scenes[g]->ws->stream) != cudaSuccess) rc = set_err(RTIOW_ERR_CUDA, "cudaEventRecord");
    }
    if (rc == RTIOW_OK) {
        Workspace& W0 = *scenes[0]->ws;
        cudaError_t e = cudaSetDevice(scenes[0]->device);
        for (uint32_t g = 1; g < G && e == cudaSuccess; ++g) e = cudaStreamWaitEvent(W0.stream, pf[g]->done_ev, 0);
        if (e == cudaSuccess) e = W0.copy_to_host(out_rgb, pf[0]->frame(0, pf[0]->epoch), static_cast<size_t>(nx) * ny * 3 * sizeof(float));
        if (e != cudaSuccess) rc = set_err(RTIOW_ERR_CUDA, std::string("render_multi gather: ") + cudaGetErrorString(e));
    }
    return rc;
}

int rtiow_b200_get_stats(rtiow_scene_t* s, rtiow_stats_t* out) {
    if (!s || !out) return set_err(RTIOW_ERR_INVALID_ARG, "null argument");
    CK(cudaSetDevice(s->device));
    CK(cudaDeviceSynchronize());
    double trace = 0, fold = 0;
    for (uint32_t i = 0; i + 4 <= s->events_used; i += 4) {  // per pass: render start/end, fold start/end
        float a = 0, b = 0;
        CK(cudaEventElapsedTime(&a, s->ws->events[i], s->ws->events[i + 1]));
        CK(cudaEventElapsedTime(&b, s->ws->events[i + 2], s->ws->events[i + 3]));
        trace += a;
        fold += b;
    }
    s->stats.trace_ms = trace;
    s->stats.reduce_ms = fold;
    unsigned long long segs = 0;
    CK(cudaMemcpy(&segs, s->ws->d_segs, sizeof(segs), cudaMemcpyDeviceToHost));
    s->stats.segments = segs;
    *out = s->stats;
    return RTIOW_OK;
}

}  // extern "C"
