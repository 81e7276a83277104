// C ABI of the B200 render path (include/rtiow_b200.h): scene validation + upload, the
// persistent megakernel launch per sample pass, the sample fold, statistics.
// No CPU fallback: every entry point fails with RTIOW_ERR_NO_DEVICE / RTIOW_ERR_CUDA if the
// device path is unavailable.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../../include/rtiow_b200.h"
#include "../device/render_kernel.cuh"
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
    DevBuf staging, accum, out, samples;
    unsigned int* d_counter = nullptr;
    unsigned long long* d_segs = nullptr;
    cudaStream_t stream = nullptr;
    std::vector<cudaEvent_t> events;  // [start, trace_end, fold_end] per pass
    // two pinned bounce buffers: device -> pinned -> caller's pageable memory, pipelined
    static constexpr size_t kBounce = 4u << 20;
    void* pinned[2] = {nullptr, nullptr};
    cudaEvent_t pin_ev[2] = {nullptr, nullptr};

    cudaError_t init(int dev) {
        device = dev;
        cudaError_t e;
        if ((e = cudaMalloc(reinterpret_cast<void**>(&d_counter), sizeof(unsigned int))) != cudaSuccess) return e;
        if ((e = cudaMalloc(reinterpret_cast<void**>(&d_segs), sizeof(unsigned long long))) != cudaSuccess) return e;
        if ((e = cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking)) != cudaSuccess) return e;
        return cudaSuccess;
    }
    void destroy() {
        cudaSetDevice(device);
        if (stream) cudaStreamSynchronize(stream);
        staging.release(); accum.release(); out.release(); samples.release();
        if (d_counter) cudaFree(d_counter);
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

std::mutex g_ws_mutex;
std::vector<Workspace*> g_ws_cache;
constexpr size_t kWsCachePerDevice = 2;

Workspace* ws_acquire(int device, cudaError_t* err) {
    {
        std::lock_guard<std::mutex> lock(g_ws_mutex);
        for (size_t i = 0; i < g_ws_cache.size(); ++i)
            if (g_ws_cache[i]->device == device) {
                Workspace* w = g_ws_cache[i];
                g_ws_cache.erase(g_ws_cache.begin() + static_cast<std::ptrdiff_t>(i));
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
    cudaStreamSynchronize(w->stream);
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

struct rtiow_scene {
    int device = 0;
    int sm_count = 0;
    int max_smem_optin = 0;
    // [0] = re-indexed Bvh subtrees (default), [1] = the plain reference-order stream
    struct Blob {
        std::vector<unsigned char> host;  // uploaded on first use
        unsigned char* d = nullptr;
        uint32_t bytes = 0;
        rtiow::BlobLayout lay{};
    } blobs[2];
    int traversal = 0;
    bool has_frames = false;
    uint32_t bg_kind = 0;
    float bg0[3] = {0, 0, 0}, bg1[3] = {0, 0, 0};

    Workspace* ws = nullptr;
    uint32_t events_used = 0;

    // tuning
    uint32_t cta_threads = 0, ctas_per_sm = 0, staging_mib = 2048, sample_chunk = 0;
    bool force_global = false;

    // last render
    rtiow_stats_t stats{};
};

namespace {

using rtiow::KParams;

typedef void (*kernel_fn)(const KParams);

struct Variant {
    kernel_fn fn;
    int threads;
};

template <bool S, bool F>
Variant pick_threads(uint32_t threads) {
    switch (threads) {
        case 128: return {rtiow::render_kernel<S, F, 128, 1>, 128};
        case 256: return {rtiow::render_kernel<S, F, 256, 1>, 256};
        case 512: return {rtiow::render_kernel<S, F, 512, 1>, 512};
        case 1024: return {rtiow::render_kernel<S, F, 1024, 1>, 1024};
        default: return {rtiow::render_kernel<S, F, 768, 1>, 768};
    }
}

Variant pick_variant(bool smem, bool frames, uint32_t threads) {
    if (smem) return frames ? pick_threads<true, true>(threads) : pick_threads<true, false>(threads);
    return frames ? pick_threads<false, true>(threads) : pick_threads<false, false>(threads);
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
    CK(cudaMalloc(reinterpret_cast<void**>(&B.d), B.bytes));
    CK(cudaMemcpy(B.d, B.host.data(), B.bytes, cudaMemcpyHostToDevice));
    return RTIOW_OK;
}

int check_render_args(rtiow_scene* s, const rtiow_camera_t* cam, uint32_t nx, uint32_t ny, uint32_t ns, uint32_t r0,
                      uint32_t r1, const void* out, uint32_t step = 1) {
    if (!s || !cam || !out) return set_err(RTIOW_ERR_INVALID_ARG, "null argument");
    if (nx == 0 || ny == 0 || ns == 0) return set_err(RTIOW_ERR_INVALID_ARG, "nx, ny and ns must be non-zero");
    if (r0 >= r1 || r1 > ny) return set_err(RTIOW_ERR_INVALID_ARG, "row range must satisfy row_begin < row_end <= ny");
    if (step == 0) return set_err(RTIOW_ERR_INVALID_ARG, "row_step must be non-zero");
    if (static_cast<uint64_t>(nx) * ny >= (1ull << 32)) return set_err(RTIOW_ERR_INVALID_ARG, "image too large");
    if (!(cam->time0 < cam->time1))  // rand's gen_range asserts low < high (camera.rs:55)
        return set_err(RTIOW_ERR_INVALID_ARG, "Uniform::sample_single called with low >= high (camera exposure)");
    return RTIOW_OK;
}

// Enqueue a full render of rows [r0, r1) into device buffer d_out (rgb floats) and/or d_samples.
int enqueue_render(rtiow_scene* s, const rtiow_camera_t* cam, uint32_t nx, uint32_t ny, uint32_t ns, uint64_t seed,
                   uint32_t r0, uint32_t r1, float* d_out, float4* d_samples, cudaStream_t stream, uint32_t step = 1) {
    CK(cudaSetDevice(s->device));
    const uint32_t n_rows = (r1 - r0 + step - 1) / step;  // rows r0, r0 + step, ... below r1
    const uint64_t npix64 = static_cast<uint64_t>(n_rows) * nx;
    const uint32_t npix = static_cast<uint32_t>(npix64);
    const uint64_t budget = static_cast<uint64_t>(s->staging_mib) << 20;
    uint32_t s_pass = static_cast<uint32_t>(std::max<uint64_t>(1, std::min<uint64_t>(ns, budget / (npix64 * 16))));
    const uint32_t n_pass = (ns + s_pass - 1) / s_pass;
    Workspace& W = *s->ws;
    CK(W.staging.reserve(npix64 * s_pass * 16));
    CK(W.accum.reserve(npix64 * 16));

    rtiow_scene::Blob& B = s->blobs[s->traversal];
    if (int rc = ensure_uploaded(s, B)) return rc;
    const bool fits = B.bytes + 1024u <= static_cast<uint32_t>(s->max_smem_optin);
    const bool smem = fits && !s->force_global;
    // 0 = automatic: 768 threads (80 registers) per CTA, one CTA per SM; scenes with wrapper frames keep 512
    const uint32_t threads = s->cta_threads ? s->cta_threads : (s->has_frames ? 512u : 768u);
    const Variant var = pick_variant(smem, s->has_frames, threads);
    const size_t dyn_smem = smem ? B.bytes : 0;
    CK(cudaFuncSetAttribute(var.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(dyn_smem)));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, var.fn, var.threads, dyn_smem));
    if (occ < 1) return set_err(RTIOW_ERR_CUDA, "render kernel does not fit on an SM");
    if (s->ctas_per_sm) occ = std::min<int>(occ, static_cast<int>(s->ctas_per_sm));
    const uint32_t n_groups = (npix + 31u) / 32u;
    uint32_t grid = static_cast<uint32_t>(s->sm_count) * static_cast<uint32_t>(occ);
    const uint32_t warps_per_cta = static_cast<uint32_t>(var.threads) / 32u;
    grid = std::max(1u, std::min(grid, (n_groups + warps_per_cta - 1) / warps_per_cta));
    cudaFuncAttributes fa{};
    CK(cudaFuncGetAttributes(&fa, var.fn));

    KParams P{};
    P.blob = B.d;
    P.blob_bytes = B.bytes;
    P.off_nodes = B.lay.off_nodes; P.off_frames = B.lay.off_frames; P.off_ops = B.lay.off_ops; P.off_mats = B.lay.off_mats;
    P.off_tex = B.lay.off_tex; P.off_pvecs = B.lay.off_pvecs; P.off_pperm = B.lay.off_pperm;
    std::memcpy(P.cam, cam, sizeof(float) * 21);
    P.nx = nx; P.ny = ny; P.row_begin = r0; P.n_rows = n_rows; P.row_step = step;
    P.npix = npix; P.n_groups = n_groups;
    P.key0 = static_cast<uint32_t>(seed); P.key1 = static_cast<uint32_t>(seed >> 32);
    P.bg_kind = s->bg_kind;
    std::memcpy(P.bg0, s->bg0, 12); std::memcpy(P.bg1, s->bg1, 12);
    P.staging = static_cast<float4*>(W.staging.p);
    P.work_counter = W.d_counter;

    const uint32_t chunk_pref = s->sample_chunk ? s->sample_chunk : 8u;
    if (int rc = ensure_events(s, 1 + 2 * n_pass)) return rc;
    s->events_used = 0;
    CK(cudaMemsetAsync(W.d_segs, 0, sizeof(unsigned long long), stream));
    CK(cudaEventRecord(W.events[s->events_used++], stream));
    uint32_t launches = 0;
    for (uint32_t pass = 0; pass < n_pass; ++pass) {
        P.s_begin = pass * s_pass;
        P.s_count = std::min(s_pass, ns - P.s_begin);
        P.s_chunk = std::min(chunk_pref, P.s_count);
        P.n_chunks = (P.s_count + P.s_chunk - 1) / P.s_chunk;
        if (static_cast<uint64_t>(n_groups) * P.n_chunks >= (1ull << 32)) {  // keep the unit counter in 32 bits
            P.s_chunk = P.s_count;
            P.n_chunks = 1;
        }
        P.n_units = n_groups * P.n_chunks;
        CK(cudaMemsetAsync(W.d_counter, 0, sizeof(unsigned int), stream));
        var.fn<<<grid, var.threads, dyn_smem, stream>>>(P);
        CK(cudaGetLastError());
        ++launches;
        CK(cudaEventRecord(W.events[s->events_used++], stream));
        if (d_samples) {
            const uint64_t n = npix64 * P.s_count;
            rtiow::export_samples_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(
                P.staging, d_samples, npix, P.s_begin, P.s_count, ns);
            CK(cudaGetLastError());
            ++launches;
        }
        if (d_out) {
            rtiow::fold_kernel<<<(npix + 255u) / 256u, 256, 0, stream>>>(
                P.staging, static_cast<float4*>(W.accum.p), d_out, npix, P.s_count, pass == 0, pass + 1 == n_pass,
                static_cast<float>(ns), W.d_segs);
            CK(cudaGetLastError());
            ++launches;
        }
        CK(cudaEventRecord(W.events[s->events_used++], stream));
    }
    s->stats = rtiow_stats_t{};
    s->stats.samples = npix64 * ns;
    s->stats.kernel_launches = launches;
    s->stats.passes = n_pass;
    s->stats.scene_in_smem = smem ? 1u : 0u;
    s->stats.scene_bytes = B.bytes;
    s->stats.accel_nodes = B.lay.n_nodes;
    s->stats.accel_subtrees = B.lay.n_accel;
    s->stats.grid = grid;
    s->stats.block = static_cast<uint32_t>(var.threads);
    s->stats.dyn_smem_bytes = static_cast<uint32_t>(dyn_smem);
    s->stats.regs_per_thread = static_cast<uint32_t>(fa.numRegs);
    return RTIOW_OK;
}

}  // namespace

extern "C" {

int rtiow_b200_abi_version(void) { return static_cast<int>(RTIOW_B200_ABI_VERSION); }
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
    cudaDeviceProp prop{};
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return set_err(RTIOW_ERR_NO_DEVICE, std::string("device is sm_") + std::to_string(prop.major) + std::to_string(prop.minor) +
                                                "; this library carries sm_100a code only");

    auto s = new rtiow_scene();
    s->device = device;
    s->sm_count = prop.multiProcessorCount;
    s->max_smem_optin = static_cast<int>(prop.sharedMemPerBlockOptin);
    s->has_frames = has_frames;
    s->bg_kind = d->background_kind;
    std::memcpy(s->bg0, d->background_c0, 12);
    std::memcpy(s->bg1, d->background_c1, 12);

    auto fail = [&](cudaError_t ce, const char* what) {
        rtiow_b200_scene_destroy(s);
        return set_err(RTIOW_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(ce));
    };
    // ---- build the device blobs: items | accel nodes | frames | ops | materials | textures | perlin
    for (int v = 0; v < 2; ++v) {
        rtiow_scene::Blob& B = s->blobs[v];
        B.host = rtiow::build_blob(d, uses_perlin, &B.lay, v == 0);
        B.bytes = static_cast<uint32_t>(B.host.size());
    }
    s->ws = ws_acquire(device, &e);
    if (!s->ws) return fail(e, "workspace");

    if (const char* env = std::getenv("RTIOW_B200_CTA_THREADS")) s->cta_threads = static_cast<uint32_t>(std::atoi(env));
    if (const char* env = std::getenv("RTIOW_B200_CTAS_PER_SM")) s->ctas_per_sm = static_cast<uint32_t>(std::atoi(env));
    if (const char* env = std::getenv("RTIOW_B200_STAGING_MIB")) s->staging_mib = static_cast<uint32_t>(std::atoi(env));
    if (const char* env = std::getenv("RTIOW_B200_FORCE_GLOBAL")) s->force_global = std::atoi(env) != 0;
    if (const char* env = std::getenv("RTIOW_B200_TRAVERSAL")) s->traversal = std::atoi(env) != 0 ? 1 : 0;
    if (const char* env = std::getenv("RTIOW_B200_SAMPLE_CHUNK")) s->sample_chunk = static_cast<uint32_t>(std::max(0, std::atoi(env)));
    if (int rc = ensure_uploaded(s, s->blobs[s->traversal])) {
        rtiow_b200_scene_destroy(s);
        return rc;
    }
    *out = s;
    return RTIOW_OK;
}

void rtiow_b200_release_cached_memory(void) {
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
    ws_release(s->ws);  // synchronises the scene's stream first
    for (auto& B : s->blobs)
        if (B.d) cudaFree(B.d);
    delete s;
}

int rtiow_b200_set_tuning(rtiow_scene_t* s, uint32_t cta_threads, uint32_t ctas_per_sm, uint32_t staging_mib, int force_global) {
    if (!s) return set_err(RTIOW_ERR_INVALID_ARG, "null scene");
    if (cta_threads) {
        if (cta_threads != 128 && cta_threads != 256 && cta_threads != 512 && cta_threads != 768 && cta_threads != 1024)
            return set_err(RTIOW_ERR_INVALID_ARG, "cta_threads must be 128, 256, 512, 768 or 1024");
    }
    s->cta_threads = cta_threads;
    s->ctas_per_sm = ctas_per_sm;
    if (staging_mib) s->staging_mib = staging_mib;
    s->force_global = force_global != 0;
    return RTIOW_OK;
}

int rtiow_b200_set_traversal(rtiow_scene_t* s, int mode) {
    if (!s) return set_err(RTIOW_ERR_INVALID_ARG, "null scene");
    if (mode != RTIOW_TRAVERSAL_REINDEXED && mode != RTIOW_TRAVERSAL_REFERENCE_ORDER)
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
                                          uint64_t seed, uint32_t r0, uint32_t r1, uint32_t step, float* d_out,
                                          void* cuda_stream) {
    if (int rc = check_render_args(s, cam, nx, ny, ns, r0, r1, d_out, step)) return rc;
    return enqueue_render(s, cam, nx, ny, ns, seed, r0, r1, d_out, nullptr, static_cast<cudaStream_t>(cuda_stream), step);
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

int rtiow_b200_get_stats(rtiow_scene_t* s, rtiow_stats_t* out) {
    if (!s || !out) return set_err(RTIOW_ERR_INVALID_ARG, "null argument");
    CK(cudaSetDevice(s->device));
    CK(cudaDeviceSynchronize());
    double trace = 0, fold = 0;
    for (uint32_t i = 0; i + 2 < s->events_used; i += 2) {  // [start, (trace_end, fold_end) per pass]
        float a = 0, b = 0;
        CK(cudaEventElapsedTime(&a, s->ws->events[i], s->ws->events[i + 1]));
        CK(cudaEventElapsedTime(&b, s->ws->events[i + 1], s->ws->events[i + 2]));
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
