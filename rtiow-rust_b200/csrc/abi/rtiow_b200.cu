// C ABI of the B200 render path (include/rtiow_b200.h): scene validation + upload, the
// persistent megakernel launch per sample pass, the sample fold, statistics.
// No CPU fallback: every entry point fails with RTIOW_ERR_NO_DEVICE / RTIOW_ERR_CUDA if the
// device path is unavailable.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
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

}  // namespace

struct rtiow_scene {
    int device = 0;
    int sm_count = 0;
    int max_smem_optin = 0;
    // [0] = re-indexed Bvh subtrees (default), [1] = the plain reference-order stream
    struct Blob {
        unsigned char* d = nullptr;
        uint32_t bytes = 0;
        rtiow::BlobLayout lay{};
    } blobs[2];
    int traversal = 0;
    bool has_frames = false;
    uint32_t bg_kind = 0;
    float bg0[3] = {0, 0, 0}, bg1[3] = {0, 0, 0};

    DevBuf staging, accum, out, samples;
    unsigned int* d_counter = nullptr;
    unsigned long long* d_segs = nullptr;
    cudaStream_t own_stream = nullptr;
    std::vector<cudaEvent_t> events;  // [start, trace_end, fold_end] per pass
    uint32_t events_used = 0;

    // tuning
    uint32_t cta_threads = 256, ctas_per_sm = 0, staging_mib = 2048;
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
        case 512: return {rtiow::render_kernel<S, F, 512, 1>, 512};
        case 768: return {rtiow::render_kernel<S, F, 768, 1>, 768};
        case 1024: return {rtiow::render_kernel<S, F, 1024, 1>, 1024};
        default: return {rtiow::render_kernel<S, F, 256, 1>, 256};
    }
}

Variant pick_variant(bool smem, bool frames, uint32_t threads) {
    if (smem) return frames ? pick_threads<true, true>(threads) : pick_threads<true, false>(threads);
    return frames ? pick_threads<false, true>(threads) : pick_threads<false, false>(threads);
}

int ensure_events(rtiow_scene* s, uint32_t n) {
    while (s->events.size() < n) {
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        s->events.push_back(e);
    }
    return RTIOW_OK;
}

int check_render_args(rtiow_scene* s, const rtiow_camera_t* cam, uint32_t nx, uint32_t ny, uint32_t ns, uint32_t r0,
                      uint32_t r1, const void* out) {
    if (!s || !cam || !out) return set_err(RTIOW_ERR_INVALID_ARG, "null argument");
    if (nx == 0 || ny == 0 || ns == 0) return set_err(RTIOW_ERR_INVALID_ARG, "nx, ny and ns must be non-zero");
    if (r0 >= r1 || r1 > ny) return set_err(RTIOW_ERR_INVALID_ARG, "row range must satisfy row_begin < row_end <= ny");
    if (static_cast<uint64_t>(nx) * ny >= (1ull << 32)) return set_err(RTIOW_ERR_INVALID_ARG, "image too large");
    if (!(cam->time0 < cam->time1))  // rand's gen_range asserts low < high (camera.rs:55)
        return set_err(RTIOW_ERR_INVALID_ARG, "Uniform::sample_single called with low >= high (camera exposure)");
    return RTIOW_OK;
}

// Enqueue a full render of rows [r0, r1) into device buffer d_out (rgb floats) and/or d_samples.
int enqueue_render(rtiow_scene* s, const rtiow_camera_t* cam, uint32_t nx, uint32_t ny, uint32_t ns, uint64_t seed,
                   uint32_t r0, uint32_t r1, float* d_out, float4* d_samples, cudaStream_t stream) {
    CK(cudaSetDevice(s->device));
    const uint32_t n_rows = r1 - r0;
    const uint64_t npix64 = static_cast<uint64_t>(n_rows) * nx;
    const uint32_t npix = static_cast<uint32_t>(npix64);
    const uint64_t budget = static_cast<uint64_t>(s->staging_mib) << 20;
    uint32_t s_pass = static_cast<uint32_t>(std::max<uint64_t>(1, std::min<uint64_t>(ns, budget / (npix64 * 16))));
    const uint32_t n_pass = (ns + s_pass - 1) / s_pass;
    CK(s->staging.reserve(npix64 * s_pass * 16));
    CK(s->accum.reserve(npix64 * 16));

    const rtiow_scene::Blob& B = s->blobs[s->traversal];
    const bool fits = B.bytes + 1024u <= static_cast<uint32_t>(s->max_smem_optin);
    const bool smem = fits && !s->force_global;
    const Variant var = pick_variant(smem, s->has_frames, s->cta_threads);
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
    P.nx = nx; P.ny = ny; P.row_begin = r0; P.n_rows = n_rows;
    P.npix = npix; P.n_groups = n_groups;
    P.key0 = static_cast<uint32_t>(seed); P.key1 = static_cast<uint32_t>(seed >> 32);
    P.bg_kind = s->bg_kind;
    std::memcpy(P.bg0, s->bg0, 12); std::memcpy(P.bg1, s->bg1, 12);
    P.staging = static_cast<float4*>(s->staging.p);
    P.work_counter = s->d_counter;

    if (int rc = ensure_events(s, 1 + 2 * n_pass)) return rc;
    s->events_used = 0;
    CK(cudaMemsetAsync(s->d_segs, 0, sizeof(unsigned long long), stream));
    CK(cudaEventRecord(s->events[s->events_used++], stream));
    uint32_t launches = 0;
    for (uint32_t pass = 0; pass < n_pass; ++pass) {
        P.s_begin = pass * s_pass;
        P.s_count = std::min(s_pass, ns - P.s_begin);
        CK(cudaMemsetAsync(s->d_counter, 0, sizeof(unsigned int), stream));
        var.fn<<<grid, var.threads, dyn_smem, stream>>>(P);
        CK(cudaGetLastError());
        ++launches;
        CK(cudaEventRecord(s->events[s->events_used++], stream));
        if (d_samples) {
            const uint64_t n = npix64 * P.s_count;
            rtiow::export_samples_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(
                P.staging, d_samples, npix, P.s_begin, P.s_count, ns);
            CK(cudaGetLastError());
            ++launches;
        }
        if (d_out) {
            rtiow::fold_kernel<<<(npix + 255u) / 256u, 256, 0, stream>>>(
                P.staging, static_cast<float4*>(s->accum.p), d_out, npix, P.s_count, pass == 0, pass + 1 == n_pass,
                static_cast<float>(ns), s->d_segs);
            CK(cudaGetLastError());
            ++launches;
        }
        CK(cudaEventRecord(s->events[s->events_used++], stream));
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
        const std::vector<unsigned char> blob = rtiow::build_blob(d, uses_perlin, &B.lay, v == 0);
        B.bytes = static_cast<uint32_t>(blob.size());
        if ((e = cudaMalloc(reinterpret_cast<void**>(&B.d), B.bytes)) != cudaSuccess) return fail(e, "cudaMalloc(blob)");
        if ((e = cudaMemcpy(B.d, blob.data(), B.bytes, cudaMemcpyHostToDevice)) != cudaSuccess) return fail(e, "cudaMemcpy(blob)");
    }
    if ((e = cudaMalloc(reinterpret_cast<void**>(&s->d_counter), sizeof(unsigned int))) != cudaSuccess) return fail(e, "cudaMalloc");
    if ((e = cudaMalloc(reinterpret_cast<void**>(&s->d_segs), sizeof(unsigned long long))) != cudaSuccess) return fail(e, "cudaMalloc");
    if ((e = cudaStreamCreateWithFlags(&s->own_stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "cudaStreamCreate");

    if (const char* env = std::getenv("RTIOW_B200_CTA_THREADS")) s->cta_threads = static_cast<uint32_t>(std::atoi(env));
    if (const char* env = std::getenv("RTIOW_B200_CTAS_PER_SM")) s->ctas_per_sm = static_cast<uint32_t>(std::atoi(env));
    if (const char* env = std::getenv("RTIOW_B200_STAGING_MIB")) s->staging_mib = static_cast<uint32_t>(std::atoi(env));
    if (const char* env = std::getenv("RTIOW_B200_FORCE_GLOBAL")) s->force_global = std::atoi(env) != 0;
    if (const char* env = std::getenv("RTIOW_B200_TRAVERSAL")) s->traversal = std::atoi(env) != 0 ? 1 : 0;
    *out = s;
    return RTIOW_OK;
}

void rtiow_b200_scene_destroy(rtiow_scene_t* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    cudaDeviceSynchronize();
    for (auto& B : s->blobs)
        if (B.d) cudaFree(B.d);
    if (s->d_counter) cudaFree(s->d_counter);
    if (s->d_segs) cudaFree(s->d_segs);
    s->staging.release(); s->accum.release(); s->out.release(); s->samples.release();
    for (cudaEvent_t ev : s->events) cudaEventDestroy(ev);
    if (s->own_stream) cudaStreamDestroy(s->own_stream);
    delete s;
}

int rtiow_b200_set_tuning(rtiow_scene_t* s, uint32_t cta_threads, uint32_t ctas_per_sm, uint32_t staging_mib, int force_global) {
    if (!s) return set_err(RTIOW_ERR_INVALID_ARG, "null scene");
    if (cta_threads) {
        if (cta_threads != 128 && cta_threads != 256 && cta_threads != 512 && cta_threads != 768 && cta_threads != 1024)
            return set_err(RTIOW_ERR_INVALID_ARG, "cta_threads must be 128, 256, 512, 768 or 1024");
        s->cta_threads = cta_threads;
    }
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

int rtiow_b200_render_rows(rtiow_scene_t* s, const rtiow_camera_t* cam, uint32_t nx, uint32_t ny, uint32_t ns, uint64_t seed,
                           uint32_t r0, uint32_t r1, float* out_rows) {
    if (int rc = check_render_args(s, cam, nx, ny, ns, r0, r1, out_rows)) return rc;
    CK(cudaSetDevice(s->device));
    const size_t bytes = static_cast<size_t>(r1 - r0) * nx * 3 * sizeof(float);
    CK(s->out.reserve(bytes));
    if (int rc = enqueue_render(s, cam, nx, ny, ns, seed, r0, r1, static_cast<float*>(s->out.p), nullptr, s->own_stream)) return rc;
    CK(cudaMemcpyAsync(out_rows, s->out.p, bytes, cudaMemcpyDeviceToHost, s->own_stream));
    CK(cudaStreamSynchronize(s->own_stream));
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
    CK(s->samples.reserve(bytes));
    if (int rc = enqueue_render(s, cam, nx, ny, ns, seed, r0, r1, nullptr, static_cast<float4*>(s->samples.p), s->own_stream)) return rc;
    CK(cudaMemcpyAsync(out_samples, s->samples.p, bytes, cudaMemcpyDeviceToHost, s->own_stream));
    CK(cudaStreamSynchronize(s->own_stream));
    return RTIOW_OK;
}

int rtiow_b200_ppm_quantise(rtiow_scene_t* s, const float* linear, size_t n, uint8_t* out) {
    if (!s || !linear || !out || n == 0) return set_err(RTIOW_ERR_INVALID_ARG, "null argument");
    CK(cudaSetDevice(s->device));
    CK(s->out.reserve(n * sizeof(float)));
    CK(s->samples.reserve(n));
    CK(cudaMemcpyAsync(s->out.p, linear, n * sizeof(float), cudaMemcpyHostToDevice, s->own_stream));
    rtiow::ppm_quantise_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s->own_stream>>>(
        static_cast<const float*>(s->out.p), static_cast<unsigned char*>(s->samples.p), n);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, s->samples.p, n, cudaMemcpyDeviceToHost, s->own_stream));
    CK(cudaStreamSynchronize(s->own_stream));
    return RTIOW_OK;
}

int rtiow_b200_get_stats(rtiow_scene_t* s, rtiow_stats_t* out) {
    if (!s || !out) return set_err(RTIOW_ERR_INVALID_ARG, "null argument");
    CK(cudaSetDevice(s->device));
    CK(cudaDeviceSynchronize());
    double trace = 0, fold = 0;
    for (uint32_t i = 0; i + 2 < s->events_used; i += 2) {  // [start, (trace_end, fold_end) per pass]
        float a = 0, b = 0;
        CK(cudaEventElapsedTime(&a, s->events[i], s->events[i + 1]));
        CK(cudaEventElapsedTime(&b, s->events[i + 1], s->events[i + 2]));
        trace += a;
        fold += b;
    }
    s->stats.trace_ms = trace;
    s->stats.reduce_ms = fold;
    unsigned long long segs = 0;
    CK(cudaMemcpy(&segs, s->d_segs, sizeof(segs), cudaMemcpyDeviceToHost));
    s->stats.segments = segs;
    *out = s->stats;
    return RTIOW_OK;
}

}  // extern "C"
