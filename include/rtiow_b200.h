/* rtiow_b200.h — C ABI of the B200 render path for cbiffle/rtiow-rust.
 *
 * This is the drop-in boundary for ONE path of the reference crate:
 *     par_cast / cast  ->  Camera::get_ray -> color -> World::hit_top -> Object::hit
 *                      ->  Material::scatter / emitted -> Texture -> perlin
 * (reference: src/lib.rs:363-397).  Everything below that call runs on the device in one
 * sm_100a megakernel; nothing here falls back to the CPU.
 *
 * The reference has no FFI today (#![forbid(unsafe_code)], src/lib.rs:1): the seam is the Rust
 * signature
 *     pub fn par_cast(nx: usize, ny: usize, ns: usize, camera: &Camera, world: impl World) -> Image
 * and this header is what a `rtiow-b200-sys` crate binds (INTEGRATION.md shows the Rust side).
 * The host language flattens its scene (Box<dyn Object> tree, Bvh, Material, Texture) into the
 * plain arrays of rtiow_scene_desc_t and hands over its Camera field for field.
 *
 * Conventions: plain pointers and sizes only; all input buffers are caller-owned HOST memory and
 * are copied during the call; every function returns 0 on success or an RTIOW_ERR_* code and
 * leaves a message readable with rtiow_b200_last_error() (thread-local).  A scene handle may be
 * used from one host thread at a time and has ONE set of device work buffers: renders of the same scene
 * enqueued on different CUDA streams (the *_device entry points) are ordered one after the other by
 * the library, they do not overlap; rtiow_b200_scene_destroy waits for the scene's last render.
 */
#ifndef RTIOW_B200_H
#define RTIOW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTIOW_B200_ABI_VERSION 2u

/* The shared library exports these entry points and nothing else. */
#if defined(__GNUC__)
#define RTIOW_API __attribute__((visibility("default")))
#else
#define RTIOW_API
#endif

enum {
    RTIOW_OK = 0,
    RTIOW_ERR_INVALID_ARG = 1,  /* null pointer, zero size, low >= high exposure ...             */
    RTIOW_ERR_INVALID_SCENE = 2,/* index out of range, backward skip link, missing END ...       */
    RTIOW_ERR_UNSUPPORTED = 3,  /* a nesting the flattened format cannot express                 */
    RTIOW_ERR_CUDA = 4,         /* any CUDA runtime failure (message carries cudaGetErrorString) */
    RTIOW_ERR_NO_DEVICE = 5     /* no sm_100 device: there is deliberately no CPU fallback       */
};

/* ---------------------------------------------------------------------------------------------
 * Traversal stream.  The object tree is stored as ONE array of 32-byte items in the reference's
 * own depth-first, left-first visiting order (Bvh::hit src/bvh.rs:85-120, And::hit
 * src/object.rs:396-410, the list loop src/lib.rs:40-45 all visit children in order with a
 * shrinking t_range.end, keeping the last accepted hit).  A box test that fails jumps to `skip`;
 * everything else advances by one item, so traversal needs no stack and cannot loop.
 *
 * word a_w = kind | (payload << 4)            word b_w = material | (flags << 24)
 * ------------------------------------------------------------------------------------------- */
enum {
    RTIOW_ITEM_END = 0,       /* end of stream (must be the last item)                               */
    RTIOW_ITEM_BBOX = 1,      /* Bvh node box (src/aabb.rs:18-29): a=min b=max payload=skip index    */
    RTIOW_ITEM_SPHERE = 2,    /* src/object.rs:82-111: a[0]=radius, b=inline Translate offset,
                                 payload=frame                                                         */
    RTIOW_ITEM_RECT = 3,      /* src/object.rs:183-218: a={k,range0.start,range0.end}
                                 b={range1.start,range1.end,-} payload=frame                           */
    RTIOW_ITEM_MEDIUM = 4,    /* ConstantMedium<O> src/object.rs:533-575: a[0]=density, a[1]=bits(medium id),
                                 a[2]=bits(index of the item after its boundary), payload=frame of the
                                 medium.  The items in between are the boundary object O flattened like any
                                 other object, in the medium's frame: a sphere, the rects of a rect_prism,
                                 or the BBOX items (skip links inside the run) and primitives of a Bvh.
                                 They are asked twice, with f32::MIN.. and hit1.t+0.0001.. (object.rs:551-553),
                                 and not visited on their own.                                          */
    RTIOW_ITEM_SET_FRAME = 5, /* following BBOX items live in frame `payload`                          */
    RTIOW_ITEM_PRISM = 6      /* rect_prism(p0, p1, material) src/object.rs:420-473 as one record:
                                 a=p0 b=p1 payload=frame; defined as its six Rect items in And order
                                 (z=p1.z, y=p1.y, x=p1.x, then FlipNormals z=p0.z, y=p0.y, x=p0.x).
                                 Optional: the library fuses such runs of six RECT items itself.       */
};
enum {
    RTIOW_FLAG_HAS_OFFSET = 1u, /* SPHERE: innermost Translate folded into b[0..2] (object.rs:267-283) */
    RTIOW_FLAG_FLIP = 2u,       /* innermost FlipNormals folded in (object.rs:241-253); PRISM: around all six  */
    RTIOW_FLAG_AXIS_SHIFT = 2u, /* RECT: (flags >> 2) & 3 = orthogonal axis 0=X 1=Y 2=Z                */
};
typedef struct rtiow_item_t {
    float a[3];
    uint32_t a_w;
    float b[3];
    uint32_t b_w;
} rtiow_item_t;

/* A frame is the chain of ray-transforming wrappers between the world and an item, applied
 * outermost first to the ray and innermost first to the hit record, exactly as the nested
 * Object::hit calls do.  Frame 0 is the world (no ops). */
enum {
    RTIOW_OP_TRANSLATE = 0,   /* v = offset            src/object.rs:267-283 */
    RTIOW_OP_SCALE = 1,       /* v = factor            src/object.rs:301-319 */
    RTIOW_OP_ROTATE_Y = 2,    /* v = {sin, cos, -}     src/object.rs:341-370 */
    RTIOW_OP_LINEAR_MOVE = 3, /* v = motion per time   src/object.rs:496-512 */
    RTIOW_OP_FLIP = 4         /* FlipNormals           src/object.rs:241-253 */
};
typedef struct rtiow_xform_op_t {
    uint32_t kind;
    float v[3];
} rtiow_xform_op_t;
typedef struct rtiow_frame_t {
    uint32_t first_op;
    uint32_t n_ops;
} rtiow_frame_t;

/* src/material.rs:11-40 */
enum {
    RTIOW_MAT_LAMBERTIAN = 0,    /* tex = albedo texture                    */
    RTIOW_MAT_METAL = 1,         /* albedo[3], param = fuzz                 */
    RTIOW_MAT_DIELECTRIC = 2,    /* param = ref_idx                         */
    RTIOW_MAT_DIFFUSE_LIGHT = 3, /* tex = emission texture, param = brightness */
    RTIOW_MAT_ISOTROPIC = 4      /* tex = albedo texture                    */
};
typedef struct rtiow_material_t {
    uint32_t kind;
    uint32_t tex;
    float albedo[3];
    float param;
    uint32_t reserved[2];
} rtiow_material_t;

/* src/texture.rs:6-26 as data instead of closures */
enum {
    RTIOW_TEX_CONSTANT = 0, /* color                                  */
    RTIOW_TEX_CHECKER = 1,  /* child0 (s >= 0) / child1 (s < 0)       */
    RTIOW_TEX_PERLIN = 2    /* Vec3::from(turb(scale * p, 7))         */
};
typedef struct rtiow_texture_t {
    uint32_t kind;
    float color[3];
    float scale;
    uint32_t child0;
    uint32_t child1;
    uint32_t reserved;
} rtiow_texture_t;

/* What color() returns when a ray escapes.  BLACK is HEAD's behaviour (src/lib.rs:100);
 * SKY_GRADIENT is the book-1 sky the README image was rendered with:
 * strength * ((1-t)*c0 + t*c1), t = 0.5*(unit(dir).y + 1). */
enum { RTIOW_BG_BLACK = 0, RTIOW_BG_SKY_GRADIENT = 1 };

typedef struct rtiow_scene_desc_t {
    uint32_t abi_version; /* RTIOW_B200_ABI_VERSION */
    uint32_t n_items;
    const rtiow_item_t* items;
    uint32_t n_frames;
    uint32_t n_ops;
    const rtiow_frame_t* frames; /* frames[0] must be {0,0} */
    const rtiow_xform_op_t* ops;
    uint32_t n_materials;
    uint32_t n_textures;
    const rtiow_material_t* materials;
    const rtiow_texture_t* textures;
    const float* perlin_vecs;   /* 256*3 floats (src/perlin.rs:15-21); may be NULL if no Perlin texture */
    const uint8_t* perlin_perm; /* 3*256 bytes: PERM_X, PERM_Y, PERM_Z (src/perlin.rs:5-13)             */
    uint32_t background_kind;
    float background_c0[3];
    float background_c1[3];
} rtiow_scene_desc_t;

/* src/camera.rs:6-15, field for field (exposure: Range<f32> -> time0, time1). */
typedef struct rtiow_camera_t {
    float origin[3];
    float lower_left_corner[3];
    float horizontal[3];
    float vertical[3];
    float u[3];
    float v[3];
    float lens_radius;
    float time0;
    float time1;
} rtiow_camera_t;

typedef struct rtiow_stats_t {
    double trace_ms;        /* device time of the path-tracing kernel(s) of the last render (CUDA events) */
    double reduce_ms;       /* device time of the sample-fold kernel(s)                                   */
    uint64_t samples;       /* pixel-samples traced by the last render                                   */
    uint64_t segments;      /* hit_top calls (path segments) of the last render, counted on device        */
    uint32_t kernel_launches; /* kernels launched by the last render                                      */
    uint32_t passes;
    uint32_t scene_in_smem; /* 1 if the scene blob was staged into shared memory                          */
    uint32_t scene_bytes;
    uint32_t grid, block, dyn_smem_bytes, regs_per_thread;
    uint32_t accel_nodes;    /* nodes of the library's own index over the scene's Bvh subtrees (0 = none) */
    uint32_t accel_subtrees; /* how many Bvh subtrees were re-indexed                                      */
    uint32_t traversal;      /* RTIOW_TRAVERSAL_*: how the last render walked Bvh subtrees                 */
    uint32_t kernel_profile; /* which compilation of the megakernel the last render used: 0 general, 1 spheres-only
                                scenes, 2 rect-list scenes, 3 general minus the features none of the reference's own
                                scenes builds (rtiow_b200_set_specialisation)                               */
} rtiow_stats_t;

typedef struct rtiow_scene rtiow_scene_t;

RTIOW_API int rtiow_b200_abi_version(void);
/* Which arithmetic this build of the library uses: "parity: ..." (the default build: every f32 operation as the
 * reference does it, results bit-identical to the oracle) or "fast: ..." (make FAST=1: fused multiply-add and
 * approximate division/sqrt; same algorithm and random numbers, image within a tolerance — mean |delta| <= 1e-3
 * linear, >= 99.5 % of 8-bit PPM values within +-1 at >= 50 spp). */
RTIOW_API const char* rtiow_b200_build_flavour(void);
RTIOW_API const char* rtiow_b200_last_error(void);

/* Checks `desc` exactly as rtiow_b200_scene_create does, without touching a GPU. */
RTIOW_API int rtiow_b200_scene_validate(const rtiow_scene_desc_t* desc);

/* Validates `desc`, copies it to `device` (CUDA ordinal) and returns a handle. */
RTIOW_API int rtiow_b200_scene_create(const rtiow_scene_desc_t* desc, int device, rtiow_scene_t** out);
RTIOW_API void rtiow_b200_scene_destroy(rtiow_scene_t* scene);
/* Destroyed scenes leave their device work buffers in a small per-device cache for the next
 * scene_create; this frees them. */
RTIOW_API void rtiow_b200_release_cached_memory(void);

/* par_cast (src/lib.rs:363-376) with an explicit seed: out_rgb receives ny*nx*3 floats, row 0 =
 * TOP scanline (lib.rs:326-330), linear (pre-gamma), already divided by ns (lib.rs:374) — i.e.
 * exactly `Image`.  Samples of a pixel are summed left to right from 0 in sample order
 * (src/vec3.rs:195-203).  out_rgb is HOST memory. */
RTIOW_API int rtiow_b200_render(rtiow_scene_t* scene, const rtiow_camera_t* camera, uint32_t nx, uint32_t ny, uint32_t ns,
                      uint64_t seed, float* out_rgb);

/* Rows [row_begin, row_end) of the same image (row 0 = top): the unit of multi-GPU sharding.
 * The result is bit-identical to the corresponding rows of rtiow_b200_render. */
RTIOW_API int rtiow_b200_render_rows(rtiow_scene_t* scene, const rtiow_camera_t* camera, uint32_t nx, uint32_t ny, uint32_t ns,
                           uint64_t seed, uint32_t row_begin, uint32_t row_end, float* out_rows);

/* Same, but `d_out_rows` is DEVICE memory on the scene's device and the work is only enqueued on
 * `cuda_stream` (a cudaStream_t, NULL = default stream); nothing is synchronised. */
RTIOW_API int rtiow_b200_render_rows_device(rtiow_scene_t* scene, const rtiow_camera_t* camera, uint32_t nx, uint32_t ny,
                                  uint32_t ns, uint64_t seed, uint32_t row_begin, uint32_t row_end,
                                  float* d_out_rows, void* cuda_stream);

/* Bands of `band_rows` consecutive rows starting at row_begin, row_begin + row_step, ... (clipped to
 * row_end), packed in that order; 1 <= band_rows <= row_step.  With band_rows = B, row_begin = rank * B
 * and row_step = n_ranks * B every GPU gets an equal mix of cheap (sky) and expensive scanlines while
 * its 8x4-pixel work tiles stay compact.  Each row is bit-identical to the same row of rtiow_b200_render. */
RTIOW_API int rtiow_b200_render_rows_strided_device(rtiow_scene_t* scene, const rtiow_camera_t* camera, uint32_t nx, uint32_t ny,
                                          uint32_t ns, uint64_t seed, uint32_t row_begin, uint32_t row_end,
                                          uint32_t row_step, uint32_t band_rows, float* d_out_rows, void* cuda_stream);

/* ---------------------------------------------------------------------------------------------
 * Several GPUs.  par_cast parallelises over scanlines (src/lib.rs:324-332); here the frame's 8x4-pixel tiles are
 * dealt round-robin to the GPUs (GPU g: tiles g, g + G, ... in row-major tile order, so every GPU's share is spread
 * evenly over the image: shares of whole scanline bands differed by 6 % in work on book-1, tiles by well under 1 %),
 * every GPU renders its tiles with its own copy of the scene, and the kernel that finishes a pixel (the in-order
 * sample fold) stores it straight into the frame of every GPU that wants it, through NVLink peer pointers, at its
 * final position: the framebuffer exchange is fused into the fold, there is no all-gather and no de-interleave pass
 * behind it.  Every pixel is bit-identical to the same pixel of rtiow_b200_render, for any number of GPUs.
 *
 * One host process: rtiow_b200_render_multi is par_cast over `ngpus` scene handles (one per device,
 * created from the same descriptor).  One process per GPU (torchrun, MPI): each rank creates a peer
 * frame, the opaque handles are exchanged by whatever transport the host has, and
 * rtiow_b200_render_peers leaves the whole image in every rank's frame.
 * ------------------------------------------------------------------------------------------- */
RTIOW_API int rtiow_b200_render_multi(rtiow_scene_t* const* scenes, int ngpus, const rtiow_camera_t* camera, uint32_t nx,
                                      uint32_t ny, uint32_t ns, uint64_t seed, float* out_rgb);

typedef struct rtiow_peer_frame rtiow_peer_frame_t;
#define RTIOW_PEER_HANDLE_BYTES 128u
/* This rank's copy of the ny*nx*3 float frame on `device` — two buffers, consecutive renders alternate between
 * them — plus the hand-shake flags; n_ranks <= 16. */
RTIOW_API int rtiow_b200_peer_frame_create(int device, uint32_t nx, uint32_t ny, uint32_t rank, uint32_t n_ranks,
                                           rtiow_peer_frame_t** out);
/* RTIOW_PEER_HANDLE_BYTES bytes that let another rank (another process, or this one) map this frame. */
RTIOW_API int rtiow_b200_peer_frame_export(rtiow_peer_frame_t* frame, uint8_t* handle);
/* `handles`: the exported handles of all n_ranks ranks, in rank order.  Maps the peers' frames (CUDA IPC across
 * processes, peer access inside one). */
RTIOW_API int rtiow_b200_peer_frame_connect(rtiow_peer_frame_t* frame, const uint8_t* handles);
/* DEVICE pointer to the frame the most recent rtiow_b200_render_peers call on `frame` assembles: ny*nx*3 floats,
 * row 0 = top.  Ask again after every render (the two buffers alternate); the pointer stays valid, and its contents
 * stay untouched, until the end of the NEXT render on this frame. */
RTIOW_API int rtiow_b200_peer_frame_ptr(rtiow_peer_frame_t* frame, float** d_frame);
RTIOW_API void rtiow_b200_peer_frame_destroy(rtiow_peer_frame_t* frame);
/* This rank's share of par_cast, enqueued on `cuda_stream`: renders this rank's tiles, folds the samples into EVERY
 * rank's frame, then one barrier: signals "my rows are there" and waits for everybody's.  When the stream reaches the
 * end the frame of this rank (rtiow_b200_peer_frame_ptr) holds the whole image.  Nothing waits before the fold: it
 * writes the buffer the ranks read two renders ago, and every rank enters a barrier only after the reads it enqueued
 * before that call — so whatever reads a frame must be enqueued on the stream of the next call, before it.  All ranks
 * must make the same sequence of calls; a rank that does not arrive within 60 s makes the next call fail instead of
 * hanging the GPU. */
RTIOW_API int rtiow_b200_render_peers(rtiow_scene_t* scene, const rtiow_camera_t* camera, uint32_t nx, uint32_t ny,
                                      uint32_t ns, uint64_t seed, rtiow_peer_frame_t* frame, void* cuda_stream);

/* Parity/debug: per-sample radiance before the fold, HOST memory,
 * (row_end-row_begin)*nx*ns*4 floats laid out [row][x][sample]{r,g,b,segments}. */
RTIOW_API int rtiow_b200_render_samples(rtiow_scene_t* scene, const rtiow_camera_t* camera, uint32_t nx, uint32_t ny,
                              uint32_t ns, uint64_t seed, uint32_t row_begin, uint32_t row_end, float* out_samples);

/* print_ppm's per-channel sqrt + to_u8 (src/lib.rs:344-361) on the device: n floats -> n bytes.
 * `linear` and `out` are HOST memory (a checker for frames that already live on the host). */
RTIOW_API int rtiow_b200_ppm_quantise(rtiow_scene_t* scene, const float* linear, size_t n, uint8_t* out);

/* The same quantiser on a frame that is already on the device (the output of rtiow_b200_render_rows_device):
 * `d_linear` (n floats) and `d_out` (n bytes) are DEVICE memory; enqueued on `cuda_stream`, nothing synchronised. */
RTIOW_API int rtiow_b200_ppm_quantise_device(rtiow_scene_t* scene, const float* d_linear, size_t n, uint8_t* d_out,
                                             void* cuda_stream);

/* par_cast followed by print_ppm's quantiser, both on the device: `out_rgb8` (HOST) receives ny*nx*3 bytes, row 0 =
 * top, exactly the numbers print_ppm writes (src/lib.rs:346-360).  The frame crosses PCIe as bytes, a quarter of
 * the float frame (C5: 46 MB instead of 184 MB), and the host does no per-pixel arithmetic. */
RTIOW_API int rtiow_b200_render_ppm(rtiow_scene_t* scene, const rtiow_camera_t* camera, uint32_t nx, uint32_t ny, uint32_t ns,
                                    uint64_t seed, uint8_t* out_rgb8);

/* Synchronises the scene's device and reports the last render. */
RTIOW_API int rtiow_b200_get_stats(rtiow_scene_t* scene, rtiow_stats_t* out);

/* Tuning knobs (0 = default / automatic; every call sets all four): threads per CTA (256, 512, 768, or 1024 for the
 * specialised kernels only), CTAs per SM, per-sample staging budget in MiB (automatic: up to 56 GiB and 70 % of the
 * free device memory, so that the BASELINE configurations are one pass), force_global != 0 keeps the scene in global
 * memory even if it fits shared memory.  None of them changes a bit of the image. */
RTIOW_API int rtiow_b200_set_tuning(rtiow_scene_t* scene, uint32_t cta_threads, uint32_t ctas_per_sm, uint32_t staging_mib,
                          int force_global);

/* The megakernel is compiled for four feature sets: any scene; scenes that hold nothing but unwrapped spheres
 * with Lambertian / Metal / Dielectric materials and constant textures (book-1's random_scene); lists of rects and
 * rect_prisms, wrapped or not, with Lambertian and DiffuseLight materials (the Cornell box); and any scene without a
 * multi-item ConstantMedium boundary, a checker texture or a Scale wrapper (the book-2 final scene) — the same
 * per-path code with everything else compiled out: the kernel is bound by instruction fetch, so code that never runs
 * still costs.  Picked automatically from the scene's content; enable = 0 forces the general kernel.  Same image. */
RTIOW_API int rtiow_b200_set_specialisation(rtiow_scene_t* scene, int enable);

/* How `Bvh` subtrees (src/bvh.rs) are walked.  All give the same image bit for bit; the choice is
 * speed only.  REINDEXED (default): the library indexes the subtree's leaves with its own tree,
 * visits the nearer child first, culls inner boxes with a cheaper test that never rejects what the
 * reference's Aabb::hit accepts, and runs the reference's Aabb::hit on each leaf's own box before
 * its primitives.  REINDEXED_EXACT: the same tree with the reference's Aabb::hit arithmetic at
 * every node (smaller device image; picked automatically when only that fits in shared memory).
 * REFERENCE_ORDER: the boxes exactly as flattened, left first, like Bvh::hit (src/bvh.rs:85-120). */
enum { RTIOW_TRAVERSAL_REINDEXED = 0, RTIOW_TRAVERSAL_REFERENCE_ORDER = 1, RTIOW_TRAVERSAL_REINDEXED_EXACT = 2 };
RTIOW_API int rtiow_b200_set_traversal(rtiow_scene_t* scene, int mode);

#ifdef __cplusplus
}
#endif
#endif /* RTIOW_B200_H */
