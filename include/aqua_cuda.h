/*
 * aqua_cuda.h — C ABI of libaqua_cuda.so, the B200 (sm_100a) render hot path for
 * aqua-engine ("Arukas Engine") scenes.
 *
 * This is the drop-in boundary.  The reference crate `arukas`
 * (Cargo.toml:1-5) keeps its host code in Rust (`src/lib.rs`, which is EMPTY in the
 * reference snapshot — src/lib.rs:0) and binds these entry points through a `-sys`
 * crate (see INTEGRATION.md and rust/arukas-cuda-sys/src/lib.rs).  Because the
 * reference has no callable interface for this path, every entry point below cites
 * the reference DATA it consumes instead of a function it replaces.
 *
 * Conventions
 *   - plain C, POD structs, pointers + counts; no C++/torch types cross the boundary
 *   - every call returns an aq_status (0 = OK, <0 = error); text via aq_last_error()
 *   - the caller owns all input arrays (copied during aq_scene_create; may be freed
 *     as soon as the call returns); the library owns the opaque handles
 *   - one aq_ctx per device; calls on one ctx are serialised by the caller
 *   - all calls are synchronous on return unless the name ends in _async
 *   - there is NO CPU fallback: every compute entry fails with AQ_ERR_CUDA when no
 *     sm_100 device is usable
 */
#ifndef AQUA_CUDA_H
#define AQUA_CUDA_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AQ_ABI_VERSION 1

typedef enum aq_status {
    AQ_OK = 0,
    AQ_ERR_BAD_ARG = -1,
    AQ_ERR_CUDA = -2,
    AQ_ERR_OOM = -3,
    AQ_ERR_UNSUPPORTED = -4,
    AQ_ERR_STATE = -5, /* e.g. render before accel build */
    AQ_ERR_NCCL = -6,
    AQ_ERR_IO = -7
} aq_status;

typedef struct aq_ctx aq_ctx;     /* one per device */
typedef struct aq_scene aq_scene; /* geometry + materials + accel, device resident */

/* ---- scene description: flat mirror of the reference's serde types ----------------
 * Scene{named_bsdfs,camera,lights,shapes}            scenes/cbox.json:1-627
 * TriangleMesh{vertices,normals,texcoords,indices}   scenes/ *.mesh (BSON)
 * The host (Rust `arukas`, or aqua-engine_b200/host/ in this repo) flattens all
 * `shapes[]` into ONE indexed triangle list in `shapes[]` order; the global primitive
 * id of a triangle is its index in `indices`. */

/* Bsdf::Principled  scenes/cbox.json:4-65.  Colours are LINEAR (the host linearises
 * Texture::Srgb).  Only `color` may be an image (Texture::Image, room.json:6).
 * clearcoat / transmission (+ior) / subsurface (+subsurface_color) are evaluated as soon as one
 * material of the scene sets one of them > 0 or textures one of them; `anisotropic` > 0
 * also selects it (GGX stretched along dp/du, rotated by `anisotropic_rotation` turns);
 * subsurface_radius is carried but not evaluated (no volumetric walk). */
typedef struct aq_material {
    float color[3];
    int32_t color_tex; /* index into aq_scene_desc.textures, or -1 */
    float metallic;
    float roughness;
    float specular;
    float specular_tint;
    float sheen;
    float sheen_tint;
    float clearcoat;
    float clearcoat_roughness;
    float ior;
    float transmission;
    float subsurface;
    float anisotropic;
    float anisotropic_rotation;
    float emission[3];
    float subsurface_color[3];
    float subsurface_radius[3];
    /* Texture::Image on a parameter other than `color` (the schema allows it on every field,
     * scenes/cbox.json:5-63): 1-based index into aq_scene_desc.textures per AQ_PTEX_* slot, 0 = the
     * constant above.  The parameter is then constant * (first channel of the bilinear, linearised
     * texel) — subsurface_color: constant * rgb — so hosts set the constant to 1 (as for color_tex).
     * emission, anisotropic* and subsurface_radius take no image. */
    uint8_t param_tex[16];
} aq_material;
enum {
    AQ_PTEX_METALLIC = 0, AQ_PTEX_ROUGHNESS, AQ_PTEX_SPECULAR, AQ_PTEX_SPECULAR_TINT, AQ_PTEX_SHEEN, AQ_PTEX_SHEEN_TINT,
    AQ_PTEX_TRANSMISSION, AQ_PTEX_CLEARCOAT, AQ_PTEX_CLEARCOAT_ROUGHNESS, AQ_PTEX_IOR, AQ_PTEX_SUBSURFACE,
    AQ_PTEX_SUBSURFACE_COLOR, AQ_PTEX_COUNT
};

/* Texture::Image  scenes/room.json:6 — decoded by the host to 8-bit sRGB RGBA,
 * row 0 = top row of the image file. */
typedef struct aq_texture {
    uint32_t width, height;
    const uint8_t* rgba8; /* width*height*4 bytes */
} aq_texture;

/* Light::Point  scenes/cbox.json:545-559.  `intensity` = linearised emission (W/sr). */
typedef struct aq_point_light {
    float pos[3];
    float intensity[3];
} aq_point_light;

/* Camera::Perspective + Transform  scenes/cbox.json:517-542 */
typedef struct aq_camera {
    uint32_t res[2];    /* JSON res; the integrator cfg may override it */
    float fov;          /* full angle, degrees, across the larger image dimension */
    float lens_radius;
    float focal;
    float translate[3];
    float rotate[3];    /* Euler radians, R = Rz*Ry*Rx */
    float scale[3];
} aq_camera;

typedef struct aq_scene_desc {
    uint32_t n_verts, n_tris;
    const float* positions;       /* n_verts*3 */
    const float* normals;         /* n_verts*3 or NULL (=> geometric normals) */
    const float* uvs;             /* n_verts*2 or NULL */
    const uint32_t* indices;      /* n_tris*3 */
    const uint32_t* tri_material; /* n_tris */
    uint32_t n_materials;
    const aq_material* materials;
    uint32_t n_textures;
    const aq_texture* textures;
    uint32_t n_lights;
    const aq_point_light* lights;
    aq_camera camera;
} aq_scene_desc;

/* integrator config  scenes/integrator.json:1-8 (spp :3, max_depth :5; the NRC-only
 * keys :4,:6-8 are ignored by this path). */
typedef struct aq_integrator_cfg {
    uint32_t width, height;      /* 0 => use camera.res */
    uint32_t spp_begin, spp_end; /* sample index range [begin,end) rendered by this call */
    uint32_t max_depth;          /* max scattering events per path */
    uint32_t seed;
    uint32_t pool_paths;         /* wavefront pool size, 0 => default */
    uint32_t flags;              /* AQ_RENDER_* */
} aq_integrator_cfg;

#define AQ_RENDER_ACCUMULATE 1u /* add to the film instead of clearing it first */
#define AQ_RENDER_DUMP_SAMPLES 2u /* also keep per-sample radiance (aq_render_samples) */
#define AQ_RENDER_MIS_NEE_ONLY 8u   /* area lights through next-event estimation only (test hook) */
#define AQ_RENDER_MIS_BSDF_ONLY 16u /* area lights through BSDF-sampled hits only (test hook) */
#define AQ_RENDER_FORCE_FULL_BSDF 32u /* run the full-Principled vertex code even when no material needs it (test hook) */
#define AQ_RENDER_NRC_TENSOR 64u /* aq_nrc_render*: run the cache's MLP on the tensor cores (bf16 tcgen05, fp32 accumulate): \
                                    faster, within a tolerance of the exact fp32 lookup instead of bit-identical to it */
#define AQ_RENDER_PROFILE 4u /* CUDA events around the launches of every 8th wave -> aq_stats.ms_<stage> (scaled) */

typedef struct aq_ray {
    float o[3];
    float tmin;
    float d[3];
    float tmax;
} aq_ray; /* 32 B */

typedef struct aq_hit {
    uint32_t prim; /* global triangle id, 0xFFFFFFFF = miss */
    float t, u, v;
} aq_hit; /* 16 B */

#define AQ_MISS 0xFFFFFFFFu

typedef struct aq_stats {
    uint64_t samples;        /* camera paths started */
    uint64_t sample_bounces; /* path vertices shaded */
    uint64_t rays_closest;   /* closest-hit queries */
    uint64_t rays_shadow;    /* any-hit queries */
    uint64_t nodes_fetched;  /* BVH8 nodes fetched (only with AQ_STATS_COUNTERS builds) */
    uint64_t tris_fetched;
    float ms_total;          /* device time of the whole render (CUDA events) */
    float ms_raygen, ms_trace, ms_shade, ms_shadow, ms_film; /* per stage, if profiled */
    uint32_t n_launches;     /* kernels launched by the call */
    uint32_t n_waves;
} aq_stats;

typedef struct aq_accel_info {
    uint32_t n_nodes;      /* BVH8 nodes (80 B each) */
    uint32_t n_tri_records;/* 48 B each */
    uint32_t max_depth;
    float sah_cost;        /* host builder only */
    float build_ms;
    uint32_t builder;      /* 0 = host binned SAH, 1 = device LBVH, 2 = hybrid: LBVH now, SAH swapped in when ready */
} aq_accel_info;

/* ---- lifecycle -------------------------------------------------------------------- */
int aq_abi_version(void);
int aq_init(int device, aq_ctx** out);
/* destroys the ctx AND every scene created on it that is still alive: those aq_scene handles
 * are invalid afterwards (destroy scenes first, or not at all) */
void aq_destroy(aq_ctx* ctx);
const char* aq_last_error(aq_ctx* ctx); /* ctx may be NULL: last error of this thread */
/* use an externally created cudaStream_t (e.g. torch's current stream); NULL = private */
int aq_set_stream(aq_ctx* ctx, void* cuda_stream);
int aq_device_info(aq_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, size_t* hbm_bytes);

/* ---- scene ------------------------------------------------------------------------ */
int aq_scene_create(aq_ctx* ctx, const aq_scene_desc* desc, aq_scene** out);
void aq_scene_destroy(aq_scene* scene);
/* BVH2 -> BVH8 collapse -> 80 B quantised nodes + 48 B triangle records.  Builder: binned SAH on
 * the host below AQ_DEVICE_BUILD_MIN_TRIS triangles, Morton/LBVH on the device from there on
 * (env AQUA_ACCEL_BUILDER=host|device overrides). */
#define AQ_DEVICE_BUILD_MIN_TRIS 1000000u
/* Between AQ_HYBRID_BUILD_MIN_TRIS and AQ_DEVICE_BUILD_MIN_TRIS triangles the build is HYBRID
 * (info->builder == 2): the device LBVH tree is built first (a few ms) and the call returns; a
 * host thread builds the binned-SAH tree beside the caller's first renders and the render loop
 * swaps it in between two waves (AQUA_ACCEL_BUILDER=hybrid forces this for any size).  Both trees
 * give identical hits; the SAH tree renders ~15 % faster on room.json.  aq_accel_wait blocks until
 * the swap has happened (benchmarks, aq_accel_download and aq_render_multi's clone call it). */
#define AQ_HYBRID_BUILD_MIN_TRIS 50000u
int aq_accel_build(aq_scene* scene, aq_accel_info* info /* may be NULL */);
int aq_accel_wait(aq_scene* scene, aq_accel_info* info /* may be NULL: the final tree's figures */);
/* copy the built BVH8 back (test hook: lets tests walk the same tree on the CPU) */
int aq_accel_download(aq_scene* scene, void* nodes80, size_t nodes_bytes, void* tris48,
                      size_t tris_bytes);

/* host-only variant: build the same BVH8 from flat arrays without touching a GPU (offline
 * baking, and lets CPU-only tests walk the tree).  Buffers are malloc'ed; free with aq_free. */
int aq_accel_build_host(const float* positions, uint32_t n_verts, const uint32_t* indices,
                        uint32_t n_tris, void** nodes80, size_t* nodes_bytes, void** tris48,
                        size_t* tris_bytes, aq_accel_info* info);
void aq_free(void* p);

/* ---- intersection test hook (hit-ID parity, config C4) --------------------------- */
/* rays/hits are HOST pointers; any_hit!=0 => hits[i].prim is 0 (occluded) or AQ_MISS */
int aq_intersect(aq_scene* scene, const aq_ray* rays, uint32_t n, aq_hit* hits, int any_hit);
/* same, DEVICE pointers, asynchronous on the ctx stream */
int aq_intersect_device_async(aq_scene* scene, const void* d_rays, uint32_t n, void* d_hits,
                              int any_hit);

/* BVH8 nodes / triangle records fetched by the aq_intersect* calls since the last reset
 * (the figure the traversal roofline is computed from: 80 B per node, 48 B per record) */
int aq_trace_counters(aq_scene* scene, uint64_t* nodes_fetched, uint64_t* tris_fetched, int reset);

/* ---- render ------------------------------------------------------------------------ */
/* film_out: HOST float4[width*height] = (sum r, sum g, sum b, sample count) */
int aq_render(aq_scene* scene, const aq_integrator_cfg* cfg, float* film_out, aq_stats* stats);
/* film stays on the device (float4[width*height], caller-provided DEVICE pointer, or NULL
 * to use the scene's internal film); asynchronous on the ctx stream */
int aq_render_device_async(aq_scene* scene, const aq_integrator_cfg* cfg, void* d_film);
/* wait for the ctx stream and fetch the counters of the last render */
int aq_render_finish(aq_scene* scene, aq_stats* stats);
/* per-sample radiance of the last render made with AQ_RENDER_DUMP_SAMPLES:
 * out = HOST float4[(spp_end-spp_begin)*width*height], index (s-spp_begin)*W*H + pixel */
int aq_render_samples(aq_scene* scene, float* out, size_t n_float4);
/* camera rays exactly as the raygen kernel emits them (test hook) */
int aq_generate_camera_rays(aq_scene* scene, const aq_integrator_cfg* cfg, uint32_t sample,
                            aq_ray* rays_out /* HOST, width*height */);

/* ---- `nrc` integrator (neural radiance cache) --------------------------------------------
 * scenes/integrator.json:1-8 names this integrator: {"type":"nrc","spp":4,"batch_size":512,
 * "max_depth":5,"training_iters":2048,"learning_rate":0.001,"visualize_cache":false}.  The
 * semantics are defined in aqua-engine_b200/csrc/aq_nrc.h: a 64-wide fp32 MLP
 * is trained on path-traced radiance records of the scene and queried at the second hit of
 * every camera path (the first hit with visualize_cache). */
typedef struct aq_nrc_cfg {
    uint32_t batch_size;      /* integrator.json:4 — records per training iteration */
    uint32_t training_iters;  /* :6 */
    float learning_rate;      /* :7 */
    uint32_t visualize_cache; /* :8 — query the cache at the first hit */
} aq_nrc_cfg;

typedef struct aq_nrc_info {
    uint32_t n_weights;  /* AQ_NRC weights of the cache (16,640) */
    uint32_t n_records;  /* training_iters * batch_size */
    uint32_t n_valid;    /* records whose path reached the record vertex */
    float loss_first, loss_last; /* mean relative squared error of the first / last iteration */
    float ms_records;    /* device time: generating the records (one wavefront pass per record depth) */
    float ms_train;      /* device time: training_iters descent steps */
} aq_nrc_info;

#define AQ_NRC_N_WEIGHTS_ABI 16640u

/* generate the records and train the scene's cache (replaces an existing one).  Uses cfg's
 * width/height/max_depth/seed; spp is irrelevant here. */
int aq_nrc_train(aq_scene* scene, const aq_integrator_cfg* cfg, const aq_nrc_cfg* nrc, aq_nrc_info* info);
/* render with the trained cache (AQ_ERR_STATE if aq_nrc_train has not run); film_out as aq_render */
int aq_nrc_render(aq_scene* scene, const aq_integrator_cfg* cfg, const aq_nrc_cfg* nrc, float* film_out,
                  aq_stats* stats);
/* same, film stays on the device (d_film: DEVICE float4[width*height], or NULL for the ctx film);
 * asynchronous on the ctx stream, finish with aq_render_finish */
int aq_nrc_render_device_async(aq_scene* scene, const aq_integrator_cfg* cfg, const aq_nrc_cfg* nrc, void* d_film);
/* test hooks: the cache's weights (AQ_NRC_N_WEIGHTS_ABI floats, layout aq_nrc.h), the
 * per-iteration training loss, and the records of the last aq_nrc_train:
 * x = n_records*64 inputs, y = n_records*4 (target.rgb / fac, valid flag).  NULL = skip. */
int aq_nrc_get_weights(aq_scene* scene, float* weights_out, size_t n);
int aq_nrc_set_weights(aq_scene* scene, const float* weights, size_t n);
int aq_nrc_get_loss(aq_scene* scene, float* loss_out, size_t n_iters);
int aq_nrc_get_records(aq_scene* scene, float* x_out, float* y_out, size_t n_records);

/* ---- output stage ---------------------------------------------------------------------- */
/* film (DEVICE pointer d_film, or HOST pointer h_film when d_film is NULL) -> RGBA8 sRGB,
 * rgb = clamp(exposure * sum / count); rgba8_out: HOST, width*height*4 bytes */
int aq_resolve(aq_ctx* ctx, const void* d_film, const float* h_film, uint32_t width, uint32_t height,
               float exposure, uint8_t* rgba8_out);

/* ---- multi-GPU (single process, one ctx per device, NCCL film reduce) -------------- */
/* Renders spp range [spp_begin,spp_end) split evenly over n_gpus devices; scene and BVH
 * are replicated; per-GPU float4 films are summed onto device 0 with ncclReduce and copied
 * to film_out (HOST).  Requires libnccl at run time (dlopen). */
int aq_render_multi(const aq_scene_desc* desc, const aq_integrator_cfg* cfg, int n_gpus,
                    const int* devices, float* film_out, aq_stats* stats);

#ifdef __cplusplus
}
#endif
#endif /* AQUA_CUDA_H */
