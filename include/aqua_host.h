/*
 * aqua_host.h — C interface of libaqua_host.so, the C++ stand-in for the reference's Rust
 * host crate `arukas` (Cargo.toml:1-5, src/lib.rs:0 — empty in the snapshot).
 *
 * The north star keeps scene loading on the host in Rust; there is no Rust toolchain in
 * this image, so the same logic (serde-JSON scene schema, BSON .mesh reader, texture decode,
 * flattening into aq_scene_desc) is written in C++ here and mirrored 1:1 by
 * rust/arukas/src/lib.rs.  Nothing in this library touches the GPU.
 *
 *   scene JSON schema      scenes/cbox.json:1-627, scenes/room.json:1-3899
 *   integrator JSON        scenes/integrator.json:1-8                      
 *   .mesh BSON             scenes/ *.mesh                                   
 *   textures               scenes/textures/ *.jpg (baseline + progressive JPEG)
 */
#ifndef AQUA_HOST_H
#define AQUA_HOST_H

#include "aqua_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct aq_host_scene aq_host_scene;

typedef struct aq_host_scene_info {
    uint32_t n_shapes;        /* shapes[] entries in the JSON */
    uint32_t n_meshes_loaded; /* .mesh files found and parsed */
    uint32_t n_meshes_missing;/* e.g. living_room_Default_43.mesh (.MISSING_LARGE_BLOBS:1) */
    uint32_t n_verts, n_tris, n_materials, n_textures, n_lights;
    float bounds_min[3], bounds_max[3];
} aq_host_scene_info;

/* parse <json_path>, load every referenced .mesh / texture relative to its directory */
int aq_host_scene_load(const char* json_path, aq_host_scene** out);
void aq_host_scene_free(aq_host_scene* s);
/* flat description, valid until aq_host_scene_free */
const aq_scene_desc* aq_host_scene_desc(const aq_host_scene* s);
int aq_host_scene_get_info(const aq_host_scene* s, aq_host_scene_info* info);
/* name of material i (named_bsdfs key, JSON order), first triangle / count of shape i */
const char* aq_host_material_name(const aq_host_scene* s, uint32_t i);
int aq_host_shape_range(const aq_host_scene* s, uint32_t shape, uint32_t* first_tri,
                        uint32_t* n_tris, uint32_t* material);

/* integrator.json -> cfg (spp -> [0,spp), max_depth); type_out receives "nrc"/"pt"/... */
int aq_host_integrator_load(const char* json_path, aq_integrator_cfg* cfg, char* type_out,
                            size_t type_cap);
/* the NRC-only keys of the same file (:4 batch_size, :6 training_iters, :7 learning_rate,
 * :8 visualize_cache) -> aq_nrc_cfg (include/aqua_cuda.h) */
int aq_host_integrator_load_nrc(const char* json_path, aq_nrc_cfg* nrc);

/* single BSON mesh; arrays are malloc'ed, free with aq_host_free.  positions and normals hold
 * n_verts * 3 floats (normals are zero vectors when the file has none), uvs n_uvs * 2 (n_uvs is 0
 * or n_verts), indices n_tris * 3 */
int aq_host_mesh_load(const char* path, char* name_out, size_t name_cap, uint32_t* n_verts,
                      uint32_t* n_tris, float** positions, float** normals, float** uvs,
                      uint32_t* n_uvs, uint32_t** indices);
/* decode a JPEG (baseline or progressive, 8-bit, 1 or 3 components) to RGBA8 */
int aq_host_jpeg_decode(const char* path, uint32_t* width, uint32_t* height, uint8_t** rgba8);
void aq_host_free(void* p);
/* film (float4 sum, count) -> 8-bit sRGB PPM */
int aq_host_write_ppm(const char* path, const float* film, uint32_t width, uint32_t height);

/* output stage: RGBA8 -> PNG (stored-deflate, no external zlib), film -> PFM */
int aq_host_write_png(const char* path, const uint8_t* rgba8, uint32_t width, uint32_t height);
int aq_host_write_pfm(const char* path, const float* film, uint32_t width, uint32_t height);
/* 255 thresholds t[k]: a linear value v maps to sRGB level k+1 iff v >= t[k] (exact, monotone) */
void aq_host_srgb_thresholds(float* t255);

/* asset import: Wavefront OBJ/MTL -> <out_dir>/<scene_name>.json + BSON
 * .mesh files named <obj>_<group>_<i>.mesh; returns the number of meshes written or an aq_status */
int aq_host_import_obj(const char* obj_path, const char* out_dir, const char* scene_name);
const char* aq_host_import_last_error(void);

float aq_host_srgb_to_linear(float c);
const char* aq_host_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
