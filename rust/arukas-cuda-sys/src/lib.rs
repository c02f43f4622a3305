//! Raw `extern "C"` bindings to `libaqua_cuda.so` — a 1:1 transcription of
//! `include/aqua_cuda.h` (ABI version 1).  No logic lives here.
//!
//! NOTE: written without a Rust toolchain in the build image (no `cargo`/`rustc`), so this
//! crate is uncompiled; the C header is authoritative and `tests/test_abi.py` checks the
//! struct sizes of the (identically laid out) ctypes mirror against it.
#![allow(non_camel_case_types)]

use std::os::raw::{c_char, c_int, c_void};

pub const AQ_ABI_VERSION: c_int = 1;
pub const AQ_OK: c_int = 0;
pub const AQ_ERR_BAD_ARG: c_int = -1;
pub const AQ_ERR_CUDA: c_int = -2;
pub const AQ_ERR_OOM: c_int = -3;
pub const AQ_ERR_UNSUPPORTED: c_int = -4;
pub const AQ_ERR_STATE: c_int = -5;
pub const AQ_ERR_NCCL: c_int = -6;
pub const AQ_ERR_IO: c_int = -7;
pub const AQ_MISS: u32 = 0xFFFF_FFFF;
pub const AQ_RENDER_ACCUMULATE: u32 = 1;
pub const AQ_RENDER_DUMP_SAMPLES: u32 = 2;
pub const AQ_RENDER_PROFILE: u32 = 4;
pub const AQ_RENDER_MIS_NEE_ONLY: u32 = 8;
pub const AQ_RENDER_MIS_BSDF_ONLY: u32 = 16;
pub const AQ_RENDER_FORCE_FULL_BSDF: u32 = 32;
pub const AQ_RENDER_NRC_TENSOR: u32 = 64;
pub const AQ_NRC_N_WEIGHTS_ABI: usize = 16640;

/// scenes/integrator.json:4,6-8 — the NRC-only keys.
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct aq_nrc_cfg {
    pub batch_size: u32,
    pub training_iters: u32,
    pub learning_rate: f32,
    pub visualize_cache: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct aq_nrc_info {
    pub n_weights: u32,
    pub n_records: u32,
    pub n_valid: u32,
    pub loss_first: f32,
    pub loss_last: f32,
    pub ms_records: f32,
    pub ms_train: f32,
}

#[repr(C)]
pub struct aq_ctx { _private: [u8; 0] }
#[repr(C)]
pub struct aq_scene { _private: [u8; 0] }

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct aq_material {
    pub color: [f32; 3],
    pub color_tex: i32,
    pub metallic: f32,
    pub roughness: f32,
    pub specular: f32,
    pub specular_tint: f32,
    pub sheen: f32,
    pub sheen_tint: f32,
    pub clearcoat: f32,
    pub clearcoat_roughness: f32,
    pub ior: f32,
    pub transmission: f32,
    pub subsurface: f32,
    pub anisotropic: f32,
    pub anisotropic_rotation: f32,
    pub emission: [f32; 3],
    pub subsurface_color: [f32; 3],
    pub subsurface_radius: [f32; 3],
    /// 1-based texture index per AQ_PTEX_* slot (0 = the constant): `Texture::Image` on a parameter
    /// other than `color`.
    pub param_tex: [u8; 16],
}
pub const AQ_PTEX_METALLIC: usize = 0;
pub const AQ_PTEX_ROUGHNESS: usize = 1;
pub const AQ_PTEX_SPECULAR: usize = 2;
pub const AQ_PTEX_SPECULAR_TINT: usize = 3;
pub const AQ_PTEX_SHEEN: usize = 4;
pub const AQ_PTEX_SHEEN_TINT: usize = 5;
pub const AQ_PTEX_TRANSMISSION: usize = 6;
pub const AQ_PTEX_CLEARCOAT: usize = 7;
pub const AQ_PTEX_CLEARCOAT_ROUGHNESS: usize = 8;
pub const AQ_PTEX_IOR: usize = 9;
pub const AQ_PTEX_SUBSURFACE: usize = 10;
pub const AQ_PTEX_SUBSURFACE_COLOR: usize = 11;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct aq_texture { pub width: u32, pub height: u32, pub rgba8: *const u8 }

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct aq_point_light { pub pos: [f32; 3], pub intensity: [f32; 3] }

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct aq_camera {
    pub res: [u32; 2],
    pub fov: f32,
    pub lens_radius: f32,
    pub focal: f32,
    pub translate: [f32; 3],
    pub rotate: [f32; 3],
    pub scale: [f32; 3],
}

#[repr(C)]
pub struct aq_scene_desc {
    pub n_verts: u32,
    pub n_tris: u32,
    pub positions: *const f32,
    pub normals: *const f32,
    pub uvs: *const f32,
    pub indices: *const u32,
    pub tri_material: *const u32,
    pub n_materials: u32,
    pub materials: *const aq_material,
    pub n_textures: u32,
    pub textures: *const aq_texture,
    pub n_lights: u32,
    pub lights: *const aq_point_light,
    pub camera: aq_camera,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct aq_integrator_cfg {
    pub width: u32,
    pub height: u32,
    pub spp_begin: u32,
    pub spp_end: u32,
    pub max_depth: u32,
    pub seed: u32,
    pub pool_paths: u32,
    pub flags: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct aq_ray { pub o: [f32; 3], pub tmin: f32, pub d: [f32; 3], pub tmax: f32 }

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct aq_hit { pub prim: u32, pub t: f32, pub u: f32, pub v: f32 }

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct aq_stats {
    pub samples: u64,
    pub sample_bounces: u64,
    pub rays_closest: u64,
    pub rays_shadow: u64,
    pub nodes_fetched: u64,
    pub tris_fetched: u64,
    pub ms_total: f32,
    pub ms_raygen: f32,
    pub ms_trace: f32,
    pub ms_shade: f32,
    pub ms_shadow: f32,
    pub ms_film: f32,
    pub n_launches: u32,
    pub n_waves: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct aq_accel_info {
    pub n_nodes: u32,
    pub n_tri_records: u32,
    pub max_depth: u32,
    pub sah_cost: f32,
    pub build_ms: f32,
    pub builder: u32,
}

extern "C" {
    pub fn aq_abi_version() -> c_int;
    pub fn aq_init(device: c_int, out: *mut *mut aq_ctx) -> c_int;
    pub fn aq_destroy(ctx: *mut aq_ctx);
    pub fn aq_last_error(ctx: *mut aq_ctx) -> *const c_char;
    pub fn aq_set_stream(ctx: *mut aq_ctx, cuda_stream: *mut c_void) -> c_int;
    pub fn aq_device_info(ctx: *mut aq_ctx, sm_count: *mut c_int, cc_major: *mut c_int,
                          cc_minor: *mut c_int, hbm_bytes: *mut usize) -> c_int;
    pub fn aq_scene_create(ctx: *mut aq_ctx, desc: *const aq_scene_desc, out: *mut *mut aq_scene) -> c_int;
    pub fn aq_scene_destroy(scene: *mut aq_scene);
    pub fn aq_accel_build(scene: *mut aq_scene, info: *mut aq_accel_info) -> c_int;
    pub fn aq_accel_wait(scene: *mut aq_scene, info: *mut aq_accel_info) -> c_int;
    pub fn aq_accel_download(scene: *mut aq_scene, nodes80: *mut c_void, nodes_bytes: usize,
                             tris48: *mut c_void, tris_bytes: usize) -> c_int;
    pub fn aq_accel_build_host(positions: *const f32, n_verts: u32, indices: *const u32, n_tris: u32,
                               nodes80: *mut *mut c_void, nodes_bytes: *mut usize,
                               tris48: *mut *mut c_void, tris_bytes: *mut usize,
                               info: *mut aq_accel_info) -> c_int;
    pub fn aq_free(p: *mut c_void);
    pub fn aq_intersect(scene: *mut aq_scene, rays: *const aq_ray, n: u32, hits: *mut aq_hit, any_hit: c_int) -> c_int;
    pub fn aq_intersect_device_async(scene: *mut aq_scene, d_rays: *const c_void, n: u32,
                                     d_hits: *mut c_void, any_hit: c_int) -> c_int;
    pub fn aq_trace_counters(scene: *mut aq_scene, nodes_fetched: *mut u64, tris_fetched: *mut u64, reset: c_int) -> c_int;
    pub fn aq_render(scene: *mut aq_scene, cfg: *const aq_integrator_cfg, film_out: *mut f32,
                     stats: *mut aq_stats) -> c_int;
    pub fn aq_render_device_async(scene: *mut aq_scene, cfg: *const aq_integrator_cfg, d_film: *mut c_void) -> c_int;
    pub fn aq_render_finish(scene: *mut aq_scene, stats: *mut aq_stats) -> c_int;
    pub fn aq_render_samples(scene: *mut aq_scene, out: *mut f32, n_float4: usize) -> c_int;
    pub fn aq_generate_camera_rays(scene: *mut aq_scene, cfg: *const aq_integrator_cfg, sample: u32,
                                   rays_out: *mut aq_ray) -> c_int;
    pub fn aq_resolve(ctx: *mut aq_ctx, d_film: *const c_void, h_film: *const f32, width: u32, height: u32,
                      exposure: f32, rgba8_out: *mut u8) -> c_int;
    pub fn aq_nrc_train(scene: *mut aq_scene, cfg: *const aq_integrator_cfg, nrc: *const aq_nrc_cfg,
                        info: *mut aq_nrc_info) -> c_int;
    pub fn aq_nrc_render(scene: *mut aq_scene, cfg: *const aq_integrator_cfg, nrc: *const aq_nrc_cfg,
                         film_out: *mut f32, stats: *mut aq_stats) -> c_int;
    pub fn aq_nrc_render_device_async(scene: *mut aq_scene, cfg: *const aq_integrator_cfg, nrc: *const aq_nrc_cfg,
                                      d_film: *mut c_void) -> c_int;
    pub fn aq_nrc_get_weights(scene: *mut aq_scene, weights_out: *mut f32, n: usize) -> c_int;
    pub fn aq_nrc_set_weights(scene: *mut aq_scene, weights: *const f32, n: usize) -> c_int;
    pub fn aq_nrc_get_loss(scene: *mut aq_scene, loss_out: *mut f32, n_iters: usize) -> c_int;
    pub fn aq_nrc_get_records(scene: *mut aq_scene, x_out: *mut f32, y_out: *mut f32, n_records: usize) -> c_int;
    pub fn aq_render_multi(desc: *const aq_scene_desc, cfg: *const aq_integrator_cfg, n_gpus: c_int,
                           devices: *const c_int, film_out: *mut f32, stats: *mut aq_stats) -> c_int;
}
