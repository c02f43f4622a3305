//! `arukas` host crate — what `/root/reference/src/lib.rs` (0 bytes in the snapshot) has to
//! contain for the render hot path: the serde scene types exactly as the shipped JSON
//! spells them (scenes/cbox.json:1-627), the BSON `.mesh` reader, texture
//! decode through the `image` crate, flattening into `aq_scene_desc`, and a safe wrapper over
//! the FFI crate.  It mirrors aqua-engine_b200/host/aq_host.cpp line for line in behaviour.
//!
//! NOTE: no Rust toolchain exists in the build image, so this file is written but has never
//! been compiled; the C++ twin (libaqua_host.so) is what the tests and benches exercise.
use arukas_cuda_sys as sys;
use serde::Deserialize;
use std::collections::BTreeMap;
use std::ffi::CStr;
use std::path::{Path, PathBuf};

// ---------------------------------------------------------------- serde scene schema
#[derive(Deserialize, Clone, Debug)]
pub enum Texture {
    Float(f32),
    Float3([f32; 3]),
    Srgb([f32; 3]),
    Image(String),
}

#[derive(Deserialize, Clone, Debug)]
pub struct Principled {
    pub color: Texture,
    pub subsurface: Texture,
    pub subsurface_radius: Texture,
    pub subsurface_color: Texture,
    pub metallic: Texture,
    pub specular: Texture,
    pub specular_tint: Texture,
    pub roughness: Texture,
    pub anisotropic: Texture,
    pub anisotropic_rotation: Texture,
    pub sheen: Texture,
    pub sheen_tint: Texture,
    pub clearcoat: Texture,
    pub clearcoat_roughness: Texture,
    pub ior: Texture,
    pub transmission: Texture,
    pub emission: Texture,
    pub hint: String,
}

#[derive(Deserialize, Clone, Debug)]
pub enum Bsdf { Principled(Principled) }

#[derive(Deserialize, Clone, Debug)]
pub enum BsdfRef { Named(String) }

#[derive(Deserialize, Clone, Debug)]
pub struct Transform { pub translate: [f32; 3], pub rotate: [f32; 3], pub scale: [f32; 3] }

#[derive(Deserialize, Clone, Debug)]
pub enum Camera {
    Perspective { res: [u32; 2], fov: f32, lens_radius: f32, focal: f32, transform: Transform },
}

#[derive(Deserialize, Clone, Debug)]
pub enum Light { Point { pos: [f32; 3], emission: Texture } }

#[derive(Deserialize, Clone, Debug)]
pub enum Shape { Mesh(String, BsdfRef) }

#[derive(Deserialize, Clone, Debug)]
pub struct Scene {
    // serde_json's `preserve_order` feature (or an IndexMap) keeps JSON order; material ids
    // are assigned in that order by the C++ twin.  BTreeMap would sort by name instead —
    // harmless for rendering (ids are internal) but different from libaqua_host.so.
    pub named_bsdfs: BTreeMap<String, Bsdf>,
    pub camera: Camera,
    pub lights: Vec<Light>,
    pub shapes: Vec<Shape>,
}

/// scenes/integrator.json:1-8 — internally tagged; NRC-only keys are accepted and ignored.
#[derive(Deserialize, Clone, Debug)]
pub struct IntegratorConfig {
    #[serde(rename = "type")]
    pub kind: String,
    pub spp: u32,
    pub max_depth: u32,
    #[serde(default)]
    pub seed: u32,
    // the NRC-only keys (scenes/integrator.json:4,6-8); defaults = the shipped file's values
    #[serde(default = "default_batch_size")]
    pub batch_size: u32,
    #[serde(default = "default_training_iters")]
    pub training_iters: u32,
    #[serde(default = "default_learning_rate")]
    pub learning_rate: f32,
    #[serde(default)]
    pub visualize_cache: bool,
}
fn default_batch_size() -> u32 { 512 }
fn default_training_iters() -> u32 { 2048 }
fn default_learning_rate() -> f32 { 1.0e-3 }

// ---------------------------------------------------------------- BSON TriangleMesh
#[derive(Default, Clone, Debug)]
pub struct TriangleMesh {
    pub name: String,
    pub vertices: Vec<[f32; 3]>,
    pub normals: Vec<[f32; 3]>,
    pub texcoords: Vec<[f32; 2]>,
    pub indices: Vec<[u32; 3]>,
}

#[derive(Debug)]
pub enum Error { Io(std::io::Error), Json(serde_json::Error), Bson(String), Image(String), Aqua(i32, String) }
impl From<std::io::Error> for Error { fn from(e: std::io::Error) -> Self { Error::Io(e) } }
impl From<serde_json::Error> for Error { fn from(e: serde_json::Error) -> Self { Error::Json(e) } }
pub type Result<T> = std::result::Result<T, Error>;

fn rd_i32(d: &[u8], p: usize) -> i32 { i32::from_le_bytes([d[p], d[p + 1], d[p + 2], d[p + 3]]) }

/// Iterate the elements of the BSON document starting at `off`: (type, key, value offset, size).
fn bson_each<F: FnMut(u8, &str, usize, usize) -> Result<()>>(d: &[u8], off: usize, mut f: F) -> Result<()> {
    if off + 5 > d.len() { return Err(Error::Bson("truncated document".into())); }
    let len = rd_i32(d, off) as usize;
    if len < 5 || off + len > d.len() || d[off + len - 1] != 0 { return Err(Error::Bson("bad document".into())); }
    let (mut p, end) = (off + 4, off + len - 1);
    while p < end {
        let ty = d[p];
        p += 1;
        let kend = d[p..end].iter().position(|&b| b == 0).ok_or_else(|| Error::Bson("bad key".into()))? + p;
        let key = std::str::from_utf8(&d[p..kend]).map_err(|_| Error::Bson("bad key".into()))?;
        p = kend + 1;
        let size = match ty {
            0x01 | 0x12 => 8,
            0x10 => 4,
            0x08 => 1,
            0x0A => 0,
            0x02 => 4 + rd_i32(d, p) as usize,
            0x03 | 0x04 => rd_i32(d, p) as usize,
            t => return Err(Error::Bson(format!("unsupported BSON element type {:#x}", t))),
        };
        if p + size > end { return Err(Error::Bson("element overruns document".into())); }
        f(ty, key, p, size)?;
        p += size;
    }
    Ok(())
}

fn bson_num(d: &[u8], ty: u8, p: usize) -> Result<f64> {
    let mut b = [0u8; 8];
    match ty {
        0x01 => { b.copy_from_slice(&d[p..p + 8]); Ok(f64::from_le_bytes(b)) }
        0x12 => { b.copy_from_slice(&d[p..p + 8]); Ok(i64::from_le_bytes(b) as f64) }
        0x10 => Ok(rd_i32(d, p) as f64),
        _ => Err(Error::Bson("expected a number".into())),
    }
}

fn bson_tuples(d: &[u8], off: usize, arity: usize) -> Result<Vec<f64>> {
    let mut out = Vec::new();
    bson_each(d, off, |ty, _, v, _| {
        if ty != 0x04 { return Err(Error::Bson("expected array of arrays".into())); }
        let before = out.len();
        bson_each(d, v, |t2, _, v2, _| { out.push(bson_num(d, t2, v2)?); Ok(()) })?;
        if out.len() - before != arity { return Err(Error::Bson("wrong tuple arity".into())); }
        Ok(())
    })?;
    Ok(out)
}

pub fn load_mesh(path: &Path) -> Result<TriangleMesh> {
    let d = std::fs::read(path)?;
    if d.len() < 5 || rd_i32(&d, 0) as usize != d.len() { return Err(Error::Bson("length != file size".into())); }
    let mut m = TriangleMesh::default();
    bson_each(&d, 0, |ty, key, p, size| {
        match (key, ty) {
            ("name", 0x02) => m.name = String::from_utf8_lossy(&d[p + 4..p + size - 1]).into_owned(),
            ("vertices", 0x04) => m.vertices = bson_tuples(&d, p, 3)?.chunks(3).map(|c| [c[0] as f32, c[1] as f32, c[2] as f32]).collect(),
            ("normals", 0x04) => m.normals = bson_tuples(&d, p, 3)?.chunks(3).map(|c| [c[0] as f32, c[1] as f32, c[2] as f32]).collect(),
            ("texcoords", 0x04) => m.texcoords = bson_tuples(&d, p, 2)?.chunks(2).map(|c| [c[0] as f32, c[1] as f32]).collect(),
            ("indices", 0x04) => m.indices = bson_tuples(&d, p, 3)?.chunks(3).map(|c| [c[0] as u32, c[1] as u32, c[2] as u32]).collect(),
            _ => {}
        }
        Ok(())
    })?;
    Ok(m)
}

// ---------------------------------------------------------------- flattening
pub fn srgb_to_linear(c: f32) -> f32 {
    let c = c as f64;
    (if c <= 0.04045 { c / 12.92 } else { ((c + 0.055) / 1.055).powf(2.4) }) as f32
}

/// Decodes `rel` (relative to the scene file) once per scene; returns its index into the texture table.
fn texture_id(f: &mut FlatScene, tex_index: &mut BTreeMap<String, i32>, base: &Path, rel: String) -> Result<i32> {
    if let Some(&i) = tex_index.get(&rel) { return Ok(i); }
    let im = image::open(base.join(&rel)).map_err(|e| Error::Image(e.to_string()))?.to_rgba8();
    f.texture_dims.push((im.width(), im.height()));
    f.texture_pixels.push(im.into_raw());
    let i = f.texture_pixels.len() as i32 - 1;
    tex_index.insert(rel, i);
    Ok(i)
}

fn tex3(t: &Texture) -> ([f32; 3], Option<String>) {
    match t {
        Texture::Float(v) => ([*v; 3], None),
        Texture::Float3(v) => (*v, None),
        Texture::Srgb(v) => ([srgb_to_linear(v[0]), srgb_to_linear(v[1]), srgb_to_linear(v[2])], None),
        Texture::Image(p) => ([1.0; 3], Some(p.replace('\\', "/"))),
    }
}

/// Host-side flattened scene: owns every array `aq_scene_desc` points into.
pub struct FlatScene {
    pub positions: Vec<f32>,
    pub normals: Vec<f32>,
    pub uvs: Vec<f32>,
    pub indices: Vec<u32>,
    pub tri_material: Vec<u32>,
    pub materials: Vec<sys::aq_material>,
    pub texture_pixels: Vec<Vec<u8>>,
    pub texture_dims: Vec<(u32, u32)>,
    pub lights: Vec<sys::aq_point_light>,
    pub camera: sys::aq_camera,
    pub missing_meshes: Vec<PathBuf>,
}

pub fn load_scene(json_path: &Path) -> Result<FlatScene> {
    let scene: Scene = serde_json::from_reader(std::fs::File::open(json_path)?)?;
    let base = json_path.parent().unwrap_or_else(|| Path::new("."));
    let mut f = FlatScene {
        positions: vec![], normals: vec![], uvs: vec![], indices: vec![], tri_material: vec![],
        materials: vec![], texture_pixels: vec![], texture_dims: vec![], lights: vec![],
        camera: unsafe { std::mem::zeroed() }, missing_meshes: vec![],
    };
    let mut mat_index = BTreeMap::new();
    let mut tex_index: BTreeMap<String, i32> = BTreeMap::new();
    for (name, Bsdf::Principled(p)) in &scene.named_bsdfs {
        let (color, img) = tex3(&p.color);
        let mut color_tex = -1;
        if let Some(rel) = img {
            color_tex = texture_id(&mut f, &mut tex_index, base, rel)?;
        }
        // Texture::Image on a non-colour parameter: constant 1 x texel (aq_material.param_tex, 1-based)
        let mut param_tex = [0u8; 16];
        let mut sp = |t: &Texture, slot: usize, f: &mut FlatScene, ti: &mut BTreeMap<String, i32>| -> Result<f32> {
            let (v, img) = tex3(t);
            match img {
                Some(rel) => {
                    let id = texture_id(f, ti, base, rel)?;
                    if id >= 255 { return Err(Error::Image("more than 255 textures on non-colour parameters".into())); }
                    param_tex[slot] = (id + 1) as u8;
                    Ok(1.0)
                }
                None => Ok(v[0]),
            }
        };
        let metallic = sp(&p.metallic, sys::AQ_PTEX_METALLIC, &mut f, &mut tex_index)?;
        let roughness = sp(&p.roughness, sys::AQ_PTEX_ROUGHNESS, &mut f, &mut tex_index)?;
        let specular = sp(&p.specular, sys::AQ_PTEX_SPECULAR, &mut f, &mut tex_index)?;
        let specular_tint = sp(&p.specular_tint, sys::AQ_PTEX_SPECULAR_TINT, &mut f, &mut tex_index)?;
        let sheen = sp(&p.sheen, sys::AQ_PTEX_SHEEN, &mut f, &mut tex_index)?;
        let sheen_tint = sp(&p.sheen_tint, sys::AQ_PTEX_SHEEN_TINT, &mut f, &mut tex_index)?;
        let transmission = sp(&p.transmission, sys::AQ_PTEX_TRANSMISSION, &mut f, &mut tex_index)?;
        let clearcoat = sp(&p.clearcoat, sys::AQ_PTEX_CLEARCOAT, &mut f, &mut tex_index)?;
        let clearcoat_roughness = sp(&p.clearcoat_roughness, sys::AQ_PTEX_CLEARCOAT_ROUGHNESS, &mut f, &mut tex_index)?;
        let ior = sp(&p.ior, sys::AQ_PTEX_IOR, &mut f, &mut tex_index)?;
        let subsurface = sp(&p.subsurface, sys::AQ_PTEX_SUBSURFACE, &mut f, &mut tex_index)?;
        drop(sp);
        let (mut subsurface_color, sc_img) = tex3(&p.subsurface_color);
        if let Some(rel) = sc_img {
            let id = texture_id(&mut f, &mut tex_index, base, rel)?;
            if id >= 255 { return Err(Error::Image("more than 255 textures on non-colour parameters".into())); }
            param_tex[sys::AQ_PTEX_SUBSURFACE_COLOR] = (id + 1) as u8;
            subsurface_color = [1.0; 3];
        }
        let s = |t: &Texture| tex3(t).0[0]; // anisotropic*: carried, not evaluated
        mat_index.insert(name.clone(), f.materials.len() as u32);
        f.materials.push(sys::aq_material {
            color, color_tex, metallic, roughness, specular, specular_tint, sheen, sheen_tint,
            clearcoat, clearcoat_roughness, ior, transmission, subsurface, anisotropic: s(&p.anisotropic),
            anisotropic_rotation: s(&p.anisotropic_rotation), emission: tex3(&p.emission).0,
            subsurface_color, subsurface_radius: tex3(&p.subsurface_radius).0, param_tex,
        });
    }
    let Camera::Perspective { res, fov, lens_radius, focal, transform } = &scene.camera;
    f.camera = sys::aq_camera { res: *res, fov: *fov, lens_radius: *lens_radius, focal: *focal,
                                translate: transform.translate, rotate: transform.rotate, scale: transform.scale };
    for Light::Point { pos, emission } in &scene.lights {
        f.lights.push(sys::aq_point_light { pos: *pos, intensity: tex3(emission).0 });
    }
    for Shape::Mesh(path, BsdfRef::Named(bsdf)) in &scene.shapes {
        let mat = *mat_index.get(bsdf).ok_or_else(|| Error::Bson(format!("unresolved bsdf '{}'", bsdf)))?;
        let p = base.join(path.replace('\\', "/"));
        if !p.exists() { f.missing_meshes.push(p); continue; } // .MISSING_LARGE_BLOBS:1
        let m = load_mesh(&p)?;
        let vbase = (f.positions.len() / 3) as u32;
        for v in &m.vertices { f.positions.extend_from_slice(v); }
        if m.normals.len() == m.vertices.len() { for n in &m.normals { f.normals.extend_from_slice(n); } }
        else { f.normals.extend(std::iter::repeat(0.0).take(3 * m.vertices.len())); }
        if m.texcoords.len() == m.vertices.len() { for t in &m.texcoords { f.uvs.extend_from_slice(t); } }
        else { f.uvs.extend(std::iter::repeat(0.0).take(2 * m.vertices.len())); }
        for t in &m.indices { f.indices.extend_from_slice(&[vbase + t[0], vbase + t[1], vbase + t[2]]); }
        f.tri_material.extend(std::iter::repeat(mat).take(m.indices.len()));
    }
    Ok(f)
}

// ---------------------------------------------------------------- safe wrapper over the FFI
pub struct Context(*mut sys::aq_ctx);
pub struct DeviceScene<'a> { raw: *mut sys::aq_scene, ctx: &'a Context, pub res: [u32; 2] }

fn check(ctx: *mut sys::aq_ctx, rc: i32) -> Result<()> {
    if rc == sys::AQ_OK { return Ok(()); }
    let msg = unsafe { CStr::from_ptr(sys::aq_last_error(ctx)) }.to_string_lossy().into_owned();
    Err(Error::Aqua(rc, msg))
}

impl Context {
    pub fn new(device: i32) -> Result<Self> {
        let mut raw = std::ptr::null_mut();
        check(std::ptr::null_mut(), unsafe { sys::aq_init(device, &mut raw) })?;
        Ok(Context(raw))
    }
    pub fn upload<'a>(&'a self, f: &FlatScene) -> Result<DeviceScene<'a>> {
        let textures: Vec<sys::aq_texture> = f.texture_pixels.iter().zip(&f.texture_dims)
            .map(|(px, &(w, h))| sys::aq_texture { width: w, height: h, rgba8: px.as_ptr() }).collect();
        let desc = sys::aq_scene_desc {
            n_verts: (f.positions.len() / 3) as u32, n_tris: (f.indices.len() / 3) as u32,
            positions: f.positions.as_ptr(), normals: f.normals.as_ptr(),
            uvs: if textures.is_empty() { std::ptr::null() } else { f.uvs.as_ptr() },
            indices: f.indices.as_ptr(), tri_material: f.tri_material.as_ptr(),
            n_materials: f.materials.len() as u32, materials: f.materials.as_ptr(),
            n_textures: textures.len() as u32, textures: textures.as_ptr(),
            n_lights: f.lights.len() as u32, lights: f.lights.as_ptr(), camera: f.camera,
        };
        let mut raw = std::ptr::null_mut();
        check(self.0, unsafe { sys::aq_scene_create(self.0, &desc, &mut raw) })?;
        check(self.0, unsafe { sys::aq_accel_build(raw, std::ptr::null_mut()) })?;
        Ok(DeviceScene { raw, ctx: self, res: f.camera.res })
    }
}
impl Drop for Context { fn drop(&mut self) { unsafe { sys::aq_destroy(self.0) } } }

impl<'a> DeviceScene<'a> {
    /// Hybrid build: blocks until the host SAH tree has replaced the device LBVH tree (a render
    /// never needs this; benchmarks do).
    pub fn accel_wait(&self) -> Result<sys::aq_accel_info> {
        let mut info: sys::aq_accel_info = unsafe { std::mem::zeroed() };
        check(self.ctx.0, unsafe { sys::aq_accel_wait(self.raw, &mut info) })?;
        Ok(info)
    }
    /// Renders `cfg.spp` samples per pixel; returns the float4 film (sum r,g,b, count).
    pub fn render(&self, cfg: &IntegratorConfig, width: u32, height: u32) -> Result<(Vec<f32>, sys::aq_stats)> {
        let (w, h) = (if width == 0 { self.res[0] } else { width }, if height == 0 { self.res[1] } else { height });
        let c = sys::aq_integrator_cfg { width: w, height: h, spp_begin: 0, spp_end: cfg.spp,
                                         max_depth: cfg.max_depth, seed: cfg.seed, pool_paths: 0, flags: 0 };
        let mut film = vec![0f32; 4 * (w as usize) * (h as usize)];
        let mut stats = sys::aq_stats::default();
        if cfg.kind == "nrc" {
            // scenes/integrator.json:2 — train the radiance cache for this view, then render with it
            let n = sys::aq_nrc_cfg { batch_size: cfg.batch_size, training_iters: cfg.training_iters,
                                      learning_rate: cfg.learning_rate, visualize_cache: cfg.visualize_cache as u32 };
            let mut info = sys::aq_nrc_info::default();
            check(self.ctx.0, unsafe { sys::aq_nrc_train(self.raw, &c, &n, &mut info) })?;
            check(self.ctx.0, unsafe { sys::aq_nrc_render(self.raw, &c, &n, film.as_mut_ptr(), &mut stats) })?;
        } else {
            check(self.ctx.0, unsafe { sys::aq_render(self.raw, &c, film.as_mut_ptr(), &mut stats) })?;
        }
        Ok((film, stats))
    }
    pub fn intersect(&self, rays: &[sys::aq_ray], any_hit: bool) -> Result<Vec<sys::aq_hit>> {
        let mut hits = vec![sys::aq_hit { prim: sys::AQ_MISS, t: 0.0, u: 0.0, v: 0.0 }; rays.len()];
        check(self.ctx.0, unsafe { sys::aq_intersect(self.raw, rays.as_ptr(), rays.len() as u32, hits.as_mut_ptr(), any_hit as i32) })?;
        Ok(hits)
    }
}
impl<'a> Drop for DeviceScene<'a> { fn drop(&mut self) { unsafe { sys::aq_scene_destroy(self.raw) } } }

/// The whole path in one call, as the CLI of the original crate would drive it.
pub fn render_files(scene_json: &Path, integrator_json: &Path, width: u32, height: u32) -> Result<Vec<f32>> {
    let flat = load_scene(scene_json)?;
    let cfg: IntegratorConfig = serde_json::from_reader(std::fs::File::open(integrator_json)?)?;
    let ctx = Context::new(0)?;
    let dev = ctx.upload(&flat)?;
    Ok(dev.render(&cfg, width, height)?.0)
}
