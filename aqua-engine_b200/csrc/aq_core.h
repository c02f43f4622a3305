/*
 * aq_core.h — the DEFINITIONAL scalar functions of the render hot path.
 *
 * The reference snapshot has no implementation of this path (src/lib.rs:0 is empty), so
 * these functions *define* the semantics.  They are written once
 * as __host__ __device__ code with a fixed floating-point operation order:
 *   - nvcc compiles them into the sm_100a kernels with -fmad=false (no implicit
 *     contraction; every fused op is an explicit fmaf),
 *   - g++ compiles them with -ffp-contract=off into the CPU oracle (oracle/),
 * so that GPU and oracle agree bit for bit on every value computed here.  Only
 * +,-,*,/,sqrtf,fmaf and comparisons are used (all IEEE-754 correctly rounded on both
 * sides); there are no libm transcendentals on the per-sample path (sin/cos are the
 * polynomials below; tan/pow are evaluated once on the host at scene-create time).
 *
 * What is NOT shared: the execution strategy.  The GPU walks a compressed BVH8 in a
 * wavefront with queues; the oracle is a depth-first per-path loop over a brute-force
 * (or BVH2) intersector.
 *
 * Reference data each function consumes is cited next to it (paths relative to
 * /root/reference).
 */
#ifndef AQ_CORE_H
#define AQ_CORE_H

#include <math.h>
#include <stdint.h>
#if !defined(__CUDA_ARCH__)
#include <vector>
#endif

#if defined(__CUDACC__)
#define AQ_HD __host__ __device__ __forceinline__
#else
#define AQ_HD inline
#endif

#define AQ_PI 3.14159265358979323846f
#define AQ_INV_PI 0.31830988618379067154f
#define AQ_INF 3.402823466e+38f
#define AQ_MISS_ID 0xFFFFFFFFu
#define AQ_RAY_EPS 3.0e-5f      /* spawn offset scale, see aq_spawn_origin */
#define AQ_SHADOW_EPS 1.0e-4f   /* relative shortening of shadow rays */
#define AQ_RR_START_DEPTH 3u    /* Russian roulette from the 4th vertex on */
#define AQ_RNG_DIMS_PER_BOUNCE 8u

/* ------------------------------------------------------------------ vectors */
struct alignas(16) aq_u4 {
    uint32_t x, y, z, w;
};
struct alignas(16) aq_f4 {
    float x, y, z, w;
};
struct aq_v3 {
    float x, y, z;
};
AQ_HD aq_v3 aq_mk(float x, float y, float z) {
    aq_v3 r;
    r.x = x;
    r.y = y;
    r.z = z;
    return r;
}
AQ_HD aq_v3 aq_add(aq_v3 a, aq_v3 b) { return aq_mk(a.x + b.x, a.y + b.y, a.z + b.z); }
AQ_HD aq_v3 aq_sub(aq_v3 a, aq_v3 b) { return aq_mk(a.x - b.x, a.y - b.y, a.z - b.z); }
AQ_HD aq_v3 aq_mul(aq_v3 a, aq_v3 b) { return aq_mk(a.x * b.x, a.y * b.y, a.z * b.z); }
AQ_HD aq_v3 aq_scale(aq_v3 a, float s) { return aq_mk(a.x * s, a.y * s, a.z * s); }
AQ_HD aq_v3 aq_neg(aq_v3 a) { return aq_mk(-a.x, -a.y, -a.z); }
/* dot = fma(a.z,b.z, fma(a.y,b.y, a.x*b.x)) — fixed order */
AQ_HD float aq_dot(aq_v3 a, aq_v3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
/* cross component = fma(a1,b2, -(a2*b1)) */
AQ_HD aq_v3 aq_cross(aq_v3 a, aq_v3 b) {
    return aq_mk(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)),
                 fmaf(a.x, b.y, -(a.y * b.x)));
}
/* a + b*s */
AQ_HD aq_v3 aq_madd(aq_v3 a, aq_v3 b, float s) {
    return aq_mk(fmaf(b.x, s, a.x), fmaf(b.y, s, a.y), fmaf(b.z, s, a.z));
}
AQ_HD float aq_max3(aq_v3 a) {
    float m = a.x > a.y ? a.x : a.y;
    return m > a.z ? m : a.z;
}
AQ_HD float aq_minf(float a, float b) { return a < b ? a : b; }
AQ_HD float aq_maxf(float a, float b) { return a > b ? a : b; }
AQ_HD float aq_clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
AQ_HD aq_v3 aq_normalize(aq_v3 a) {
    float l = sqrtf(aq_dot(a, a));
    float inv = 1.0f / l;
    return aq_scale(a, inv);
}
AQ_HD float aq_lum(aq_v3 c) { return fmaf(0.0722f, c.z, fmaf(0.7152f, c.y, 0.2126f * c.x)); }

/* ------------------------------------------------------------------ RNG
 * Counter-based PCG hash keyed on (seed, pixel, sample, dimension); replaces the
 * reference's `rand 0.8.3` (Cargo.toml:14) whose stream is unknowable. */
AQ_HD uint32_t aq_pcg(uint32_t v) {
    uint32_t s = v * 747796405u + 2891336453u;
    uint32_t w = ((s >> ((s >> 28u) + 4u)) ^ s) * 277803737u;
    return (w >> 22u) ^ w;
}
/* key for one path; dims are then hashed off it */
AQ_HD uint32_t aq_rng_key(uint32_t seed, uint32_t pixel, uint32_t sample) {
    return aq_pcg(sample + aq_pcg(pixel + aq_pcg(seed)));
}
AQ_HD uint32_t aq_rng_u32(uint32_t key, uint32_t dim) { return aq_pcg(key + dim); }
AQ_HD float aq_u32_to_unit(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
AQ_HD float aq_rng(uint32_t key, uint32_t dim) { return aq_u32_to_unit(aq_rng_u32(key, dim)); }

/* ------------------------------------------------------------------ sin/cos(2*pi*u)
 * Polynomial (Cephes single-precision kernels on [-pi/4, pi/4]) so that host and device
 * agree bit for bit; |error| < 2e-7. */
AQ_HD void aq_sincos_2pi(float u, float* s_out, float* c_out) {
    float x = u * 4.0f;                  /* quadrants */
    float qf = floorf(x + 0.5f);
    float f = x - qf;                    /* [-0.5, 0.5] */
    float th = f * 1.57079632679489661923f;
    float z = th * th;
    float sp = fmaf(fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f), z, -1.6666654611e-1f);
    float s = fmaf(sp * z, th, th);
    float cp = fmaf(fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f), z,
                    4.166664568298827e-2f);
    float c = fmaf(cp, z * z, fmaf(-0.5f, z, 1.0f));
    int q = ((int)qf) & 3;
    float so = (q == 0) ? s : (q == 1) ? c : (q == 2) ? -s : -c;
    float co = (q == 0) ? c : (q == 1) ? -s : (q == 2) ? -c : s;
    *s_out = so;
    *c_out = co;
}

/* ------------------------------------------------------------------ camera
 * Camera::Perspective{res,fov,lens_radius,focal,transform}  scenes/cbox.json:517-542.
 * Conventions (builder-defined): right-handed, looks down -z, +y up, fov is
 * the full angle in degrees across the larger image dimension, pixel (0,0) top-left,
 * R = Rz*Ry*Rx.  aq_cam is the derived form (host evaluates tan/sin/cos once, in double). */
struct aq_cam {
    aq_v3 pos;
    aq_v3 right, up, back; /* columns of R (camera x,y,z axes in world space) */
    float tan_x, tan_y;
    float lens_radius, focal;
    uint32_t width, height;
};

#if 1 /* host-only helpers (plain inline: host in both nvcc passes) */
/* host only */
inline aq_cam aq_cam_derive(const float translate[3], const float rotate[3], float fov_deg,
                            float lens_radius, float focal, uint32_t w, uint32_t h) {
    aq_cam c;
    double cx = cos((double)rotate[0]), sx = sin((double)rotate[0]);
    double cy = cos((double)rotate[1]), sy = sin((double)rotate[1]);
    double cz = cos((double)rotate[2]), sz = sin((double)rotate[2]);
    /* R = Rz*Ry*Rx */
    double r00 = cz * cy, r01 = cz * sy * sx - sz * cx, r02 = cz * sy * cx + sz * sx;
    double r10 = sz * cy, r11 = sz * sy * sx + cz * cx, r12 = sz * sy * cx - cz * sx;
    double r20 = -sy, r21 = cy * sx, r22 = cy * cx;
    c.right = aq_mk((float)r00, (float)r10, (float)r20);
    c.up = aq_mk((float)r01, (float)r11, (float)r21);
    c.back = aq_mk((float)r02, (float)r12, (float)r22);
    c.pos = aq_mk(translate[0], translate[1], translate[2]);
    double t = tan(0.5 * (double)fov_deg * 3.14159265358979323846 / 180.0);
    if (w >= h) {
        c.tan_x = (float)t;
        c.tan_y = (float)(t * (double)h / (double)w);
    } else {
        c.tan_y = (float)t;
        c.tan_x = (float)(t * (double)w / (double)h);
    }
    c.lens_radius = lens_radius;
    c.focal = focal;
    c.width = w;
    c.height = h;
    return c;
}
#endif

struct aq_rayf {
    aq_v3 o;
    float tmin;
    aq_v3 d;
    float tmax;
};

/* dims 0,1 = pixel jitter; 2,3 = lens */
AQ_HD aq_rayf aq_camera_ray(const aq_cam& c, uint32_t px, uint32_t py, uint32_t key) {
    float u0 = aq_rng(key, 0u), u1 = aq_rng(key, 1u);
    float sx = fmaf(((float)px + u0) / (float)c.width, 2.0f, -1.0f);
    float sy = fmaf(((float)py + u1) / (float)c.height, -2.0f, 1.0f);
    aq_v3 dc = aq_mk(sx * c.tan_x, sy * c.tan_y, -1.0f); /* camera space, z = -1 plane */
    aq_v3 oc = aq_mk(0.0f, 0.0f, 0.0f);
    if (c.lens_radius > 0.0f) {
        float u2 = aq_rng(key, 2u), u3 = aq_rng(key, 3u);
        float r = sqrtf(u2) * c.lens_radius, sn, cs;
        aq_sincos_2pi(u3, &sn, &cs);
        oc = aq_mk(r * cs, r * sn, 0.0f);
        aq_v3 pf = aq_scale(dc, c.focal); /* point on the plane of focus */
        dc = aq_sub(pf, oc);
    }
    dc = aq_normalize(dc);
    aq_rayf r;
    r.d = aq_madd(aq_madd(aq_scale(c.right, dc.x), c.up, dc.y), c.back, dc.z);
    r.o = aq_madd(aq_madd(aq_madd(c.pos, c.right, oc.x), c.up, oc.y), c.back, oc.z);
    r.tmin = 0.0f;
    r.tmax = AQ_INF;
    return r;
}

/* ------------------------------------------------------------------ ray / triangle
 * Geometry: TriangleMesh  scenes/ *.mesh.  Moller-Trumbore, two-sided, with
 * a sign-folded determinant so the division happens only for candidates that pass the
 * barycentric tests.  Edges are inclusive (a ray through a shared edge hits both
 * triangles; the (t, prim) tie-break below picks one).  e1 = v1-v0, e2 = v2-v0 are plain
 * float subtractions (the GPU stores them precomputed in its 48 B triangle records).
 * Returns true and (t,u,v) if the ray crosses the triangle's plane inside it with
 * t > tmin; the caller applies its own upper bound / tie-break. */
AQ_HD bool aq_tri_test(aq_v3 o, aq_v3 d, float tmin, aq_v3 v0, aq_v3 e1, aq_v3 e2, float* t_out,
                       float* u_out, float* v_out) {
    aq_v3 pvec = aq_cross(d, e2);
    float det = aq_dot(e1, pvec);
    float adet = fabsf(det);
    if (!(adet > 0.0f)) return false;
    float sg = det < 0.0f ? -1.0f : 1.0f;
    aq_v3 tvec = aq_sub(o, v0);
    float U = aq_dot(tvec, pvec) * sg;
    if (U < 0.0f || U > adet) return false;
    aq_v3 qvec = aq_cross(tvec, e1);
    float V = aq_dot(d, qvec) * sg;
    if (V < 0.0f || U + V > adet) return false;
    float T = aq_dot(e2, qvec) * sg;
    float inv = 1.0f / adet;
    float t = T * inv;
    if (!(t > tmin)) return false;
    *t_out = t;
    *u_out = U * inv;
    *v_out = V * inv;
    return true;
}

/* closest-hit ordering: lexicographic min of (t, prim).  Coincident / duplicated
 * triangles exist in both scenes, so the tie-break is observable. */
AQ_HD bool aq_hit_closer(float t, uint32_t prim, float best_t, uint32_t best_prim) {
    return t < best_t || (t == best_t && prim < best_prim);
}

/* spawn a ray origin off the surface: p + ng * (eps * (1 + max|p|)), ng on the side of w */
AQ_HD aq_v3 aq_spawn_origin(aq_v3 p, aq_v3 ng_out) {
    float m = fabsf(p.x);
    m = aq_maxf(m, fabsf(p.y));
    m = aq_maxf(m, fabsf(p.z));
    float e = AQ_RAY_EPS * (1.0f + m);
    return aq_madd(p, ng_out, e);
}

/* ------------------------------------------------------------------ shading frame */
struct aq_frame {
    aq_v3 t, b, n;
};
/* Duff et al. 2017, branchless ONB */
AQ_HD aq_frame aq_make_frame(aq_v3 n) {
    float sg = n.z >= 0.0f ? 1.0f : -1.0f;
    float a = -1.0f / (sg + n.z);
    float b = n.x * n.y * a;
    aq_frame f;
    f.t = aq_mk(fmaf(sg * n.x, n.x * a, 1.0f), sg * b, -sg * n.x);
    f.b = aq_mk(b, fmaf(n.y, n.y * a, sg), -n.y);
    f.n = n;
    return f;
}
AQ_HD aq_v3 aq_to_local(const aq_frame& f, aq_v3 w) {
    return aq_mk(aq_dot(w, f.t), aq_dot(w, f.b), aq_dot(w, f.n));
}
AQ_HD aq_v3 aq_to_world(const aq_frame& f, aq_v3 w) {
    return aq_madd(aq_madd(aq_scale(f.t, w.x), f.b, w.y), f.n, w.z);
}

/* ------------------------------------------------------------------ Principled BSDF
 * Bsdf::Principled (17 Texture inputs)  scenes/cbox.json:4-65.  Burley/Cycles-style:
 *   diffuse  = Burley diffuse (+ sheen), weight (1-metallic)(1-transmission)
 *   specular = GGX (alpha = max(roughness^2, 1e-4)), height-correlated Smith G,
 *              Schlick Fresnel with F0 = lerp(0.08*specular*tint, base, metallic)
 * Two instantiations of the vertex code use this section:
 *   fast (aq_bsdf_setup/eval/sample)            the two lobes above; what both shipped scenes run
 *   FULL (aq_bsdf_setup_full/eval_full/sample_full)  adds the lobes whose inputs are zero in every
 *        shipped material (scenes/cbox.json:15-17,48-58): clearcoat, transmission (rough dielectric
 *        refraction through `ior`), the Disney flattened-diffuse `subsurface` blend.  A scene
 *        uses FULL as soon as one material sets one of them.  With those inputs at zero the FULL
 *        functions execute the same operations in the same order as the fast ones, so the two are
 *        bit-identical there (tests/test_full_bsdf.py).
 * Anisotropy (round 2, FULL instantiation only): `anisotropic` a in (0,1] stretches the GGX lobe of the
 * specular reflection and refraction along the surface tangent dp/du (from the triangle's UVs; the
 * shading frame's own tangent when the mesh has none), rotated about the normal by
 * `anisotropic_rotation` turns: alpha_x = max(r^2 / s, 1e-4), alpha_y = max(r^2 * s, 1e-4) with
 * s = sqrt(1 - 0.9 a) (Disney / Cycles).  a = 0 takes exactly the isotropic code.
 * All directions are in the local shading frame (n = +z), wo.z > 0. */
struct aq_bsdf_params {
    aq_v3 base;
    float metallic, roughness, specular, specular_tint, sheen, sheen_tint, transmission;
    /* read by the FULL instantiation only */
    float clearcoat, clearcoat_roughness, ior, subsurface;
    aq_v3 subsurface_color;
    float anisotropic, anisotropic_rotation;
};

struct aq_bsdf_ctx {
    aq_v3 base, f0, sheen_col;
    float alpha, diff_w, rough, p_spec;
    float lam_o; /* Smith Lambda(wo), shared by every eval at this vertex */
    float k_o;   /* 1 / ((1 + Lambda(wo)) * 4 wo.z): VNDF pdf = D * k_o */
    float fv;    /* (1 - wo.z)^5 */
};

AQ_HD float aq_pow5(float x) {
    float x2 = x * x;
    return x2 * x2 * x;
}

AQ_HD float aq_ggx_lambda(float alpha, float cz) {
    float c2 = cz * cz;
    float tan2 = (1.0f - c2) / c2;
    return 0.5f * (sqrtf(fmaf(alpha * alpha, tan2, 1.0f)) - 1.0f);
}

/* the part of the setup that does not depend on the lobe pick probabilities */
AQ_HD void aq_bsdf_setup_base(const aq_bsdf_params& m, aq_v3 wo, aq_bsdf_ctx* c) {
    c->base = m.base;
    c->rough = m.roughness;
    c->alpha = aq_maxf(m.roughness * m.roughness, 1.0e-4f);
    c->diff_w = (1.0f - m.metallic) * (1.0f - m.transmission);
    float l = aq_lum(m.base);
    aq_v3 tint = l > 0.0f ? aq_scale(m.base, 1.0f / l) : aq_mk(1.0f, 1.0f, 1.0f);
    aq_v3 one = aq_mk(1.0f, 1.0f, 1.0f);
    aq_v3 spec_col = aq_madd(one, aq_sub(tint, one), m.specular_tint); /* lerp(1,tint,st) */
    aq_v3 dielectric = aq_scale(spec_col, 0.08f * m.specular);
    c->f0 = aq_madd(dielectric, aq_sub(m.base, dielectric), m.metallic); /* lerp(diel,base,metallic) */
    c->sheen_col = aq_scale(aq_madd(one, aq_sub(tint, one), m.sheen_tint), m.sheen);
    c->fv = aq_pow5(1.0f - aq_clampf(wo.z, 0.0f, 1.0f));
    c->lam_o = 0.0f;
    c->k_o = 0.0f;
}

AQ_HD aq_bsdf_ctx aq_bsdf_setup(const aq_bsdf_params& m, aq_v3 wo) {
    aq_bsdf_ctx c;
    aq_bsdf_setup_base(m, wo, &c);
    /* lobe selection probability from the Fresnel-weighted specular albedo at wo */
    aq_v3 one = aq_mk(1.0f, 1.0f, 1.0f);
    aq_v3 Fo = aq_madd(c.f0, aq_sub(one, c.f0), c.fv);
    float ws = aq_max3(Fo);
    float wd = c.diff_w * aq_max3(m.base);
    float sum = ws + wd;
    c.p_spec = sum > 0.0f ? ws / sum : 0.0f;
    if (c.p_spec > 0.0f && wo.z > 0.0f) {
        c.lam_o = aq_ggx_lambda(c.alpha, wo.z);
        c.k_o = 1.0f / ((1.0f + c.lam_o) * (4.0f * wo.z));
    }
    return c;
}

/* GGX normal distribution D(h) * pi^0: a2 / (pi * ((n.h)^2 (a2-1) + 1)^2), the bracket written as
 * hz^2*a2 + (hx^2+hy^2): no cancellation at h = n */
AQ_HD float aq_ggx_d(float alpha, aq_v3 h) {
    float a2 = alpha * alpha;
    float dd = fmaf(h.z * h.z, a2, fmaf(h.x, h.x, h.y * h.y));
    return a2 / (AQ_PI * dd * dd);
}

/* f * |cos(theta_i)| and the one-sample-MIS pdf over both lobes; returns false if the
 * pair carries no energy */
AQ_HD bool aq_bsdf_eval(const aq_bsdf_ctx& c, aq_v3 wo, aq_v3 wi, aq_v3* f_cos, float* pdf) {
    if (!(wi.z > 0.0f) || !(wo.z > 0.0f)) return false;
    aq_v3 h = aq_add(wo, wi);
    float hl2 = aq_dot(h, h);
    if (!(hl2 > 0.0f)) return false;
    h = aq_scale(h, 1.0f / sqrtf(hl2));
    float ldh = aq_dot(wi, h);
    aq_v3 f = aq_mk(0.0f, 0.0f, 0.0f);
    float p = 0.0f;
    if (c.diff_w > 0.0f) {
        float fl = aq_pow5(1.0f - wi.z), fv = c.fv;
        float fd90 = fmaf(2.0f * c.rough, ldh * ldh, 0.5f);
        float fd = fmaf(fd90 - 1.0f, fl, 1.0f) * fmaf(fd90 - 1.0f, fv, 1.0f);
        float fh = aq_pow5(1.0f - ldh);
        aq_v3 d = aq_madd(aq_scale(c.base, AQ_INV_PI * fd), c.sheen_col, fh);
        f = aq_scale(d, c.diff_w * wi.z); /* f * cos */
        p = (1.0f - c.p_spec) * (wi.z * AQ_INV_PI);
    }
    if (c.p_spec > 0.0f) {
        float D = aq_ggx_d(c.alpha, h);
        float li = aq_ggx_lambda(c.alpha, wi.z);
        float fh = aq_pow5(1.0f - ldh);
        aq_v3 one = aq_mk(1.0f, 1.0f, 1.0f);
        aq_v3 F = aq_madd(c.f0, aq_sub(one, c.f0), fh);
        /* F D G / (4 wo.z wi.z) * wi.z with G = 1/(1 + Lambda_o + Lambda_i): one division */
        float sc = D / ((1.0f + c.lam_o + li) * (4.0f * wo.z));
        f = aq_madd(f, F, sc);
        p = fmaf(c.p_spec, D * c.k_o, p);
    }
    if (!(p > 0.0f)) return false;
    *f_cos = f;
    *pdf = p;
    return true;
}

/* GGX visible-normal sampling (Heitz 2018): the half vector for wo from u1 and (sin, cos)(2 pi u2) */
AQ_HD aq_v3 aq_sample_vndf(float alpha, aq_v3 wo, float u1, float sn, float cs) {
    aq_v3 vh = aq_normalize(aq_mk(alpha * wo.x, alpha * wo.y, wo.z));
    float lensq = fmaf(vh.x, vh.x, vh.y * vh.y);
    aq_v3 T1 = aq_mk(1.0f, 0.0f, 0.0f);
    if (lensq > 0.0f) {
        float il = 1.0f / sqrtf(lensq);
        T1 = aq_mk(-vh.y * il, vh.x * il, 0.0f);
    }
    aq_v3 T2 = aq_cross(vh, T1);
    float r = sqrtf(u1);
    float t1 = r * cs, t2 = r * sn;
    float s = 0.5f * (1.0f + vh.z);
    t2 = fmaf(s, t2, (1.0f - s) * sqrtf(aq_maxf(0.0f, 1.0f - t1 * t1)));
    float nz = sqrtf(aq_maxf(0.0f, 1.0f - t1 * t1 - t2 * t2));
    aq_v3 nh = aq_madd(aq_madd(aq_scale(T1, t1), T2, t2), vh, nz);
    return aq_normalize(aq_mk(alpha * nh.x, alpha * nh.y, aq_maxf(0.0f, nh.z)));
}
/* the same construction with two stretch factors (wo and the result in lobe axes) */
AQ_HD aq_v3 aq_sample_vndf_aniso(float ax, float ay, aq_v3 wo, float u1, float sn, float cs) {
    aq_v3 vh = aq_normalize(aq_mk(ax * wo.x, ay * wo.y, wo.z));
    float lensq = fmaf(vh.x, vh.x, vh.y * vh.y);
    aq_v3 T1 = aq_mk(1.0f, 0.0f, 0.0f);
    if (lensq > 0.0f) {
        float il = 1.0f / sqrtf(lensq);
        T1 = aq_mk(-vh.y * il, vh.x * il, 0.0f);
    }
    aq_v3 T2 = aq_cross(vh, T1);
    float r = sqrtf(u1);
    float t1 = r * cs, t2 = r * sn;
    float s = 0.5f * (1.0f + vh.z);
    t2 = fmaf(s, t2, (1.0f - s) * sqrtf(aq_maxf(0.0f, 1.0f - t1 * t1)));
    float nz = sqrtf(aq_maxf(0.0f, 1.0f - t1 * t1 - t2 * t2));
    aq_v3 nh = aq_madd(aq_madd(aq_scale(T1, t1), T2, t2), vh, nz);
    return aq_normalize(aq_mk(ax * nh.x, ay * nh.y, aq_maxf(0.0f, nh.z)));
}
AQ_HD aq_v3 aq_reflect(aq_v3 wo, aq_v3 h) {
    float odh = aq_dot(wo, h);
    return aq_sub(aq_scale(h, 2.0f * odh), wo);
}
AQ_HD aq_v3 aq_sample_cosine(float u1, float sn, float cs) {
    float r = sqrtf(u1);
    return aq_mk(r * cs, r * sn, sqrtf(aq_maxf(0.0f, 1.0f - u1)));
}

/* sample wi: u_lobe picks the lobe, (u1,u2) the direction.  Returns false if the sample
 * carries no energy.  weight = f*cos/pdf. */
AQ_HD bool aq_bsdf_sample(const aq_bsdf_ctx& c, aq_v3 wo, float u_lobe, float u1, float u2,
                          aq_v3* wi_out, aq_v3* weight, float* pdf_out) {
    if (!(wo.z > 0.0f)) return false;
    float sn, cs;
    aq_sincos_2pi(u2, &sn, &cs);
    aq_v3 wi;
    if (u_lobe < c.p_spec)
        wi = aq_reflect(wo, aq_sample_vndf(c.alpha, wo, u1, sn, cs));
    else
        wi = aq_sample_cosine(u1, sn, cs);
    aq_v3 fc;
    float p;
    if (!aq_bsdf_eval(c, wo, wi, &fc, &p)) return false;
    *wi_out = wi;
    *weight = aq_scale(fc, 1.0f / p);
    *pdf_out = p;
    return true;
}

/* ---- FULL: clearcoat + transmission + subsurface blend on top of the two lobes above.
 *   f = diff_w * diffuse(base_d) + [lerp(Schlick(F0), Fd, tw)] * GGX reflection
 *       + tw * (1 - Fd) * base * GGX refraction + 0.25 * clearcoat * Schlick(0.04) * GGX_cc reflection
 *   tw      = (1 - metallic) * transmission          (Cycles' "final transmission")
 *   Fd      = unpolarised dielectric Fresnel for the relative index eta = n_t / n_i
 *             (eta = ior entering through a front face, 1/ior leaving; ior clamped to >= 1.0001)
 *   base_d  = lerp(base, subsurface_color, subsurface); diffuse shape = lerp(Burley, Disney
 *             flattened Hanrahan-Krueger term, subsurface).  `subsurface_radius` has no meaning
 *             without a volumetric walk and is ignored.
 *   refraction is the rough-dielectric BTDF of Walter et al. 2007 in radiance-transport form
 *   (the 1/eta^2 scaling of radiance crossing the interface is applied, as in pbrt);
 *   clearcoat: GGX with alpha = max(clearcoat_roughness^2, 1e-4), fixed F0 = 0.04 (ior 1.5).
 * Lobe pick probabilities are proportional to max3 of each lobe's Fresnel-weighted albedo at wo
 * (refraction: (1 - Fd(wo.z)), floored at min(alpha, 0.5)). */
struct aq_bsdf_full {
    aq_bsdf_ctx c; /* c.base is the diffuse albedo base_d */
    aq_v3 tcol;    /* refraction tint */
    float tw, eta, ss;
    float cc_w, cc_alpha, cc_lam_o, cc_k_o;
    float p_cc, p_tr, p_diff;
    /* anisotropic GGX (aniso = false: the isotropic alpha of c is used everywhere): lobe axes in the
     * shading frame = (ct, st, 0) and (-st, ct, 0) */
    bool aniso;
    float ax, ay, ct, st;
};

/* ---- anisotropic GGX in the lobe's own axes (Heitz 2014/2018) */
AQ_HD aq_v3 aq_aniso_in(const aq_v3 w, float ct, float st) { return aq_mk(fmaf(ct, w.x, st * w.y), fmaf(ct, w.y, -(st * w.x)), w.z); }
AQ_HD aq_v3 aq_aniso_out(const aq_v3 w, float ct, float st) { return aq_mk(fmaf(ct, w.x, -(st * w.y)), fmaf(ct, w.y, st * w.x), w.z); }
AQ_HD float aq_ggx_d_aniso(float ax, float ay, aq_v3 h) { /* h in lobe axes */
    float hx = h.x / ax, hy = h.y / ay;
    float dd = fmaf(h.z, h.z, fmaf(hx, hx, hy * hy));
    return 1.0f / (AQ_PI * ax * ay * dd * dd);
}
AQ_HD float aq_ggx_lambda_aniso(float ax, float ay, aq_v3 w) { /* w in lobe axes, w.z != 0 */
    float sx = ax * w.x, sy = ay * w.y;
    float t2 = fmaf(sx, sx, sy * sy) / (w.z * w.z);
    return 0.5f * (sqrtf(1.0f + t2) - 1.0f);
}

AQ_HD float aq_fresnel_dielectric(float cos_i, float eta) {
    float c = aq_clampf(cos_i, 0.0f, 1.0f);
    float sin2_t = (1.0f - c * c) / (eta * eta);
    if (!(sin2_t < 1.0f)) return 1.0f; /* total internal reflection */
    float cos_t = sqrtf(1.0f - sin2_t);
    float r_par = (eta * c - cos_t) / (eta * c + cos_t);
    float r_per = (c - eta * cos_t) / (c + eta * cos_t);
    return 0.5f * (r_par * r_par + r_per * r_per);
}

/* D, Lambda and the half-vector sample of the (possibly anisotropic) specular / refraction lobe */
AQ_HD float aq_full_d(const aq_bsdf_full& b, aq_v3 h) {
    return b.aniso ? aq_ggx_d_aniso(b.ax, b.ay, aq_aniso_in(h, b.ct, b.st)) : aq_ggx_d(b.c.alpha, h);
}
AQ_HD float aq_full_lambda(const aq_bsdf_full& b, aq_v3 w) { /* w.z > 0 */
    return b.aniso ? aq_ggx_lambda_aniso(b.ax, b.ay, aq_aniso_in(w, b.ct, b.st)) : aq_ggx_lambda(b.c.alpha, w.z);
}
AQ_HD aq_v3 aq_full_sample_h(const aq_bsdf_full& b, aq_v3 wo, float u1, float sn, float cs) {
    if (!b.aniso) return aq_sample_vndf(b.c.alpha, wo, u1, sn, cs);
    return aq_aniso_out(aq_sample_vndf_aniso(b.ax, b.ay, aq_aniso_in(wo, b.ct, b.st), u1, sn, cs), b.ct, b.st);
}

/* (tan_c, tan_s): unit direction of the surface tangent dp/du in the shading frame's xy plane
 * (1, 0 when unknown); only read when the material is anisotropic */
AQ_HD aq_bsdf_full aq_bsdf_setup_full(const aq_bsdf_params& m, aq_v3 wo, float eta, float tan_c = 1.0f, float tan_s = 0.0f) {
    aq_bsdf_full b;
    aq_bsdf_setup_base(m, wo, &b.c);
    aq_bsdf_ctx& c = b.c;
    b.aniso = m.anisotropic > 0.0f;
    b.ax = b.ay = c.alpha;
    b.ct = 1.0f;
    b.st = 0.0f;
    if (b.aniso) {
        float asp = sqrtf(1.0f - 0.9f * aq_minf(m.anisotropic, 1.0f));
        float r2 = m.roughness * m.roughness;
        b.ax = aq_maxf(r2 / asp, 1.0e-4f);
        b.ay = aq_maxf(r2 * asp, 1.0e-4f);
        float rs, rc; /* rotate the tangent by anisotropic_rotation turns about the normal */
        aq_sincos_2pi(m.anisotropic_rotation - floorf(m.anisotropic_rotation), &rs, &rc);
        b.ct = fmaf(tan_c, rc, -(tan_s * rs));
        b.st = fmaf(tan_s, rc, tan_c * rs);
    }
    aq_v3 one = aq_mk(1.0f, 1.0f, 1.0f);
    b.tw = (1.0f - m.metallic) * m.transmission;
    b.eta = eta;
    b.tcol = m.base;
    b.ss = m.subsurface;
    if (b.ss > 0.0f) c.base = aq_madd(m.base, aq_sub(m.subsurface_color, m.base), b.ss);
    b.cc_w = 0.25f * m.clearcoat;
    b.cc_alpha = aq_maxf(m.clearcoat_roughness * m.clearcoat_roughness, 1.0e-4f);
    b.cc_lam_o = 0.0f;
    b.cc_k_o = 0.0f;
    aq_v3 Fo = aq_madd(c.f0, aq_sub(one, c.f0), c.fv);
    float fdo = 0.0f;
    if (b.tw > 0.0f) {
        fdo = aq_fresnel_dielectric(wo.z, eta);
        Fo = aq_madd(Fo, aq_sub(aq_mk(fdo, fdo, fdo), Fo), b.tw);
    }
    float ws = aq_max3(Fo);
    float wd = c.diff_w * aq_max3(c.base);
    /* the macro-surface Fresnel term may be 1 (total internal reflection) while microfacets still
     * refract: never let the pick probability of a lobe that carries energy fall to zero */
    float wt = b.tw > 0.0f ? b.tw * aq_maxf(1.0f - fdo, aq_minf(c.alpha, 0.5f)) * aq_max3(m.base) : 0.0f;
    float wc = b.cc_w > 0.0f ? b.cc_w * fmaf(0.96f, c.fv, 0.04f) : 0.0f;
    float sum = ws + wd;
    sum = sum + wt;
    sum = sum + wc;
    c.p_spec = sum > 0.0f ? ws / sum : 0.0f;
    b.p_tr = sum > 0.0f ? wt / sum : 0.0f;
    b.p_cc = sum > 0.0f ? wc / sum : 0.0f;
    b.p_diff = aq_maxf(0.0f, ((1.0f - c.p_spec) - b.p_cc) - b.p_tr);
    if ((c.p_spec > 0.0f || b.p_tr > 0.0f) && wo.z > 0.0f) {
        c.lam_o = aq_full_lambda(b, wo);
        c.k_o = 1.0f / ((1.0f + c.lam_o) * (4.0f * wo.z));
    }
    if (b.p_cc > 0.0f && wo.z > 0.0f) {
        b.cc_lam_o = aq_ggx_lambda(b.cc_alpha, wo.z);
        b.cc_k_o = 1.0f / ((1.0f + b.cc_lam_o) * (4.0f * wo.z));
    }
    return b;
}

/* wi.z > 0: the reflection lobes; wi.z < 0: refraction (only when tw > 0) */
AQ_HD bool aq_bsdf_eval_full(const aq_bsdf_full& b, aq_v3 wo, aq_v3 wi, aq_v3* f_cos, float* pdf) {
    const aq_bsdf_ctx& c = b.c;
    if (!(wo.z > 0.0f)) return false;
    aq_v3 f = aq_mk(0.0f, 0.0f, 0.0f);
    float p = 0.0f;
    if (wi.z > 0.0f) {
        aq_v3 h = aq_add(wo, wi);
        float hl2 = aq_dot(h, h);
        if (!(hl2 > 0.0f)) return false;
        h = aq_scale(h, 1.0f / sqrtf(hl2));
        float ldh = aq_dot(wi, h);
        if (c.diff_w > 0.0f) {
            float fl = aq_pow5(1.0f - wi.z), fv = c.fv;
            float fd90 = fmaf(2.0f * c.rough, ldh * ldh, 0.5f);
            float fd = fmaf(fd90 - 1.0f, fl, 1.0f) * fmaf(fd90 - 1.0f, fv, 1.0f);
            if (b.ss > 0.0f) {
                float fss90 = ldh * ldh * c.rough;
                float fss = fmaf(fss90 - 1.0f, fl, 1.0f) * fmaf(fss90 - 1.0f, fv, 1.0f);
                float sst = 1.25f * fmaf(fss, 1.0f / (wi.z + wo.z) - 0.5f, 0.5f);
                fd = fmaf(sst - fd, b.ss, fd);
            }
            float fh = aq_pow5(1.0f - ldh);
            aq_v3 d = aq_madd(aq_scale(c.base, AQ_INV_PI * fd), c.sheen_col, fh);
            f = aq_scale(d, c.diff_w * wi.z);
            p = b.p_diff * (wi.z * AQ_INV_PI);
        }
        if (c.p_spec > 0.0f) {
            float D = aq_full_d(b, h);
            float li = aq_full_lambda(b, wi);
            float fh = aq_pow5(1.0f - ldh);
            aq_v3 one = aq_mk(1.0f, 1.0f, 1.0f);
            aq_v3 F = aq_madd(c.f0, aq_sub(one, c.f0), fh);
            if (b.tw > 0.0f) {
                float fd = aq_fresnel_dielectric(ldh, b.eta);
                F = aq_madd(F, aq_sub(aq_mk(fd, fd, fd), F), b.tw);
            }
            float sc = D / ((1.0f + c.lam_o + li) * (4.0f * wo.z));
            f = aq_madd(f, F, sc);
            p = fmaf(c.p_spec, D * c.k_o, p);
        }
        if (b.p_cc > 0.0f) {
            float D = aq_ggx_d(b.cc_alpha, h);
            float li = aq_ggx_lambda(b.cc_alpha, wi.z);
            float Fc = fmaf(0.96f, aq_pow5(1.0f - ldh), 0.04f);
            float sc = b.cc_w * Fc * D / ((1.0f + b.cc_lam_o + li) * (4.0f * wo.z));
            f = aq_add(f, aq_mk(sc, sc, sc));
            p = fmaf(b.p_cc, D * b.cc_k_o, p);
        }
    } else if (wi.z < 0.0f && b.p_tr > 0.0f) {
        aq_v3 h = aq_madd(wo, wi, b.eta);
        float hl2 = aq_dot(h, h);
        if (!(hl2 > 0.0f)) return false;
        h = aq_scale(h, 1.0f / sqrtf(hl2));
        if (h.z < 0.0f) h = aq_neg(h);
        float odh = aq_dot(wo, h), idh = aq_dot(wi, h);
        if (!(odh > 0.0f) || !(idh < 0.0f)) return false;
        float F = aq_fresnel_dielectric(odh, b.eta);
        float den = fmaf(b.eta, idh, odh);
        float den2 = den * den;
        if (!(den2 > 0.0f)) return false;
        float D = aq_full_d(b, h);
        float li = aq_full_lambda(b, aq_neg(wi));
        float jac = D * odh * (-idh) / den2; /* D |wo.h| |wi.h| / (wo.h + eta wi.h)^2 */
        float sc = b.tw * (1.0f - F) * jac / ((1.0f + c.lam_o + li) * wo.z);
        f = aq_scale(b.tcol, sc);
        p = b.p_tr * (jac * (4.0f * c.k_o)) * (b.eta * b.eta);
    } else {
        return false;
    }
    if (!(p > 0.0f)) return false;
    *f_cos = f;
    *pdf = p;
    return true;
}

AQ_HD bool aq_bsdf_sample_full(const aq_bsdf_full& b, aq_v3 wo, float u_lobe, float u1, float u2,
                               aq_v3* wi_out, aq_v3* weight, float* pdf_out) {
    const aq_bsdf_ctx& c = b.c;
    if (!(wo.z > 0.0f)) return false;
    float sn, cs;
    aq_sincos_2pi(u2, &sn, &cs);
    aq_v3 wi;
    const float e1 = c.p_spec, e2 = e1 + b.p_cc, e3 = e2 + b.p_tr;
    if (u_lobe < e1) {
        /* a reflection sample below the horizon is lost (it must not be read as a refraction) */
        wi = aq_reflect(wo, aq_full_sample_h(b, wo, u1, sn, cs));
        if (!(wi.z > 0.0f)) return false;
    } else if (u_lobe < e2) {
        wi = aq_reflect(wo, aq_sample_vndf(b.cc_alpha, wo, u1, sn, cs));
        if (!(wi.z > 0.0f)) return false;
    } else if (u_lobe < e3) {
        aq_v3 h = aq_full_sample_h(b, wo, u1, sn, cs);
        float odh = aq_dot(wo, h);
        float sin2_t = (1.0f - odh * odh) / (b.eta * b.eta);
        if (!(sin2_t < 1.0f)) return false;
        float cos_t = sqrtf(1.0f - sin2_t);
        float inv_eta = 1.0f / b.eta;
        wi = aq_madd(aq_scale(wo, -inv_eta), h, odh * inv_eta - cos_t);
        if (!(wi.z < 0.0f)) return false;
    } else {
        wi = aq_sample_cosine(u1, sn, cs);
    }
    aq_v3 fc;
    float p;
    if (!aq_bsdf_eval_full(b, wo, wi, &fc, &p)) return false;
    *wi_out = wi;
    *weight = aq_scale(fc, 1.0f / p);
    *pdf_out = p;
    return true;
}

/* ------------------------------------------------------------------ textures
 * Texture::Image  scenes/room.json:6.  RGBA8 sRGB texels -> linear through a 256-entry
 * table built on the host; v' = 1-v, repeat wrap, bilinear. */
AQ_HD float aq_wrap01(float x) { return x - floorf(x); }

template <class TexelFn>
AQ_HD aq_v3 aq_tex_bilinear(uint32_t w, uint32_t h, float u, float v, TexelFn texel) {
    float fx = fmaf(aq_wrap01(u), (float)w, -0.5f);
    float fy = fmaf(aq_wrap01(1.0f - v), (float)h, -0.5f);
    float x0f = floorf(fx), y0f = floorf(fy);
    float ax = fx - x0f, ay = fy - y0f;
    int x0 = (int)x0f, y0 = (int)y0f;
    int x1 = x0 + 1, y1 = y0 + 1;
    if (x0 < 0) x0 += (int)w;
    if (y0 < 0) y0 += (int)h;
    if (x1 >= (int)w) x1 -= (int)w;
    if (y1 >= (int)h) y1 -= (int)h;
    aq_v3 c00 = texel(x0, y0), c10 = texel(x1, y0), c01 = texel(x0, y1), c11 = texel(x1, y1);
    aq_v3 a = aq_madd(c00, aq_sub(c10, c00), ax);
    aq_v3 b = aq_madd(c01, aq_sub(c11, c01), ax);
    return aq_madd(a, aq_sub(b, a), ay);
}

/* read-only scene loads: the non-coherent path (LDG.CONSTANT) on the device */
#if defined(__CUDA_ARCH__)
#define AQ_RO(p) __ldg(p)
__device__ __forceinline__ aq_f4 aq_ro_f4(const aq_f4* p) {
    float4 v = __ldg(reinterpret_cast<const float4*>(p));
    aq_f4 r;
    r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w;
    return r;
}
__device__ __forceinline__ aq_u4 aq_ro_u4(const aq_u4* p) {
    uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    aq_u4 r;
    r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w;
    return r;
}
#else
#define AQ_RO(p) (*(p))
inline aq_f4 aq_ro_f4(const aq_f4* p) { return *p; }
inline aq_u4 aq_ro_u4(const aq_u4* p) { return *p; }
#endif

/* ------------------------------------------------------------------ lights
 * Light::Point{pos,emission}  scenes/cbox.json:545-559 (Li = I/d^2), plus every triangle whose
 * material has emission > 0 (Bsdf::Principled.emission, scenes/cbox.json:61-63) as a two-sided
 * diffuse area light.  One table entry = AQ_LIGHT_WORDS x 16 B:
 *   w0 = p0.xyz | type (0 point, 1 triangle; stored as float)      point: position, tri: v0
 *   w1 = e1.xyz | pick probability of this light
 *   w2 = e2.xyz | area
 *   w3 = radiance.rgb (point: intensity) | cdf (cumulative pick probability, inclusive)
 *   w4 = unit normal.xyz | prim id bits
 * Lights are picked with probability proportional to power (point: 4 pi lum(I); triangle:
 * 2 pi area lum(Le)).  NEE on an area light and BSDF-sampled hits of it are combined with the
 * power heuristic; a point light is a delta: MIS weight 1. */
#define AQ_LIGHT_WORDS 5
#define AQ_MIS_BOTH 0u      /* NEE + BSDF hits, power heuristic (default) */
#define AQ_MIS_NEE_ONLY 1u  /* area lights through NEE only (camera-visible emission still counts) */
#define AQ_MIS_BSDF_ONLY 2u /* area lights through BSDF-sampled hits only */

AQ_HD float aq_power_heuristic(float pa, float pb) {
    float a2 = pa * pa;
    return a2 / fmaf(pb, pb, a2);
}

/* ------------------------------------------------------------------ path vertex
 * One scattering event, shared by the GPU shade kernel and the oracle's path loop
 * (call stack SURVEY §3 (3)).  Inputs: the hit geometry and material; outputs: the NEE
 * shadow ray + its pending contribution, the continuation ray + new throughput. */
struct aq_vertex_in {
    aq_v3 p;      /* hit point */
    aq_v3 ng;     /* unit geometric normal (any side) */
    aq_v3 ns;     /* interpolated shading normal, unnormalised; (0,0,0) => use ng */
    aq_v3 wo;     /* unit, pointing away from the surface (= -ray.d) */
    aq_bsdf_params mat;
    aq_v3 emission;
    aq_v3 dpdu;           /* FULL + anisotropic material: surface tangent dp/du, unnormalised; (0,0,0) = unknown */
    float t_hit;          /* ray parameter of the hit (|d| = 1: distance) */
    float prev_pdf;       /* BSDF pdf that generated this ray; 0 for camera rays */
    float light_pdf_area; /* pick probability / area if this triangle is a light, else 0 */
};

struct aq_vertex_out {
    bool has_shadow;
    aq_rayf shadow;
    aq_v3 shadow_contrib; /* beta * f*cos * Li * w_mis / pdf, added to L if unoccluded */
    bool has_next;
    aq_rayf next;
    float next_pdf; /* BSDF pdf of the continuation direction (for the next vertex's MIS) */
    aq_v3 beta;     /* throughput after the bounce (incl. RR compensation) */
    aq_v3 emitted;  /* beta_in * emission * w_mis, always added */
};

/* emitted radiance of the surface that was hit (two-sided), weighted against the NEE of the
 * previous vertex when the hit triangle is a light that NEE could have picked */
template <bool AREA>
AQ_HD aq_v3 aq_vertex_emitted(const aq_vertex_in& vi, aq_v3 beta, uint32_t mis_mode) {
    if (vi.emission.x != 0.0f || vi.emission.y != 0.0f || vi.emission.z != 0.0f) {
        float w = 1.0f;
        if (AREA && vi.prev_pdf > 0.0f && vi.light_pdf_area > 0.0f) {
            if (mis_mode == AQ_MIS_NEE_ONLY) {
                w = 0.0f;
            } else if (mis_mode == AQ_MIS_BOTH) {
                float cosl = fabsf(aq_dot(vi.ng, vi.wo));
                /* solid-angle pdf with which NEE at the previous vertex would have picked this point */
                float pl = cosl > 0.0f ? vi.light_pdf_area * vi.t_hit * vi.t_hit / cosl : 0.0f;
                w = aq_power_heuristic(vi.prev_pdf, pl);
            }
        }
        return aq_scale(aq_mul(beta, vi.emission), w);
    }
    return aq_mk(0.0f, 0.0f, 0.0f);
}

/* geometric and shading normal on the side of wo.  Returns whether wo is on the side of the
 * winding-defined normal (it faces the outside of a transmissive object, so true = the ray
 * arrives from outside).  ns = normalised interpolated normal, flipped to ng's side; ng itself
 * when the mesh has no normals or the interpolated one faces away from wo. */
AQ_HD bool aq_orient_normals(const aq_vertex_in& vi, aq_v3* ng_out, aq_v3* ns_out) {
    aq_v3 ng = vi.ng;
    const bool front = !(aq_dot(ng, vi.wo) < 0.0f);
    if (!front) ng = aq_neg(ng);
    aq_v3 ns = vi.ns;
    float nl2 = aq_dot(ns, ns);
    if (nl2 > 0.0f) {
        ns = aq_scale(ns, 1.0f / sqrtf(nl2));
        if (aq_dot(ns, ng) < 0.0f) ns = aq_neg(ns);
        if (!(aq_dot(ns, vi.wo) > 0.0f)) ns = ng;
    } else {
        ns = ng;
    }
    *ng_out = ng;
    *ns_out = ns;
    return front;
}

/* dims used per bounce b (base = 4 + 8*b): +0 light pick, +1 lobe, +2,+3 direction, +4 RR,
 * +5,+6 point on an area light */
/* AREA = the scene has emissive triangles.  AREA=false is the same function with the
 * area-light branches compiled out (they can never be taken then), so that scenes lit by
 * point lights only do not pay registers and instructions for them.
 * FULL = some material has clearcoat, transmission or subsurface > 0: the vertex uses the
 * aq_bsdf_*_full functions, light can arrive from and paths can continue to the far side of a
 * transmissive surface (rays spawned there start on the far side of the geometric plane). */
template <bool AREA, bool FULL>
AQ_HD void aq_shade_vertex(const aq_vertex_in& vi, aq_v3 beta, uint32_t key, uint32_t depth,
                           uint32_t max_depth, uint32_t n_lights, const aq_f4* lights, uint32_t mis_mode,
                           aq_vertex_out* vo) {
    vo->has_shadow = false;
    vo->has_next = false;
    vo->next_pdf = 0.0f;
    vo->beta = beta;
    uint32_t dim0 = 4u + AQ_RNG_DIMS_PER_BOUNCE * depth;

    /* emitted radiance of the surface that was hit (two-sided) */
    vo->emitted = aq_vertex_emitted<AREA>(vi, beta, mis_mode);

    /* orient normals to the side of wo */
    aq_v3 ng, ns;
    const bool front = aq_orient_normals(vi, &ng, &ns);
    aq_frame fr = aq_make_frame(ns);
    aq_v3 wo = aq_to_local(fr, vi.wo);
    if (!(wo.z > 0.0f)) return; /* exactly grazing */
    aq_bsdf_ctx bc;
    aq_bsdf_full bf;
    if (FULL) {
        float ior = aq_maxf(vi.mat.ior, 1.0001f);
        float tc = 1.0f, ts = 0.0f;
        if (vi.mat.anisotropic > 0.0f) { /* the tangent's direction in the shading frame's xy plane */
            float tx = aq_dot(vi.dpdu, fr.t), ty = aq_dot(vi.dpdu, fr.b);
            float tl2 = fmaf(tx, tx, ty * ty);
            if (tl2 > 0.0f) {
                float il = 1.0f / sqrtf(tl2);
                tc = tx * il;
                ts = ty * il;
            }
        }
        bf = aq_bsdf_setup_full(vi.mat, wo, front ? ior : 1.0f / ior, tc, ts);
    } else {
        bc = aq_bsdf_setup(vi.mat, wo);
    }
    const bool two_sided = FULL && bf.p_tr > 0.0f; /* energy can cross the surface */
    aq_v3 org = aq_spawn_origin(vi.p, ng);

    /* next-event estimation on one light picked by power */
    if (n_lights > 0u) {
        uint32_t li = 0u;
        if (n_lights > 1u) { /* binary search of the cdf (w3.w) */
            float ul = aq_rng(key, dim0 + 0u);
            uint32_t lo = 0u, hi = n_lights - 1u;
            while (lo < hi) {
                uint32_t mid = (lo + hi) >> 1;
                if (ul < aq_ro_f4(lights + (size_t)mid * AQ_LIGHT_WORDS + 3).w)
                    hi = mid;
                else
                    lo = mid + 1u;
            }
            li = lo;
        }
        const aq_f4* Lp = lights + (size_t)li * AQ_LIGHT_WORDS;
        aq_f4 l0 = aq_ro_f4(Lp), l1 = aq_ro_f4(Lp + 1), l3 = aq_ro_f4(Lp + 3);
        float pick = l1.w;
        aq_v3 rad = aq_mk(l3.x, l3.y, l3.z);
        bool is_tri = AREA && l0.w != 0.0f;
        aq_v3 lp = aq_mk(l0.x, l0.y, l0.z);
        float pdf_area = 0.0f; /* tri: pick / area */
        aq_v3 nl = aq_mk(0.0f, 0.0f, 0.0f);
        bool use = pick > 0.0f;
        if (AREA && is_tri) {
            if (mis_mode == AQ_MIS_BSDF_ONLY) use = false;
            aq_f4 l2 = aq_ro_f4(Lp + 2), l4 = aq_ro_f4(Lp + 4);
            float u1 = aq_rng(key, dim0 + 5u), u2 = aq_rng(key, dim0 + 6u);
            float su = sqrtf(u1);
            float b1 = 1.0f - su, b2 = u2 * su;
            lp = aq_madd(aq_madd(lp, aq_mk(l1.x, l1.y, l1.z), b1), aq_mk(l2.x, l2.y, l2.z), b2);
            pdf_area = pick / l2.w;
            nl = aq_mk(l4.x, l4.y, l4.z);
        }
        aq_v3 so = org; /* shadow-ray origin: on the light's side of the surface */
        if (two_sided && aq_dot(aq_sub(lp, vi.p), ng) < 0.0f) so = aq_spawn_origin(vi.p, aq_neg(ng));
        aq_v3 dl = aq_sub(lp, so);
        float d2 = aq_dot(dl, dl);
        if (use && d2 > 0.0f) {
            float dist = sqrtf(d2);
            float inv_dist = 1.0f / dist;
            aq_v3 wiw = aq_scale(dl, inv_dist);
            float side = aq_dot(wiw, ng);
            if (side > 0.0f || (two_sided && side < 0.0f)) {
                aq_v3 wi = aq_to_local(fr, wiw);
                aq_v3 fc;
                float pdf;
                bool ok;
                if (FULL) /* the shading frame and the geometric plane must agree on the side */
                    ok = ((side > 0.0f) == (wi.z > 0.0f)) && aq_bsdf_eval_full(bf, wo, wi, &fc, &pdf);
                else
                    ok = aq_bsdf_eval(bc, wo, wi, &fc, &pdf);
                if (ok) {
                    aq_v3 Li;
                    if (AREA && is_tri) {
                        float cosl = fabsf(aq_dot(nl, wiw));
                        float pl = cosl > 0.0f ? pdf_area * d2 / cosl : 0.0f; /* solid-angle pdf */
                        /* at the last vertex there is no continuation ray that could find the
                         * light by BSDF sampling: NEE carries the whole direct term there */
                        bool last = depth + 1u >= max_depth;
                        float w = (mis_mode == AQ_MIS_BOTH && !last) ? aq_power_heuristic(pl, pdf) : 1.0f;
                        Li = pl > 0.0f ? aq_scale(rad, w / pl) : aq_mk(0.0f, 0.0f, 0.0f);
                    } else {
                        Li = aq_scale(rad, inv_dist * inv_dist / pick);
                    }
                    vo->shadow_contrib = aq_mul(beta, aq_mul(fc, Li));
                    vo->shadow.o = so;
                    vo->shadow.d = wiw;
                    vo->shadow.tmin = 0.0f;
                    vo->shadow.tmax = dist * (1.0f - AQ_SHADOW_EPS);
                    vo->has_shadow = aq_max3(vo->shadow_contrib) > 0.0f;
                }
            }
        }
    }

    /* continuation */
    if (depth + 1u >= max_depth) return;
    float ulobe = aq_rng(key, dim0 + 1u), u1 = aq_rng(key, dim0 + 2u), u2 = aq_rng(key, dim0 + 3u);
    aq_v3 wi, w;
    float pdf;
    if (FULL) {
        if (!aq_bsdf_sample_full(bf, wo, ulobe, u1, u2, &wi, &w, &pdf)) return;
    } else {
        if (!aq_bsdf_sample(bc, wo, ulobe, u1, u2, &wi, &w, &pdf)) return;
    }
    aq_v3 wiw = aq_to_world(fr, wi);
    aq_v3 no = org; /* origin of the continuation ray */
    if (FULL && wi.z < 0.0f) {
        if (!(aq_dot(wiw, ng) < 0.0f)) return;
        no = aq_spawn_origin(vi.p, aq_neg(ng));
    } else {
        if (!(aq_dot(wiw, ng) > 0.0f)) return; /* shading-normal light leak guard */
    }
    aq_v3 nb = aq_mul(beta, w);
    if (depth + 1u >= AQ_RR_START_DEPTH) {
        float q = aq_minf(aq_max3(nb), 0.95f);
        float ur = aq_rng(key, dim0 + 4u);
        if (!(ur < q)) return;
        nb = aq_scale(nb, 1.0f / q);
    }
    if (!(aq_max3(nb) > 0.0f)) return;
    vo->beta = nb;
    vo->next.o = no;
    vo->next.d = wiw; /* unit to 1e-7: reflection / refraction / cosine sample of unit vectors in an orthonormal frame */
    vo->next.tmin = 0.0f;
    vo->next.tmax = AQ_INF;
    vo->next_pdf = pdf;
    vo->has_next = true;
}

/* ------------------------------------------------------------------ scene fetch
 * Flat device/host view of the scene arrays (aq_scene_desc, include/aqua_cuda.h) in the
 * form both the shade kernel and the oracle read them.
 *   materials: AQ_MAT_WORDS x 16 B per material
 *     m0 = base.rgb | color_tex (int bits, -1 = none)
 *     m1 = metallic roughness specular specular_tint
 *     m2 = sheen sheen_tint transmission 0
 *     m3 = emission.rgb 0
 *     m4 = clearcoat clearcoat_roughness ior subsurface      (read by the FULL instantiation only)
 *     m5 = subsurface_color.rgb 0                            (   "   )
 *   textures: desc[i] = {width, height, first texel, 0}; texels are packed RGBA8 words
 *   srgb_lut: 256 floats, sRGB byte -> linear (built on the host)
 *   lights: AQ_LIGHT_WORDS x 16 B per light (point lights first, then emissive triangles)
 *   prim_light_pdf: pick probability / area per triangle (0 = not a light), or null */
/* slots of aq_material.param_tex (the AQ_PTEX_* enum of include/aqua_cuda.h; aq_cuda.cu asserts they agree) */
#define AQ_PT_METALLIC 0
#define AQ_PT_ROUGHNESS 1
#define AQ_PT_SPECULAR 2
#define AQ_PT_SPECULAR_TINT 3
#define AQ_PT_SHEEN 4
#define AQ_PT_SHEEN_TINT 5
#define AQ_PT_TRANSMISSION 6
#define AQ_PT_CLEARCOAT 7
#define AQ_PT_CLEARCOAT_ROUGHNESS 8
#define AQ_PT_IOR 9
#define AQ_PT_SUBSURFACE 10
#define AQ_PT_SUBSURFACE_COLOR 11
#define AQ_MAT_WORDS 7 /* row 6: the 16 one-based texture indices of aq_material.param_tex; row 2 .w != 0 when any is set */
struct aq_scene_view {
    const float* pos;
    const float* nrm; /* may be null */
    const float* uv;  /* may be null */
    const uint32_t* idx;
    const uint32_t* tri_mat;
    const aq_f4* mats;
    const aq_u4* tex_desc;
    const uint32_t* texels;
    const float* srgb_lut;
    const aq_f4* lights;
    uint32_t n_lights;
    const float* prim_light_pdf;
    const aq_f4* shade_recs; /* AQ_SHADE_REC_WORDS x 16 B per triangle, or null */
};

#if defined(AQUA_CUDA_H)
/* host only: aq_material (include/aqua_cuda.h) -> AQ_MAT_WORDS packed rows; sRGB byte -> linear table */
inline void aq_pack_material(const aq_material& m, aq_f4* row) {
    union {
        float f;
        int32_t i;
    } t;
    t.i = m.color_tex;
    row[0].x = m.color[0]; row[0].y = m.color[1]; row[0].z = m.color[2]; row[0].w = t.f;
    row[1].x = m.metallic; row[1].y = m.roughness; row[1].z = m.specular; row[1].w = m.specular_tint;
    bool any_ptex = false;
    uint32_t pw[4] = {0u, 0u, 0u, 0u};
    for (int k = 0; k < 16; ++k) {
        any_ptex = any_ptex || m.param_tex[k] != 0;
        pw[k >> 2] |= (uint32_t)m.param_tex[k] << (8 * (k & 3));
    }
    row[2].x = m.sheen; row[2].y = m.sheen_tint; row[2].z = m.transmission; row[2].w = any_ptex ? 1.0f : 0.0f;
    std::memcpy(&row[6], pw, sizeof pw);
    row[3].x = m.emission[0]; row[3].y = m.emission[1]; row[3].z = m.emission[2]; row[3].w = m.anisotropic;
    row[4].x = m.clearcoat; row[4].y = m.clearcoat_roughness; row[4].z = m.ior; row[4].w = m.subsurface;
    row[5].x = m.subsurface_color[0]; row[5].y = m.subsurface_color[1]; row[5].z = m.subsurface_color[2];
    row[5].w = m.anisotropic_rotation;
}
/* does this material need the FULL instantiation of the vertex code? */
inline bool aq_material_needs_full(const aq_material& m) {
    return m.clearcoat > 0.0f || m.transmission > 0.0f || m.subsurface > 0.0f || m.anisotropic > 0.0f || m.param_tex[AQ_PTEX_CLEARCOAT] ||
           m.param_tex[AQ_PTEX_TRANSMISSION] || m.param_tex[AQ_PTEX_SUBSURFACE];
}
/* host: light table (layout above) + per-triangle pick probability / area.  Point lights
 * first (in desc order), then emissive triangles in prim order. */
inline void aq_build_light_table(const aq_scene_desc& d, std::vector<aq_f4>* table, std::vector<float>* prim_pdf) {
    struct L {
        int type;
        uint32_t prim;
        aq_v3 p0, e1, e2, n, rad;
        float area;
        double weight;
    };
    std::vector<L> ls;
    for (uint32_t i = 0; i < d.n_lights; ++i) {
        L l{};
        l.type = 0;
        l.p0 = aq_mk(d.lights[i].pos[0], d.lights[i].pos[1], d.lights[i].pos[2]);
        l.rad = aq_mk(d.lights[i].intensity[0], d.lights[i].intensity[1], d.lights[i].intensity[2]);
        l.weight = 4.0 * 3.14159265358979323846 * (double)aq_lum(l.rad);
        ls.push_back(l);
    }
    prim_pdf->assign(d.n_tris, 0.0f);
    for (uint32_t t = 0; t < d.n_tris; ++t) {
        uint32_t m = d.tri_material ? d.tri_material[t] : 0u;
        if (m >= d.n_materials) continue;
        const float* em = d.materials[m].emission;
        if (!(em[0] > 0.0f || em[1] > 0.0f || em[2] > 0.0f)) continue;
        aq_v3 v0 = aq_mk(d.positions[3 * (size_t)d.indices[3 * (size_t)t]], d.positions[3 * (size_t)d.indices[3 * (size_t)t] + 1],
                         d.positions[3 * (size_t)d.indices[3 * (size_t)t] + 2]);
        aq_v3 v1 = aq_mk(d.positions[3 * (size_t)d.indices[3 * (size_t)t + 1]], d.positions[3 * (size_t)d.indices[3 * (size_t)t + 1] + 1],
                         d.positions[3 * (size_t)d.indices[3 * (size_t)t + 1] + 2]);
        aq_v3 v2 = aq_mk(d.positions[3 * (size_t)d.indices[3 * (size_t)t + 2]], d.positions[3 * (size_t)d.indices[3 * (size_t)t + 2] + 1],
                         d.positions[3 * (size_t)d.indices[3 * (size_t)t + 2] + 2]);
        L l{};
        l.type = 1;
        l.prim = t;
        l.p0 = v0;
        l.e1 = aq_sub(v1, v0);
        l.e2 = aq_sub(v2, v0);
        aq_v3 n = aq_cross(l.e1, l.e2);
        float len = sqrtf(aq_dot(n, n));
        if (!(len > 0.0f)) continue; /* zero-area triangles cannot be sampled */
        l.n = aq_scale(n, 1.0f / len);
        l.area = 0.5f * len;
        l.rad = aq_mk(em[0], em[1], em[2]);
        l.weight = 2.0 * 3.14159265358979323846 * (double)l.area * (double)aq_lum(l.rad);
        ls.push_back(l);
    }
    double total = 0.0;
    for (auto& l : ls) total += l.weight;
    table->assign(ls.size() * AQ_LIGHT_WORDS, aq_f4{0.f, 0.f, 0.f, 0.f});
    double acc = 0.0;
    for (size_t i = 0; i < ls.size(); ++i) {
        const L& l = ls[i];
        double pr = total > 0.0 ? l.weight / total : 1.0 / (double)ls.size();
        acc += pr;
        aq_f4* w = &(*table)[i * AQ_LIGHT_WORDS];
        union {
            float f;
            uint32_t u;
        } pid;
        pid.u = l.prim;
        w[0] = aq_f4{l.p0.x, l.p0.y, l.p0.z, l.type ? 1.0f : 0.0f};
        w[1] = aq_f4{l.e1.x, l.e1.y, l.e1.z, (float)pr};
        w[2] = aq_f4{l.e2.x, l.e2.y, l.e2.z, l.area};
        w[3] = aq_f4{l.rad.x, l.rad.y, l.rad.z, i + 1 == ls.size() ? 2.0f : (float)acc};
        w[4] = aq_f4{l.n.x, l.n.y, l.n.z, pid.f};
        if (l.type) (*prim_pdf)[l.prim] = (float)pr / l.area;
    }
}
inline void aq_build_srgb_lut(float* lut256) {
    for (int i = 0; i < 256; ++i) {
        double c = (double)i / 255.0;
        lut256[i] = (float)(c <= 0.04045 ? c / 12.92 : pow((c + 0.055) / 1.055, 2.4));
    }
}
#endif

AQ_HD aq_v3 aq_ld3(const float* p, uint32_t i) {
    const float* q = p + 3 * (size_t)i;
    return aq_mk(AQ_RO(q), AQ_RO(q + 1), AQ_RO(q + 2));
}
/* a*(1-u-v) + b*u + c*v with a fixed op order */
AQ_HD float aq_bary(float a, float b, float c, float w, float u, float v) {
    return fmaf(c, v, fmaf(b, u, a * w));
}

struct aq_texel_fetch {
    const uint32_t* texels;
    const float* lut;
    uint32_t w;
    AQ_HD aq_v3 operator()(int x, int y) const {
        uint32_t t = AQ_RO(texels + ((size_t)y * w + (size_t)x));
        return aq_mk(AQ_RO(lut + (t & 0xFFu)), AQ_RO(lut + ((t >> 8) & 0xFFu)), AQ_RO(lut + ((t >> 16) & 0xFFu)));
    }
};

/* per-triangle shading inputs, however they were fetched */
struct aq_tri_shading {
    aq_v3 v0, e1, e2;   /* e = plain float subtraction v1-v0, v2-v0 */
    aq_v3 ng;           /* unit geometric normal normalize(e1 x e2), or 0 when degenerate */
    aq_v3 n0, n1, n2;   /* vertex normals (unnormalised); all zero when the mesh has none */
    float uv[6];        /* u0 v0 u1 v1 u2 v2 */
    uint32_t material;
    float light_pdf_area; /* pick probability / area when the triangle is a light, else 0 */
};

/* the arithmetic of a path vertex's geometry + material lookup: one definition, two fetchers */
template <bool FULL>
AQ_HD void aq_finish_vertex(const aq_scene_view& s, const aq_tri_shading& g, float u, float v,
                            aq_v3 ray_d, aq_vertex_in* vi) {
    vi->p = aq_madd(aq_madd(g.v0, g.e1, u), g.e2, v);
    vi->wo = aq_neg(ray_d);
    vi->ng = (g.ng.x != 0.0f || g.ng.y != 0.0f || g.ng.z != 0.0f) ? g.ng : vi->wo;
    float w = 1.0f - u - v;
    vi->ns = aq_mk(aq_bary(g.n0.x, g.n1.x, g.n2.x, w, u, v), aq_bary(g.n0.y, g.n1.y, g.n2.y, w, u, v),
                   aq_bary(g.n0.z, g.n1.z, g.n2.z, w, u, v));
    const aq_f4* mp = s.mats + AQ_MAT_WORDS * (size_t)g.material;
    aq_f4 m0 = aq_ro_f4(mp), m1 = aq_ro_f4(mp + 1), m2 = aq_ro_f4(mp + 2), m3 = aq_ro_f4(mp + 3);
    aq_v3 base = aq_mk(m0.x, m0.y, m0.z);
    union {
        float f;
        int32_t i;
    } tid;
    tid.f = m0.w;
    if (tid.i >= 0 && s.uv) {
        float tu = aq_bary(g.uv[0], g.uv[2], g.uv[4], w, u, v);
        float tv = aq_bary(g.uv[1], g.uv[3], g.uv[5], w, u, v);
        aq_u4 td = aq_ro_u4(s.tex_desc + tid.i);
        aq_texel_fetch tf;
        tf.texels = s.texels + td.z;
        tf.lut = s.srgb_lut;
        tf.w = td.x;
        base = aq_mul(base, aq_tex_bilinear(td.x, td.y, tu, tv, tf));
    }
    vi->mat.base = base;
    vi->mat.metallic = m1.x;
    vi->mat.roughness = m1.y;
    vi->mat.specular = m1.z;
    vi->mat.specular_tint = m1.w;
    vi->mat.sheen = m2.x;
    vi->mat.sheen_tint = m2.y;
    vi->mat.transmission = m2.z;
    if (FULL) {
        aq_f4 m4 = aq_ro_f4(mp + 4), m5 = aq_ro_f4(mp + 5);
        vi->mat.clearcoat = m4.x;
        vi->mat.clearcoat_roughness = m4.y;
        vi->mat.ior = m4.z;
        vi->mat.subsurface = m4.w;
        vi->mat.subsurface_color = aq_mk(m5.x, m5.y, m5.z);
        vi->mat.anisotropic = m3.w;
        vi->mat.anisotropic_rotation = m5.w;
        vi->dpdu = aq_mk(0.0f, 0.0f, 0.0f);
        if (m3.w > 0.0f && s.uv) { /* dp/du of the triangle's uv parametrisation */
            float du1 = g.uv[2] - g.uv[0], dv1 = g.uv[3] - g.uv[1];
            float du2 = g.uv[4] - g.uv[0], dv2 = g.uv[5] - g.uv[1];
            float det = fmaf(du1, dv2, -(du2 * dv1));
            if (det != 0.0f) vi->dpdu = aq_scale(aq_sub(aq_scale(g.e1, dv2), aq_scale(g.e2, dv1)), 1.0f / det);
        }
    }
    if (m2.w != 0.0f && s.uv) { /* Texture::Image on other parameters: constant * texel (rare path) */
        const aq_u4 pt = aq_ro_u4(reinterpret_cast<const aq_u4*>(mp + 6));
        const float tu = aq_bary(g.uv[0], g.uv[2], g.uv[4], w, u, v);
        const float tv = aq_bary(g.uv[1], g.uv[3], g.uv[5], w, u, v);
        auto sample = [&](int slot, aq_v3* rgb) -> bool {
            const uint32_t word = slot < 4 ? pt.x : (slot < 8 ? pt.y : (slot < 12 ? pt.z : pt.w));
            const uint32_t t1 = (word >> (8 * (slot & 3))) & 0xFFu;
            if (t1 == 0u) return false;
            const aq_u4 td = aq_ro_u4(s.tex_desc + (t1 - 1u));
            aq_texel_fetch tf;
            tf.texels = s.texels + td.z;
            tf.lut = s.srgb_lut;
            tf.w = td.x;
            *rgb = aq_tex_bilinear(td.x, td.y, tu, tv, tf);
            return true;
        };
        aq_v3 c;
        if (sample(AQ_PT_METALLIC, &c)) vi->mat.metallic = vi->mat.metallic * c.x;
        if (sample(AQ_PT_ROUGHNESS, &c)) vi->mat.roughness = vi->mat.roughness * c.x;
        if (sample(AQ_PT_SPECULAR, &c)) vi->mat.specular = vi->mat.specular * c.x;
        if (sample(AQ_PT_SPECULAR_TINT, &c)) vi->mat.specular_tint = vi->mat.specular_tint * c.x;
        if (sample(AQ_PT_SHEEN, &c)) vi->mat.sheen = vi->mat.sheen * c.x;
        if (sample(AQ_PT_SHEEN_TINT, &c)) vi->mat.sheen_tint = vi->mat.sheen_tint * c.x;
        if (sample(AQ_PT_TRANSMISSION, &c)) vi->mat.transmission = vi->mat.transmission * c.x;
        if (FULL) {
            if (sample(AQ_PT_CLEARCOAT, &c)) vi->mat.clearcoat = vi->mat.clearcoat * c.x;
            if (sample(AQ_PT_CLEARCOAT_ROUGHNESS, &c)) vi->mat.clearcoat_roughness = vi->mat.clearcoat_roughness * c.x;
            if (sample(AQ_PT_IOR, &c)) vi->mat.ior = vi->mat.ior * c.x;
            if (sample(AQ_PT_SUBSURFACE, &c)) vi->mat.subsurface = vi->mat.subsurface * c.x;
            if (sample(AQ_PT_SUBSURFACE_COLOR, &c)) vi->mat.subsurface_color = aq_mul(vi->mat.subsurface_color, c);
        }
    }
    vi->emission = aq_mk(m3.x, m3.y, m3.z);
    vi->light_pdf_area = g.light_pdf_area;
}

/* fetcher 1: the indexed mesh arrays of aq_scene_desc (what the oracle reads) */
AQ_HD void aq_gather_tri(const aq_scene_view& s, uint32_t prim, aq_tri_shading* g) {
    const uint32_t* ip = s.idx + 3 * (size_t)prim;
    uint32_t i0 = AQ_RO(ip), i1 = AQ_RO(ip + 1), i2 = AQ_RO(ip + 2);
    aq_v3 v0 = aq_ld3(s.pos, i0), v1 = aq_ld3(s.pos, i1), v2 = aq_ld3(s.pos, i2);
    g->v0 = v0;
    g->e1 = aq_sub(v1, v0);
    g->e2 = aq_sub(v2, v0);
    aq_v3 n = aq_cross(g->e1, g->e2);
    float l2 = aq_dot(n, n);
    g->ng = l2 > 0.0f ? aq_scale(n, 1.0f / sqrtf(l2)) : aq_mk(0.0f, 0.0f, 0.0f);
    if (s.nrm) {
        g->n0 = aq_ld3(s.nrm, i0);
        g->n1 = aq_ld3(s.nrm, i1);
        g->n2 = aq_ld3(s.nrm, i2);
    } else {
        g->n0 = g->n1 = g->n2 = aq_mk(0.0f, 0.0f, 0.0f);
    }
    if (s.uv) {
        const float *u0 = s.uv + 2 * (size_t)i0, *u1 = s.uv + 2 * (size_t)i1, *u2 = s.uv + 2 * (size_t)i2;
        g->uv[0] = AQ_RO(u0); g->uv[1] = AQ_RO(u0 + 1);
        g->uv[2] = AQ_RO(u1); g->uv[3] = AQ_RO(u1 + 1);
        g->uv[4] = AQ_RO(u2); g->uv[5] = AQ_RO(u2 + 1);
    } else {
        for (int k = 0; k < 6; ++k) g->uv[k] = 0.0f;
    }
    g->material = AQ_RO(s.tri_mat + prim);
    g->light_pdf_area = s.prim_light_pdf ? AQ_RO(s.prim_light_pdf + prim) : 0.0f;
}

/* fetcher 2: the 128-byte per-triangle shading record the GPU builds at scene-create time
 * from the same arrays (the same float subtractions, stored instead of recomputed): one
 * cache line and no index indirection instead of ~11 scattered sectors.
 *   w0 v0.xyz e1.x | w1 e1.yz e2.xy | w2 e2.z n0.xyz | w3 n1.xyz n2.x | w4 n2.yz uv0 |
 *   w5 uv1 uv2 | w6 material ng.xyz | w7 light pick probability / area (0 = not a light) - - - */
#define AQ_SHADE_REC_WORDS 8
AQ_HD void aq_unpack_shade_rec(const aq_f4* rec, aq_tri_shading* g) {
    aq_f4 w0 = aq_ro_f4(rec + 0), w1 = aq_ro_f4(rec + 1), w2 = aq_ro_f4(rec + 2), w3 = aq_ro_f4(rec + 3),
          w4 = aq_ro_f4(rec + 4), w5 = aq_ro_f4(rec + 5), w6 = aq_ro_f4(rec + 6);
    g->v0 = aq_mk(w0.x, w0.y, w0.z);
    g->e1 = aq_mk(w0.w, w1.x, w1.y);
    g->e2 = aq_mk(w1.z, w1.w, w2.x);
    g->n0 = aq_mk(w2.y, w2.z, w2.w);
    g->n1 = aq_mk(w3.x, w3.y, w3.z);
    g->n2 = aq_mk(w3.w, w4.x, w4.y);
    g->uv[0] = w4.z; g->uv[1] = w4.w;
    g->uv[2] = w5.x; g->uv[3] = w5.y;
    g->uv[4] = w5.z; g->uv[5] = w5.w;
    union {
        float f;
        uint32_t u;
    } m;
    m.f = w6.x;
    g->material = m.u;
    g->ng = aq_mk(w6.y, w6.z, w6.w);
    g->light_pdf_area = aq_ro_f4(rec + 7).x;
}
/* build one record (scene-create time; runs as a kernel over all triangles on the GPU) */
AQ_HD void aq_pack_shade_rec(const aq_tri_shading& g, aq_f4* rec) {
    union {
        float f;
        uint32_t u;
    } m;
    m.u = g.material;
    rec[0].x = g.v0.x; rec[0].y = g.v0.y; rec[0].z = g.v0.z; rec[0].w = g.e1.x;
    rec[1].x = g.e1.y; rec[1].y = g.e1.z; rec[1].z = g.e2.x; rec[1].w = g.e2.y;
    rec[2].x = g.e2.z; rec[2].y = g.n0.x; rec[2].z = g.n0.y; rec[2].w = g.n0.z;
    rec[3].x = g.n1.x; rec[3].y = g.n1.y; rec[3].z = g.n1.z; rec[3].w = g.n2.x;
    rec[4].x = g.n2.y; rec[4].y = g.n2.z; rec[4].z = g.uv[0]; rec[4].w = g.uv[1];
    rec[5].x = g.uv[2]; rec[5].y = g.uv[3]; rec[5].z = g.uv[4]; rec[5].w = g.uv[5];
    rec[6].x = m.f; rec[6].y = g.ng.x; rec[6].z = g.ng.y; rec[6].w = g.ng.z;
    rec[7].x = g.light_pdf_area;
    rec[7].y = rec[7].z = rec[7].w = 0.0f;
}

/* hit (prim,u,v) + incoming direction -> everything aq_shade_vertex needs */
template <bool FULL>
AQ_HD void aq_fetch_vertex(const aq_scene_view& s, uint32_t prim, float u, float v, aq_v3 ray_d,
                           aq_vertex_in* vi) {
    aq_tri_shading g;
    if (s.shade_recs)
        aq_unpack_shade_rec(s.shade_recs + (size_t)prim * AQ_SHADE_REC_WORDS, &g);
    else
        aq_gather_tri(s, prim, &g);
    aq_finish_vertex<FULL>(s, g, u, v, ray_d, vi);
}

#endif /* AQ_CORE_H */
