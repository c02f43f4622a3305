/*
 * aq_nrc.h — DEFINITIONAL functions of the `nrc` integrator (neural radiance cache), the
 * integrator type scenes/integrator.json names (:2 "type":"nrc", :4 batch_size, :6
 * training_iters, :7 learning_rate, :8 visualize_cache).
 *
 * The reference snapshot holds no implementation (src/lib.rs:0), so — as for the path tracer in
 * aq_core.h — this header DEFINES the semantics once, as __host__ __device__ code with a fixed
 * floating-point operation order; the sm_100a kernels (aq_nrc.cuh) and the CPU oracle
 * (oracle/) both instantiate it and agree bit for bit.  The method follows Mueller
 * et al., "Real-time Neural Radiance Caching for Path Tracing" (SIGGRAPH 2021), simplified:
 *
 *   cache    a small multilayer perceptron  x in R^64 -> y in R^3,  AQ_NRC_HIDDEN_LAYERS hidden
 *            layers of 64 ReLU neurons, no biases (x carries a constant 1), fp32
 *   query    L_s(p, wo) ~= fac * max(y, 0): radiance SCATTERED at surface point p towards wo
 *            (direct + indirect, without p's own emission), factored by the albedo-like
 *            fac = max(diffuse albedo + specular F0, 0.02)
 *   records  training_iters * batch_size path suffixes: record r starts at the camera through a
 *            hashed pixel, is traced to its first (r even) or second (r odd) hit with the
 *            ordinary vertex code, and from there — throughput reset to 1, radiance to 0, the
 *            vertex's own emission skipped — to the integrator's max_depth with next-event
 *            estimation and Russian roulette as usual.  The radiance it collects is an unbiased
 *            estimate of L_s at the record vertex; target y* = L / fac.
 *   training plain minibatch descent with Adam on the relative squared error
 *            sum_c (y_c - y*_c)^2 / (y_c^2 + 0.01)  (the denominator is treated as a constant),
 *            iteration i uses records [i*B, (i+1)*B).  The records do not depend on the weights
 *            (no self-training), so they are generated once.
 *   render   camera ray, one ordinary vertex with next-event estimation, one bounce, then the
 *            cache is queried at the second hit:  L += beta * fac * max(y, 0)  (plus that
 *            vertex's own emission, MIS-weighted as usual).  visualize_cache = true queries at
 *            the first hit instead, so the image shows the cache itself.
 *
 * Every sum has a fixed order: dot products accumulate with fmaf in ascending index order
 * starting from 0; weight gradients accumulate over the samples of a 64-sample chunk in
 * ascending order and then over the chunks in ascending order; only + - * / sqrt fma are used.
 */
#ifndef AQ_NRC_H
#define AQ_NRC_H

#include "aq_core.h"

#define AQ_NRC_IN 64            /* input features */
#define AQ_NRC_WIDTH 64         /* neurons per hidden layer */
#define AQ_NRC_HIDDEN_LAYERS 4  /* hidden layers => 5 weight matrices */
#define AQ_NRC_OUT 3
#define AQ_NRC_OUT_PAD 4        /* output matrix is 64 x 4, column 3 unused (kept at 0) */
#ifndef AQ_NRC_CHUNK
#define AQ_NRC_CHUNK 16         /* samples whose gradient is accumulated by one CTA / in one pass (round 1: 64 = 8 busy
                                 * SMs at the reference's batch of 512; 16 = 32 CTAs, same arithmetic per record,
                                 * the per-chunk summation order is part of the definition and the oracle follows it) */
#endif
#define AQ_NRC_N_MATS (AQ_NRC_HIDDEN_LAYERS + 1)
/* weights are stored input-major: W_l[i][j] at l*4096 + i*64 + j for the 64x64 matrices,
 * then the 64x4 output matrix W_out[i][c] at 4*4096 + i*4 + c */
#define AQ_NRC_N_WEIGHTS (AQ_NRC_HIDDEN_LAYERS * AQ_NRC_WIDTH * AQ_NRC_WIDTH + AQ_NRC_WIDTH * AQ_NRC_OUT_PAD)
#define AQ_NRC_MAT_OFF(l) ((l) * AQ_NRC_WIDTH * AQ_NRC_WIDTH)
#define AQ_NRC_MAT_COLS(l) ((l) < AQ_NRC_HIDDEN_LAYERS ? AQ_NRC_WIDTH : AQ_NRC_OUT_PAD)

#define AQ_NRC_POS_OCTAVES 8
#define AQ_NRC_BLOB_BINS 4
#define AQ_NRC_FAC_MIN 0.02f
#define AQ_NRC_LOSS_EPS 0.01f
#define AQ_NRC_ADAM_B1 0.9f
#define AQ_NRC_ADAM_B2 0.99f
#define AQ_NRC_ADAM_EPS 1.0e-8f
#define AQ_NRC_SEED_SALT 0xA5A5A5A5u

/* ------------------------------------------------------------------ records: which path */
/* pixel of training record r: uniform over the image from the counter-based hash */
AQ_HD uint32_t aq_nrc_record_pixel(uint32_t seed, uint32_t r, uint32_t npix) {
    uint32_t h = aq_rng_u32(aq_rng_key(seed ^ AQ_NRC_SEED_SALT, r, 0x4E5243u), 0u);
    return (uint32_t)(((uint64_t)h * (uint64_t)npix) >> 32);
}
/* RNG key of the path behind training record r (sample index = r, seed salted) */
AQ_HD uint32_t aq_nrc_record_key(uint32_t seed, uint32_t pixel, uint32_t r) {
    return aq_rng_key(seed ^ AQ_NRC_SEED_SALT, pixel, r);
}
/* depth index of the vertex record r describes: 0 = first hit, 1 = second hit */
AQ_HD uint32_t aq_nrc_record_depth(uint32_t r) { return r & 1u; }

/* ------------------------------------------------------------------ input encoding
 *   [0..2]    p normalised to the scene bounds
 *   [3..26]   triangle wave |2 frac(2^k p) - 1|, k = 0..7, per axis
 *   [27..38]  one-blob (4 bins, quartic kernel) of (wo + 1)/2 per component
 *   [39..50]  one-blob of (ns + 1)/2 per component (shading normal on wo's side)
 *   [51..54]  one-blob of roughness
 *   [55..57]  diffuse albedo  base * (1-metallic)(1-transmission)
 *   [58..60]  specular F0
 *   [61]      1 (bias input)     [62],[63]  0
 * fac = max(diffuse albedo + F0, AQ_NRC_FAC_MIN) */
AQ_HD float aq_nrc_frac(float x) { return x - floorf(x); }
AQ_HD void aq_nrc_oneblob(float v, float* out, int stride) {
    float x = aq_clampf(v, 0.0f, 1.0f);
#pragma unroll
    for (int b = 0; b < AQ_NRC_BLOB_BINS; ++b) {
        float t = (x - ((float)b + 0.5f) * (1.0f / AQ_NRC_BLOB_BINS)) * (float)AQ_NRC_BLOB_BINS;
        float q = 1.0f - t * t;
        out[b * stride] = q > 0.0f ? q * q : 0.0f;
    }
}

struct aq_nrc_bounds {
    aq_v3 lo, inv_ext; /* inv_ext = 1 / max(hi - lo, tiny) per axis (host-computed) */
};

/* x: 64 features written at x[k*stride] */
AQ_HD void aq_nrc_encode(const aq_vertex_in& vi, const aq_nrc_bounds& bb, float* x, int stride, aq_v3* fac) {
    aq_v3 ng, ns;
    aq_orient_normals(vi, &ng, &ns);
    aq_v3 wo0 = aq_mk(0.0f, 0.0f, 1.0f);
    aq_bsdf_ctx c;
    aq_bsdf_setup_base(vi.mat, wo0, &c);
    aq_v3 alb = aq_scale(vi.mat.base, c.diff_w);
    float pn[3] = {aq_clampf((vi.p.x - bb.lo.x) * bb.inv_ext.x, 0.0f, 1.0f),
                   aq_clampf((vi.p.y - bb.lo.y) * bb.inv_ext.y, 0.0f, 1.0f),
                   aq_clampf((vi.p.z - bb.lo.z) * bb.inv_ext.z, 0.0f, 1.0f)};
#pragma unroll
    for (int a = 0; a < 3; ++a) x[a * stride] = pn[a];
    float f = 1.0f;
    for (int k = 0; k < AQ_NRC_POS_OCTAVES; ++k) {
#pragma unroll
        for (int a = 0; a < 3; ++a) x[(3 + 3 * k + a) * stride] = fabsf(fmaf(2.0f, aq_nrc_frac(pn[a] * f), -1.0f));
        f = f * 2.0f;
    }
    aq_nrc_oneblob(fmaf(vi.wo.x, 0.5f, 0.5f), x + 27 * stride, stride);
    aq_nrc_oneblob(fmaf(vi.wo.y, 0.5f, 0.5f), x + 31 * stride, stride);
    aq_nrc_oneblob(fmaf(vi.wo.z, 0.5f, 0.5f), x + 35 * stride, stride);
    aq_nrc_oneblob(fmaf(ns.x, 0.5f, 0.5f), x + 39 * stride, stride);
    aq_nrc_oneblob(fmaf(ns.y, 0.5f, 0.5f), x + 43 * stride, stride);
    aq_nrc_oneblob(fmaf(ns.z, 0.5f, 0.5f), x + 47 * stride, stride);
    aq_nrc_oneblob(vi.mat.roughness, x + 51 * stride, stride);
    x[55 * stride] = alb.x; x[56 * stride] = alb.y; x[57 * stride] = alb.z;
    x[58 * stride] = c.f0.x; x[59 * stride] = c.f0.y; x[60 * stride] = c.f0.z;
    x[61 * stride] = 1.0f;
    x[62 * stride] = 0.0f;
    x[63 * stride] = 0.0f;
    *fac = aq_mk(aq_maxf(alb.x + c.f0.x, AQ_NRC_FAC_MIN), aq_maxf(alb.y + c.f0.y, AQ_NRC_FAC_MIN),
                 aq_maxf(alb.z + c.f0.z, AQ_NRC_FAC_MIN));
}

/* ------------------------------------------------------------------ the arithmetic of the MLP */
/* sum_{k<n} a[k*sa] * b[k*sb], ascending k, fmaf chain from 0: THE summation order of every
 * matrix product of the cache (forward, backward-data and weight gradient) */
AQ_HD float aq_nrc_dot(const float* a, int sa, const float* b, int sb, int n) {
    float acc = 0.0f;
    for (int k = 0; k < n; ++k) acc = fmaf(a[k * sa], b[k * sb], acc);
    return acc;
}
AQ_HD float aq_nrc_relu(float v) { return v > 0.0f ? v : 0.0f; }

/* d loss / d y_c for one sample: loss = sum_c (y_c - t_c)^2 / (y_c^2 + eps) * inv_norm,
 * denominator constant; inv_norm = 1 / (3 * batch_size) */
AQ_HD float aq_nrc_loss_grad(float y, float t, float inv_norm) {
    return 2.0f * (y - t) / fmaf(y, y, AQ_NRC_LOSS_EPS) * inv_norm;
}
AQ_HD float aq_nrc_loss_term(float y, float t, float inv_norm) {
    float d = y - t;
    return d * d / fmaf(y, y, AQ_NRC_LOSS_EPS) * inv_norm;
}

/* Adam step for one weight; bc1 = 1/(1-b1^t), bc2 = 1/(1-b2^t) are computed on the host */
AQ_HD void aq_nrc_adam(float g, float lr, float bc1, float bc2, float* w, float* m, float* v) {
    float mm = fmaf(AQ_NRC_ADAM_B1, *m, (1.0f - AQ_NRC_ADAM_B1) * g);
    float vv = fmaf(AQ_NRC_ADAM_B2, *v, (1.0f - AQ_NRC_ADAM_B2) * (g * g));
    *m = mm;
    *v = vv;
    *w = *w - lr * (mm * bc1) / (sqrtf(vv * bc2) + AQ_NRC_ADAM_EPS);
}

/* initial weight k (flat index): uniform in +-sqrt(6 / (fan_in + fan_out)) from the hash; the
 * unused 4th output column is 0 */
AQ_HD float aq_nrc_init_weight(uint32_t seed, uint32_t k) {
    const uint32_t hid = AQ_NRC_HIDDEN_LAYERS * AQ_NRC_WIDTH * AQ_NRC_WIDTH;
    float bound = 0.21650635f; /* sqrt(6/128) */
    if (k >= hid) {
        if (((k - hid) & 3u) == 3u) return 0.0f;
        bound = 0.29924238f; /* sqrt(6/67) */
    }
    float u = aq_rng(aq_rng_key(seed ^ AQ_NRC_SEED_SALT, k, 0x57454947u), 0u);
    return fmaf(2.0f * bound, u, -bound);
}

/* ---- host-only helpers (plain inline: host code in both nvcc passes) */
/* host: bias corrections of Adam at iteration t (1-based), evaluated in double */
inline void aq_nrc_adam_bias(uint32_t t, float* bc1, float* bc2) {
    *bc1 = (float)(1.0 / (1.0 - pow((double)AQ_NRC_ADAM_B1, (double)t)));
    *bc2 = (float)(1.0 / (1.0 - pow((double)AQ_NRC_ADAM_B2, (double)t)));
}
/* host: bounds of the vertex positions -> aq_nrc_bounds */
inline aq_nrc_bounds aq_nrc_bounds_of(const float* positions, uint32_t n_verts) {
    float lo[3] = {0.f, 0.f, 0.f}, hi[3] = {1.f, 1.f, 1.f};
    for (uint32_t v = 0; v < n_verts; ++v)
        for (int a = 0; a < 3; ++a) {
            float p = positions[3 * (size_t)v + a];
            if (v == 0 || p < lo[a]) lo[a] = p;
            if (v == 0 || p > hi[a]) hi[a] = p;
        }
    aq_nrc_bounds b;
    b.lo = aq_mk(lo[0], lo[1], lo[2]);
    float e[3];
    for (int a = 0; a < 3; ++a) e[a] = hi[a] - lo[a] > 1.0e-20f ? 1.0f / (hi[a] - lo[a]) : 1.0f;
    b.inv_ext = aq_mk(e[0], e[1], e[2]);
    return b;
}

#endif /* AQ_NRC_H */
