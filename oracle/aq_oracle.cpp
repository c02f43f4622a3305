/*
 * aq_oracle.cpp — CPU ORACLE for the render hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  The product (libaqua_cuda.so) never links or calls it.
 *
 * PARITY UNPINNED: the reference snapshot has no implementation of this path
 * (/root/reference/src/lib.rs:0 is an empty file), no tests, no golden vectors and no
 * reference images.  The oracle therefore *defines* the semantics, from
 * the reference's DATA contract:
 *   camera      Camera::Perspective      scenes/cbox.json:517-542
 *   light       Light::Point             scenes/cbox.json:545-559
 *   material    Bsdf::Principled         scenes/cbox.json:4-65
 *   geometry    TriangleMesh (.mesh)     scenes/ *.mesh
 *   integrator  spp / max_depth          scenes/integrator.json:3,5
 * and is pinned only by the loader facts the reference data provides (tests/test_loader.py)
 * and by analytic checks (tests/test_bsdf.py: furnace, pdf normalisation, closed-form
 * one-bounce radiance).
 *
 * The scalar definitions (triangle test, camera, BSDF, RNG, vertex shading) are the
 * single-sourced host/device functions of aqua-engine_b200/csrc/aq_core.h, compiled here
 * with -ffp-contract=off.  Everything about HOW they are driven is separate from the GPU:
 * a depth-first loop per path (the `rayon` par_iter of the original CPU renderer,
 * Cargo.toml:12), a brute-force all-triangles intersector (the definitive one) and an
 * independent median-split BVH2 for larger runs.  The GPU's BVH8 builder and wavefront
 * are not used.
 */
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

#include "aqua_cuda.h"
/* aqua_cuda.h first: aq_core.h then also defines the host-only material packing */
#include "aq_core.h"
#include "aq_nrc.h"
#include "aq_bvh.h" /* only for the optional host walk of a downloaded BVH8 (aqo_bvh8_intersect) */

namespace {

struct Tri {
    aq_v3 v0, e1, e2;
};

/* ---- independent BVH2: median split on the largest centroid axis, leaves <= 4 */
struct ONode {
    float lo[3], hi[3];
    uint32_t left, right, first, count;
};

struct OBvh {
    std::vector<ONode> nodes;
    std::vector<uint32_t> order;
};

struct Oracle {
    aq_scene_desc d;
    std::vector<Tri> tris;
    std::vector<aq_f4> mats;
    std::vector<aq_u4> tex_desc;
    std::vector<uint32_t> texels;
    std::vector<float> lut, prim_light_pdf;
    std::vector<aq_f4> lights;
    aq_scene_view view;
    OBvh bvh;
    bool has_bvh = false;
    bool full_bsdf = false; /* some material needs the FULL vertex code (aq_material_needs_full) */
    /* copies of the caller's arrays so the handle outlives them */
    std::vector<float> pos, nrm, uv;
    std::vector<uint32_t> idx, tri_mat;
};

void build_obvh(Oracle& O) {
    uint32_t n = (uint32_t)O.tris.size();
    OBvh& B = O.bvh;
    B.order.resize(n);
    std::vector<float> lo(3 * (size_t)n), hi(3 * (size_t)n), ce(3 * (size_t)n);
    float scale = 0.f;
    for (uint32_t i = 0; i < n; ++i) {
        const Tri& t = O.tris[i];
        aq_v3 v1 = aq_add(t.v0, t.e1), v2 = aq_add(t.v0, t.e2);
        /* use the original vertices for the bounds, not v0+e */
        const float* p0 = &O.pos[3 * (size_t)O.idx[3 * (size_t)i]];
        const float* p1 = &O.pos[3 * (size_t)O.idx[3 * (size_t)i + 1]];
        const float* p2 = &O.pos[3 * (size_t)O.idx[3 * (size_t)i + 2]];
        (void)v1;
        (void)v2;
        for (int a = 0; a < 3; ++a) {
            float l = std::min(p0[a], std::min(p1[a], p2[a]));
            float h = std::max(p0[a], std::max(p1[a], p2[a]));
            lo[3 * (size_t)i + a] = l;
            hi[3 * (size_t)i + a] = h;
            ce[3 * (size_t)i + a] = 0.5f * (l + h);
            scale = std::max(scale, std::max(std::fabs(l), std::fabs(h)));
        }
        B.order[i] = i;
    }
    const float pad = 2.0e-5f * std::max(scale, 1e-30f);
    B.nodes.clear();
    B.nodes.reserve(2 * (size_t)n + 1);
    struct Job {
        uint32_t node, b, e;
    };
    std::vector<Job> stack;
    B.nodes.push_back(ONode{});
    stack.push_back({0u, 0u, n});
    while (!stack.empty()) {
        Job j = stack.back();
        stack.pop_back();
        float nlo[3] = {INFINITY, INFINITY, INFINITY}, nhi[3] = {-INFINITY, -INFINITY, -INFINITY};
        float clo[3] = {INFINITY, INFINITY, INFINITY}, chi[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (uint32_t i = j.b; i < j.e; ++i) {
            uint32_t p = B.order[i];
            for (int a = 0; a < 3; ++a) {
                nlo[a] = std::min(nlo[a], lo[3 * (size_t)p + a]);
                nhi[a] = std::max(nhi[a], hi[3 * (size_t)p + a]);
                clo[a] = std::min(clo[a], ce[3 * (size_t)p + a]);
                chi[a] = std::max(chi[a], ce[3 * (size_t)p + a]);
            }
        }
        ONode nd{};
        for (int a = 0; a < 3; ++a) {
            nd.lo[a] = nlo[a] - pad;
            nd.hi[a] = nhi[a] + pad;
        }
        uint32_t cnt = j.e - j.b;
        if (cnt <= 4) {
            nd.first = j.b;
            nd.count = cnt;
            B.nodes[j.node] = nd;
            continue;
        }
        int ax = 0;
        if (chi[1] - clo[1] > chi[ax] - clo[ax]) ax = 1;
        if (chi[2] - clo[2] > chi[ax] - clo[ax]) ax = 2;
        uint32_t mid = j.b + cnt / 2;
        std::nth_element(B.order.begin() + j.b, B.order.begin() + mid, B.order.begin() + j.e,
                         [&](uint32_t a, uint32_t b) {
                             return ce[3 * (size_t)a + ax] < ce[3 * (size_t)b + ax];
                         });
        nd.count = 0;
        nd.left = (uint32_t)B.nodes.size();
        nd.right = nd.left + 1;
        B.nodes[j.node] = nd;
        B.nodes.push_back(ONode{});
        B.nodes.push_back(ONode{});
        stack.push_back({nd.left, j.b, mid});
        stack.push_back({nd.right, mid, j.e});
    }
    O.has_bvh = true;
}

inline bool slab(const ONode& n, aq_v3 o, aq_v3 id, float tmin, float tmax) {
    float t0 = tmin, t1 = tmax;
    const float oo[3] = {o.x, o.y, o.z}, ii[3] = {id.x, id.y, id.z};
    for (int a = 0; a < 3; ++a) {
        float ta = (n.lo[a] - oo[a]) * ii[a], tb = (n.hi[a] - oo[a]) * ii[a];
        if (ta > tb) std::swap(ta, tb);
        t0 = std::max(t0, ta);
        t1 = std::min(t1, tb);
    }
    return t0 <= t1 * 1.0000004f;
}

/* closest hit: lexicographic min of (t, prim), tmin < t <= tmax */
void closest_brute(const Oracle& O, aq_v3 o, aq_v3 d, float tmin, float tmax, aq_hit* h) {
    uint32_t bp = AQ_MISS_ID;
    float bt = tmax, bu = 0.f, bv = 0.f;
    uint32_t n = (uint32_t)O.tris.size();
    for (uint32_t i = 0; i < n; ++i) {
        const Tri& T = O.tris[i];
        float t, u, v;
        if (aq_tri_test(o, d, tmin, T.v0, T.e1, T.e2, &t, &u, &v) && t <= tmax &&
            aq_hit_closer(t, i, bt, bp)) {
            bt = t;
            bp = i;
            bu = u;
            bv = v;
        }
    }
    h->prim = bp;
    h->t = bp == AQ_MISS_ID ? tmax : bt;
    h->u = bu;
    h->v = bv;
}

bool any_brute(const Oracle& O, aq_v3 o, aq_v3 d, float tmin, float tmax) {
    uint32_t n = (uint32_t)O.tris.size();
    for (uint32_t i = 0; i < n; ++i) {
        const Tri& T = O.tris[i];
        float t, u, v;
        if (aq_tri_test(o, d, tmin, T.v0, T.e1, T.e2, &t, &u, &v) && t < tmax) return true;
    }
    return false;
}

template <bool ANY>
bool trace_bvh(const Oracle& O, aq_v3 o, aq_v3 d, float tmin, float tmax, aq_hit* h) {
    aq_v3 id = aq_mk(aq_safe_rcp_dir(d.x), aq_safe_rcp_dir(d.y), aq_safe_rcp_dir(d.z));
    uint32_t bp = AQ_MISS_ID;
    float bt = tmax, bu = 0.f, bv = 0.f;
    uint32_t st[128];
    int sp = 0;
    st[sp++] = 0;
    while (sp) {
        const ONode& n = O.bvh.nodes[st[--sp]];
        if (!slab(n, o, id, tmin, bt)) continue;
        if (n.count) {
            for (uint32_t k = 0; k < n.count; ++k) {
                uint32_t i = O.bvh.order[n.first + k];
                const Tri& T = O.tris[i];
                float t, u, v;
                if (!aq_tri_test(o, d, tmin, T.v0, T.e1, T.e2, &t, &u, &v)) continue;
                if (ANY) {
                    if (t < tmax) return true;
                } else if (t <= tmax && aq_hit_closer(t, i, bt, bp)) {
                    bt = t;
                    bp = i;
                    bu = u;
                    bv = v;
                }
            }
        } else {
            st[sp++] = n.left;
            st[sp++] = n.right;
        }
    }
    if (!ANY) {
        h->prim = bp;
        h->t = bp == AQ_MISS_ID ? tmax : bt;
        h->u = bu;
        h->v = bv;
    }
    return bp != AQ_MISS_ID;
}

void closest(const Oracle& O, bool use_bvh, aq_v3 o, aq_v3 d, float tmin, float tmax, aq_hit* h) {
    if (use_bvh)
        trace_bvh<false>(O, o, d, tmin, tmax, h);
    else
        closest_brute(O, o, d, tmin, tmax, h);
}
bool occluded(const Oracle& O, bool use_bvh, aq_v3 o, aq_v3 d, float tmin, float tmax) {
    aq_hit h;
    return use_bvh ? trace_bvh<true>(O, o, d, tmin, tmax, &h) : any_brute(O, o, d, tmin, tmax);
}

template <class F>
void parallel_for(uint64_t n, int n_threads, uint64_t chunk, F&& f) {
    if (n_threads <= 0) {
        n_threads = (int)std::thread::hardware_concurrency();
        if (n_threads <= 0) n_threads = 1;
    }
    std::atomic<uint64_t> next{0};
    auto worker = [&](int tid) {
        for (;;) {
            uint64_t b = next.fetch_add(chunk);
            if (b >= n) break;
            uint64_t e = std::min(n, b + chunk);
            f(b, e, tid);
        }
    };
    if (n_threads == 1) {
        worker(0);
        return;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(worker, t);
    for (auto& t : th) t.join();
}

/* walk a BVH8 (as built by the product, downloaded with aq_accel_download) on the CPU with
 * the same traversal template the kernel instantiates — isolates builder bugs from kernel
 * bugs in the tests */
template <bool STEP2>
static int bvh8_intersect_impl(const void* nodes, const void* tris, const aq_ray* rays, uint32_t n,
                       aq_hit* hits, int any_hit, int n_threads, uint64_t* nodes_fetched,
                       uint64_t* tris_fetched) {
    std::atomic<uint64_t> nn{0}, nt{0};
    parallel_for(n, n_threads, 256, [&](uint64_t b, uint64_t e, int) {
        aq_local_stack st;
        aq_trav_counters c{0, 0};
        uint64_t ln = 0, lt = 0;
        for (uint64_t i = b; i < e; ++i) {
            const aq_ray& r = rays[i];
            aq_v3 o = aq_mk(r.o[0], r.o[1], r.o[2]), d = aq_mk(r.d[0], r.d[1], r.d[2]);
            uint32_t prim;
            float t, u, v;
            c.nodes = c.tris = 0;
            if (any_hit) {
                bool occ = aq_bvh8_trace<true, true, STEP2>((const aq_u4*)nodes, (const aq_f4*)tris, o, d,
                                                     r.tmin, r.tmax, st, prim, t, u, v, &c);
                hits[i].prim = occ ? 0u : AQ_MISS_ID;
                hits[i].t = hits[i].u = hits[i].v = 0.f;
            } else {
                aq_bvh8_trace<false, true, STEP2>((const aq_u4*)nodes, (const aq_f4*)tris, o, d, r.tmin,
                                           r.tmax, st, prim, t, u, v, &c);
                hits[i].prim = prim;
                hits[i].t = prim == AQ_MISS_ID ? r.tmax : t;
                hits[i].u = u;
                hits[i].v = v;
            }
            ln += c.nodes;
            lt += c.tris;
        }
        nn += ln;
        nt += lt;
    });
    if (nodes_fetched) *nodes_fetched = nn.load();
    if (tris_fetched) *tris_fetched = nt.load();
    return AQ_OK;
}


/* ---- nrc integrator (aq_nrc.h): plain loops in the defined summation order */
struct NrcActs {
    float a[AQ_NRC_N_MATS][AQ_NRC_WIDTH]; /* a[0] = input, a[1..4] = hidden (post-ReLU) */
    float y[AQ_NRC_OUT];
};
void nrc_forward(const float* W, const float* x, NrcActs* A) {
    for (int i = 0; i < AQ_NRC_IN; ++i) A->a[0][i] = x[i];
    for (int l = 0; l < AQ_NRC_HIDDEN_LAYERS; ++l) {
        const float* Wl = W + AQ_NRC_MAT_OFF(l);
        for (int j = 0; j < AQ_NRC_WIDTH; ++j)
            A->a[l + 1][j] = aq_nrc_relu(aq_nrc_dot(A->a[l], 1, Wl + j, AQ_NRC_WIDTH, AQ_NRC_WIDTH));
    }
    const float* Wo = W + AQ_NRC_MAT_OFF(AQ_NRC_HIDDEN_LAYERS);
    for (int c = 0; c < AQ_NRC_OUT; ++c)
        A->y[c] = aq_nrc_dot(A->a[AQ_NRC_HIDDEN_LAYERS], 1, Wo + c, AQ_NRC_OUT_PAD, AQ_NRC_WIDTH);
}

template <bool FULL>
void nrc_shade_dispatch(bool has_area, const aq_vertex_in& vi, aq_v3 beta, uint32_t key, uint32_t depth,
                        uint32_t max_depth, const aq_scene_view& V, uint32_t mis_mode, aq_vertex_out* vo) {
    if (has_area)
        aq_shade_vertex<true, FULL>(vi, beta, key, depth, max_depth, V.n_lights, V.lights, mis_mode, vo);
    else
        aq_shade_vertex<false, FULL>(vi, beta, key, depth, max_depth, V.n_lights, V.lights, mis_mode, vo);
}
}  // namespace

extern "C" {

typedef struct aqo_scene aqo_scene;

int aqo_threads(void) {
    int n = (int)std::thread::hardware_concurrency();
    return n > 0 ? n : 1;
}

/* copies everything it needs out of desc */
int aqo_scene_create(const aq_scene_desc* d, int build_bvh, aqo_scene** out) {
    if (!d || !out) return AQ_ERR_BAD_ARG;
    Oracle* O = new Oracle;
    O->d = *d;
    O->pos.assign(d->positions, d->positions + 3 * (size_t)d->n_verts);
    if (d->normals) O->nrm.assign(d->normals, d->normals + 3 * (size_t)d->n_verts);
    if (d->uvs) O->uv.assign(d->uvs, d->uvs + 2 * (size_t)d->n_verts);
    O->idx.assign(d->indices, d->indices + 3 * (size_t)d->n_tris);
    if (d->tri_material)
        O->tri_mat.assign(d->tri_material, d->tri_material + d->n_tris);
    else
        O->tri_mat.assign(d->n_tris, 0u);
    O->tris.resize(d->n_tris);
    for (uint32_t i = 0; i < d->n_tris; ++i) {
        aq_v3 v0 = aq_ld3(O->pos.data(), O->idx[3 * (size_t)i]);
        aq_v3 v1 = aq_ld3(O->pos.data(), O->idx[3 * (size_t)i + 1]);
        aq_v3 v2 = aq_ld3(O->pos.data(), O->idx[3 * (size_t)i + 2]);
        O->tris[i].v0 = v0;
        O->tris[i].e1 = aq_sub(v1, v0);
        O->tris[i].e2 = aq_sub(v2, v0);
    }
    O->mats.assign(AQ_MAT_WORDS * (size_t)std::max(1u, d->n_materials), aq_f4{0.f, 0.f, 0.f, 0.f});
    for (uint32_t m = 0; m < d->n_materials; ++m) {
        aq_pack_material(d->materials[m], &O->mats[AQ_MAT_WORDS * (size_t)m]);
        O->full_bsdf = O->full_bsdf || aq_material_needs_full(d->materials[m]);
    }
    size_t off = 0;
    for (uint32_t t = 0; t < d->n_textures; ++t) {
        aq_u4 td;
        td.x = d->textures[t].width;
        td.y = d->textures[t].height;
        td.z = (uint32_t)off;
        td.w = 0;
        O->tex_desc.push_back(td);
        size_t n = (size_t)td.x * td.y;
        O->texels.resize(off + n);
        std::memcpy(&O->texels[off], d->textures[t].rgba8, n * 4);
        off += n;
    }
    O->lut.resize(256);
    aq_build_srgb_lut(O->lut.data());
    {
        aq_scene_desc dd = *d;
        dd.tri_material = O->tri_mat.data();
        aq_build_light_table(dd, &O->lights, &O->prim_light_pdf);
    }
    aq_scene_view& V = O->view;
    V.pos = O->pos.data();
    V.nrm = O->nrm.empty() ? nullptr : O->nrm.data();
    V.uv = O->uv.empty() ? nullptr : O->uv.data();
    V.idx = O->idx.data();
    V.tri_mat = O->tri_mat.data();
    V.mats = O->mats.data();
    V.tex_desc = O->tex_desc.data();
    V.texels = O->texels.data();
    V.srgb_lut = O->lut.data();
    V.lights = O->lights.data();
    V.n_lights = (uint32_t)(O->lights.size() / AQ_LIGHT_WORDS);
    V.prim_light_pdf = O->prim_light_pdf.empty() ? nullptr : O->prim_light_pdf.data();
    V.shade_recs = nullptr; /* the oracle reads the indexed mesh arrays */
    if (build_bvh) build_obvh(*O);
    *out = reinterpret_cast<aqo_scene*>(O);
    return AQ_OK;
}

void aqo_scene_destroy(aqo_scene* s) { delete reinterpret_cast<Oracle*>(s); }

/* mode 0 = brute force over all triangles, 1 = oracle BVH2 */
int aqo_intersect(aqo_scene* s, const aq_ray* rays, uint32_t n, aq_hit* hits, int any_hit, int mode,
                  int n_threads) {
    Oracle* O = reinterpret_cast<Oracle*>(s);
    if (!O || !rays || !hits) return AQ_ERR_BAD_ARG;
    if (mode == 1 && !O->has_bvh) build_obvh(*O);
    parallel_for(n, n_threads, 256, [&](uint64_t b, uint64_t e, int) {
        for (uint64_t i = b; i < e; ++i) {
            const aq_ray& r = rays[i];
            aq_v3 o = aq_mk(r.o[0], r.o[1], r.o[2]), d = aq_mk(r.d[0], r.d[1], r.d[2]);
            if (any_hit) {
                bool occ = occluded(*O, mode == 1, o, d, r.tmin, r.tmax);
                hits[i].prim = occ ? 0u : AQ_MISS_ID;
                hits[i].t = 0.f;
                hits[i].u = 0.f;
                hits[i].v = 0.f;
            } else {
                closest(*O, mode == 1, o, d, r.tmin, r.tmax, &hits[i]);
            }
        }
    });
    return AQ_OK;
}

/* step = 0: aq_trav_step (node visit, then all of its triangles), 1: aq_trav_step2 (interleaved) */
int aqo_bvh8_intersect_step(const void* nodes, const void* tris, const aq_ray* rays, uint32_t n, aq_hit* hits,
                            int any_hit, int n_threads, uint64_t* nodes_fetched, uint64_t* tris_fetched, int step) {
    return step ? bvh8_intersect_impl<true>(nodes, tris, rays, n, hits, any_hit, n_threads, nodes_fetched, tris_fetched)
                : bvh8_intersect_impl<false>(nodes, tris, rays, n, hits, any_hit, n_threads, nodes_fetched, tris_fetched);
}
int aqo_bvh8_intersect(const void* nodes, const void* tris, const aq_ray* rays, uint32_t n, aq_hit* hits,
                       int any_hit, int n_threads, uint64_t* nodes_fetched, uint64_t* tris_fetched) {
    return aqo_bvh8_intersect_step(nodes, tris, rays, n, hits, any_hit, n_threads, nodes_fetched, tris_fetched, 0);
}

int aqo_camera_rays(aqo_scene* s, const aq_integrator_cfg* cfg, uint32_t sample, aq_ray* out) {
    Oracle* O = reinterpret_cast<Oracle*>(s);
    if (!O || !cfg || !out) return AQ_ERR_BAD_ARG;
    uint32_t W = cfg->width ? cfg->width : O->d.camera.res[0];
    uint32_t H = cfg->height ? cfg->height : O->d.camera.res[1];
    aq_cam cam = aq_cam_derive(O->d.camera.translate, O->d.camera.rotate, O->d.camera.fov,
                               O->d.camera.lens_radius, O->d.camera.focal, W, H);
    for (uint32_t p = 0; p < W * H; ++p) {
        uint32_t key = aq_rng_key(cfg->seed, p, sample);
        aq_rayf r = aq_camera_ray(cam, p % W, p / W, key);
        out[p].o[0] = r.o.x; out[p].o[1] = r.o.y; out[p].o[2] = r.o.z;
        out[p].d[0] = r.d.x; out[p].d[1] = r.d.y; out[p].d[2] = r.d.z;
        out[p].tmin = r.tmin;
        out[p].tmax = r.tmax;
    }
    return AQ_OK;
}

/* film: float4[W*H] (sum rgb, count), accumulated in ascending sample order per pixel.
 * samples (optional): float4[(spp_end-spp_begin)*W*H], index (s-spp_begin)*W*H + pixel.
 * mode 0 = brute force, 1 = oracle BVH2. */
int aqo_render(aqo_scene* s, const aq_integrator_cfg* cfg, float* film, float* samples,
               aq_stats* stats, int mode, int n_threads) {
    Oracle* O = reinterpret_cast<Oracle*>(s);
    if (!O || !cfg || !film) return AQ_ERR_BAD_ARG;
    if (mode == 1 && !O->has_bvh) build_obvh(*O);
    uint32_t W = cfg->width ? cfg->width : O->d.camera.res[0];
    uint32_t H = cfg->height ? cfg->height : O->d.camera.res[1];
    aq_cam cam = aq_cam_derive(O->d.camera.translate, O->d.camera.rotate, O->d.camera.fov,
                               O->d.camera.lens_radius, O->d.camera.focal, W, H);
    const uint64_t npix = (uint64_t)W * H;
    if (!(cfg->flags & AQ_RENDER_ACCUMULATE)) std::memset(film, 0, npix * 16);
    std::atomic<uint64_t> c_samples{0}, c_sb{0}, c_rc{0}, c_rs{0};
    auto t0 = std::chrono::steady_clock::now();
    const bool use_bvh = mode == 1;
    const uint32_t mis_mode = (cfg->flags & AQ_RENDER_MIS_NEE_ONLY)    ? AQ_MIS_NEE_ONLY
                              : (cfg->flags & AQ_RENDER_MIS_BSDF_ONLY) ? AQ_MIS_BSDF_ONLY
                                                                       : AQ_MIS_BOTH;
    const bool has_area = O->view.n_lights > O->d.n_lights;
    const bool full = O->full_bsdf || (cfg->flags & AQ_RENDER_FORCE_FULL_BSDF);
    parallel_for(npix, n_threads, 64, [&](uint64_t b, uint64_t e, int) {
        uint64_t ls = 0, lsb = 0, lrc = 0, lrs = 0;
        for (uint64_t p = b; p < e; ++p) {
            float* fp = film + 4 * p;
            for (uint32_t sidx = cfg->spp_begin; sidx < cfg->spp_end; ++sidx) {
                uint32_t key = aq_rng_key(cfg->seed, (uint32_t)p, sidx);
                aq_rayf ray = aq_camera_ray(cam, (uint32_t)(p % W), (uint32_t)(p / W), key);
                aq_v3 beta = aq_mk(1.f, 1.f, 1.f), L = aq_mk(0.f, 0.f, 0.f);
                float prev_pdf = 0.f;
                ++ls;
                for (uint32_t depth = 0; depth < cfg->max_depth; ++depth) {
                    aq_hit h;
                    closest(*O, use_bvh, ray.o, ray.d, ray.tmin, ray.tmax, &h);
                    ++lrc;
                    if (h.prim == AQ_MISS_ID) break;
                    ++lsb;
                    aq_vertex_in vi;
                    if (full)
                        aq_fetch_vertex<true>(O->view, h.prim, h.u, h.v, ray.d, &vi);
                    else
                        aq_fetch_vertex<false>(O->view, h.prim, h.u, h.v, ray.d, &vi);
                    vi.t_hit = h.t;
                    vi.prev_pdf = prev_pdf;
                    aq_vertex_out vo;
#define AQO_SHADE(A, F)                                                                          \
    aq_shade_vertex<A, F>(vi, beta, key, depth, cfg->max_depth, O->view.n_lights, O->view.lights, \
                          mis_mode, &vo)
                    if (has_area) {
                        if (full) AQO_SHADE(true, true); else AQO_SHADE(true, false);
                    } else {
                        if (full) AQO_SHADE(false, true); else AQO_SHADE(false, false);
                    }
#undef AQO_SHADE
                    L = aq_add(L, vo.emitted);
                    if (vo.has_shadow) {
                        ++lrs;
                        if (!occluded(*O, use_bvh, vo.shadow.o, vo.shadow.d, vo.shadow.tmin,
                                      vo.shadow.tmax))
                            L = aq_add(L, vo.shadow_contrib);
                    }
                    if (!vo.has_next) break;
                    ray = vo.next;
                    beta = vo.beta;
                    prev_pdf = vo.next_pdf;
                }
                fp[0] += L.x;
                fp[1] += L.y;
                fp[2] += L.z;
                fp[3] += 1.0f;
                if (samples) {
                    float* sp = samples + 4 * ((uint64_t)(sidx - cfg->spp_begin) * npix + p);
                    sp[0] = L.x;
                    sp[1] = L.y;
                    sp[2] = L.z;
                    sp[3] = 1.0f;
                }
            }
        }
        c_samples += ls;
        c_sb += lsb;
        c_rc += lrc;
        c_rs += lrs;
    });
    auto t1 = std::chrono::steady_clock::now();
    if (stats) {
        std::memset(stats, 0, sizeof *stats);
        stats->samples = c_samples.load();
        stats->sample_bounces = c_sb.load();
        stats->rays_closest = c_rc.load();
        stats->rays_shadow = c_rs.load();
        stats->ms_total = (float)std::chrono::duration<double, std::milli>(t1 - t0).count();
    }
    return AQ_OK;
}

/* ---- scalar probes so tests can check the definitional functions one by one */
void aqo_sincos_2pi(float u, float* s, float* c) { aq_sincos_2pi(u, s, c); }
uint32_t aqo_rng_key(uint32_t seed, uint32_t pixel, uint32_t sample) { return aq_rng_key(seed, pixel, sample); }
float aqo_rng(uint32_t key, uint32_t dim) { return aq_rng(key, dim); }
int aqo_tri_test(const float* o, const float* d, float tmin, const float* v0, const float* v1,
                 const float* v2, float* tuv) {
    aq_v3 a = aq_mk(v0[0], v0[1], v0[2]);
    aq_v3 e1 = aq_sub(aq_mk(v1[0], v1[1], v1[2]), a), e2 = aq_sub(aq_mk(v2[0], v2[1], v2[2]), a);
    return aq_tri_test(aq_mk(o[0], o[1], o[2]), aq_mk(d[0], d[1], d[2]), tmin, a, e1, e2, &tuv[0],
                       &tuv[1], &tuv[2])
               ? 1
               : 0;
}
/* params: base.rgb metallic roughness specular specular_tint sheen sheen_tint transmission */
static aq_bsdf_params mk_params(const float* p) {
    aq_bsdf_params m{};
    m.base = aq_mk(p[0], p[1], p[2]);
    m.metallic = p[3];
    m.roughness = p[4];
    m.specular = p[5];
    m.specular_tint = p[6];
    m.sheen = p[7];
    m.sheen_tint = p[8];
    m.transmission = p[9];
    return m;
}
int aqo_bsdf_eval(const float* params, const float* wo, const float* wi, float* f_cos, float* pdf) {
    aq_v3 o = aq_mk(wo[0], wo[1], wo[2]), i = aq_mk(wi[0], wi[1], wi[2]);
    aq_bsdf_ctx c = aq_bsdf_setup(mk_params(params), o);
    aq_v3 f;
    if (!aq_bsdf_eval(c, o, i, &f, pdf)) return 0;
    f_cos[0] = f.x; f_cos[1] = f.y; f_cos[2] = f.z;
    return 1;
}
int aqo_bsdf_sample(const float* params, const float* wo, const float* u3, float* wi, float* weight,
                    float* pdf) {
    aq_v3 o = aq_mk(wo[0], wo[1], wo[2]);
    aq_bsdf_ctx c = aq_bsdf_setup(mk_params(params), o);
    aq_v3 w, wt;
    if (!aq_bsdf_sample(c, o, u3[0], u3[1], u3[2], &w, &wt, pdf)) return 0;
    wi[0] = w.x; wi[1] = w.y; wi[2] = w.z;
    weight[0] = wt.x; weight[1] = wt.y; weight[2] = wt.z;
    return 1;
}

/* FULL lobes.  params: the 10 above + clearcoat clearcoat_roughness ior subsurface
 * subsurface_color.rgb (17 floats); eta = relative index n_t/n_i seen from wo's side */
static aq_bsdf_params mk_params_full(const float* p) {
    aq_bsdf_params m = mk_params(p);
    m.clearcoat = p[10];
    m.clearcoat_roughness = p[11];
    m.ior = p[12];
    m.subsurface = p[13];
    m.subsurface_color = aq_mk(p[14], p[15], p[16]);
    m.anisotropic = 0.0f;
    m.anisotropic_rotation = 0.0f;
    return m;
}
/* anisotropic variants of the batched hooks: aniso[3] = anisotropic, anisotropic_rotation (turns),
 * angle of the surface tangent in the shading frame (radians) */
void aqo_bsdf_eval_aniso_n(const float* params, const float* aniso, float eta, const float* wo, const float* wi,
                           uint32_t n, float* f_cos, float* pdf, uint8_t* ok) {
    aq_v3 o = aq_mk(wo[0], wo[1], wo[2]);
    aq_bsdf_params m = mk_params_full(params);
    m.anisotropic = aniso[0];
    m.anisotropic_rotation = aniso[1];
    aq_bsdf_full b = aq_bsdf_setup_full(m, o, eta, cosf(aniso[2]), sinf(aniso[2]));
    for (uint32_t k = 0; k < n; ++k) {
        aq_v3 f = aq_mk(0.f, 0.f, 0.f);
        float p = 0.f;
        ok[k] = aq_bsdf_eval_full(b, o, aq_mk(wi[3 * k], wi[3 * k + 1], wi[3 * k + 2]), &f, &p) ? 1 : 0;
        f_cos[3 * k] = ok[k] ? f.x : 0.f;
        f_cos[3 * k + 1] = ok[k] ? f.y : 0.f;
        f_cos[3 * k + 2] = ok[k] ? f.z : 0.f;
        pdf[k] = ok[k] ? p : 0.f;
    }
}
void aqo_bsdf_sample_aniso_n(const float* params, const float* aniso, float eta, const float* wo, const float* u3,
                             uint32_t n, float* wi, float* weight, float* pdf, uint8_t* ok) {
    aq_v3 o = aq_mk(wo[0], wo[1], wo[2]);
    aq_bsdf_params m = mk_params_full(params);
    m.anisotropic = aniso[0];
    m.anisotropic_rotation = aniso[1];
    aq_bsdf_full b = aq_bsdf_setup_full(m, o, eta, cosf(aniso[2]), sinf(aniso[2]));
    for (uint32_t k = 0; k < n; ++k) {
        aq_v3 w = aq_mk(0.f, 0.f, 0.f), wt = w;
        float p = 0.f;
        ok[k] = aq_bsdf_sample_full(b, o, u3[3 * k], u3[3 * k + 1], u3[3 * k + 2], &w, &wt, &p) ? 1 : 0;
        wi[3 * k] = w.x; wi[3 * k + 1] = w.y; wi[3 * k + 2] = w.z;
        weight[3 * k] = ok[k] ? wt.x : 0.f;
        weight[3 * k + 1] = ok[k] ? wt.y : 0.f;
        weight[3 * k + 2] = ok[k] ? wt.z : 0.f;
        pdf[k] = ok[k] ? p : 0.f;
    }
}
int aqo_bsdf_eval_full(const float* params, float eta, const float* wo, const float* wi, float* f_cos,
                       float* pdf) {
    aq_v3 o = aq_mk(wo[0], wo[1], wo[2]), i = aq_mk(wi[0], wi[1], wi[2]);
    aq_bsdf_full b = aq_bsdf_setup_full(mk_params_full(params), o, eta);
    aq_v3 f;
    if (!aq_bsdf_eval_full(b, o, i, &f, pdf)) return 0;
    f_cos[0] = f.x; f_cos[1] = f.y; f_cos[2] = f.z;
    return 1;
}
int aqo_bsdf_sample_full(const float* params, float eta, const float* wo, const float* u3, float* wi,
                         float* weight, float* pdf) {
    aq_v3 o = aq_mk(wo[0], wo[1], wo[2]);
    aq_bsdf_full b = aq_bsdf_setup_full(mk_params_full(params), o, eta);
    aq_v3 w, wt;
    if (!aq_bsdf_sample_full(b, o, u3[0], u3[1], u3[2], &w, &wt, pdf)) return 0;
    wi[0] = w.x; wi[1] = w.y; wi[2] = w.z;
    weight[0] = wt.x; weight[1] = wt.y; weight[2] = wt.z;
    return 1;
}
float aqo_fresnel_dielectric(float cos_i, float eta) { return aq_fresnel_dielectric(cos_i, eta); }
/* batched forms for the statistical tests: one wo, n directions / n random triples */
void aqo_bsdf_eval_full_n(const float* params, float eta, const float* wo, const float* wi, uint32_t n,
                          float* f_cos, float* pdf, uint8_t* ok) {
    aq_v3 o = aq_mk(wo[0], wo[1], wo[2]);
    aq_bsdf_full b = aq_bsdf_setup_full(mk_params_full(params), o, eta);
    for (uint32_t k = 0; k < n; ++k) {
        aq_v3 f = aq_mk(0.f, 0.f, 0.f);
        float p = 0.f;
        ok[k] = aq_bsdf_eval_full(b, o, aq_mk(wi[3 * k], wi[3 * k + 1], wi[3 * k + 2]), &f, &p) ? 1 : 0;
        f_cos[3 * k] = ok[k] ? f.x : 0.f;
        f_cos[3 * k + 1] = ok[k] ? f.y : 0.f;
        f_cos[3 * k + 2] = ok[k] ? f.z : 0.f;
        pdf[k] = ok[k] ? p : 0.f;
    }
}
void aqo_bsdf_sample_full_n(const float* params, float eta, const float* wo, const float* u3, uint32_t n,
                            float* wi, float* weight, float* pdf, uint8_t* ok) {
    aq_v3 o = aq_mk(wo[0], wo[1], wo[2]);
    aq_bsdf_full b = aq_bsdf_setup_full(mk_params_full(params), o, eta);
    for (uint32_t k = 0; k < n; ++k) {
        aq_v3 w = aq_mk(0.f, 0.f, 0.f), wt = w;
        float p = 0.f;
        ok[k] = aq_bsdf_sample_full(b, o, u3[3 * k], u3[3 * k + 1], u3[3 * k + 2], &w, &wt, &p) ? 1 : 0;
        wi[3 * k] = w.x; wi[3 * k + 1] = w.y; wi[3 * k + 2] = w.z;
        weight[3 * k] = ok[k] ? wt.x : 0.f;
        weight[3 * k + 1] = ok[k] ? wt.y : 0.f;
        weight[3 * k + 2] = ok[k] ? wt.z : 0.f;
        pdf[k] = ok[k] ? p : 0.f;
    }
}

/* ---- nrc integrator ------------------------------------------------------------------- */
/* records: x_out[R*64], y_out[R*4] (target.rgb / fac, valid).  mode 0 = brute force, 1 = BVH2 */
int aqo_nrc_records(aqo_scene* s, const aq_integrator_cfg* cfg, const aq_nrc_cfg* nrc, float* x_out,
                    float* y_out, int mode, int n_threads) {
    Oracle* O = reinterpret_cast<Oracle*>(s);
    if (!O || !cfg || !nrc || !x_out || !y_out) return AQ_ERR_BAD_ARG;
    if (mode == 1 && !O->has_bvh) build_obvh(*O);
    const uint32_t W = cfg->width ? cfg->width : O->d.camera.res[0];
    const uint32_t H = cfg->height ? cfg->height : O->d.camera.res[1];
    const uint32_t npix = W * H;
    aq_cam cam = aq_cam_derive(O->d.camera.translate, O->d.camera.rotate, O->d.camera.fov,
                               O->d.camera.lens_radius, O->d.camera.focal, W, H);
    const aq_nrc_bounds bb = aq_nrc_bounds_of(O->pos.data(), O->d.n_verts);
    const uint64_t R = (uint64_t)nrc->batch_size * nrc->training_iters;
    const bool use_bvh = mode == 1;
    const uint32_t mis_mode = (cfg->flags & AQ_RENDER_MIS_NEE_ONLY)    ? AQ_MIS_NEE_ONLY
                              : (cfg->flags & AQ_RENDER_MIS_BSDF_ONLY) ? AQ_MIS_BSDF_ONLY
                                                                       : AQ_MIS_BOTH;
    const bool has_area = O->view.n_lights > O->d.n_lights;
    const bool full = O->full_bsdf || (cfg->flags & AQ_RENDER_FORCE_FULL_BSDF);
    parallel_for(R, n_threads, 64, [&](uint64_t b, uint64_t e, int) {
        for (uint64_t r = b; r < e; ++r) {
            float* x = x_out + AQ_NRC_IN * r;
            float* y = y_out + 4 * r;
            for (int k = 0; k < AQ_NRC_IN; ++k) x[k] = 0.f;
            y[0] = y[1] = y[2] = y[3] = 0.f;
            const uint32_t pixel = aq_nrc_record_pixel(cfg->seed, (uint32_t)r, npix);
            const uint32_t key = aq_nrc_record_key(cfg->seed, pixel, (uint32_t)r);
            const uint32_t D = aq_nrc_record_depth((uint32_t)r);
            aq_rayf ray = aq_camera_ray(cam, pixel % W, pixel / W, key);
            aq_v3 beta = aq_mk(1.f, 1.f, 1.f), L = aq_mk(0.f, 0.f, 0.f), fac = aq_mk(1.f, 1.f, 1.f);
            float prev_pdf = 0.f;
            bool valid = false;
            for (uint32_t depth = 0; depth < cfg->max_depth; ++depth) {
                aq_hit h;
                closest(*O, use_bvh, ray.o, ray.d, ray.tmin, ray.tmax, &h);
                if (h.prim == AQ_MISS_ID) break;
                aq_vertex_in vi;
                if (full)
                    aq_fetch_vertex<true>(O->view, h.prim, h.u, h.v, ray.d, &vi);
                else
                    aq_fetch_vertex<false>(O->view, h.prim, h.u, h.v, ray.d, &vi);
                vi.t_hit = h.t;
                vi.prev_pdf = prev_pdf;
                if (depth == D) { /* the record vertex: describe it, restart the estimate here */
                    aq_nrc_encode(vi, bb, x, 1, &fac);
                    valid = true;
                    beta = aq_mk(1.f, 1.f, 1.f);
                    L = aq_mk(0.f, 0.f, 0.f);
                }
                aq_vertex_out vo;
                if (full)
                    nrc_shade_dispatch<true>(has_area, vi, beta, key, depth, cfg->max_depth, O->view, mis_mode, &vo);
                else
                    nrc_shade_dispatch<false>(has_area, vi, beta, key, depth, cfg->max_depth, O->view, mis_mode, &vo);
                if (depth != D) L = aq_add(L, vo.emitted); /* the record vertex's own emission is not cached */
                if (depth >= D && vo.has_shadow &&
                    !occluded(*O, use_bvh, vo.shadow.o, vo.shadow.d, vo.shadow.tmin, vo.shadow.tmax))
                    L = aq_add(L, vo.shadow_contrib);
                if (!vo.has_next) break;
                ray = vo.next;
                beta = vo.beta;
                prev_pdf = vo.next_pdf;
            }
            if (valid) {
                y[0] = L.x / fac.x;
                y[1] = L.y / fac.y;
                y[2] = L.z / fac.z;
                y[3] = 1.0f;
            }
        }
    });
    return AQ_OK;
}

void aqo_nrc_init_weights(uint32_t seed, float* w) {
    for (uint32_t k = 0; k < AQ_NRC_N_WEIGHTS; ++k) w[k] = aq_nrc_init_weight(seed, k);
}

void aqo_nrc_forward(const float* weights, const float* x, uint32_t n, float* y_out) {
    NrcActs A;
    for (uint32_t q = 0; q < n; ++q) {
        nrc_forward(weights, x + AQ_NRC_IN * (size_t)q, &A);
        for (int c = 0; c < AQ_NRC_OUT; ++c) y_out[3 * (size_t)q + c] = A.y[c];
    }
}

/* training_iters descent steps over records (x, y) as laid out by aqo_nrc_records; weights in/out
 * (AQ_NRC_N_WEIGHTS), loss_out[training_iters] (may be NULL) */
int aqo_nrc_fit(const aq_nrc_cfg* nrc, const float* x, const float* y, float* weights, float* loss_out) {
    if (!nrc || !x || !y || !weights) return AQ_ERR_BAD_ARG;
    const uint32_t B = nrc->batch_size;
    const uint32_t n_chunks = (B + AQ_NRC_CHUNK - 1) / AQ_NRC_CHUNK;
    const float inv_norm = 1.0f / (3.0f * (float)B);
    std::vector<float> m(AQ_NRC_N_WEIGHTS, 0.f), v(AQ_NRC_N_WEIGHTS, 0.f), g(AQ_NRC_N_WEIGHTS);
    std::vector<float> gc((size_t)n_chunks * AQ_NRC_N_WEIGHTS);
    std::vector<NrcActs> acts(AQ_NRC_CHUNK);
    /* delta[l][s][j]: d loss / d (pre-activation of layer l+1's neuron j); l = 4 is the output */
    std::vector<float> delta((size_t)AQ_NRC_N_MATS * AQ_NRC_CHUNK * AQ_NRC_WIDTH);
    auto dl = [&](int l, int s2) { return &delta[((size_t)l * AQ_NRC_CHUNK + s2) * AQ_NRC_WIDTH]; };
    for (uint32_t it = 0; it < nrc->training_iters; ++it) {
        float loss = 0.f;
        for (uint32_t c = 0; c < n_chunks; ++c) {
            float part = 0.f;
            for (int s2 = 0; s2 < AQ_NRC_CHUNK; ++s2) {
                const uint32_t bi = c * AQ_NRC_CHUNK + s2;
                const bool live = bi < B && y[4 * ((size_t)it * B + bi) + 3] != 0.f;
                const size_t r = (size_t)it * B + (bi < B ? bi : 0);
                NrcActs& A = acts[s2];
                if (live) {
                    nrc_forward(weights, x + AQ_NRC_IN * r, &A);
                } else {
                    std::memset(&A, 0, sizeof A);
                }
                for (int ch = 0; ch < AQ_NRC_OUT; ++ch) {
                    dl(AQ_NRC_HIDDEN_LAYERS, s2)[ch] = live ? aq_nrc_loss_grad(A.y[ch], y[4 * r + ch], inv_norm) : 0.f;
                    if (live) part += aq_nrc_loss_term(A.y[ch], y[4 * r + ch], inv_norm);
                }
                /* backward-data through the output matrix, then the hidden ones */
                for (int l = AQ_NRC_HIDDEN_LAYERS; l >= 1; --l) {
                    const float* Wl = weights + AQ_NRC_MAT_OFF(l);
                    const int cols = AQ_NRC_MAT_COLS(l), n = l == AQ_NRC_HIDDEN_LAYERS ? AQ_NRC_OUT : AQ_NRC_WIDTH;
                    for (int i = 0; i < AQ_NRC_WIDTH; ++i)
                        dl(l - 1, s2)[i] = A.a[l][i] > 0.f ? aq_nrc_dot(Wl + (size_t)i * cols, 1, dl(l, s2), 1, n) : 0.f;
                }
            }
            loss += part;
            /* weight gradient of this chunk: sum over its samples in ascending order */
            float* G = &gc[(size_t)c * AQ_NRC_N_WEIGHTS];
            for (int l = 0; l < AQ_NRC_N_MATS; ++l) {
                const int cols = AQ_NRC_MAT_COLS(l), n = l == AQ_NRC_HIDDEN_LAYERS ? AQ_NRC_OUT : AQ_NRC_WIDTH;
                for (int i = 0; i < AQ_NRC_WIDTH; ++i)
                    for (int j = 0; j < cols; ++j)
                        G[AQ_NRC_MAT_OFF(l) + i * cols + j] =
                            j < n ? aq_nrc_dot(&acts[0].a[l][i], (int)(sizeof(NrcActs) / sizeof(float)),
                                               dl(l, 0) + j, AQ_NRC_WIDTH, AQ_NRC_CHUNK)
                                  : 0.f;
            }
        }
        float bc1, bc2;
        aq_nrc_adam_bias(it + 1, &bc1, &bc2);
        for (uint32_t k = 0; k < AQ_NRC_N_WEIGHTS; ++k) {
            float gs = 0.f;
            for (uint32_t c = 0; c < n_chunks; ++c) gs = gs + gc[(size_t)c * AQ_NRC_N_WEIGHTS + k];
            aq_nrc_adam(gs, nrc->learning_rate, bc1, bc2, &weights[k], &m[k], &v[k]);
        }
        if (loss_out) loss_out[it] = loss;
    }
    return AQ_OK;
}

/* render with the cache (weights: AQ_NRC_N_WEIGHTS floats) */
int aqo_nrc_render(aqo_scene* s, const aq_integrator_cfg* cfg, const aq_nrc_cfg* nrc, const float* weights,
                   float* film, float* samples, aq_stats* stats, int mode, int n_threads) {
    Oracle* O = reinterpret_cast<Oracle*>(s);
    if (!O || !cfg || !nrc || !weights || !film) return AQ_ERR_BAD_ARG;
    if (mode == 1 && !O->has_bvh) build_obvh(*O);
    const uint32_t W = cfg->width ? cfg->width : O->d.camera.res[0];
    const uint32_t H = cfg->height ? cfg->height : O->d.camera.res[1];
    aq_cam cam = aq_cam_derive(O->d.camera.translate, O->d.camera.rotate, O->d.camera.fov,
                               O->d.camera.lens_radius, O->d.camera.focal, W, H);
    const aq_nrc_bounds bb = aq_nrc_bounds_of(O->pos.data(), O->d.n_verts);
    const uint64_t npix = (uint64_t)W * H;
    if (!(cfg->flags & AQ_RENDER_ACCUMULATE)) std::memset(film, 0, npix * 16);
    std::atomic<uint64_t> c_samples{0}, c_sb{0}, c_rc{0}, c_rs{0};
    const bool use_bvh = mode == 1;
    const uint32_t mis_mode = (cfg->flags & AQ_RENDER_MIS_NEE_ONLY)    ? AQ_MIS_NEE_ONLY
                              : (cfg->flags & AQ_RENDER_MIS_BSDF_ONLY) ? AQ_MIS_BSDF_ONLY
                                                                       : AQ_MIS_BOTH;
    const bool has_area = O->view.n_lights > O->d.n_lights;
    const bool full = O->full_bsdf || (cfg->flags & AQ_RENDER_FORCE_FULL_BSDF);
    const uint32_t Dq = nrc->visualize_cache ? 0u : 1u;
    parallel_for(npix, n_threads, 64, [&](uint64_t b, uint64_t e, int) {
        uint64_t ls = 0, lsb = 0, lrc = 0, lrs = 0;
        NrcActs A;
        float x[AQ_NRC_IN];
        for (uint64_t p = b; p < e; ++p) {
            float* fp = film + 4 * p;
            for (uint32_t sidx = cfg->spp_begin; sidx < cfg->spp_end; ++sidx) {
                uint32_t key = aq_rng_key(cfg->seed, (uint32_t)p, sidx);
                aq_rayf ray = aq_camera_ray(cam, (uint32_t)(p % W), (uint32_t)(p / W), key);
                aq_v3 beta = aq_mk(1.f, 1.f, 1.f), L = aq_mk(0.f, 0.f, 0.f);
                float prev_pdf = 0.f;
                ++ls;
                for (uint32_t depth = 0; depth < cfg->max_depth; ++depth) {
                    aq_hit h;
                    closest(*O, use_bvh, ray.o, ray.d, ray.tmin, ray.tmax, &h);
                    ++lrc;
                    if (h.prim == AQ_MISS_ID) break;
                    ++lsb;
                    aq_vertex_in vi;
                    if (full)
                        aq_fetch_vertex<true>(O->view, h.prim, h.u, h.v, ray.d, &vi);
                    else
                        aq_fetch_vertex<false>(O->view, h.prim, h.u, h.v, ray.d, &vi);
                    vi.t_hit = h.t;
                    vi.prev_pdf = prev_pdf;
                    if (depth == Dq) { /* terminate into the cache */
                        aq_v3 em = has_area ? aq_vertex_emitted<true>(vi, beta, mis_mode)
                                            : aq_vertex_emitted<false>(vi, beta, mis_mode);
                        aq_v3 fac;
                        aq_nrc_encode(vi, bb, x, 1, &fac);
                        nrc_forward(weights, x, &A);
                        aq_v3 yr = aq_mk(aq_nrc_relu(A.y[0]), aq_nrc_relu(A.y[1]), aq_nrc_relu(A.y[2]));
                        L = aq_add(aq_add(L, em), aq_mul(beta, aq_mul(fac, yr)));
                        break;
                    }
                    aq_vertex_out vo;
                    if (full)
                        nrc_shade_dispatch<true>(has_area, vi, beta, key, depth, cfg->max_depth, O->view, mis_mode, &vo);
                    else
                        nrc_shade_dispatch<false>(has_area, vi, beta, key, depth, cfg->max_depth, O->view, mis_mode, &vo);
                    L = aq_add(L, vo.emitted);
                    if (vo.has_shadow) {
                        ++lrs;
                        if (!occluded(*O, use_bvh, vo.shadow.o, vo.shadow.d, vo.shadow.tmin, vo.shadow.tmax))
                            L = aq_add(L, vo.shadow_contrib);
                    }
                    if (!vo.has_next) break;
                    ray = vo.next;
                    beta = vo.beta;
                    prev_pdf = vo.next_pdf;
                }
                fp[0] += L.x;
                fp[1] += L.y;
                fp[2] += L.z;
                fp[3] += 1.0f;
                if (samples) {
                    float* sp = samples + 4 * ((uint64_t)(sidx - cfg->spp_begin) * npix + p);
                    sp[0] = L.x; sp[1] = L.y; sp[2] = L.z; sp[3] = 1.0f;
                }
            }
        }
        c_samples += ls; c_sb += lsb; c_rc += lrc; c_rs += lrs;
    });
    if (stats) {
        std::memset(stats, 0, sizeof *stats);
        stats->samples = c_samples.load();
        stats->sample_bounces = c_sb.load();
        stats->rays_closest = c_rc.load();
        stats->rays_shadow = c_rs.load();
    }
    return AQ_OK;
}

}  // extern "C"
