/*
 * aq_cuda.cu — implementation of the C ABI in include/aqua_cuda.h (libaqua_cuda.so).
 *
 * The drop-in boundary of SURVEY §8b: the Rust host crate `arukas` (Cargo.toml:1-5; its
 * src/lib.rs is empty in the reference snapshot) hands over flat scene arrays and gets a
 * float4 film back.  Everything below the boundary is CUDA for sm_100a; there is no CPU
 * fallback — every compute entry fails with AQ_ERR_CUDA when no device is usable.
 */
#include <cuda_runtime.h>

#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <mutex>
#include <thread>
#include <new>
#include <string>
#include <vector>

#include "aqua_cuda.h"
/* aqua_cuda.h first: aq_core.h then also defines the host-only material packing */
#include "aq_bvh_build.h"
#include "aq_internal.h"
#include "aq_kernels.cuh"
#include "aq_nrc.cuh"


/* default wavefront pool: 2^24 path slots = 2.95 GB of queues.  Per-launch overhead (launch
 * latency, ramp-up, tail) is ~13 us; at 2^21 slots it cost 18 % of the cbox render, at 2^23 4 %;
 * 2^24 is another 3.4 % (cbox) / 3.9 % (room) faster, 2^25 is slower again on cbox (measured on
 * B200, profiles/r01_pool_sweep.log). */
#define AQ_DEFAULT_POOL (1u << 24)
#define AQ_PROF_STRIDE 8u
#define AQ_CTRL_ALLOC_BYTES (256u * 1024u)

static_assert(AQ_PT_METALLIC == AQ_PTEX_METALLIC && AQ_PT_ROUGHNESS == AQ_PTEX_ROUGHNESS && AQ_PT_SPECULAR == AQ_PTEX_SPECULAR &&
                  AQ_PT_SPECULAR_TINT == AQ_PTEX_SPECULAR_TINT && AQ_PT_SHEEN == AQ_PTEX_SHEEN && AQ_PT_SHEEN_TINT == AQ_PTEX_SHEEN_TINT &&
                  AQ_PT_TRANSMISSION == AQ_PTEX_TRANSMISSION && AQ_PT_CLEARCOAT == AQ_PTEX_CLEARCOAT &&
                  AQ_PT_CLEARCOAT_ROUGHNESS == AQ_PTEX_CLEARCOAT_ROUGHNESS && AQ_PT_IOR == AQ_PTEX_IOR &&
                  AQ_PT_SUBSURFACE == AQ_PTEX_SUBSURFACE && AQ_PT_SUBSURFACE_COLOR == AQ_PTEX_SUBSURFACE_COLOR,
              "aq_core.h's AQ_PT_* slots must mirror include/aqua_cuda.h's AQ_PTEX_* enum");

namespace {

thread_local std::string g_thread_err;

}  // namespace

struct aq_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    int sm_count = 0, cc_major = 0, cc_minor = 0;
    size_t hbm = 0;
    std::string err;
    /* wavefront pool: owned by the ctx, shared by all of its scenes (calls on a ctx are
     * serialised), so a new scene does not re-allocate 1.4 GB of queues */
    uint32_t pool = 0;
    float4* d_pool = nullptr; /* one allocation carved into the queues below */
    aq_queue q[2], shq;
    uint4* d_hits = nullptr;
    float4* d_L = nullptr;
    /* film / per-sample dump / aq_intersect staging: also ctx-owned and grown on demand —
     * cudaMalloc + cudaFree of a 16 MiB film per scene cost 25-240 ms per create/destroy cycle */
    float4* d_film = nullptr;
    size_t film_pixels = 0;
    float4* d_samples = nullptr;
    size_t samples_count = 0;
    void* d_scratch_rays = nullptr;
    void* d_scratch_hits = nullptr;
    size_t scratch_n = 0;
    /* block-structured queues (aq_kernels.cuh): per-block entry counts of the two ray queues and
     * the shadow queue; the queue arrays hold pool + q_slack entries because every producing warp
     * may leave two partly filled blocks per queue */
    aq_qcounts qc{};
    uint32_t* d_qcnt = nullptr;
    uint32_t q_blocks = 0; /* capacity of each queue in blocks */
    /* scenes created on this ctx and not yet destroyed: aq_destroy destroys them first */
    std::vector<aq_scene*> scenes;
};

struct aq_scene {
    aq_ctx* ctx = nullptr;
    /* host copies needed by the builder */
    std::vector<float> h_pos;
    std::vector<uint32_t> h_idx;
    uint32_t n_verts = 0, n_tris = 0;
    aq_camera camera{};
    /* device scene */
    float *d_pos = nullptr, *d_nrm = nullptr, *d_uv = nullptr, *d_lut = nullptr;
    aq_f4* d_lights = nullptr; /* AQ_LIGHT_WORDS x 16 B per light */
    float* d_prim_light_pdf = nullptr;
    uint32_t n_area_lights = 0;
    bool full_bsdf = false; /* some material needs the FULL vertex code (aq_material_needs_full) */
    /* nrc integrator (aq_nrc.cuh): weights + Adam moments, records of the last training, loss curve */
    float *d_nrc_w = nullptr, *d_nrc_m = nullptr, *d_nrc_v = nullptr;
    float* d_nrc_x = nullptr;
    float4* d_nrc_y = nullptr;
    float *d_nrc_g = nullptr, *d_nrc_loss = nullptr, *d_nrc_loss_chunk = nullptr;
    uint8_t* d_nrc_wt = nullptr; /* bf16 operand tiles of the weights (AQ_RENDER_NRC_TENSOR) */
    uint64_t nrc_records = 0;
    uint32_t nrc_iters = 0;
    bool nrc_trained = false;
    aq_nrc_bounds nrc_bb{};
    uint32_t *d_idx = nullptr, *d_tri_mat = nullptr, *d_texels = nullptr;
    aq_f4* d_mats = nullptr;
    aq_f4* d_shade_recs = nullptr; /* 128 B per triangle */
    aq_u4* d_tex_desc = nullptr;
    uint32_t n_lights = 0;
    /* accel */
    aq_u4* d_nodes = nullptr;
    aq_f4* d_tris = nullptr;
    size_t n_node_words = 0, n_tri_words = 0;
    bool built = false;
    aq_accel_info accel{};
    /* hybrid build (AQ_HYBRID_BUILD_MIN_TRIS <= n_tris < AQ_DEVICE_BUILD_MIN_TRIS): the device LBVH tree
     * makes the scene usable after a few ms; a host thread builds the SAH tree meanwhile, uploads it
     * on its own stream and the render loop swaps it in between two waves.  Both trees obey the
     * conservativeness contract, so results do not depend on which one a wave used. */
    struct upgrade_t {
        std::thread th;
        std::atomic<int> state{0}; /* 1 running, 2 ready (d_nodes/d_tris below), 3 failed */
        aq_u4* d_nodes = nullptr;
        aq_f4* d_tris = nullptr;
        size_t n_node_words = 0, n_tri_words = 0;
        aq_accel_info info{};
    };
    upgrade_t* upg = nullptr;
    std::vector<void*> garbage; /* replaced trees: freed once the stream has drained */
    uint32_t* d_ctrl = nullptr; /* AQC_WORDS counters, each on its own 128-byte line */
    cudaEvent_t ev_pace[2] = {nullptr, nullptr}; /* hybrid build: waves in flight while the SAH tree is pending */
    void* d_ctrl_alloc = nullptr;
    unsigned long long* d_stats = nullptr;
    /* film / samples */
    /* last render */
    aq_integrator_cfg last_cfg{};
    uint32_t last_launches = 0, last_waves = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool render_pending = false;
    /* AQ_RENDER_PROFILE: one event after every launch, tagged with its stage */
    std::vector<cudaEvent_t> prof_ev;
    std::vector<uint8_t> prof_stage;
    size_t prof_n = 0;
    uint32_t prof_waves = 0;
    /* scratch for aq_intersect */
};

namespace {

int set_err(aq_ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_thread_err = buf;
    if (c) c->err = buf;
    return code;
}

#define AQ_CK(ctx, call)                                                                     \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess)                                                               \
            return set_err(ctx, e_ == cudaErrorMemoryAllocation ? AQ_ERR_OOM : AQ_ERR_CUDA,  \
                           "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                           __LINE__);                                                        \
    } while (0)

template <class T>
int upload(aq_ctx* c, T** dst, const void* src, size_t count) {
    *dst = nullptr;
    if (count == 0) return AQ_OK;
    AQ_CK(c, cudaMallocAsync((void**)dst, count * sizeof(T), c->stream));
    AQ_CK(c, cudaMemcpyAsync(*dst, src, count * sizeof(T), cudaMemcpyHostToDevice, c->stream));
    return AQ_OK;
}

aq_scene_view make_view(const aq_scene* s) {
    aq_scene_view v;
    v.pos = s->d_pos;
    v.nrm = s->d_nrm;
    v.uv = s->d_uv;
    v.idx = s->d_idx;
    v.tri_mat = s->d_tri_mat;
    v.mats = s->d_mats;
    v.tex_desc = s->d_tex_desc;
    v.texels = s->d_texels;
    v.srgb_lut = s->d_lut;
    v.lights = s->d_lights;
    v.n_lights = s->n_lights;
    v.prim_light_pdf = s->d_prim_light_pdf;
    v.shade_recs = s->d_shade_recs;
    return v;
}

/* the most warps a producing (shade / nrc) kernel runs with: 8 CTAs of 128 threads per SM */
uint32_t max_producer_warps(const aq_ctx* c) { return (uint32_t)c->sm_count * 8u * (AQ_SHADE_THREADS / 32u); }

int ensure_pool(aq_scene* scene, uint32_t pool) {
    aq_ctx* c = scene->ctx;
    aq_ctx* s = c; /* the pool lives in the ctx */
    if (s->pool >= pool && s->d_pool) return AQ_OK;
    cudaStreamSynchronize(c->stream);
    if (s->d_pool) cudaFree(s->d_pool);
    if (s->d_qcnt) cudaFree(s->d_qcnt);
    s->d_pool = nullptr;
    s->d_qcnt = nullptr;
    s->pool = 0;
    /* queue capacity: the pool in blocks + two static blocks per producing warp */
    const uint32_t blocks = (pool + AQ_QBLK - 1u) / AQ_QBLK + 2u * max_producer_warps(c);
    const size_t n = (size_t)blocks * AQ_QBLK;
    /* 2 ray queues x 3 + shadow queue x 3 + hits (n entries each) + L (pool entries) */
    AQ_CK(c, cudaMalloc((void**)&s->d_pool, (n * 10 + (size_t)pool) * sizeof(float4)));
    AQ_CK(c, cudaMalloc((void**)&s->d_qcnt, (size_t)blocks * 3 * sizeof(uint32_t)));
    float4* p = s->d_pool;
    for (int k = 0; k < 2; ++k) {
        s->q[k].o_tmin = p;  p += n;
        s->q[k].d_tmax = p;  p += n;
        s->q[k].beta_id = p; p += n;
    }
    s->shq.o_tmin = p;  p += n;
    s->shq.d_tmax = p;  p += n;
    s->shq.beta_id = p; p += n;
    s->d_hits = reinterpret_cast<uint4*>(p); p += n;
    s->d_L = p;
    s->qc.ray[0] = s->d_qcnt;
    s->qc.ray[1] = s->d_qcnt + blocks;
    s->qc.shadow = s->d_qcnt + 2 * (size_t)blocks;
    s->q_blocks = blocks;
    s->pool = pool;
    return AQ_OK;
}

int ensure_scratch(aq_scene* s, size_t n) {
    aq_ctx* c = s->ctx;
    if (s->ctx->scratch_n >= n) return AQ_OK;
    if (s->ctx->d_scratch_rays) cudaFree(s->ctx->d_scratch_rays);
    if (s->ctx->d_scratch_hits) cudaFree(s->ctx->d_scratch_hits);
    s->ctx->d_scratch_rays = s->ctx->d_scratch_hits = nullptr;
    s->ctx->scratch_n = 0;
    AQ_CK(c, cudaMalloc(&s->ctx->d_scratch_rays, n * sizeof(aq_ray)));
    AQ_CK(c, cudaMalloc(&s->ctx->d_scratch_hits, n * sizeof(aq_hit)));
    s->ctx->scratch_n = n;
    return AQ_OK;
}

/* persistent kernels: exactly the number of CTAs that are co-resident (occupancy query per
 * kernel instantiation, cached) so no CTA waits for an SM slot only to find the queue empty */
template <class K>
int resident_grid(const aq_ctx* c, K kernel, int threads) {
    static std::vector<std::pair<const void*, int>> cache;
    static std::mutex mu; /* several ctxs may be driven from different host threads */
    std::lock_guard<std::mutex> lock(mu);
    for (auto& e : cache)
        if (e.first == (const void*)kernel) return e.second * c->sm_count;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0) != cudaSuccess || per_sm < 1)
        per_sm = 4;
    cache.emplace_back((const void*)kernel, per_sm);
    return per_sm * c->sm_count;
}



/* ---- hybrid accel build */
void accel_upgrade_thread(aq_scene* s, int device) {
    aq_scene::upgrade_t* u = s->upg;
    auto t0 = std::chrono::steady_clock::now();
    aq_bvh8 bvh;
    int nt = (int)std::thread::hardware_concurrency() - 2; /* leave cores to the thread that launches the renders */
    bool ok = aq_build_bvh8(s->h_pos.data(), s->h_idx.data(), s->n_tris, nt < 1 ? 1 : nt, &bvh) == 0;
    cudaStream_t st = nullptr;
    ok = ok && cudaSetDevice(device) == cudaSuccess && cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess;
    if (ok) {
        if (bvh.tris.empty()) bvh.tris.resize(AQ_TRI_WORDS);
        ok = cudaMallocAsync((void**)&u->d_nodes, bvh.nodes.size() * sizeof(aq_u4), st) == cudaSuccess &&
             cudaMallocAsync((void**)&u->d_tris, bvh.tris.size() * sizeof(aq_f4), st) == cudaSuccess &&
             cudaMemcpyAsync(u->d_nodes, bvh.nodes.data(), bvh.nodes.size() * sizeof(aq_u4), cudaMemcpyHostToDevice, st) == cudaSuccess &&
             cudaMemcpyAsync(u->d_tris, bvh.tris.data(), bvh.tris.size() * sizeof(aq_f4), cudaMemcpyHostToDevice, st) == cudaSuccess &&
             cudaStreamSynchronize(st) == cudaSuccess;
    }
    if (std::getenv("AQ_BUILD_VERBOSE"))
        std::fprintf(stderr, "[aq hybrid] background SAH build + upload: %.1f ms (%s)\n",
                     std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(), ok ? "ok" : "failed");
    if (!ok && st) {
        if (u->d_nodes) cudaFreeAsync(u->d_nodes, st);
        if (u->d_tris) cudaFreeAsync(u->d_tris, st);
        u->d_nodes = nullptr;
        u->d_tris = nullptr;
        cudaStreamSynchronize(st);
    }
    if (st) cudaStreamDestroy(st);
    if (ok) {
        u->n_node_words = bvh.nodes.size();
        u->n_tri_words = bvh.tris.size();
        u->info.n_nodes = (uint32_t)(bvh.nodes.size() / AQ_NODE_WORDS);
        u->info.n_tri_records = s->n_tris;
        u->info.max_depth = bvh.max_depth;
        u->info.sah_cost = bvh.sah_cost;
        u->info.builder = 0;
        u->info.build_ms = (float)std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    u->state.store(ok ? 2 : 3, std::memory_order_release);
}

/* swap the host-built tree in if it is ready (`wait`: block until the background build is over).
 * Kernels already enqueued keep the old pointers, so the old tree goes to the garbage list. */
void accel_try_upgrade(aq_scene* s, bool wait) {
    aq_scene::upgrade_t* u = s->upg;
    if (!u) return;
    if (!wait && u->state.load(std::memory_order_acquire) < 2) return;
    if (u->th.joinable()) u->th.join();
    if (u->state.load(std::memory_order_acquire) == 2) {
        s->garbage.push_back(s->d_nodes);
        s->garbage.push_back(s->d_tris);
        s->d_nodes = u->d_nodes;
        s->d_tris = u->d_tris;
        s->n_node_words = u->n_node_words;
        s->n_tri_words = u->n_tri_words;
        const float first_ms = s->accel.build_ms;
        s->accel = u->info;
        s->accel.build_ms = first_ms; /* time until the scene was usable; the SAH build ran beside the renders */
    }
    delete u;
    s->upg = nullptr;
}

void accel_free_garbage(aq_scene* s) { /* stream-ordered: behind every kernel enqueued so far */
    for (void* p : s->garbage)
        if (p) cudaFreeAsync(p, s->ctx->stream);
    s->garbage.clear();
}

/* the launches of one wavefront stage, shared by the path tracer, the nrc record passes and the
 * cache render */
struct wave_launcher {
    aq_scene* s;
    aq_ctx* c;
    cudaStream_t st;
    aq_scene_view sv;
    bool area, full;
    int tgrid, tgrid_sh, ggrid, sgrid;
    bool dyn; /* balanced triangle phase from depth 1 on (AQ_TRI_DYN_MIN_NODES; AQUA_TRI_DYN=0|1 overrides) */
    uint32_t launches = 0;
    void (*shade_fn)(aq_scene_view, aq_wave_params, int, aq_queue, const uint4*, aq_queue, aq_queue, float4*,
                     uint32_t*, aq_qcounts, unsigned long long*);

    wave_launcher(aq_scene* scene, uint32_t flags) : s(scene), c(scene->ctx), st(scene->ctx->stream) {
        sv = make_view(s);
        area = s->n_area_lights > 0;
        full = s->full_bsdf || (flags & AQ_RENDER_FORCE_FULL_BSDF);
        /* the shade instantiation this scene needs: <emissive triangles, full Principled lobes> */
        shade_fn = area ? (full ? aq_k_shade<true, true> : aq_k_shade<true, false>)
                        : (full ? aq_k_shade<false, true> : aq_k_shade<false, false>);
        tgrid = std::min(resident_grid(c, aq_k_trace<3, false, 0>, AQ_TRACE_THREADS),
                         resident_grid(c, aq_k_trace<3, false, AQ_TRI_DYN>, AQ_TRACE_THREADS));
        tgrid_sh = std::min(resident_grid(c, aq_k_trace<1, false, 0>, AQ_TRACE_THREADS),
                            resident_grid(c, aq_k_trace<1, false, AQ_TRI_DYN>, AQ_TRACE_THREADS));
        ggrid = c->sm_count * 8;
        dyn = s->n_node_words / AQ_NODE_WORDS >= AQ_TRI_DYN_MIN_NODES;
        if (const char* e = std::getenv("AQUA_TRI_DYN")) dyn = e[0] != '0';
        sgrid = resident_grid(c, shade_fn, AQ_SHADE_THREADS);
        if (sgrid > c->sm_count * 8) sgrid = c->sm_count * 8; /* the queues' slack is sized for this (max_producer_warps) */
    }
    /* static output blocks of the shade pass: two per warp and queue */
    uint32_t shade_static_blocks() const { return 2u * (uint32_t)sgrid * (AQ_SHADE_THREADS / 32u); }
    void raygen(const aq_wave_params& wp) {
        aq_k_raygen<<<ggrid, AQ_GEN_THREADS, 0, st>>>(wp, c->q[0], c->d_L, s->d_ctrl, c->qc.ray[0], s->d_stats);
        ++launches;
    }
    void closest(uint32_t depth) {
        const aq_queue& cur = c->q[depth & 1];
        auto k = (dyn && depth >= 1) ? aq_k_trace<3, false, AQ_TRI_DYN> : aq_k_trace<3, false, 0>;
        k<<<tgrid, AQ_TRACE_THREADS, 0, st>>>(
            s->d_nodes, s->d_tris, cur.o_tmin, cur.d_tmax, 1, nullptr, c->qc.ray[depth & 1],
            &s->d_ctrl[aqc_blocks_ray((int)depth)], 0, &s->d_ctrl[AQC_FETCH_CLOSEST], c->d_hits, nullptr, s->d_ctrl,
            (int)depth, shade_static_blocks(), s->d_stats);
        ++launches;
    }
    void shade(const aq_wave_params& wp, uint32_t depth) {
        shade_fn<<<sgrid, AQ_SHADE_THREADS, 0, st>>>(sv, wp, (int)depth, c->q[depth & 1], c->d_hits,
                                                     c->q[(depth & 1) ^ 1], c->shq, c->d_L, s->d_ctrl, c->qc, s->d_stats);
        ++launches;
    }
    void shadow(uint32_t depth) {
        auto k = (dyn && depth >= 1) ? aq_k_trace<1, false, AQ_TRI_DYN> : aq_k_trace<1, false, 0>;
        k<<<tgrid_sh, AQ_TRACE_THREADS, 0, st>>>(
            s->d_nodes, s->d_tris, c->shq.o_tmin, c->shq.d_tmax, 1, c->shq.beta_id, c->qc.shadow,
            &s->d_ctrl[AQC_BLOCKS_SHADOW], 0, &s->d_ctrl[AQC_FETCH_SHADOW], nullptr, c->d_L, s->d_ctrl, (int)depth, 0u,
            s->d_stats);
        ++launches;
    }
    void film(const aq_wave_params& wp, float4* film_buf, float4* samples) {
        aq_k_film<<<ggrid, AQ_GEN_THREADS, 0, st>>>(wp, c->d_L, film_buf, samples);
        ++launches;
    }
};
}  // namespace

extern "C" {

int aq_abi_version(void) { return AQ_ABI_VERSION; }

const char* aq_last_error(aq_ctx* ctx) { return ctx ? ctx->err.c_str() : g_thread_err.c_str(); }

int aq_init(int device, aq_ctx** out) {
    if (!out) return set_err(nullptr, AQ_ERR_BAD_ARG, "aq_init: out is null");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return set_err(nullptr, AQ_ERR_CUDA, "aq_init: no CUDA device (%s); this library has no CPU path",
                       e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) return set_err(nullptr, AQ_ERR_BAD_ARG, "aq_init: device %d out of range (%d devices)", device, n);
    cudaDeviceProp prop;
    AQ_CK(nullptr, cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return set_err(nullptr, AQ_ERR_UNSUPPORTED,
                       "aq_init: device %d is sm_%d%d; this library is built for sm_100a only", device,
                       prop.major, prop.minor);
    AQ_CK(nullptr, cudaSetDevice(device));
    aq_ctx* c = new (std::nothrow) aq_ctx;
    if (!c) return set_err(nullptr, AQ_ERR_OOM, "aq_init: out of memory");
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->cc_major = prop.major;
    c->cc_minor = prop.minor;
    c->hbm = prop.totalGlobalMem;
    { /* scene-lifetime memory is stream-ordered (cudaMallocAsync); never hand it back to the OS between scenes */
        cudaMemPool_t mp = nullptr;
        unsigned long long keep = ~0ull;
        if (cudaDeviceGetDefaultMemPool(&mp, device) == cudaSuccess && mp)
            cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaError_t se = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (se != cudaSuccess) {
        delete c;
        return set_err(nullptr, AQ_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(se));
    }
    *out = c;
    return AQ_OK;
}

void aq_destroy(aq_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    /* scenes hold a pointer to their ctx (stream, pool): a ctx takes the scenes that are still
     * alive with it; their handles are invalid afterwards (include/aqua_cuda.h) */
    while (!ctx->scenes.empty()) aq_scene_destroy(ctx->scenes.back());
    void* owned[] = {ctx->d_pool, ctx->d_film, ctx->d_samples, ctx->d_scratch_rays, ctx->d_scratch_hits, ctx->d_qcnt};
    for (void* p : owned)
        if (p) cudaFree(p);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int aq_set_stream(aq_ctx* ctx, void* cuda_stream) {
    if (!ctx) return set_err(nullptr, AQ_ERR_BAD_ARG, "aq_set_stream: ctx is null");
    AQ_CK(ctx, cudaSetDevice(ctx->device));
    AQ_CK(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    if (cuda_stream) {
        ctx->stream = (cudaStream_t)cuda_stream;
        ctx->own_stream = false;
    } else {
        AQ_CK(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->own_stream = true;
    }
    return AQ_OK;
}

int aq_device_info(aq_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, size_t* hbm_bytes) {
    if (!ctx) return set_err(nullptr, AQ_ERR_BAD_ARG, "aq_device_info: ctx is null");
    if (sm_count) *sm_count = ctx->sm_count;
    if (cc_major) *cc_major = ctx->cc_major;
    if (cc_minor) *cc_minor = ctx->cc_minor;
    if (hbm_bytes) *hbm_bytes = ctx->hbm;
    return AQ_OK;
}

/* ------------------------------------------------------------------ scene */
int aq_scene_create(aq_ctx* c, const aq_scene_desc* d, aq_scene** out) {
    if (!c || !d || !out) return set_err(c, AQ_ERR_BAD_ARG, "aq_scene_create: null argument");
    *out = nullptr;
    if (d->n_tris && (!d->positions || !d->indices))
        return set_err(c, AQ_ERR_BAD_ARG, "aq_scene_create: positions/indices missing");
    for (size_t i = 0; i < 3 * (size_t)d->n_tris; ++i)
        if (d->indices[i] >= d->n_verts)
            return set_err(c, AQ_ERR_BAD_ARG, "aq_scene_create: index %zu out of range", i);
    if (d->tri_material)
        for (uint32_t i = 0; i < d->n_tris; ++i)
            if (d->tri_material[i] >= d->n_materials)
                return set_err(c, AQ_ERR_BAD_ARG, "aq_scene_create: tri_material[%u] out of range", i);
    for (uint32_t m = 0; m < d->n_materials; ++m) {
        if (d->materials[m].color_tex >= (int32_t)d->n_textures)
            return set_err(c, AQ_ERR_BAD_ARG, "aq_scene_create: material %u texture out of range", m);
        for (int k = 0; k < 16; ++k)
            if ((k >= AQ_PTEX_COUNT && d->materials[m].param_tex[k] != 0) || d->materials[m].param_tex[k] > d->n_textures)
                return set_err(c, AQ_ERR_BAD_ARG, "aq_scene_create: material %u param_tex[%d] out of range", m, k);
    }
    AQ_CK(c, cudaSetDevice(c->device));
    aq_scene* s = new (std::nothrow) aq_scene;
    if (!s) return set_err(c, AQ_ERR_OOM, "aq_scene_create: out of memory");
    s->ctx = c;
    c->scenes.push_back(s);
    s->n_verts = d->n_verts;
    s->n_tris = d->n_tris;
    s->camera = d->camera;
    s->h_pos.assign(d->positions, d->positions + 3 * (size_t)d->n_verts);
    s->h_idx.assign(d->indices, d->indices + 3 * (size_t)d->n_tris);
    s->nrc_bb = aq_nrc_bounds_of(d->positions, d->n_verts);
    int rc;
#define AQ_TRY(x)                \
    if ((rc = (x)) != AQ_OK) {   \
        aq_scene_destroy(s);     \
        return rc;               \
    }
    AQ_TRY(upload(c, &s->d_pos, d->positions, 3 * (size_t)d->n_verts));
    if (d->normals) AQ_TRY(upload(c, &s->d_nrm, d->normals, 3 * (size_t)d->n_verts));
    if (d->uvs) AQ_TRY(upload(c, &s->d_uv, d->uvs, 2 * (size_t)d->n_verts));
    AQ_TRY(upload(c, &s->d_idx, d->indices, 3 * (size_t)d->n_tris));
    std::vector<uint32_t> tm;
    if (d->tri_material)
        tm.assign(d->tri_material, d->tri_material + d->n_tris);
    else
        tm.assign(d->n_tris, 0u);
    AQ_TRY(upload(c, &s->d_tri_mat, tm.data(), tm.size()));
    { /* light table, then the per-triangle shading records (one 128 B line each), built by
       * a kernel from the arrays just uploaded */
        aq_scene_desc dd = *d;
        dd.tri_material = tm.data();
        std::vector<aq_f4> ltab;
        std::vector<float> lpdf;
        aq_build_light_table(dd, &ltab, &lpdf);
        s->n_lights = (uint32_t)(ltab.size() / AQ_LIGHT_WORDS);
        s->n_area_lights = s->n_lights - d->n_lights;
        AQ_TRY(upload(c, &s->d_lights, ltab.data(), ltab.size()));
        if (s->n_area_lights) AQ_TRY(upload(c, &s->d_prim_light_pdf, lpdf.data(), lpdf.size()));
        if (d->n_tris) {
            cudaError_t me = cudaMallocAsync((void**)&s->d_shade_recs, (size_t)d->n_tris * AQ_SHADE_REC_WORDS * sizeof(aq_f4), c->stream);
            if (me != cudaSuccess) {
                aq_scene_destroy(s);
                return set_err(c, AQ_ERR_OOM, "aq_scene_create: shading records: %s", cudaGetErrorString(me));
            }
            aq_scene_view dv = make_view(s);
            dv.shade_recs = nullptr; /* gather from the mesh arrays */
            aq_k_build_shade_recs<<<(d->n_tris + 255) / 256, 256, 0, c->stream>>>(dv, d->n_tris, s->d_shade_recs);
        }
        cudaError_t se = cudaStreamSynchronize(c->stream); /* ltab / lpdf die at the end of this block */
        if (se != cudaSuccess) {
            aq_scene_destroy(s);
            return set_err(c, AQ_ERR_CUDA, "aq_scene_create: %s", cudaGetErrorString(se));
        }
    }
    std::vector<aq_f4> mats(AQ_MAT_WORDS * (size_t)(d->n_materials ? d->n_materials : 1));
    std::memset(mats.data(), 0, mats.size() * sizeof(aq_f4));
    if (d->n_materials == 0) { /* default grey diffuse */
        aq_material dm;
        std::memset(&dm, 0, sizeof dm);
        dm.color[0] = dm.color[1] = dm.color[2] = 0.5f;
        dm.color_tex = -1;
        dm.roughness = 0.5f;
        aq_pack_material(dm, mats.data());
    }
    for (uint32_t m = 0; m < d->n_materials; ++m) {
        aq_pack_material(d->materials[m], &mats[AQ_MAT_WORDS * (size_t)m]);
        s->full_bsdf = s->full_bsdf || aq_material_needs_full(d->materials[m]);
    }
    AQ_TRY(upload(c, &s->d_mats, mats.data(), mats.size()));
    /* texture atlas: every texture goes straight from the caller's memory into its place of one device
     * allocation (no host-side repack: 62 MiB for room.json) */
    std::vector<aq_u4> tdesc;
    size_t n_texels = 0;
    for (uint32_t t = 0; t < d->n_textures; ++t) {
        aq_u4 td;
        td.x = d->textures[t].width;
        td.y = d->textures[t].height;
        td.z = (uint32_t)n_texels;
        td.w = 0;
        tdesc.push_back(td);
        n_texels += (size_t)td.x * td.y;
    }
    if (n_texels > 0xFFFFFFFFull) {
        aq_scene_destroy(s);
        return set_err(c, AQ_ERR_UNSUPPORTED, "aq_scene_create: more than 2^32 texels");
    }
    AQ_TRY(upload(c, &s->d_tex_desc, tdesc.data(), tdesc.size()));
    if (n_texels) {
        cudaError_t te = cudaMallocAsync((void**)&s->d_texels, n_texels * sizeof(uint32_t), c->stream);
        for (uint32_t t = 0; t < d->n_textures && te == cudaSuccess; ++t)
            te = cudaMemcpyAsync(s->d_texels + tdesc[t].z, d->textures[t].rgba8, (size_t)tdesc[t].x * tdesc[t].y * 4,
                                 cudaMemcpyHostToDevice, c->stream);
        if (te != cudaSuccess) {
            aq_scene_destroy(s);
            return set_err(c, te == cudaErrorMemoryAllocation ? AQ_ERR_OOM : AQ_ERR_CUDA, "aq_scene_create: textures: %s", cudaGetErrorString(te));
        }
    }
    float lut[256];
    aq_build_srgb_lut(lut);
    AQ_TRY(upload(c, &s->d_lut, lut, 256));
    cudaError_t e = cudaMallocAsync((void**)&s->d_stats, AQS_WORDS * sizeof(unsigned long long), c->stream);
    if (e == cudaSuccess) e = cudaMallocAsync((void**)&s->d_ctrl_alloc, AQ_CTRL_ALLOC_BYTES, c->stream);
    if (e == cudaSuccess) {
        /* AQUA_DEBUG_CTRL_OFFSET (bytes, multiple of 128): where in its allocation the control block
         * sits — tools/ctrl_sweep.py uses it to show that render time no longer depends on the
         * address of the counters */
        size_t off = 0;
        if (const char* ev = std::getenv("AQUA_DEBUG_CTRL_OFFSET")) off = (size_t)std::strtoull(ev, nullptr, 0) & ~(size_t)127;
        if (off + AQC_WORDS * sizeof(uint32_t) > AQ_CTRL_ALLOC_BYTES) off = 0;
        s->d_ctrl = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(s->d_ctrl_alloc) + off);
    }
    if (e == cudaSuccess) e = cudaMemsetAsync(s->d_ctrl, 0, AQC_WORDS * sizeof(uint32_t), c->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(s->d_stats, 0, AQS_WORDS * sizeof(unsigned long long), c->stream);
    if (e == cudaSuccess) e = cudaEventCreate(&s->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&s->ev1);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream); /* staging vectors die here */
    if (e != cudaSuccess) {
        aq_scene_destroy(s);
        return set_err(c, AQ_ERR_CUDA, "aq_scene_create: %s", cudaGetErrorString(e));
    }
#undef AQ_TRY
    *out = s;
    return AQ_OK;
}

void aq_scene_destroy(aq_scene* s) {
    if (!s) return;
    cudaSetDevice(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    accel_try_upgrade(s, true); /* joins the background builder */
    accel_free_garbage(s);
    auto& live = s->ctx->scenes;
    live.erase(std::remove(live.begin(), live.end(), s), live.end());
    void* ptrs[] = {s->d_pos, s->d_nrm, s->d_uv, s->d_lut, s->d_lights, s->d_prim_light_pdf, s->d_idx, s->d_tri_mat,
                    s->d_texels, s->d_mats, s->d_shade_recs, s->d_tex_desc, s->d_nodes, s->d_tris,
                    s->d_stats, s->d_ctrl_alloc, s->d_nrc_w, s->d_nrc_m, s->d_nrc_v, s->d_nrc_x, s->d_nrc_y,
                    s->d_nrc_g, s->d_nrc_loss, s->d_nrc_loss_chunk, s->d_nrc_wt};
    /* scene memory comes from the device's stream-ordered pool (aq_init keeps the pool's memory
     * cached): cudaMalloc / cudaFree of a scene cost 25-250 ms per create/destroy cycle on this
     * platform, the pool makes both a few microseconds */
    for (void* p : ptrs)
        if (p) cudaFreeAsync(p, s->ctx->stream);
    for (cudaEvent_t e : s->ev_pace)
        if (e) cudaEventDestroy(e);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    for (cudaEvent_t e : s->prof_ev) cudaEventDestroy(e);
    delete s;
}

int aq_accel_build(aq_scene* s, aq_accel_info* info) {
    if (!s) return set_err(nullptr, AQ_ERR_BAD_ARG, "aq_accel_build: scene is null");
    aq_ctx* c = s->ctx;
    AQ_CK(c, cudaSetDevice(c->device));
    accel_try_upgrade(s, true); /* a rebuild first ends the previous background build */
    AQ_CK(c, cudaStreamSynchronize(c->stream));
    accel_free_garbage(s);
    auto t0 = std::chrono::steady_clock::now();
    if (s->d_nodes) cudaFreeAsync(s->d_nodes, c->stream);
    if (s->d_tris) cudaFreeAsync(s->d_tris, c->stream);
    s->d_nodes = nullptr;
    s->d_tris = nullptr;
    s->built = false;
    bool device = s->n_tris >= AQ_DEVICE_BUILD_MIN_TRIS;
    bool hybrid = !device && s->n_tris >= AQ_HYBRID_BUILD_MIN_TRIS;
    if (const char* e = std::getenv("AQUA_ACCEL_BUILDER")) {
        if (!std::strcmp(e, "device")) device = s->n_tris > 0, hybrid = false;
        if (!std::strcmp(e, "host")) device = false, hybrid = false;
        if (!std::strcmp(e, "hybrid")) device = false, hybrid = s->n_tris > 0;
    }
    device = device || hybrid;
    std::memset(&s->accel, 0, sizeof s->accel);
    if (device) {
        cudaError_t ce = cudaSuccess;
        uint32_t depth = 0;
        int rc = aq_build_bvh8_device(c->stream, s->d_pos, s->d_idx, s->n_tris, &s->d_nodes, &s->n_node_words,
                                      &s->d_tris, &depth, &ce, hybrid ? 1 : 0); /* hybrid: the 3 ms radix tree first */
        if (rc == -2) return set_err(c, ce == cudaErrorMemoryAllocation ? AQ_ERR_OOM : AQ_ERR_CUDA,
                                     "aq_accel_build (device): %s", cudaGetErrorString(ce));
        if (rc == 0) {
            s->n_tri_words = (size_t)s->n_tris * AQ_TRI_WORDS;
            s->accel.n_nodes = (uint32_t)(s->n_node_words / AQ_NODE_WORDS);
            s->accel.max_depth = depth;
            s->accel.builder = 1;
        } else {
            device = false; /* degenerate input for the LBVH: fall back to the host SAH builder */
            hybrid = false;
        }
    }
    if (device && hybrid) { /* usable now; the SAH tree follows (accel_try_upgrade) */
        s->accel.builder = 2;
        s->upg = new (std::nothrow) aq_scene::upgrade_t;
        if (s->upg) {
            s->upg->state.store(1);
            try {
                s->upg->th = std::thread(accel_upgrade_thread, s, c->device);
            } catch (...) { /* no thread: stay on the LBVH tree */
                delete s->upg;
                s->upg = nullptr;
                s->accel.builder = 1;
            }
        } else {
            s->accel.builder = 1;
        }
    }
    if (!device) {
        aq_bvh8 bvh;
        if (aq_build_bvh8(s->h_pos.data(), s->h_idx.data(), s->n_tris, 0, &bvh) != 0)
            return set_err(c, AQ_ERR_UNSUPPORTED, "aq_accel_build: BVH deeper than %d levels", AQ_STACK_MAX);
        int rc;
        if ((rc = upload(c, &s->d_nodes, bvh.nodes.data(), bvh.nodes.size())) != AQ_OK) return rc;
        /* keep at least one record so the pointer is valid */
        if (bvh.tris.empty()) bvh.tris.resize(AQ_TRI_WORDS);
        if ((rc = upload(c, &s->d_tris, bvh.tris.data(), bvh.tris.size())) != AQ_OK) return rc;
        AQ_CK(c, cudaStreamSynchronize(c->stream));
        s->n_node_words = bvh.nodes.size();
        s->n_tri_words = bvh.tris.size();
        s->accel.n_nodes = (uint32_t)(bvh.nodes.size() / AQ_NODE_WORDS);
        s->accel.max_depth = bvh.max_depth;
        s->accel.sah_cost = bvh.sah_cost;
        s->accel.builder = 0;
    }
    s->built = true;
    auto t1 = std::chrono::steady_clock::now();
    s->accel.n_tri_records = s->n_tris;
    s->accel.build_ms = (float)std::chrono::duration<double, std::milli>(t1 - t0).count();
    if (info) *info = s->accel;
    return AQ_OK;
}

int aq_accel_wait(aq_scene* s, aq_accel_info* info) {
    if (!s) return set_err(nullptr, AQ_ERR_BAD_ARG, "aq_accel_wait: scene is null");
    if (!s->built) return set_err(s->ctx, AQ_ERR_STATE, "aq_accel_wait: call aq_accel_build first");
    AQ_CK(s->ctx, cudaSetDevice(s->ctx->device));
    accel_try_upgrade(s, true);
    if (info) *info = s->accel;
    return AQ_OK;
}

int aq_accel_build_host(const float* positions, uint32_t n_verts, const uint32_t* indices,
                        uint32_t n_tris, void** nodes80, size_t* nodes_bytes, void** tris48,
                        size_t* tris_bytes, aq_accel_info* info) {
    if (!nodes80 || !nodes_bytes || !tris48 || !tris_bytes || (n_tris && (!positions || !indices)))
        return set_err(nullptr, AQ_ERR_BAD_ARG, "aq_accel_build_host: null argument");
    for (size_t i = 0; i < 3 * (size_t)n_tris; ++i)
        if (indices[i] >= n_verts) return set_err(nullptr, AQ_ERR_BAD_ARG, "aq_accel_build_host: index %zu out of range", i);
    auto t0 = std::chrono::steady_clock::now();
    aq_bvh8 bvh;
    if (aq_build_bvh8(positions, indices, n_tris, 0, &bvh) != 0)
        return set_err(nullptr, AQ_ERR_UNSUPPORTED, "aq_accel_build_host: BVH deeper than %d levels", AQ_STACK_MAX);
    auto t1 = std::chrono::steady_clock::now();
    *nodes_bytes = bvh.nodes.size() * sizeof(aq_u4);
    *tris_bytes = (size_t)n_tris * AQ_TRI_WORDS * sizeof(aq_f4);
    *nodes80 = std::malloc(*nodes_bytes ? *nodes_bytes : 1);
    *tris48 = std::malloc(*tris_bytes ? *tris_bytes : 1);
    if (!*nodes80 || !*tris48) return set_err(nullptr, AQ_ERR_OOM, "aq_accel_build_host: out of memory");
    std::memcpy(*nodes80, bvh.nodes.data(), *nodes_bytes);
    std::memcpy(*tris48, bvh.tris.data(), *tris_bytes);
    if (info) {
        info->n_nodes = (uint32_t)(bvh.nodes.size() / AQ_NODE_WORDS);
        info->n_tri_records = n_tris;
        info->max_depth = bvh.max_depth;
        info->sah_cost = bvh.sah_cost;
        info->build_ms = (float)std::chrono::duration<double, std::milli>(t1 - t0).count();
        info->builder = 0;
    }
    return AQ_OK;
}

void aq_free(void* p) { std::free(p); }

int aq_accel_download(aq_scene* s, void* nodes80, size_t nodes_bytes, void* tris48, size_t tris_bytes) {
    if (!s) return set_err(nullptr, AQ_ERR_BAD_ARG, "aq_accel_download: scene is null");
    aq_ctx* c = s->ctx;
    if (!s->built) return set_err(c, AQ_ERR_STATE, "aq_accel_download: accel not built");
    AQ_CK(c, cudaSetDevice(c->device));
    accel_try_upgrade(s, true);
    size_t nb = s->n_node_words * sizeof(aq_u4), tb = (size_t)s->n_tris * AQ_TRI_WORDS * sizeof(aq_f4);
    if (nodes80) {
        if (nodes_bytes < nb) return set_err(c, AQ_ERR_BAD_ARG, "aq_accel_download: node buffer too small (%zu < %zu)", nodes_bytes, nb);
        AQ_CK(c, cudaMemcpy(nodes80, s->d_nodes, nb, cudaMemcpyDeviceToHost));
    }
    if (tris48 && tb) {
        if (tris_bytes < tb) return set_err(c, AQ_ERR_BAD_ARG, "aq_accel_download: triangle buffer too small (%zu < %zu)", tris_bytes, tb);
        AQ_CK(c, cudaMemcpy(tris48, s->d_tris, tb, cudaMemcpyDeviceToHost));
    }
    return AQ_OK;
}

/* ------------------------------------------------------------------ intersection hook */
int aq_intersect_device_async(aq_scene* s, const void* d_rays, uint32_t n, void* d_hits, int any_hit) {
    if (!s) return set_err(nullptr, AQ_ERR_BAD_ARG, "aq_intersect: scene is null");
    aq_ctx* c = s->ctx;
    if (!s->built) return set_err(c, AQ_ERR_STATE, "aq_intersect: call aq_accel_build first");
    if (n == 0) return AQ_OK;
    if (!d_rays || !d_hits) return set_err(c, AQ_ERR_BAD_ARG, "aq_intersect: null buffer");
    AQ_CK(c, cudaSetDevice(c->device));
    accel_try_upgrade(s, false);
    uint32_t* fetch = &s->d_ctrl[any_hit ? AQC_FETCH_SHADOW : AQC_FETCH_CLOSEST];
    AQ_CK(c, cudaMemsetAsync(fetch, 0, sizeof(uint32_t), c->stream));
    const float4* r = (const float4*)d_rays;
    /* caller-supplied ray sets are taken to be incoherent: balanced triangle phase (AQUA_TRI_DYN=0 turns it off) */
    bool dyn = true;
    if (const char* e = std::getenv("AQUA_TRI_DYN")) dyn = e[0] != '0';
    auto k = any_hit ? (dyn ? aq_k_trace<2, true, AQ_TRI_DYN> : aq_k_trace<2, true, 0>)
                     : (dyn ? aq_k_trace<0, true, AQ_TRI_DYN> : aq_k_trace<0, true, 0>);
    k<<<resident_grid(c, k, AQ_TRACE_THREADS), AQ_TRACE_THREADS, 0, c->stream>>>(
        s->d_nodes, s->d_tris, r, r + 1, 2, nullptr, nullptr, nullptr, n, fetch, (uint4*)d_hits, nullptr,
        nullptr, 0, 0u, s->d_stats);
    AQ_CK(c, cudaGetLastError());
    return AQ_OK;
}

int aq_intersect(aq_scene* s, const aq_ray* rays, uint32_t n, aq_hit* hits, int any_hit) {
    if (!s) return set_err(nullptr, AQ_ERR_BAD_ARG, "aq_intersect: scene is null");
    aq_ctx* c = s->ctx;
    if (n == 0) return s->built ? AQ_OK : set_err(c, AQ_ERR_STATE, "aq_intersect: call aq_accel_build first");
    if (!rays || !hits) return set_err(c, AQ_ERR_BAD_ARG, "aq_intersect: null buffer");
    AQ_CK(c, cudaSetDevice(c->device));
    int rc = ensure_scratch(s, n);
    if (rc != AQ_OK) return rc;
    AQ_CK(c, cudaMemcpyAsync(s->ctx->d_scratch_rays, rays, (size_t)n * sizeof(aq_ray), cudaMemcpyHostToDevice, c->stream));
    rc = aq_intersect_device_async(s, s->ctx->d_scratch_rays, n, s->ctx->d_scratch_hits, any_hit);
    if (rc != AQ_OK) return rc;
    AQ_CK(c, cudaMemcpyAsync(hits, s->ctx->d_scratch_hits, (size_t)n * sizeof(aq_hit), cudaMemcpyDeviceToHost, c->stream));
    AQ_CK(c, cudaStreamSynchronize(c->stream));
    return AQ_OK;
}

int aq_trace_counters(aq_scene* s, uint64_t* nodes_fetched, uint64_t* tris_fetched, int reset) {
    if (!s) return set_err(nullptr, AQ_ERR_BAD_ARG, "aq_trace_counters: scene is null");
    aq_ctx* c = s->ctx;
    AQ_CK(c, cudaSetDevice(c->device));
    AQ_CK(c, cudaStreamSynchronize(c->stream));
    unsigned long long h[AQS_WORDS];
    AQ_CK(c, cudaMemcpy(h, s->d_stats, sizeof h, cudaMemcpyDeviceToHost));
    if (nodes_fetched) *nodes_fetched = h[AQS_NODES];
    if (tris_fetched) *tris_fetched = h[AQS_TRIS];
    if (reset) AQ_CK(c, cudaMemset(s->d_stats, 0, sizeof h));
    return AQ_OK;
}

/* ------------------------------------------------------------------ render */
int aq_render_device_async(aq_scene* s, const aq_integrator_cfg* cfg, void* d_film_ext) {
    if (!s || !cfg) return set_err(s ? s->ctx : nullptr, AQ_ERR_BAD_ARG, "aq_render: null argument");
    aq_ctx* c = s->ctx;
    if (!s->built) return set_err(c, AQ_ERR_STATE, "aq_render: call aq_accel_build first");
    uint32_t W = cfg->width ? cfg->width : s->camera.res[0];
    uint32_t H = cfg->height ? cfg->height : s->camera.res[1];
    if (W == 0 || H == 0 || (uint64_t)W * H > 0x7FFFFFFFull)
        return set_err(c, AQ_ERR_BAD_ARG, "aq_render: bad resolution %ux%u", W, H);
    if (cfg->spp_end < cfg->spp_begin) return set_err(c, AQ_ERR_BAD_ARG, "aq_render: spp_end < spp_begin");
    if (cfg->max_depth == 0 || cfg->max_depth > 64) return set_err(c, AQ_ERR_BAD_ARG, "aq_render: max_depth must be in 1..64");
    AQ_CK(c, cudaSetDevice(c->device));
    const uint64_t npix = (uint64_t)W * H;
    uint32_t pool = cfg->pool_paths ? cfg->pool_paths : AQ_DEFAULT_POOL;
    if (pool < 1024) pool = 1024;
    int rc = ensure_pool(s, pool);
    if (rc != AQ_OK) return rc;
    /* (a ctx keeps its largest pool; waves are still sized by what this call asked for) */
    float4* film = (float4*)d_film_ext;
    if (!film) {
        if (s->ctx->film_pixels < npix) {
            if (s->ctx->d_film) cudaFree(s->ctx->d_film);
            s->ctx->d_film = nullptr;
            s->ctx->film_pixels = 0;
            AQ_CK(c, cudaMalloc((void**)&s->ctx->d_film, npix * sizeof(float4)));
            s->ctx->film_pixels = npix;
        }
        film = s->ctx->d_film;
    }
    const uint32_t nspp = cfg->spp_end - cfg->spp_begin;
    float4* samples = nullptr;
    if (cfg->flags & AQ_RENDER_DUMP_SAMPLES) {
        size_t need = (size_t)nspp * npix;
        if (s->ctx->samples_count < need) {
            if (s->ctx->d_samples) cudaFree(s->ctx->d_samples);
            s->ctx->d_samples = nullptr;
            s->ctx->samples_count = 0;
            AQ_CK(c, cudaMalloc((void**)&s->ctx->d_samples, (need ? need : 1) * sizeof(float4)));
            s->ctx->samples_count = need;
        }
        samples = s->ctx->d_samples;
    }
    cudaStream_t st = c->stream;
    AQ_CK(c, cudaEventRecord(s->ev0, st));
    if (!(cfg->flags & AQ_RENDER_ACCUMULATE)) AQ_CK(c, cudaMemsetAsync(film, 0, npix * sizeof(float4), st));
    AQ_CK(c, cudaMemsetAsync(s->d_stats, 0, AQS_WORDS * sizeof(unsigned long long), st));

    aq_wave_params wp;
    wp.cam = aq_cam_derive(s->camera.translate, s->camera.rotate, s->camera.fov, s->camera.lens_radius,
                           s->camera.focal, W, H);
    wp.seed = cfg->seed;
    wp.max_depth = cfg->max_depth;
    wp.spp_begin = cfg->spp_begin;
    wp.npix = npix;
    wp.mis_mode = (cfg->flags & AQ_RENDER_MIS_NEE_ONLY)    ? AQ_MIS_NEE_ONLY
                  : (cfg->flags & AQ_RENDER_MIS_BSDF_ONLY) ? AQ_MIS_BSDF_ONLY
                                                           : AQ_MIS_BOTH;
    wp.skip_emit_depth = 0xFFFFFFFFu;
    const uint32_t tile_pixels = (uint32_t)(npix < pool ? npix : pool);
    uint32_t S = pool / tile_pixels;
    if (S < 1) S = 1;
    wave_launcher wv(s, cfg->flags);
    uint32_t waves = 0;
    /* AQ_RENDER_PROFILE brackets every launch of every AQ_PROF_STRIDE-th wave with events (an
     * event after every launch of every wave cost 2 % of the render); stage times are scaled
     * by waves / profiled waves in aq_render_finish */
    const bool prof_on = (cfg->flags & AQ_RENDER_PROFILE) != 0;
    bool prof = false;
    s->prof_waves = 0;
    s->prof_n = 0;
    s->prof_stage.clear();
    auto mark = [&](uint8_t stage) { /* stage: 0 raygen 1 closest 2 shade 3 shadow 4 film, 255 start */
        if (!prof) return;
        if (s->prof_n == s->prof_ev.size()) {
            cudaEvent_t e;
            if (cudaEventCreate(&e) != cudaSuccess) return;
            s->prof_ev.push_back(e);
        }
        cudaEventRecord(s->prof_ev[s->prof_n++], st);
        s->prof_stage.push_back(stage);
    };
    for (uint64_t tb = 0; tb < npix; tb += tile_pixels) {
        uint32_t tp = (uint32_t)((npix - tb) < tile_pixels ? (npix - tb) : tile_pixels);
        for (uint32_t s0 = cfg->spp_begin; s0 < cfg->spp_end; s0 += S) {
            uint32_t ns = cfg->spp_end - s0 < S ? cfg->spp_end - s0 : S;
            wp.tile_base = (uint32_t)tb;
            wp.tile_pixels = tp;
            wp.s0 = s0;
            wp.ns = ns;
            wp.n_paths = tp * ns;
            prof = prof_on && (waves % AQ_PROF_STRIDE) == 0;
            if (prof) ++s->prof_waves;
            if (s->upg) { /* hybrid build: the SAH tree takes over between two waves.  While it is pending the
                           * host enqueues a wave only when the one before has finished, or every wave would
                           * already be enqueued (with the LBVH pointers) by the time the tree arrives; the GPU
                           * idles for one launch latency per wave, for the ~4 waves the build lasts */
                if (waves >= 1 && s->ev_pace[(waves - 1) & 1]) cudaEventSynchronize(s->ev_pace[(waves - 1) & 1]);
                accel_try_upgrade(s, false);
            }
            mark(255);
            wv.raygen(wp);
            mark(0);
            for (uint32_t depth = 0; depth < cfg->max_depth; ++depth) {
                wv.closest(depth);
                mark(1);
                wv.shade(wp, depth);
                mark(2);
                wv.shadow(depth);
                mark(3);
            }
            wv.film(wp, film, samples);
            mark(4);
            if (s->upg) {
                cudaEvent_t& e = s->ev_pace[waves & 1];
                if (!e) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
                if (e) cudaEventRecord(e, st);
            }
            ++waves;
        }
    }
    const uint32_t launches = wv.launches;
    AQ_CK(c, cudaGetLastError());
    AQ_CK(c, cudaEventRecord(s->ev1, st));
    s->last_cfg = *cfg;
    s->last_cfg.width = W;
    s->last_cfg.height = H;
    s->last_launches = launches;
    s->last_waves = waves;
    s->render_pending = true;
    return AQ_OK;
}

int aq_render_finish(aq_scene* s, aq_stats* stats) {
    if (!s) return set_err(nullptr, AQ_ERR_BAD_ARG, "aq_render_finish: scene is null");
    aq_ctx* c = s->ctx;
    AQ_CK(c, cudaSetDevice(c->device));
    AQ_CK(c, cudaStreamSynchronize(c->stream));
    accel_free_garbage(s);
    if (stats) {
        std::memset(stats, 0, sizeof *stats);
        unsigned long long h[AQS_WORDS];
        AQ_CK(c, cudaMemcpy(h, s->d_stats, sizeof h, cudaMemcpyDeviceToHost));
        stats->samples = h[AQS_SAMPLES];
        stats->sample_bounces = h[AQS_BOUNCES];
        stats->rays_closest = h[AQS_RAYS_CLOSEST];
        stats->rays_shadow = h[AQS_RAYS_SHADOW];
        stats->nodes_fetched = h[AQS_NODES];
        stats->tris_fetched = h[AQS_TRIS];
        if (s->render_pending) {
            float ms = 0.f;
            AQ_CK(c, cudaEventElapsedTime(&ms, s->ev0, s->ev1));
            stats->ms_total = ms;
        }
        if (s->render_pending && s->prof_n > 1) {
            float acc[5] = {0, 0, 0, 0, 0};
            for (size_t i = 1; i < s->prof_n; ++i) {
                float ms = 0.f;
                if (cudaEventElapsedTime(&ms, s->prof_ev[i - 1], s->prof_ev[i]) == cudaSuccess &&
                    s->prof_stage[i] < 5)
                    acc[s->prof_stage[i]] += ms;
            }
            const float k = s->prof_waves ? (float)s->last_waves / (float)s->prof_waves : 0.f;
            stats->ms_raygen = acc[0] * k;
            stats->ms_trace = acc[1] * k;
            stats->ms_shade = acc[2] * k;
            stats->ms_shadow = acc[3] * k;
            stats->ms_film = acc[4] * k;
        }
        stats->n_launches = s->last_launches;
        stats->n_waves = s->last_waves;
    }
    s->render_pending = false;
    return AQ_OK;
}

int aq_render(aq_scene* s, const aq_integrator_cfg* cfg, float* film_out, aq_stats* stats) {
    if (!s || !cfg || !film_out) return set_err(s ? s->ctx : nullptr, AQ_ERR_BAD_ARG, "aq_render: null argument");
    aq_ctx* c = s->ctx;
    uint32_t W = cfg->width ? cfg->width : s->camera.res[0];
    uint32_t H = cfg->height ? cfg->height : s->camera.res[1];
    size_t bytes = (size_t)W * H * sizeof(float4);
    AQ_CK(c, cudaSetDevice(c->device));
    if ((cfg->flags & AQ_RENDER_ACCUMULATE)) {
        /* host film is the accumulator: push it first */
        if (s->ctx->film_pixels < (size_t)W * H) {
            if (s->ctx->d_film) cudaFree(s->ctx->d_film);
            s->ctx->d_film = nullptr;
            s->ctx->film_pixels = 0;
            AQ_CK(c, cudaMalloc((void**)&s->ctx->d_film, bytes));
            s->ctx->film_pixels = (size_t)W * H;
        }
        AQ_CK(c, cudaMemcpyAsync(s->ctx->d_film, film_out, bytes, cudaMemcpyHostToDevice, c->stream));
    }
    int rc = aq_render_device_async(s, cfg, nullptr);
    if (rc != AQ_OK) return rc;
    AQ_CK(c, cudaMemcpyAsync(film_out, s->ctx->d_film, bytes, cudaMemcpyDeviceToHost, c->stream));
    return aq_render_finish(s, stats);
}

int aq_render_samples(aq_scene* s, float* out, size_t n_float4) {
    if (!s || !out) return set_err(s ? s->ctx : nullptr, AQ_ERR_BAD_ARG, "aq_render_samples: null argument");
    aq_ctx* c = s->ctx;
    if (!s->ctx->d_samples || !(s->last_cfg.flags & AQ_RENDER_DUMP_SAMPLES))
        return set_err(c, AQ_ERR_STATE, "aq_render_samples: last render did not use AQ_RENDER_DUMP_SAMPLES");
    size_t have = (size_t)(s->last_cfg.spp_end - s->last_cfg.spp_begin) * s->last_cfg.width * s->last_cfg.height;
    if (n_float4 < have) return set_err(c, AQ_ERR_BAD_ARG, "aq_render_samples: buffer too small (%zu < %zu)", n_float4, have);
    AQ_CK(c, cudaSetDevice(c->device));
    AQ_CK(c, cudaStreamSynchronize(c->stream));
    AQ_CK(c, cudaMemcpy(out, s->ctx->d_samples, have * sizeof(float4), cudaMemcpyDeviceToHost));
    return AQ_OK;
}

int aq_generate_camera_rays(aq_scene* s, const aq_integrator_cfg* cfg, uint32_t sample, aq_ray* rays_out) {
    if (!s || !cfg || !rays_out) return set_err(s ? s->ctx : nullptr, AQ_ERR_BAD_ARG, "aq_generate_camera_rays: null argument");
    aq_ctx* c = s->ctx;
    uint32_t W = cfg->width ? cfg->width : s->camera.res[0];
    uint32_t H = cfg->height ? cfg->height : s->camera.res[1];
    size_t n = (size_t)W * H;
    AQ_CK(c, cudaSetDevice(c->device));
    int rc = ensure_scratch(s, n);
    if (rc != AQ_OK) return rc;
    aq_cam cam = aq_cam_derive(s->camera.translate, s->camera.rotate, s->camera.fov, s->camera.lens_radius,
                               s->camera.focal, W, H);
    aq_k_camera_rays<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(cam, cfg->seed, sample,
                                                                         (float4*)s->ctx->d_scratch_rays);
    AQ_CK(c, cudaGetLastError());
    AQ_CK(c, cudaMemcpyAsync(rays_out, s->ctx->d_scratch_rays, n * sizeof(aq_ray), cudaMemcpyDeviceToHost, c->stream));
    AQ_CK(c, cudaStreamSynchronize(c->stream));
    return AQ_OK;
}

}  // extern "C"

#include "aq_nrc_host.inl"

/* ---- hooks for aq_multi.cu (aq_internal.h) */
void* aq_internal_film(aq_scene* s) { return s->ctx->d_film; }
cudaStream_t aq_internal_stream(aq_scene* s) { return s->ctx->stream; }
int aq_internal_device(aq_scene* s) { return s->ctx->device; }
cudaStream_t aq_internal_ctx_stream(aq_ctx* c) { return c->stream; }
int aq_internal_ctx_device(aq_ctx* c) { return c->device; }
int aq_internal_set_error(aq_ctx* c, int code, const char* msg) { return set_err(c, code, "%s", msg); }
int aq_internal_clone_accel(aq_scene* dst, aq_scene* src) {
    aq_ctx* c = dst->ctx;
    if (!src->built) return set_err(c, AQ_ERR_STATE, "clone_accel: source accel not built");
    accel_try_upgrade(src, true); /* clone the final (SAH) tree */
    AQ_CK(c, cudaSetDevice(c->device));
    if (dst->d_nodes) cudaFreeAsync(dst->d_nodes, c->stream);
    if (dst->d_tris) cudaFreeAsync(dst->d_tris, c->stream);
    dst->d_nodes = nullptr;
    dst->d_tris = nullptr;
    dst->built = false;
    size_t nb = src->n_node_words * sizeof(aq_u4), tb = src->n_tri_words * sizeof(aq_f4);
    AQ_CK(c, cudaMallocAsync((void**)&dst->d_nodes, nb, c->stream));
    AQ_CK(c, cudaMallocAsync((void**)&dst->d_tris, tb, c->stream));
    AQ_CK(c, cudaStreamSynchronize(c->stream)); /* cudaMemcpyPeer below is not ordered against this stream */
    AQ_CK(c, cudaMemcpyPeer(dst->d_nodes, c->device, src->d_nodes, src->ctx->device, nb));
    AQ_CK(c, cudaMemcpyPeer(dst->d_tris, c->device, src->d_tris, src->ctx->device, tb));
    dst->n_node_words = src->n_node_words;
    dst->n_tri_words = src->n_tri_words;
    dst->accel = src->accel;
    dst->built = true;
    return AQ_OK;
}
