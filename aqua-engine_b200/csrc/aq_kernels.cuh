/*
 * aq_kernels.cuh — sm_100a wavefront kernels of the render hot path (SURVEY §8 rows
 * a4,a6..a13): raygen, persistent BVH8 closest-hit / any-hit traversal, shade (material +
 * NEE + BSDF sampling + queue compaction), film accumulation.
 *
 * Data flow of one wave (P path slots = tile_pixels x samples_per_wave):
 *
 *   raygen  -> rayq[0] (o|tmin, d|tmax, beta|slot), L[slot]=0
 *   for depth in 0..max_depth-1:
 *     closest : rayq[cur][i]            -> hit[i] (prim,t,u,v)             persistent warps
 *     shade   : rayq[cur][i], hit[i]    -> shq[j] (o|tmax, d|-, contrib|slot)  compacted
 *                                          rayq[nxt][k]                    compacted
 *     shadow  : shq[j]                  -> L[slot] += contrib if visible   persistent warps
 *   film    : film[pixel] += sum_s L[s*tile_pixels + pixel]  (ascending s: deterministic)
 *
 * Queue entries carry the ray and throughput, so the only gather/scatter is L[slot].
 * All counters live in a device control block; the host never reads them inside a wave.
 */
#ifndef AQ_KERNELS_CUH
#define AQ_KERNELS_CUH

#include <cuda_runtime.h>

#include "aq_bvh.h"
#include "aq_core.h"

#define AQ_TRACE_THREADS 128
#define AQ_SHADE_THREADS 256
#define AQ_SMEM_STACK 8 /* per-thread traversal stack entries held in shared memory */

/* control block (uint32 words) */
enum {
    AQC_NRAY0 = 0,
    AQC_NRAY1 = 1,
    AQC_NSHADOW = 2,
    AQC_FETCH_CLOSEST = 3,
    AQC_FETCH_SHADOW = 4,
    AQC_WORDS = 8
};
/* stats block (uint64 words) */
enum {
    AQS_SAMPLES = 0,
    AQS_BOUNCES = 1,
    AQS_RAYS_CLOSEST = 2,
    AQS_RAYS_SHADOW = 3,
    AQS_NODES = 4,
    AQS_TRIS = 5,
    AQS_WORDS = 8
};

struct aq_queue {
    float4* o_tmin;  /* origin.xyz, tmin   (shadow queue: origin.xyz, tmax) */
    float4* d_tmax;  /* dir.xyz, tmax      (shadow queue: dir.xyz, unused)  */
    float4* beta_id; /* throughput.rgb, slot bits  (shadow queue: contribution.rgb, slot bits) */
};

struct aq_wave_params {
    aq_cam cam;
    uint32_t tile_base, tile_pixels; /* pixels [tile_base, tile_base+tile_pixels) */
    uint32_t s0, ns;                 /* samples [s0, s0+ns) */
    uint32_t spp_begin;
    uint32_t seed, max_depth;
    uint32_t n_paths;                /* tile_pixels * ns */
    uint64_t npix;
};

/* ------------------------------------------------------------------ traversal stack:
 * first AQ_SMEM_STACK entries in shared memory (entry-major => conflict-free), the rest
 * spills to thread-local memory */
struct aq_smem_stack {
    uint2* sm; /* &smem[threadIdx.x], stride blockDim.x */
    uint32_t stride;
    uint2 spill[AQ_STACK_MAX - AQ_SMEM_STACK];
    int n;
    __device__ __forceinline__ void reset() { n = 0; }
    __device__ __forceinline__ bool empty() const { return n == 0; }
    __device__ __forceinline__ void push(uint32_t x, uint32_t y) {
        if (n < AQ_SMEM_STACK)
            sm[n * stride] = make_uint2(x, y);
        else
            spill[n - AQ_SMEM_STACK] = make_uint2(x, y);
        ++n;
    }
    __device__ __forceinline__ void pop(uint32_t& x, uint32_t& y) {
        --n;
        uint2 v = n < AQ_SMEM_STACK ? sm[n * stride] : spill[n - AQ_SMEM_STACK];
        x = v.x;
        y = v.y;
    }
};

/* ------------------------------------------------------------------ raygen (row a4) */
__global__ void __launch_bounds__(AQ_SHADE_THREADS)
aq_k_raygen(aq_wave_params wp, aq_queue q, float4* __restrict__ L, uint32_t* __restrict__ ctrl,
            unsigned long long* __restrict__ stats) {
    uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid == 0) {
        ctrl[AQC_NRAY0] = wp.n_paths;
        ctrl[AQC_NRAY1] = 0;
        ctrl[AQC_NSHADOW] = 0;
        ctrl[AQC_FETCH_CLOSEST] = 0;
        ctrl[AQC_FETCH_SHADOW] = 0;
        atomicAdd(&stats[AQS_SAMPLES], (unsigned long long)wp.n_paths);
    }
    for (uint32_t slot = gid; slot < wp.n_paths; slot += gridDim.x * blockDim.x) {
        uint32_t si = slot / wp.tile_pixels;
        uint32_t pixel = wp.tile_base + (slot - si * wp.tile_pixels);
        uint32_t key = aq_rng_key(wp.seed, pixel, wp.s0 + si);
        aq_rayf r = aq_camera_ray(wp.cam, pixel % wp.cam.width, pixel / wp.cam.width, key);
        q.o_tmin[slot] = make_float4(r.o.x, r.o.y, r.o.z, r.tmin);
        q.d_tmax[slot] = make_float4(r.d.x, r.d.y, r.d.z, r.tmax);
        q.beta_id[slot] = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(slot));
        L[slot] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
}

/* ------------------------------------------------------------------ traversal (rows a6,a7)
 * Persistent warps: each warp claims 32 consecutive queue entries with one atomicAdd.
 * MODE 0: closest hit -> hits[i]            (render + aq_intersect)
 * MODE 1: any hit, render: L[slot] += contrib when unoccluded
 * MODE 2: any hit, aq_intersect: hits[i].prim = 0 / AQ_MISS
 * ray arrays are addressed as ro[i*stride], rd[i*stride] (stride 2 = AoS aq_ray). */
template <int MODE, bool COUNT>
__global__ void __launch_bounds__(AQ_TRACE_THREADS)
aq_k_trace(const aq_u4* __restrict__ nodes, const aq_f4* __restrict__ tris,
           const float4* __restrict__ ro, const float4* __restrict__ rd, uint32_t stride,
           const float4* __restrict__ payload, const uint32_t* __restrict__ n_ptr, uint32_t n_imm,
           uint32_t* __restrict__ fetch_ctr, uint4* __restrict__ hits, float4* __restrict__ L,
           uint32_t* __restrict__ ctrl, int depth, unsigned long long* __restrict__ stats) {
    __shared__ uint2 s_stack[AQ_SMEM_STACK * AQ_TRACE_THREADS];
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n = n_ptr ? *n_ptr : n_imm;
    if (ctrl && blockIdx.x == 0 && threadIdx.x == 0) {
        if (MODE == 0) { /* closest(b): nobody uses these until shade(b) */
            ctrl[(depth & 1) ? AQC_NRAY0 : AQC_NRAY1] = 0;
            ctrl[AQC_NSHADOW] = 0;
            ctrl[AQC_FETCH_SHADOW] = 0;
            atomicAdd(&stats[AQS_RAYS_CLOSEST], (unsigned long long)n);
        } else {
            atomicAdd(&stats[AQS_RAYS_SHADOW], (unsigned long long)n);
        }
    }
    aq_smem_stack st;
    st.sm = s_stack + threadIdx.x;
    st.stride = AQ_TRACE_THREADS;
    aq_trav_counters cnt;
    cnt.nodes = 0;
    cnt.tris = 0;
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(fetch_ctr, 32u);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (base >= n) break;
        uint32_t i = base + lane;
        if (i < n) {
            float4 a = __ldg(ro + (size_t)i * stride), b = __ldg(rd + (size_t)i * stride);
            aq_v3 o = aq_mk(a.x, a.y, a.z), d = aq_mk(b.x, b.y, b.z);
            uint32_t prim;
            float t, u, v;
            if (MODE == 0) {
                aq_bvh8_trace<false, COUNT>(nodes, tris, o, d, a.w, b.w, st, prim, t, u, v, &cnt);
                hits[i] = make_uint4(prim, __float_as_uint(prim == AQ_MISS_ID ? b.w : t),
                                     __float_as_uint(u), __float_as_uint(v));
            } else if (MODE == 1) {
                /* shadow queue: a.w = tmax, tmin = 0 */
                bool occ = aq_bvh8_trace<true, COUNT>(nodes, tris, o, d, 0.0f, a.w, st, prim, t, u,
                                                      v, &cnt);
                if (!occ) {
                    float4 c = __ldg(payload + i);
                    uint32_t slot = __float_as_uint(c.w);
                    float4 l = L[slot];
                    l.x += c.x;
                    l.y += c.y;
                    l.z += c.z;
                    L[slot] = l;
                }
            } else {
                bool occ = aq_bvh8_trace<true, COUNT>(nodes, tris, o, d, a.w, b.w, st, prim, t, u,
                                                      v, &cnt);
                hits[i] = make_uint4(occ ? 0u : AQ_MISS_ID, 0u, 0u, 0u);
            }
        }
    }
    if (COUNT) {
        unsigned long long cn = cnt.nodes, ct = cnt.tris;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            cn += __shfl_xor_sync(0xFFFFFFFFu, cn, o);
            ct += __shfl_xor_sync(0xFFFFFFFFu, ct, o);
        }
        if (lane == 0 && stats) {
            atomicAdd(&stats[AQS_NODES], cn);
            atomicAdd(&stats[AQS_TRIS], ct);
        }
    }
}

/* ------------------------------------------------------------------ shade (rows a8-a11)
 * One thread per queue entry; block-aggregated compaction: warp ballots + one atomicAdd per
 * block and queue per iteration (a per-warp atomic on one address serialises in L2). */
__global__ void __launch_bounds__(AQ_SHADE_THREADS)
aq_k_shade(aq_scene_view sv, aq_wave_params wp, int depth, aq_queue cur, const uint4* __restrict__ hits,
           aq_queue nxt, aq_queue shq, float4* __restrict__ L, uint32_t* __restrict__ ctrl,
           unsigned long long* __restrict__ stats) {
    __shared__ uint32_t s_cnt[2][AQ_SHADE_THREADS / 32];
    __shared__ uint32_t s_base[2];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t n = ctrl[(depth & 1) ? AQC_NRAY1 : AQC_NRAY0];
    uint32_t* n_next = &ctrl[(depth & 1) ? AQC_NRAY0 : AQC_NRAY1];
    uint32_t* n_shadow = &ctrl[AQC_NSHADOW];
    if (blockIdx.x == 0 && threadIdx.x == 0) ctrl[AQC_FETCH_CLOSEST] = 0;
    uint32_t my_bounces = 0;
    const uint32_t step = gridDim.x * blockDim.x;
    for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += step) {
        uint32_t i = base + threadIdx.x;
        aq_vertex_out vo;
        vo.has_next = false;
        vo.has_shadow = false;
        uint32_t slot = 0, key = 0;
        if (i < n) {
            uint4 h = hits[i];
            if (h.x != AQ_MISS_ID) {
                float4 rdv = cur.d_tmax[i], bi = cur.beta_id[i];
                slot = __float_as_uint(bi.w);
                uint32_t si = slot / wp.tile_pixels;
                uint32_t pixel = wp.tile_base + (slot - si * wp.tile_pixels);
                key = aq_rng_key(wp.seed, pixel, wp.s0 + si);
                aq_vertex_in vi;
                aq_fetch_vertex(sv, h.x, __uint_as_float(h.z), __uint_as_float(h.w),
                                aq_mk(rdv.x, rdv.y, rdv.z), &vi);
                aq_shade_vertex(vi, aq_mk(bi.x, bi.y, bi.z), key, (uint32_t)depth, wp.max_depth,
                                sv.n_lights, sv.lights, &vo);
                ++my_bounces;
                if (vo.emitted.x != 0.0f || vo.emitted.y != 0.0f || vo.emitted.z != 0.0f) {
                    float4 l = L[slot];
                    l.x += vo.emitted.x;
                    l.y += vo.emitted.y;
                    l.z += vo.emitted.z;
                    L[slot] = l;
                }
            }
        }
        /* ---- compaction of both output queues */
        uint32_t bn = __ballot_sync(0xFFFFFFFFu, vo.has_next);
        uint32_t bs = __ballot_sync(0xFFFFFFFFu, vo.has_shadow);
        if (lane == 0) {
            s_cnt[0][warp] = __popc(bn);
            s_cnt[1][warp] = __popc(bs);
        }
        __syncthreads();
        if (warp == 0) {
            uint32_t cn = lane < AQ_SHADE_THREADS / 32 ? s_cnt[0][lane] : 0u;
            uint32_t cs = lane < AQ_SHADE_THREADS / 32 ? s_cnt[1][lane] : 0u;
            uint32_t pn = cn, ps = cs; /* inclusive warp scan over the 8 warp counts */
#pragma unroll
            for (int o = 1; o < AQ_SHADE_THREADS / 32; o <<= 1) {
                uint32_t tn = __shfl_up_sync(0xFFFFFFFFu, pn, o);
                uint32_t ts = __shfl_up_sync(0xFFFFFFFFu, ps, o);
                if (lane >= (uint32_t)o) {
                    pn += tn;
                    ps += ts;
                }
            }
            uint32_t tot_n = __shfl_sync(0xFFFFFFFFu, pn, AQ_SHADE_THREADS / 32 - 1);
            uint32_t tot_s = __shfl_sync(0xFFFFFFFFu, ps, AQ_SHADE_THREADS / 32 - 1);
            if (lane == 0) {
                s_base[0] = tot_n ? atomicAdd(n_next, tot_n) : 0u;
                s_base[1] = tot_s ? atomicAdd(n_shadow, tot_s) : 0u;
            }
            if (lane < AQ_SHADE_THREADS / 32) {
                s_cnt[0][lane] = pn - cn; /* exclusive */
                s_cnt[1][lane] = ps - cs;
            }
        }
        __syncthreads();
        if (vo.has_next) {
            uint32_t k = s_base[0] + s_cnt[0][warp] + __popc(bn & ((1u << lane) - 1u));
            nxt.o_tmin[k] = make_float4(vo.next.o.x, vo.next.o.y, vo.next.o.z, vo.next.tmin);
            nxt.d_tmax[k] = make_float4(vo.next.d.x, vo.next.d.y, vo.next.d.z, vo.next.tmax);
            nxt.beta_id[k] = make_float4(vo.beta.x, vo.beta.y, vo.beta.z, __uint_as_float(slot));
        }
        if (vo.has_shadow) {
            uint32_t k = s_base[1] + s_cnt[1][warp] + __popc(bs & ((1u << lane) - 1u));
            shq.o_tmin[k] = make_float4(vo.shadow.o.x, vo.shadow.o.y, vo.shadow.o.z, vo.shadow.tmax);
            shq.d_tmax[k] = make_float4(vo.shadow.d.x, vo.shadow.d.y, vo.shadow.d.z, 0.0f);
            shq.beta_id[k] = make_float4(vo.shadow_contrib.x, vo.shadow_contrib.y,
                                         vo.shadow_contrib.z, __uint_as_float(slot));
        }
        __syncthreads(); /* s_cnt / s_base are reused by the next iteration */
    }
    /* bounce counter: one atomic per block */
    uint32_t wsum = my_bounces;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wsum += __shfl_xor_sync(0xFFFFFFFFu, wsum, o);
    __shared__ uint32_t s_b[AQ_SHADE_THREADS / 32];
    if (lane == 0) s_b[warp] = wsum;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < AQ_SHADE_THREADS / 32; ++w) t += s_b[w];
        if (t) atomicAdd(&stats[AQS_BOUNCES], (unsigned long long)t);
    }
}

/* ------------------------------------------------------------------ film (row a12)
 * film[p] += (sum over the wave's samples in ascending order, count); no atomics: one
 * thread owns one pixel and waves are serialised on the stream. */
__global__ void __launch_bounds__(AQ_SHADE_THREADS)
aq_k_film(aq_wave_params wp, const float4* __restrict__ L, float4* __restrict__ film,
          float4* __restrict__ samples) {
    for (uint32_t pi = blockIdx.x * blockDim.x + threadIdx.x; pi < wp.tile_pixels;
         pi += gridDim.x * blockDim.x) {
        size_t pixel = (size_t)wp.tile_base + pi;
        float4 f = film[pixel];
        for (uint32_t si = 0; si < wp.ns; ++si) {
            float4 l = L[(size_t)si * wp.tile_pixels + pi];
            f.x += l.x;
            f.y += l.y;
            f.z += l.z;
            f.w += 1.0f;
            if (samples)
                samples[(size_t)(wp.s0 + si - wp.spp_begin) * wp.npix + pixel] =
                    make_float4(l.x, l.y, l.z, 1.0f);
        }
        film[pixel] = f;
    }
}

/* camera rays only (test hook aq_generate_camera_rays) */
__global__ void aq_k_camera_rays(aq_cam cam, uint32_t seed, uint32_t sample, float4* __restrict__ out) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= cam.width * cam.height) return;
    uint32_t key = aq_rng_key(seed, p, sample);
    aq_rayf r = aq_camera_ray(cam, p % cam.width, p / cam.width, key);
    out[2 * (size_t)p] = make_float4(r.o.x, r.o.y, r.o.z, r.tmin);
    out[2 * (size_t)p + 1] = make_float4(r.d.x, r.d.y, r.d.z, r.tmax);
}

#endif /* AQ_KERNELS_CUH */
