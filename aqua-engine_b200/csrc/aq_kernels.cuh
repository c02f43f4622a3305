/*
 * aq_kernels.cuh — sm_100a wavefront kernels of the render hot path (SURVEY §8 rows
 * a4,a6..a13): raygen, persistent BVH8 closest-hit / any-hit traversal, shade (material +
 * NEE + BSDF sampling + queue compaction), film accumulation.
 *
 * Data flow of one wave (P path slots = tile_pixels x samples_per_wave):
 *
 *   raygen  -> rayq[0] (o|-, d|key, beta|slot), L[slot]=0
 *   for depth in 0..max_depth-1:
 *     closest : rayq[cur][i]            -> hit[i] (prim,t,u,v)             persistent warps
 *     shade   : rayq[cur][i], hit[i]    -> shq[j] (o|tmax, d|-, contrib|slot)  compacted
 *                                          rayq[nxt][k]                    compacted
 *     shadow  : shq[j]                  -> L[slot] += contrib if visible   persistent warps
 *   film    : film[pixel] += sum_s L[s*tile_pixels + pixel]  (ascending s: deterministic)
 *
 * Queue entries carry the ray, the throughput and the path's RNG key, so the only
 * gather/scatter is L[slot].  All counters live in a device control block; the host never
 * reads them inside a wave.
 */
#ifndef AQ_KERNELS_CUH
#define AQ_KERNELS_CUH

#include <cuda_runtime.h>

#include "aq_bvh.h"
#include "aq_core.h"

#define AQ_TRACE_THREADS 128
#define AQ_SHADE_THREADS 128
#ifndef AQ_SHADE_MIN_BLOCKS
#define AQ_SHADE_MIN_BLOCKS 6
#endif
#ifndef AQ_SHADE_MIN_BLOCKS_FULL
#define AQ_SHADE_MIN_BLOCKS_FULL 4 /* the full-Principled vertex code needs more registers */
#endif
#define AQ_GEN_THREADS 256
#define AQ_SMEM_STACK 8 /* per-thread traversal stack entries held in shared memory */
#ifndef AQ_CLAIM
#define AQ_CLAIM 32 /* rays claimed per atomicAdd by a traversal warp (A/B on B200: 32 < 64 < 128 in time) */
#endif

/* control block (uint32 words): two (n_rays, n_shadow) pairs alternating by depth parity,
 * each pair one 8-byte word so shade bumps both queue tails with ONE 64-bit atomic per warp */
enum {
    AQC_PAIR0 = 0, /* [0] rays entering an even depth, [1] shadow rays written at odd depths */
    AQC_PAIR1 = 2, /* [2] rays entering an odd depth,  [3] shadow rays written at even depths */
    AQC_FETCH_CLOSEST = 4,
    AQC_FETCH_SHADOW = 5,
    AQC_WORDS = 8
};
/* word index of the ray count consumed at `depth` / of the shadow count produced at `depth` */
__host__ __device__ inline int aqc_nray(int depth) { return (depth & 1) ? AQC_PAIR1 : AQC_PAIR0; }
__host__ __device__ inline int aqc_nshadow(int depth) { return ((depth & 1) ? AQC_PAIR0 : AQC_PAIR1) + 1; }

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
    float4* o_tmin;  /* origin.xyz, previous BSDF pdf (0 = camera ray)   (shadow queue: origin.xyz, tmax) */
    float4* d_tmax;  /* dir.xyz, RNG key   (shadow queue: dir.xyz, -)       */
    float4* beta_id; /* throughput.rgb, slot bits  (shadow queue: contribution.rgb, slot bits) */
};

/* queue traffic is write-once / read-once: stream it past the caches (ld/st.global.cs) so that
 * it does not evict the BVH and shading records, which are re-read by every ray (measured: +0.6 %
 * on cbox, neutral on room; an L2 persisting window on nodes / triangles / shading records: no
 * effect — traversal is issue-bound, not L2-miss bound) */
#define AQ_QST(ptr, val) __stcs((ptr), (val))
#define AQ_QLD(ptr) __ldcs(ptr)

struct aq_wave_params {
    aq_cam cam;
    uint32_t tile_base, tile_pixels; /* pixels [tile_base, tile_base+tile_pixels) */
    uint32_t s0, ns;                 /* samples [s0, s0+ns) */
    uint32_t spp_begin;
    uint32_t seed, max_depth;
    uint32_t n_paths;                /* tile_pixels * ns */
    uint32_t mis_mode;               /* AQ_MIS_* */
    uint32_t skip_emit_depth;        /* nrc records: the vertex at this depth does not add its own emission (else ~0u) */
    uint64_t npix;
};

#ifndef AQ_TRAV_STEP2_CLOSEST_ONLY
#define AQ_TRAV_STEP2_CLOSEST_ONLY 0
#endif
#ifndef AQ_TRAV_STEP2_ANYHIT_ONLY
#define AQ_TRAV_STEP2_ANYHIT_ONLY 0
#endif

/* ------------------------------------------------------------------ traversal stack:
 * first AQ_SMEM_STACK entries in shared memory (entry-major => conflict-free), the rest
 * spills to thread-local memory */
struct aq_smem_stack {
    uint2* sm; /* &smem[threadIdx.x], stride blockDim.x */
    uint2 spill[AQ_STACK_CAP - AQ_SMEM_STACK];
    int n;
    __device__ __forceinline__ void reset() { n = 0; }
    __device__ __forceinline__ bool empty() const { return n == 0; }
    __device__ __forceinline__ uint32_t top_y() const {
        return n <= AQ_SMEM_STACK ? sm[(n - 1) * AQ_TRACE_THREADS].y : spill[n - 1 - AQ_SMEM_STACK].y;
    }
    __device__ __forceinline__ void push(uint32_t x, uint32_t y) {
        if (n < AQ_SMEM_STACK)
            sm[n * AQ_TRACE_THREADS] = make_uint2(x, y);
        else
            spill[n - AQ_SMEM_STACK] = make_uint2(x, y);
        ++n;
    }
    __device__ __forceinline__ void pop(uint32_t& x, uint32_t& y) {
        --n;
        uint2 v = n < AQ_SMEM_STACK ? sm[n * AQ_TRACE_THREADS] : spill[n - AQ_SMEM_STACK];
        x = v.x;
        y = v.y;
    }
};

/* ------------------------------------------------------------------ raygen (row a4) */
__global__ void __launch_bounds__(AQ_GEN_THREADS)
aq_k_raygen(aq_wave_params wp, aq_queue q, float4* __restrict__ L, uint32_t* __restrict__ ctrl,
            unsigned long long* __restrict__ stats) {
    uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid == 0) {
        ctrl[AQC_PAIR0] = wp.n_paths;
        ctrl[AQC_PAIR0 + 1] = 0;
        ctrl[AQC_PAIR1] = 0;
        ctrl[AQC_PAIR1 + 1] = 0;
        ctrl[AQC_FETCH_CLOSEST] = 0;
        ctrl[AQC_FETCH_SHADOW] = 0;
        atomicAdd(&stats[AQS_SAMPLES], (unsigned long long)wp.n_paths);
    }
    for (uint32_t slot = gid; slot < wp.n_paths; slot += gridDim.x * blockDim.x) {
        uint32_t si = slot / wp.tile_pixels;
        uint32_t pixel = wp.tile_base + (slot - si * wp.tile_pixels);
        uint32_t key = aq_rng_key(wp.seed, pixel, wp.s0 + si);
        aq_rayf r = aq_camera_ray(wp.cam, pixel % wp.cam.width, pixel / wp.cam.width, key);
        AQ_QST(&q.o_tmin[slot], make_float4(r.o.x, r.o.y, r.o.z, 0.0f));
        /* wavefront rays always have tmin = 0, tmax = inf: the .w lane carries the RNG key */
        AQ_QST(&q.d_tmax[slot], make_float4(r.d.x, r.d.y, r.d.z, __uint_as_float(key)));
        AQ_QST(&q.beta_id[slot], make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(slot)));
        L[slot] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
}

/* ------------------------------------------------------------------ traversal (rows a6,a7)
 * Persistent warps with per-lane work refill.
 *   MODE 0: closest hit, aq_intersect    (tmin/tmax from the ray, AoS stride 2)
 *   MODE 1: any hit, render              (L[slot] += contribution when unoccluded)
 *   MODE 2: any hit, aq_intersect        (hits[i].prim = 0 / AQ_MISS)
 *   MODE 3: closest hit, render          (tmin = 0, tmax = inf; d.w is the RNG key)
 * ray arrays are addressed as ro[i*stride], rd[i*stride].
 *
 * A lane whose ray is finished takes the next ray at once instead of idling until the
 * slowest ray of a 32-ray chunk is done (with chunk-at-a-time scheduling only ~15 of 32
 * lanes executed per instruction after the first bounce, ncu r01 v0; per-lane refill made
 * the closest-hit pass on room.json 1.8x faster).
 *
 * Each warp owns two 32-ray pools in shared memory.  Pools are filled with cp.async (LDGSTS,
 * no register staging) one pool ahead, and the atomicAdd that claims the pool after that is
 * issued one rotation before its result is needed, so neither the claim nor the ray fetch is
 * ever on the critical path; idle lanes pick their ray up with three LDS.128. */
__device__ __forceinline__ void aq_cp_async16(void* smem, const void* gmem) {
    uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void aq_cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void aq_cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

template <int MODE, bool COUNT>
__global__ void __launch_bounds__(AQ_TRACE_THREADS, (COUNT || MODE == 1) ? 7 : 8) /* 72 registers, or 64 for the render closest-hit pass (A/B on B200) */
aq_k_trace(const aq_u4* __restrict__ nodes, const aq_f4* __restrict__ tris,
              const float4* __restrict__ ro, const float4* __restrict__ rd, uint32_t stride,
              const float4* __restrict__ payload, const uint32_t* __restrict__ n_ptr, uint32_t n_imm,
              uint32_t* __restrict__ fetch_ctr, uint4* __restrict__ hits, float4* __restrict__ L,
              uint32_t* __restrict__ ctrl, int depth, unsigned long long* __restrict__ stats) {
    constexpr int NW = AQ_TRACE_THREADS / 32;
    constexpr int NP = (MODE == 1) ? 3 : 2; /* float4 words per pooled ray */
    __shared__ uint2 s_stack[AQ_SMEM_STACK * AQ_TRACE_THREADS];
    __shared__ float4 s_pool[NW][2][NP][32];
    const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t n = n_ptr ? *n_ptr : n_imm;
    if (ctrl && blockIdx.x == 0 && threadIdx.x == 0) {
        if (MODE == 3) {
            ctrl[aqc_nray(depth + 1)] = 0;
            ctrl[aqc_nshadow(depth)] = 0;
            ctrl[AQC_FETCH_SHADOW] = 0;
            atomicAdd(&stats[AQS_RAYS_CLOSEST], (unsigned long long)n);
        } else if (MODE == 1) {
            atomicAdd(&stats[AQS_RAYS_SHADOW], (unsigned long long)n);
        }
    }
    aq_smem_stack st;
    st.sm = s_stack + threadIdx.x;
    aq_trav_counters cnt;
    cnt.nodes = 0;
    cnt.tris = 0;
    aq_trav T;
    bool active = false;
    uint32_t idx = 0;
    float4 pay = make_float4(0.f, 0.f, 0.f, 0.f), lacc = pay;

    auto fill = [&](uint32_t buf, uint32_t c, uint32_t count) {
        if (lane < count) {
            aq_cp_async16(&s_pool[wib][buf][0][lane], ro + (size_t)(c + lane) * stride);
            aq_cp_async16(&s_pool[wib][buf][1][lane], rd + (size_t)(c + lane) * stride);
            if (MODE == 1) aq_cp_async16(&s_pool[wib][buf][NP - 1][lane], payload + (c + lane));
        }
        aq_cp_async_commit();
    };
    auto count_of = [&](uint32_t c) { return c < n ? (n - c < 32u ? n - c : 32u) : 0u; };

    /* ---- prologue: two pools in flight, third claim pending */
    const uint32_t n_warps = gridDim.x * NW;
    const uint32_t warp_id = blockIdx.x * NW + wib;
    uint32_t dyn0 = 0u, c0 = 0u, c1 = 0u;
    if (MODE == 0 || MODE == 3) { /* static first two pools: no atomic storm at start-up */
        dyn0 = 2u * n_warps * 32u;
        c0 = warp_id * 32u;
        c1 = (n_warps + warp_id) * 32u;
    } else { /* shadow pass: time-ordered claims keep the L[slot] updates in a compact window */
        if (lane == 0) {
            c0 = atomicAdd(fetch_ctr, 32u);
            c1 = atomicAdd(fetch_ctr, 32u);
        }
        c0 = __shfl_sync(0xFFFFFFFFu, c0, 0);
        c1 = __shfl_sync(0xFFFFFFFFu, c1, 0);
    }
    uint32_t cur = 0, cur_base = c0, cur_cnt = count_of(c0), pool_pos = 0;
    uint32_t nxt_base = c1, nxt_cnt = count_of(c1);
    fill(0, c0, cur_cnt);
    fill(1, c1, nxt_cnt);
    /* a claim is AQ_CLAIM rays per atomicAdd, issued one claim ahead of its use */
    uint32_t blk_next = 0, blk_end = 0; /* unconsumed part of the last claim */
    uint32_t pending = 0;               /* lane 0: result of the claim after that */
    bool claiming = nxt_cnt != 0u;
    if (lane == 0 && claiming) pending = dyn0 + atomicAdd(fetch_ctr, (uint32_t)AQ_CLAIM);
    bool cur_ready = false; /* cp.async of the current pool waited for */

    for (;;) {
        /* ---- (a) rotate pools when the current one is used up */
        if (pool_pos >= cur_cnt && nxt_cnt != 0u) {
            aq_cp_async_wait_all();
            __syncwarp();
            cur ^= 1u;
            cur_base = nxt_base;
            cur_cnt = nxt_cnt;
            pool_pos = 0;
            cur_ready = true;
            uint32_t c;
            if (blk_next < blk_end) {
                c = blk_next;
                blk_next += 32u;
            } else if (claiming) {
                c = __shfl_sync(0xFFFFFFFFu, pending, 0);
                blk_next = c + 32u;
                blk_end = c + (uint32_t)AQ_CLAIM;
                claiming = c < n;
                if (lane == 0 && claiming) pending = dyn0 + atomicAdd(fetch_ctr, (uint32_t)AQ_CLAIM);
            } else {
                c = n;
            }
            nxt_base = c;
            nxt_cnt = count_of(c);
            fill(cur ^ 1u, c, nxt_cnt);
        }
        /* ---- (b) hand pool entries to idle lanes */
        const uint32_t idle = __ballot_sync(0xFFFFFFFFu, !active);
        if (idle != 0u && pool_pos < cur_cnt) {
            if (!cur_ready) { /* first pool of the kernel */
                aq_cp_async_wait_all();
                __syncwarp();
                cur_ready = true;
            }
            const uint32_t avail = cur_cnt - pool_pos;
            const uint32_t need = __popc(idle);
            const uint32_t take = need < avail ? need : avail;
            const uint32_t rank = __popc(idle & lt);
            if (!active && rank < take) {
                const uint32_t e = pool_pos + rank;
                const float4 ra = s_pool[wib][cur][0][e], rb = s_pool[wib][cur][1][e];
                idx = cur_base + e;
                float tmin, tmax;
                if (MODE == 3) {
                    tmin = 0.0f;
                    tmax = AQ_INF;
                } else if (MODE == 1) {
                    tmin = 0.0f;
                    tmax = ra.w;
                    pay = s_pool[wib][cur][NP - 1][e];
                    /* fetch the accumulator now: by the time the ray is known to be
                     * unoccluded the value has long arrived (one writer per slot per pass) */
                    lacc = L[__float_as_uint(pay.w)];
                } else {
                    tmin = ra.w;
                    tmax = rb.w;
                }
                aq_trav_init(T, aq_mk(ra.x, ra.y, ra.z), aq_mk(rb.x, rb.y, rb.z), tmin, tmax, st);
                active = true;
            }
            pool_pos += take;
            __syncwarp(); /* pool reads done before a later rotation overwrites the buffer */
        }
        /* ---- (c) one traversal step for every lane that holds a ray */
        if (__ballot_sync(0xFFFFFFFFu, active) == 0u) {
            if (pool_pos >= cur_cnt && nxt_cnt == 0u) break;
            continue;
        }
        if (active) {
            bool done;
            if (MODE == 0 || MODE == 3) {
                if (AQ_TRAV_STEP2 && !AQ_TRAV_STEP2_ANYHIT_ONLY)
                    done = aq_trav_step2<false, COUNT>(nodes, tris, T, st, &cnt);
                else
                    done = aq_trav_step<false, COUNT>(nodes, tris, T, st, &cnt);
            } else {
                if (AQ_TRAV_STEP2 && !AQ_TRAV_STEP2_CLOSEST_ONLY)
                    done = aq_trav_step2<true, COUNT>(nodes, tris, T, st, &cnt);
                else
                    done = aq_trav_step<true, COUNT>(nodes, tris, T, st, &cnt);
            }
            if (done) {
                active = false;
                if (MODE == 0 || MODE == 3) {
                    AQ_QST(&hits[idx], make_uint4(T.best_prim,
                                                  __float_as_uint(T.best_prim == AQ_MISS_ID ? T.tmax : T.best_t),
                                                  __float_as_uint(T.bu), __float_as_uint(T.bv)));
                } else if (MODE == 1) {
                    if (T.best_prim == AQ_MISS_ID) {
                        /* plain add, not atomicAdd: RED.F32 flushes denormals (FTZ) and would
                         * break bit parity with the oracle */
                        lacc.x += pay.x;
                        lacc.y += pay.y;
                        lacc.z += pay.z;
                        L[__float_as_uint(pay.w)] = lacc;
                    }
                } else {
                    hits[idx] = make_uint4(T.best_prim == AQ_MISS_ID ? AQ_MISS_ID : 0u, 0u, 0u, 0u);
                }
            }
        }
    }
    aq_cp_async_wait_all();
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
 * One thread per queue entry, grid-stride over the queue.  The three queue words of an
 * entry are loaded back to back before anything depends on them.  Compaction of the two
 * output queues (continuation rays, shadow rays) is warp-local: two ballots, and ONE packed
 * 64-bit atomicAdd per warp that advances both tails — no block barrier (the barrier-based
 * block aggregation of v0 cost 17 % of the kernel's stall samples at 2 CTAs/SM).
 * AREA / FULL select the instantiation of the vertex code (aq_core.h): emissive triangles in the
 * scene / a material with clearcoat, transmission or subsurface.  Both shipped scenes run
 * <false,false>. */
template <bool AREA, bool FULL>
__global__ void __launch_bounds__(AQ_SHADE_THREADS, FULL ? AQ_SHADE_MIN_BLOCKS_FULL : AQ_SHADE_MIN_BLOCKS)
aq_k_shade(aq_scene_view sv, aq_wave_params wp, int depth, aq_queue cur, const uint4* __restrict__ hits,
           aq_queue nxt, aq_queue shq, float4* __restrict__ L, uint32_t* __restrict__ ctrl,
           unsigned long long* __restrict__ stats) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n = ctrl[aqc_nray(depth)];
    /* (n_next, n_shadow) live in one aligned 8-byte word: low = next rays, high = shadow rays */
    unsigned long long* tails = reinterpret_cast<unsigned long long*>(&ctrl[aqc_nray(depth + 1)]);
    if (blockIdx.x == 0 && threadIdx.x == 0) ctrl[AQC_FETCH_CLOSEST] = 0;
    uint32_t my_bounces = 0;
    const uint32_t step = gridDim.x * blockDim.x;
    for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += step) {
        const uint32_t i = base + threadIdx.x;
        aq_vertex_out vo;
        vo.has_next = false;
        vo.has_shadow = false;
        uint32_t slot = 0, key = 0;
        if (i < n) {
            const uint4 h = AQ_QLD(&hits[i]);
            const float4 rdv = AQ_QLD(&cur.d_tmax[i]), bi = AQ_QLD(&cur.beta_id[i]);
            if (h.x != AQ_MISS_ID) {
                slot = __float_as_uint(bi.w);
                key = __float_as_uint(rdv.w);
                aq_vertex_in vi;
                aq_fetch_vertex<FULL>(sv, h.x, __uint_as_float(h.z), __uint_as_float(h.w),
                                aq_mk(rdv.x, rdv.y, rdv.z), &vi);
                vi.t_hit = __uint_as_float(h.y);
                /* the pdf of the BSDF sample that produced this ray is only needed for MIS
                 * against emissive triangles: scenes without them never read o.w */
                vi.prev_pdf = AREA ? cur.o_tmin[i].w : 0.0f;
                aq_shade_vertex<AREA, FULL>(vi, aq_mk(bi.x, bi.y, bi.z), key, (uint32_t)depth, wp.max_depth,
                                sv.n_lights, sv.lights, wp.mis_mode, &vo);
                ++my_bounces;
                if ((vo.emitted.x != 0.0f || vo.emitted.y != 0.0f || vo.emitted.z != 0.0f) &&
                    (uint32_t)depth != wp.skip_emit_depth) {
                    float4 l = L[slot];
                    l.x += vo.emitted.x;
                    l.y += vo.emitted.y;
                    l.z += vo.emitted.z;
                    L[slot] = l;
                }
            }
        }
        /* ---- warp-local compaction of both output queues */
        const uint32_t bn = __ballot_sync(0xFFFFFFFFu, vo.has_next);
        const uint32_t bs = __ballot_sync(0xFFFFFFFFu, vo.has_shadow);
        if ((bn | bs) != 0u) {
            unsigned long long basev = 0ull;
            if (lane == 0)
                basev = atomicAdd(tails, ((unsigned long long)__popc(bs) << 32) | (unsigned long long)__popc(bn));
            basev = __shfl_sync(0xFFFFFFFFu, basev, 0);
            const uint32_t lt = (1u << lane) - 1u;
            if (vo.has_next) {
                uint32_t k = (uint32_t)(basev & 0xFFFFFFFFull) + __popc(bn & lt);
                AQ_QST(&nxt.o_tmin[k], make_float4(vo.next.o.x, vo.next.o.y, vo.next.o.z, vo.next_pdf));
                AQ_QST(&nxt.d_tmax[k], make_float4(vo.next.d.x, vo.next.d.y, vo.next.d.z, __uint_as_float(key)));
                AQ_QST(&nxt.beta_id[k], make_float4(vo.beta.x, vo.beta.y, vo.beta.z, __uint_as_float(slot)));
            }
            if (vo.has_shadow) {
                uint32_t k = (uint32_t)(basev >> 32) + __popc(bs & lt);
                AQ_QST(&shq.o_tmin[k], make_float4(vo.shadow.o.x, vo.shadow.o.y, vo.shadow.o.z, vo.shadow.tmax));
                AQ_QST(&shq.d_tmax[k], make_float4(vo.shadow.d.x, vo.shadow.d.y, vo.shadow.d.z, 0.0f));
                AQ_QST(&shq.beta_id[k], make_float4(vo.shadow_contrib.x, vo.shadow_contrib.y,
                                                    vo.shadow_contrib.z, __uint_as_float(slot)));
            }
        }
    }
    /* bounce counter: one atomic per warp at the very end (persistent grid => few warps) */
    uint32_t wsum = my_bounces;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wsum += __shfl_xor_sync(0xFFFFFFFFu, wsum, o);
    if (lane == 0 && wsum) atomicAdd(&stats[AQS_BOUNCES], (unsigned long long)wsum);
}

/* ------------------------------------------------------------------ film (row a12)
 * film[p] += (sum over the wave's samples in ascending order, count); no atomics: one
 * thread owns one pixel and waves are serialised on the stream. */
__global__ void __launch_bounds__(AQ_GEN_THREADS)
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

/* per-triangle 128 B shading records from the uploaded mesh arrays (scene-create time) */
__global__ void aq_k_build_shade_recs(aq_scene_view sv, uint32_t n_tris, aq_f4* __restrict__ recs) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tris) return;
    aq_tri_shading g;
    aq_gather_tri(sv, t, &g);
    aq_pack_shade_rec(g, recs + (size_t)t * AQ_SHADE_REC_WORDS);
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
