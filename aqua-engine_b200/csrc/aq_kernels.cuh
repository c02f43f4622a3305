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
 * gather/scatter is L[slot].  All counters live in device memory; the host never reads them
 * inside a wave.
 *
 * Queues are BLOCK-STRUCTURED (round 2): a queue is a list of blocks of AQ_QBLK entries, block b
 * = entries [b*AQ_QBLK, b*AQ_QBLK + cnt[b]).  A producing warp (shade) owns one open block and one
 * spare per output queue, compacts its survivors into them with __ballot_sync + a warp prefix
 * sum, and takes a new spare with ONE atomicAdd on the queue's block allocator whose result is
 * first needed AQ_QBLK/32 iterations later; a consuming warp claims whole blocks, one claim
 * ahead.  No kernel waits on a contended atomic any more: the per-32-entry claim / compaction
 * atomics of round 1 (one L2 line, up to 21 % of the shade kernel depending on the line's
 * ADDRESS) are gone, and with them the control-block placement calibration.  All blocks but the
 * <= 2 per producing warp that are still open at the end of a pass are full, so consumers see a
 * dense queue.
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
#define AQ_CLAIM 32 /* queue entries claimed per atomicAdd by a traversal warp (A/B on B200: 32 < 64 < 128 in time) */
#endif

#ifndef AQ_QBLK
#define AQ_QBLK 128 /* entries per queue block (power of two, multiple of 32) */
#endif

/* control block (uint32 words); every counter on its own 128-byte line */
enum {
    AQC_BLOCKS_RAY0 = 0,    /* blocks of the ray queue entering an even depth */
    AQC_BLOCKS_RAY1 = 32,   /* blocks of the ray queue entering an odd depth */
    AQC_BLOCKS_SHADOW = 64, /* blocks of the shadow queue */
    AQC_FETCH_CLOSEST = 96, /* unit claims of the three consumers */
    AQC_FETCH_SHADE = 128,
    AQC_FETCH_SHADOW = 160,
    AQC_SPARE = 192,
    AQC_WORDS = 224
};
#ifndef AQ_SHADE_CLAIM
#define AQ_SHADE_CLAIM AQ_QBLK /* queue entries a shade warp claims per atomicAdd */
#endif
static_assert(AQ_QBLK >= 32 && (AQ_QBLK & (AQ_QBLK - 1)) == 0, "AQ_QBLK: power of two >= 32");
static_assert(AQ_CLAIM >= 32 && AQ_CLAIM <= AQ_QBLK && AQ_QBLK % AQ_CLAIM == 0 && AQ_CLAIM % 32 == 0,
              "AQ_CLAIM: multiple of 32 that divides AQ_QBLK");
/* word index of the block count of the ray queue consumed at `depth` */
__host__ __device__ inline int aqc_blocks_ray(int depth) { return (depth & 1) ? AQC_BLOCKS_RAY1 : AQC_BLOCKS_RAY0; }

/* per-block entry counts of the three queues (device arrays owned by the aq_ctx) */
struct aq_qcounts {
    uint32_t* ray[2]; /* by depth parity */
    uint32_t* shadow;
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

/* balanced triangle phase of the traversal kernels (template parameter DYN of aq_k_trace; 0 = off: the
 * whole aq_trav_step per lane and iteration): the number of lanes with a triangle pending below which
 * the warp goes back to the nodes.  B200 A/B (profiles/r02b_ab_tri_dyn*.log): 6 and 8 equal, 4 / 10 / 12 worse. */
#ifndef AQ_TRI_DYN
#define AQ_TRI_DYN 6
#endif
/* the render passes use the balanced form from depth 1 on (and when the tree has at least this many
 * nodes): coherent camera rays, and the shadow rays of their hits, find their triangles together and only
 * pay the vote (measured: depth 0 included, cbox +1 %, room's first closest-hit launch +4 %; depth >= 1 only,
 * cbox -1.5 %, room -5.7 % of the render) */
#ifndef AQ_TRI_DYN_MIN_NODES
#define AQ_TRI_DYN_MIN_NODES 0
#endif
/* shade: L2 prefetch of the next iteration's queue words (B200: shade pass -6.7 % on cbox, -4.4 % on
 * room; also prefetching the next shading record into L1 through an early load of its primitive id
 * measured +4 % / +4.4 % instead and is not in the file, profiles/r02b_ab_shade_prefetch.log) */
#ifndef AQ_SHADE_PREFETCH
#define AQ_SHADE_PREFETCH 1
#endif
#ifndef AQ_TRAV_STEP2_CLOSEST_ONLY
#define AQ_TRAV_STEP2_CLOSEST_ONLY 0
#endif
#ifndef AQ_TRAV_STEP2_ANYHIT_ONLY
#define AQ_TRAV_STEP2_ANYHIT_ONLY 0
#endif

/* ------------------------------------------------------------------ traversal stack:
 * first AQ_SMEM_STACK entries in shared memory (entry-major => conflict-free), the rest
 * spills to thread-local memory.  The shared part is addressed through its 32-bit shared-window
 * address with st/ld.shared (a generic pointer member cost ~20 instructions per push or pop:
 * generic-address arithmetic plus a local-memory copy of the fill count, because the struct with
 * the spill array in it was never split into registers); the spill array lives outside the struct.
 * The asm statements are volatile (kept in program order among themselves) without a memory clobber:
 * nothing else touches the stack words, and the node loads stay free to move across a push. */
struct aq_smem_stack {
    uint32_t sm;  /* shared-window address of this thread's entry 0; entry k is AQ_TRACE_THREADS * 8 * k further */
    uint2* spill; /* AQ_STACK_CAP - AQ_SMEM_STACK thread-local entries */
    int n;
    __device__ __forceinline__ void bind(const uint2* smem_entry0, uint2* local_spill) {
        sm = (uint32_t)__cvta_generic_to_shared(smem_entry0);
        /* opaque copy: otherwise the compiler re-derives the address (S2R CgaCtaId, LEA, IMAD) at every push and pop */
        asm volatile("mov.u32 %0, %0;" : "+r"(sm));
        spill = local_spill;
    }
    __device__ __forceinline__ void reset() { n = 0; }
    __device__ __forceinline__ bool empty() const { return n == 0; }
    __device__ __forceinline__ uint2 lds(int k) const {
        uint2 v;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];\n" : "=r"(v.x), "=r"(v.y) : "r"(sm + (uint32_t)k * (AQ_TRACE_THREADS * 8u)));
        return v;
    }
    __device__ __forceinline__ uint32_t top_y() const {
        return n <= AQ_SMEM_STACK ? lds(n - 1).y : spill[n - 1 - AQ_SMEM_STACK].y;
    }
    __device__ __forceinline__ void push(uint32_t x, uint32_t y) {
        if (n < AQ_SMEM_STACK)
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};\n" ::"r"(sm + (uint32_t)n * (AQ_TRACE_THREADS * 8u)), "r"(x), "r"(y));
        else
            spill[n - AQ_SMEM_STACK] = make_uint2(x, y);
        ++n;
    }
    __device__ __forceinline__ void pop(uint32_t& x, uint32_t& y) {
        --n;
        uint2 v = n < AQ_SMEM_STACK ? lds(n) : spill[n - AQ_SMEM_STACK];
        x = v.x;
        y = v.y;
    }
};

/* a wave's first queue: n entries in slot order = ceil(n / AQ_QBLK) blocks, all full but the last */
__device__ __forceinline__ void aq_queue_start(uint32_t n, uint32_t* __restrict__ ctrl, uint32_t* __restrict__ qcnt0,
                                               uint32_t gid, uint32_t gsize) {
    const uint32_t nb = (n + AQ_QBLK - 1u) / AQ_QBLK;
    if (gid == 0) {
        ctrl[AQC_BLOCKS_RAY0] = nb;
        ctrl[AQC_FETCH_CLOSEST] = 0;
        ctrl[AQC_FETCH_SHADE] = 0;
        ctrl[AQC_FETCH_SHADOW] = 0;
    }
    for (uint32_t b = gid; b < nb; b += gsize) qcnt0[b] = n - b * AQ_QBLK < AQ_QBLK ? n - b * AQ_QBLK : AQ_QBLK;
}

/* flat walk of a block-structured queue (kernels that visit every entry once in any order):
 * is capacity index e in [0, n_blocks * AQ_QBLK) a live entry? */
__device__ __forceinline__ bool aq_queue_live(const uint32_t* __restrict__ cnt, uint32_t n_blocks, uint32_t e) {
    const uint32_t b = e / AQ_QBLK;
    return b < n_blocks && (e & (AQ_QBLK - 1u)) < __ldg(cnt + b);
}

/* ------------------------------------------------------------------ raygen (row a4) */
__global__ void __launch_bounds__(AQ_GEN_THREADS)
aq_k_raygen(aq_wave_params wp, aq_queue q, float4* __restrict__ L, uint32_t* __restrict__ ctrl,
            uint32_t* __restrict__ qcnt0, unsigned long long* __restrict__ stats) {
    uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    aq_queue_start(wp.n_paths, ctrl, qcnt0, gid, gridDim.x * blockDim.x);
    if (gid == 0) atomicAdd(&stats[AQS_SAMPLES], (unsigned long long)wp.n_paths);
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
 * no register staging) one pool ahead; idle lanes pick their ray up with three LDS.128.
 *
 * Work supply: the queue is a list of blocks of AQ_QBLK entries (cnt[b] valid entries each; a flat
 * array — the aq_intersect hook — is the same thing with every block full), claimed in units of
 * AQ_CLAIM entries (aq_unit_feed). */
__device__ __forceinline__ void aq_cp_async16(void* smem, const void* gmem) {
    uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void aq_cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void aq_cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

/* a consuming warp's work supply.  The queue's capacity space [0, n_blocks * AQ_QBLK) is cut into
 * units of UNIT entries (a unit never straddles a block); unit u holds the entries of its block
 * that fall into it: clamp(cnt[block] - offset, 0, UNIT).  A warp starts on unit warp_id,
 * holds the unit after that with its entry count already requested, and has the claim for the
 * third in flight — one atomicAdd per unit, issued a whole unit before its result is used, and
 * the dependent count load issued when the claim is consumed, again one unit early.
 * The queue description (cnt: per-block entry counts, or null for a flat array of n_flat entries;
 * fetch: the claim counter) is passed to every call so that it stays in the constant bank. */
template <uint32_t UNIT>
struct aq_unit_feed {
    static_assert(UNIT >= 32 && UNIT <= AQ_QBLK && AQ_QBLK % UNIT == 0 && UNIT % 32 == 0, "unit: multiple of 32 that divides AQ_QBLK");
    uint32_t nxt, nxt_cnt; /* the unit after the current one (count requested early) */
    uint32_t pend;         /* lane 0: the claim after that */
    static __device__ __forceinline__ uint32_t count_of(const uint32_t* cnt, uint32_t n_units, uint32_t n_flat, uint32_t u) {
        if (u >= n_units) return 0u;
        const uint32_t e = u * UNIT;
        const uint32_t have = cnt ? __ldcg(cnt + e / AQ_QBLK) + (e / AQ_QBLK) * AQ_QBLK : n_flat; /* end of the live entries */
        return have > e ? (have - e < UNIT ? have - e : UNIT) : 0u;
    }
    /* yields the first unit; afterwards next() yields the following ones (unit >= n_units: done) */
    __device__ __forceinline__ void start(const uint32_t* cnt, uint32_t* fetch, uint32_t n_units, uint32_t n_flat,
                                          uint32_t warp_id, uint32_t n_warps, uint32_t lane, uint32_t& u, uint32_t& n) {
        u = warp_id;
        n = count_of(cnt, n_units, n_flat, u);
        nxt = n_warps + warp_id;
        nxt_cnt = count_of(cnt, n_units, n_flat, nxt);
        pend = 0u;
        if (lane == 0 && nxt < n_units) pend = 2u * n_warps + atomicAdd(fetch, 1u);
    }
    __device__ __forceinline__ void next(const uint32_t* cnt, uint32_t* fetch, uint32_t n_units, uint32_t n_flat,
                                         uint32_t n_warps, uint32_t lane, uint32_t& u, uint32_t& n) {
        u = nxt;
        n = nxt_cnt;
        if (u >= n_units) return;
        nxt = __shfl_sync(0xFFFFFFFFu, pend, 0);
        nxt_cnt = count_of(cnt, n_units, n_flat, nxt);
        if (lane == 0 && nxt < n_units) pend = 2u * n_warps + atomicAdd(fetch, 1u);
    }
};

/* Two warp-cooperative forms of the triangle phase were built, verified bit-exact and measured in
 * round 2, and are NOT in this file any more (git history: "Experiment ... pipelined triangle tests";
 * DESIGN.md section 5, profiles/r02_ab_coop_tris.log, r02_ab_tri_pipeline*.{log,txt}): pooling the
 * (ray lane, triangle) pairs of one node visit, or of several visits, in shared memory and testing them
 * 32 at a time.  Both raise threads/instruction and both are slower: the first pays a fixed
 * pooling/sync cost per visit, the second lets rays walk on with a stale best-t and opens 17-40 % more
 * nodes and triangles. */
#ifndef AQ_TRACE_CLOSEST_MIN_BLOCKS
#define AQ_TRACE_CLOSEST_MIN_BLOCKS 8
#endif
#ifndef AQ_TRACE_ANY_MIN_BLOCKS
#define AQ_TRACE_ANY_MIN_BLOCKS 7
#endif
template <int MODE, bool COUNT, int DYN = 0>
__global__ void __launch_bounds__(AQ_TRACE_THREADS, COUNT ? 7 : MODE == 1 ? AQ_TRACE_ANY_MIN_BLOCKS : AQ_TRACE_CLOSEST_MIN_BLOCKS) /* 72 registers, or 64 for the render closest-hit pass (A/B on B200) */
aq_k_trace(const aq_u4* __restrict__ nodes, const aq_f4* __restrict__ tris,
              const float4* __restrict__ ro, const float4* __restrict__ rd, uint32_t stride,
              const float4* __restrict__ payload, const uint32_t* __restrict__ blk_cnt,
              const uint32_t* __restrict__ n_blocks_ptr, uint32_t n_imm,
              uint32_t* __restrict__ fetch_ctr, uint4* __restrict__ hits, float4* __restrict__ L,
              uint32_t* __restrict__ ctrl, int depth, uint32_t out_static_blocks,
              unsigned long long* __restrict__ stats) {
    constexpr int NW = AQ_TRACE_THREADS / 32;
    constexpr int NP = (MODE == 1) ? 3 : 2; /* float4 words per pooled ray */
    __shared__ uint2 s_stack[AQ_SMEM_STACK * AQ_TRACE_THREADS];
    __shared__ float4 s_pool[NW][2][NP][32];
    const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t n_blocks = n_blocks_ptr ? *n_blocks_ptr : (n_imm + AQ_QBLK - 1u) / AQ_QBLK;
    if (ctrl && blockIdx.x == 0 && threadIdx.x == 0) {
        if (MODE == 3) {
            /* the shade pass that follows starts its two output queues with 2 static blocks per warp */
            ctrl[aqc_blocks_ray(depth + 1)] = out_static_blocks;
            ctrl[AQC_BLOCKS_SHADOW] = out_static_blocks;
            ctrl[AQC_FETCH_SHADE] = 0;
            ctrl[AQC_FETCH_SHADOW] = 0;
        }
    }
    uint2 st_spill[AQ_STACK_CAP - AQ_SMEM_STACK];
    aq_smem_stack st;
    st.bind(s_stack + threadIdx.x, st_spill);
    aq_trav_counters cnt;
    cnt.nodes = 0;
    cnt.tris = 0;
    aq_trav T;
    T.tg_y = 0u;
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

    /* ---- chunk supply: 32-entry chunks of the current unit, then of the next non-empty unit */
    const uint32_t n_units = n_blocks * (AQ_QBLK / AQ_CLAIM);
    aq_unit_feed<AQ_CLAIM> feed;
    uint32_t unit, unit_n;       /* current unit and its entry count */
    uint32_t unit_pos = 0;       /* entries of it already handed out as chunks */
    uint32_t n_rays = 0;         /* entries this warp took (stats) */
    feed.start(blk_cnt, fetch_ctr, n_units, n_imm, blockIdx.x * NW + wib, gridDim.x * NW, lane, unit, unit_n);
    if (unit >= n_units) return; /* more warps than units */
    auto next_chunk = [&](uint32_t& c, uint32_t& count) {
        if (AQ_CLAIM == 32) { /* a unit is a chunk */
            if (unit_pos != 0u) feed.next(blk_cnt, fetch_ctr, n_units, n_imm, gridDim.x * NW, lane, unit, unit_n);
            while (unit_n == 0u && unit < n_units) feed.next(blk_cnt, fetch_ctr, n_units, n_imm, gridDim.x * NW, lane, unit, unit_n);
            unit_pos = 1u;
            c = unit * 32u;
            count = unit_n;
            n_rays += count;
            return;
        }
        while (unit_pos >= unit_n) {
            if (unit >= n_units) {
                c = 0u;
                count = 0u;
                return;
            }
            feed.next(blk_cnt, fetch_ctr, n_units, n_imm, gridDim.x * NW, lane, unit, unit_n);
            unit_pos = 0u;
        }
        c = unit * AQ_CLAIM + unit_pos;
        count = unit_n - unit_pos < 32u ? unit_n - unit_pos : 32u;
        unit_pos += 32u;
        n_rays += count;
    };

    /* ---- prologue: two pools in flight */
    uint32_t cur = 0, cur_base, cur_cnt, pool_pos = 0, nxt_base, nxt_cnt;
    next_chunk(cur_base, cur_cnt);
    next_chunk(nxt_base, nxt_cnt);
    fill(0, cur_base, cur_cnt);
    fill(1, nxt_base, nxt_cnt);
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
            next_chunk(nxt_base, nxt_cnt);
            fill(cur ^ 1u, nxt_base, nxt_cnt);
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
        /* ---- DYN > 0, balanced triangle phase: the step is taken apart at warp level.  Lanes without pending
         * triangles open their next node; then triangle tests run one per lane and iteration, and after the
         * first iteration only while at least DYN lanes still have one pending.  A lane that keeps triangles
         * does not open a node in the next step, so every ray sees exactly the operation order of
         * aq_trav_step (same culling, same hits) — only the warp stops waiting for the few lanes that hit
         * several leaves at once.  Room.json, depth >= 1: warp instructions -10 %, threads per instruction
         * 16.1 -> 20.6, closest-hit pass -8 %, any-hit -4 % (profiles/README.md, round 2b). */
        bool done = false;
        if (DYN > 0) {
            constexpr bool ANYM = (MODE == 1 || MODE == 2);
            /* T.tg_y != 0 only on lanes that hold a ray with triangles pending (idle lanes: 0) */
            if (active && T.tg_y == 0u) aq_trav_open_node<COUNT>(nodes, T, st, &cnt, T.tg_x, T.tg_y);
            if (__any_sync(0xFFFFFFFFu, T.tg_y != 0u)) {
                do { /* one vote per iteration is the whole cost of the policy */
                    if (T.tg_y != 0u) {
                        const uint32_t i = aq_msb(T.tg_y);
                        T.tg_y &= ~(1u << i);
                        if (aq_trav_test_tri<ANYM, COUNT>(tris, T, T.tg_x + i, &cnt)) {
                            done = true;
                            T.tg_y = 0u;
                        }
                    }
                } while (__popc(__ballot_sync(0xFFFFFFFFu, T.tg_y != 0u)) >= DYN);
            }
            if (active && !done && T.tg_y == 0u && T.ng_y <= 0x00FFFFFFu) {
                if (st.empty())
                    done = true;
                else
                    st.pop(T.ng_x, T.ng_y);
            }
        }
        if (active) {
            if (DYN > 0) {
            } else if (MODE == 0 || MODE == 3) {
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
    /* ray counter: one atomic per warp at the very end */
    if ((MODE == 1 || MODE == 3) && lane == 0 && n_rays && stats)
        atomicAdd(&stats[MODE == 3 ? AQS_RAYS_CLOSEST : AQS_RAYS_SHADOW], (unsigned long long)n_rays);
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
 * One thread per queue entry; a warp claims units of AQ_SHADE_CLAIM entries (aq_unit_feed) and
 * walks them 32 at a time (a static stride over the queue measured 18 % slower on cbox).
 * The three queue words of an entry are loaded back to back before anything
 * depends on them.  Compaction of the two output queues (continuation rays, shadow rays) is
 * warp-local: a ballot and a prefix popc place the survivors of an iteration behind the warp's
 * fill pointer in its open output block (aq_block_writer); no atomic result is waited for and
 * there is no block barrier.
 * AREA / FULL select the instantiation of the vertex code (aq_core.h): emissive triangles in the
 * scene / a material with clearcoat, transmission or subsurface.  Both shipped scenes run
 * <false,false>. */

/* a producing warp's side of one block-structured queue: an open block, a spare block and the
 * allocator it takes further spares from.  Blocks [0, 2*n_warps) are static (warp w: w and
 * n_warps + w); the allocator word was preset to 2*n_warps by the previous kernel and ends up as
 * the queue's block count. */
struct aq_block_writer {
    uint32_t pos;   /* warp-uniform write pointer: open block * AQ_QBLK + its fill */
    uint32_t spare; /* lane 0 only: the block that follows (possibly an atomic still in flight) */
    __device__ __forceinline__ void start(uint32_t warp_id, uint32_t n_warps) {
        pos = warp_id * AQ_QBLK;
        spare = n_warps + warp_id;
    }
    /* queue index for the survivor of rank `rank` among the `c` (> 0) survivors of this iteration;
     * call with the whole warp converged.  cnt: the queue's per-block entry counts; alloc: its
     * block allocator (= block count) */
    __device__ __forceinline__ uint32_t place(uint32_t* cnt, uint32_t* alloc, uint32_t rank, uint32_t c, uint32_t lane) {
        uint32_t k = pos + rank;
        const uint32_t fill = (pos & (AQ_QBLK - 1u)) + c;
        if (fill >= AQ_QBLK) { /* warp-uniform: the open block fills up in this iteration */
            const uint32_t sp = __shfl_sync(0xFFFFFFFFu, spare, 0);
            const uint32_t cur_base = pos & ~(AQ_QBLK - 1u);
            if (lane == 0) {
                cnt[cur_base / AQ_QBLK] = AQ_QBLK;
                spare = atomicAdd(alloc, 1u); /* needed again one block (>= AQ_QBLK/32 iterations) later */
            }
            if (k >= cur_base + AQ_QBLK) k += sp * AQ_QBLK - (cur_base + AQ_QBLK);
            pos = sp * AQ_QBLK + (fill - AQ_QBLK);
        } else {
            pos += c;
        }
        return k;
    }
    __device__ __forceinline__ void finish(uint32_t* cnt, uint32_t lane) {
        if (lane == 0) {
            cnt[pos / AQ_QBLK] = pos & (AQ_QBLK - 1u);
            cnt[spare] = 0u;
        }
    }
};

template <bool AREA, bool FULL>
__global__ void __launch_bounds__(AQ_SHADE_THREADS, FULL ? AQ_SHADE_MIN_BLOCKS_FULL : AQ_SHADE_MIN_BLOCKS)
aq_k_shade(aq_scene_view sv, aq_wave_params wp, int depth, aq_queue cur, const uint4* __restrict__ hits,
           aq_queue nxt, aq_queue shq, float4* __restrict__ L, uint32_t* __restrict__ ctrl, aq_qcounts qc,
           unsigned long long* __restrict__ stats) {
    constexpr uint32_t NW = AQ_SHADE_THREADS / 32;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t warp_id = blockIdx.x * NW + (threadIdx.x >> 5), n_warps = gridDim.x * NW;
    const uint32_t n_blocks = ctrl[aqc_blocks_ray(depth)];
    if (blockIdx.x == 0 && threadIdx.x == 0) ctrl[AQC_FETCH_CLOSEST] = 0;
    const uint32_t* cnt_in = qc.ray[depth & 1];
    uint32_t* cnt_next = qc.ray[(depth & 1) ^ 1];
    uint32_t* alloc_next = &ctrl[aqc_blocks_ray(depth + 1)];
    aq_block_writer wn, ws;
    wn.start(warp_id, n_warps);
    ws.start(warp_id, n_warps);
    uint32_t my_bounces = 0;
    /* input: whole blocks, claimed one ahead (AQ_SHADE_CLAIM entries per atomicAdd) */
    const uint32_t n_units = n_blocks * (AQ_QBLK / AQ_SHADE_CLAIM);
    aq_unit_feed<AQ_SHADE_CLAIM> feed;
    uint32_t unit, unit_n;
    feed.start(cnt_in, &ctrl[AQC_FETCH_SHADE], n_units, 0u, warp_id, n_warps, lane, unit, unit_n);
    for (; unit < n_units; feed.next(cnt_in, &ctrl[AQC_FETCH_SHADE], n_units, 0u, n_warps, lane, unit, unit_n)) {
        for (uint32_t k0 = 0; k0 < unit_n; k0 += 32u) {
            const uint32_t cn = unit_n - k0;
            const uint32_t i = unit * AQ_SHADE_CLAIM + k0 + lane;
            aq_vertex_out vo;
            vo.has_next = false;
            vo.has_shadow = false;
            uint32_t slot = 0, key = 0;
#if AQ_SHADE_PREFETCH
            /* the queue words of the warp's next iteration: DRAM -> L2 while this vertex is shaded (the entries
             * were written by the previous kernel and are long gone from the caches at pool 2^24) */
            const uint32_t ni = k0 + 32u < unit_n ? i + 32u : feed.nxt * AQ_SHADE_CLAIM + lane;
            const bool have_next = k0 + 32u < unit_n || feed.nxt < n_units;
            if (have_next) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(&hits[ni]));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(&cur.d_tmax[ni]));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(&cur.beta_id[ni]));
                if (AREA) asm volatile("prefetch.global.L2 [%0];" ::"l"(&cur.o_tmin[ni]));
            }
#endif
            if (lane < cn) {
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
            if (bn != 0u) {
                const uint32_t k = wn.place(cnt_next, alloc_next, __popc(bn & lt), __popc(bn), lane);
                if (vo.has_next) {
                    AQ_QST(&nxt.o_tmin[k], make_float4(vo.next.o.x, vo.next.o.y, vo.next.o.z, vo.next_pdf));
                    AQ_QST(&nxt.d_tmax[k], make_float4(vo.next.d.x, vo.next.d.y, vo.next.d.z, __uint_as_float(key)));
                    AQ_QST(&nxt.beta_id[k], make_float4(vo.beta.x, vo.beta.y, vo.beta.z, __uint_as_float(slot)));
                }
            }
            if (bs != 0u) {
                const uint32_t k = ws.place(qc.shadow, &ctrl[AQC_BLOCKS_SHADOW], __popc(bs & lt), __popc(bs), lane);
                if (vo.has_shadow) {
                    AQ_QST(&shq.o_tmin[k], make_float4(vo.shadow.o.x, vo.shadow.o.y, vo.shadow.o.z, vo.shadow.tmax));
                    AQ_QST(&shq.d_tmax[k], make_float4(vo.shadow.d.x, vo.shadow.d.y, vo.shadow.d.z, 0.0f));
                    AQ_QST(&shq.beta_id[k], make_float4(vo.shadow_contrib.x, vo.shadow_contrib.y,
                                                        vo.shadow_contrib.z, __uint_as_float(slot)));
                }
            }
        }
    }
    wn.finish(cnt_next, lane);
    ws.finish(qc.shadow, lane);
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
