/*
 * aq_nrc.cuh — sm_100a kernels of the `nrc` integrator (scenes/integrator.json:2; semantics in
 * aq_nrc.h).  They plug into the wavefront of aq_kernels.cuh:
 *
 *   records   nrc_raygen -> closest [-> shade -> closest] -> nrc_record -> (shade -> shadow ->
 *             closest)* -> nrc_targets              one pass per record depth (0: first hit, 1: second)
 *   training  per iteration: nrc_train_chunk (one CTA per AQ_NRC_CHUNK = 16 records: forward, loss, backward,
 *             per-chunk weight gradient, all in shared memory) -> nrc_adam (sum the chunks in
 *             order, Adam step)
 *   render    raygen -> closest -> shade -> shadow -> closest -> nrc_query (encode + MLP +
 *             L[slot] += emission + beta*fac*max(y,0)) -> film
 *
 * Every product is the ascending fmaf chain aq_nrc_dot defines (the register-blocked loops keep
 * each accumulator's order), so the weights after training, the loss curve and the rendered film
 * are bit-identical to the CPU oracle's.  The MLP runs on the fp32 pipes (B200, room.json at the
 * integrator's own sizes: training 159 ms once, query kernel ~10 ms per 8.2 M queries, ~27
 * TFLOP/s); a tcgen05 query kernel would trade the bit-exact parity of the render for a
 * tolerance — DESIGN.md §8.
 */
#ifndef AQ_NRC_CUH
#define AQ_NRC_CUH

#include <cuda_bf16.h>

#include "aq_kernels.cuh"
#include "aq_nrc.h"

#ifndef AQ_NRC_TRAIN_PER
#define AQ_NRC_TRAIN_PER 2 /* neurons per thread and layer: 512 threads per 16-record chunk (A/B on B200, 2048 iterations: 4 -> 76 ms, 2 -> 63 ms, 1 -> 86 ms) */
#endif
#define AQ_NRC_TRAIN_THREADS ((AQ_NRC_WIDTH / AQ_NRC_TRAIN_PER) * AQ_NRC_CHUNK) /* thread = (sample of the chunk, group of PER neurons) */
#define AQ_NRC_TRAIN_GROUPS (AQ_NRC_TRAIN_THREADS / AQ_NRC_CHUNK) /* thread = (sample, group) */
#define AQ_NRC_QUERY_THREADS 128
#define AQ_NRC_LD (AQ_NRC_CHUNK + 1) /* padded row of the [feature][sample] tiles: conflict-free both ways */

/* shared memory of nrc_train_chunk (floats) */
#define AQ_NRC_TRAIN_SMEM_FLOATS                                                                      \
    (AQ_NRC_N_MATS * AQ_NRC_WIDTH * AQ_NRC_LD /* activations a[0..4] */ + 2 * AQ_NRC_WIDTH * AQ_NRC_LD /* deltas */ + \
     AQ_NRC_WIDTH * AQ_NRC_WIDTH /* one matrix */ + 4 * AQ_NRC_CHUNK /* targets + live */ + 3 * AQ_NRC_CHUNK /* loss terms */)
#define AQ_NRC_QUERY_SMEM_FLOATS (2 * AQ_NRC_WIDTH * AQ_NRC_QUERY_THREADS + AQ_NRC_WIDTH * AQ_NRC_WIDTH)

/* ------------------------------------------------------------------ records */
/* slot k of the wave -> training record r = rec_first + 2k (all records of one wave share the
 * record depth); the camera ray of r's hashed pixel */
__global__ void __launch_bounds__(AQ_GEN_THREADS)
aq_k_nrc_raygen(aq_wave_params wp, uint32_t rec_first, aq_queue q, float4* __restrict__ L,
                uint32_t* __restrict__ ctrl, uint32_t* __restrict__ qcnt0, unsigned long long* __restrict__ stats) {
    uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    aq_queue_start(wp.n_paths, ctrl, qcnt0, gid, gridDim.x * blockDim.x);
    if (gid == 0) atomicAdd(&stats[AQS_SAMPLES], (unsigned long long)wp.n_paths);
    for (uint32_t slot = gid; slot < wp.n_paths; slot += gridDim.x * blockDim.x) {
        uint32_t r = rec_first + 2u * slot;
        uint32_t pixel = aq_nrc_record_pixel(wp.seed, r, (uint32_t)wp.npix);
        uint32_t key = aq_nrc_record_key(wp.seed, pixel, r);
        aq_rayf ray = aq_camera_ray(wp.cam, pixel % wp.cam.width, pixel / wp.cam.width, key);
        AQ_QST(&q.o_tmin[slot], make_float4(ray.o.x, ray.o.y, ray.o.z, 0.0f));
        AQ_QST(&q.d_tmax[slot], make_float4(ray.d.x, ray.d.y, ray.d.z, __uint_as_float(key)));
        AQ_QST(&q.beta_id[slot], make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(slot)));
        L[slot] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
}

/* the record vertex (after the closest-hit pass of the record depth, before its shade pass):
 * encode it into x[r], remember fac, restart the path's estimate (beta = 1, L = 0) */
template <bool FULL>
__global__ void __launch_bounds__(AQ_SHADE_THREADS)
aq_k_nrc_record(aq_scene_view sv, aq_nrc_bounds bb, int depth, uint32_t rec_first, aq_queue cur,
                const uint4* __restrict__ hits, float4* __restrict__ L, const uint32_t* __restrict__ ctrl,
                const uint32_t* __restrict__ qcnt, float* __restrict__ x, float4* __restrict__ y) {
    const uint32_t n_blocks = ctrl[aqc_blocks_ray(depth)];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_blocks * AQ_QBLK; i += gridDim.x * blockDim.x) {
        if (!aq_queue_live(qcnt, n_blocks, i)) continue;
        const uint4 h = hits[i];
        if (h.x == AQ_MISS_ID) continue;
        const float4 rdv = cur.d_tmax[i], bi = cur.beta_id[i];
        const uint32_t slot = __float_as_uint(bi.w);
        const uint32_t r = rec_first + 2u * slot;
        aq_vertex_in vi;
        aq_fetch_vertex<FULL>(sv, h.x, __uint_as_float(h.z), __uint_as_float(h.w), aq_mk(rdv.x, rdv.y, rdv.z), &vi);
        aq_v3 fac;
        aq_nrc_encode(vi, bb, x + (size_t)r * AQ_NRC_IN, 1, &fac);
        y[r] = make_float4(fac.x, fac.y, fac.z, 1.0f); /* fac for now; nrc_targets turns it into the target */
        cur.beta_id[i] = make_float4(1.0f, 1.0f, 1.0f, bi.w);
        L[slot] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
}

/* after the wave: y[r] = (L / fac, 1) for the records whose path reached the record vertex */
__global__ void __launch_bounds__(AQ_GEN_THREADS)
aq_k_nrc_targets(uint32_t n_paths, uint32_t rec_first, const float4* __restrict__ L, float4* __restrict__ y) {
    for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < n_paths; slot += gridDim.x * blockDim.x) {
        const uint32_t r = rec_first + 2u * slot;
        float4 f = y[r];
        if (f.w == 0.0f) continue;
        const float4 l = L[slot];
        y[r] = make_float4(l.x / f.x, l.y / f.y, l.z / f.z, 1.0f);
    }
}

/* ------------------------------------------------------------------ training */
__global__ void aq_k_nrc_init(uint32_t seed, float* __restrict__ w, float* __restrict__ m, float* __restrict__ v) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= AQ_NRC_N_WEIGHTS) return;
    w[k] = aq_nrc_init_weight(seed, k);
    m[k] = 0.0f;
    v[k] = 0.0f;
}

/* one CTA = one chunk of AQ_NRC_CHUNK records of iteration `it`: forward, loss, backward;
 * writes the chunk's weight gradient g[chunk][AQ_NRC_N_WEIGHTS] and its loss partial */
/* CG: the weights are rewritten by other CTAs of the SAME launch between iterations (persistent
 * training kernel): read them past the non-coherent L1 */
template <bool CG>
__device__ __forceinline__ void aq_nrc_train_chunk_dev(const float* W, const float* __restrict__ x,
                                                       const float4* __restrict__ y, uint32_t it, uint32_t batch,
                                                       float inv_norm, float* __restrict__ g,
                                                       float* __restrict__ loss_chunk, uint32_t chunk) {
    extern __shared__ float sm[];
    float* acts = sm;                                              /* [5][64][LD] */
    float* delta = acts + AQ_NRC_N_MATS * AQ_NRC_WIDTH * AQ_NRC_LD; /* [2][64][LD] */
    float* Wl = delta + 2 * AQ_NRC_WIDTH * AQ_NRC_LD;               /* [64][cols] */
    float* tgt = Wl + AQ_NRC_WIDTH * AQ_NRC_WIDTH;                  /* [4][64]: target rgb, live */
    float* lterm = tgt + 4 * AQ_NRC_CHUNK;                          /* [64][3] */
    const uint32_t tid = threadIdx.x;
    auto A = [&](int l, int i) { return acts + ((size_t)l * AQ_NRC_WIDTH + i) * AQ_NRC_LD; };
    auto Dl = [&](int b, int j) { return delta + ((size_t)b * AQ_NRC_WIDTH + j) * AQ_NRC_LD; };
    auto load_matrix = [&](int l) {
        const int n = AQ_NRC_WIDTH * AQ_NRC_MAT_COLS(l);
        for (int k = tid; k < n; k += AQ_NRC_TRAIN_THREADS) Wl[k] = CG ? __ldcg(W + AQ_NRC_MAT_OFF(l) + k) : W[AQ_NRC_MAT_OFF(l) + k];
    };

    /* ---- inputs and targets of the chunk (records that do not exist or are not valid: zeros) */
    if (tid < AQ_NRC_CHUNK) {
        const uint32_t bi = chunk * AQ_NRC_CHUNK + tid;
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bi < batch) t = y[(size_t)it * batch + bi];
        const bool live = t.w != 0.0f;
        tgt[0 * AQ_NRC_CHUNK + tid] = t.x;
        tgt[1 * AQ_NRC_CHUNK + tid] = t.y;
        tgt[2 * AQ_NRC_CHUNK + tid] = t.z;
        tgt[3 * AQ_NRC_CHUNK + tid] = live ? 1.0f : 0.0f;
    }
    __syncthreads();
    for (int k = tid; k < AQ_NRC_CHUNK * AQ_NRC_IN; k += AQ_NRC_TRAIN_THREADS) {
        const int s = k / AQ_NRC_IN, i = k % AQ_NRC_IN;
        const uint32_t bi = chunk * AQ_NRC_CHUNK + s;
        float v = 0.0f;
        if (tgt[3 * AQ_NRC_CHUNK + s] != 0.0f) v = x[((size_t)it * batch + bi) * AQ_NRC_IN + i];
        A(0, i)[s] = v;
    }
    /* ---- forward: thread (s, group) computes AQ_NRC_TRAIN_PER neurons of sample s per layer.  The
     * neurons of a thread share the load of a_l[k][s]; each accumulator is still the ascending fmaf
     * chain of aq_nrc_dot, so the values are the ones aq_nrc.h defines. */
    static_assert(AQ_NRC_TRAIN_PER == 4 || AQ_NRC_TRAIN_PER == 2 || AQ_NRC_TRAIN_PER == 1, "neurons per thread: 1, 2 or 4");
    static_assert(AQ_NRC_TRAIN_GROUPS >= AQ_NRC_OUT, "the output layer needs one group per output neuron");
    const int s = tid & (AQ_NRC_CHUNK - 1), q4 = tid / AQ_NRC_CHUNK;
    for (int l = 0; l < AQ_NRC_HIDDEN_LAYERS; ++l) {
        __syncthreads();
        load_matrix(l);
        __syncthreads();
        const float* al = A(l, 0) + s;
        const int j0 = q4 * AQ_NRC_TRAIN_PER;
        float c[AQ_NRC_TRAIN_PER];
#pragma unroll
        for (int q = 0; q < AQ_NRC_TRAIN_PER; ++q) c[q] = 0.0f;
#pragma unroll 8
        for (int k = 0; k < AQ_NRC_WIDTH; ++k) {
            const float a = al[k * AQ_NRC_LD];
            const float* wr = Wl + k * AQ_NRC_WIDTH + j0;
#pragma unroll
            for (int q = 0; q < AQ_NRC_TRAIN_PER; ++q) c[q] = fmaf(a, wr[q], c[q]);
        }
#pragma unroll
        for (int q = 0; q < AQ_NRC_TRAIN_PER; ++q) A(l + 1, j0 + q)[s] = aq_nrc_relu(c[q]);
    }
    __syncthreads();
    load_matrix(AQ_NRC_HIDDEN_LAYERS);
    __syncthreads();
    int cur = 0;
    if (q4 < AQ_NRC_OUT) { /* output neuron q4 of sample s, loss gradient and loss term */
        const float yv = aq_nrc_dot(A(AQ_NRC_HIDDEN_LAYERS, 0) + s, AQ_NRC_LD, Wl + q4, AQ_NRC_OUT_PAD, AQ_NRC_WIDTH);
        const bool live = tgt[3 * AQ_NRC_CHUNK + s] != 0.0f;
        const float t = tgt[q4 * AQ_NRC_CHUNK + s];
        Dl(cur, q4)[s] = live ? aq_nrc_loss_grad(yv, t, inv_norm) : 0.0f;
        lterm[s * 3 + q4] = live ? aq_nrc_loss_term(yv, t, inv_norm) : 0.0f;
    }
    __syncthreads();
    if (tid == 0) { /* the chunk's loss: samples in order, channels in order */
        float part = 0.0f;
        for (int k = 0; k < AQ_NRC_CHUNK * 3; ++k)
            if (tgt[3 * AQ_NRC_CHUNK + k / 3] != 0.0f) part += lterm[k];
        loss_chunk[chunk] = part;
    }
    /* ---- backward: matrix l = 4 (output) down to 0; Wl holds matrix l at the top of each round */
    float* G = g + (size_t)chunk * AQ_NRC_N_WEIGHTS;
    for (int l = AQ_NRC_HIDDEN_LAYERS; l >= 0; --l) {
        const int cols = AQ_NRC_MAT_COLS(l), n = l == AQ_NRC_HIDDEN_LAYERS ? AQ_NRC_OUT : AQ_NRC_WIDTH;
        /* weight gradient G_l[i][j] = sum_s a_l[i][s] * delta_l[j][s] */
        for (int p = tid; p < AQ_NRC_WIDTH * cols; p += AQ_NRC_TRAIN_THREADS) {
            const int i = p / cols, j = p % cols;
            G[AQ_NRC_MAT_OFF(l) + p] = j < n ? aq_nrc_dot(A(l, i), 1, Dl(cur, j), 1, AQ_NRC_CHUNK) : 0.0f;
        }
        if (l == 0) break;
        /* delta_{l-1}[i][s] = relu'(a_l[i][s]) * sum_j W_l[i][j] * delta_l[j][s]; the 4 inputs i of a
         * thread share the load of delta_l[j][s] */
        {
            const int i0 = q4 * AQ_NRC_TRAIN_PER;
            const float* dj = Dl(cur, 0) + s;
            float c[AQ_NRC_TRAIN_PER];
#pragma unroll
            for (int q = 0; q < AQ_NRC_TRAIN_PER; ++q) c[q] = 0.0f;
            for (int j = 0; j < n; ++j) {
                const float d = dj[j * AQ_NRC_LD];
#pragma unroll
                for (int q = 0; q < AQ_NRC_TRAIN_PER; ++q) c[q] = fmaf(Wl[(size_t)(i0 + q) * cols + j], d, c[q]);
            }
#pragma unroll
            for (int q = 0; q < AQ_NRC_TRAIN_PER; ++q) Dl(cur ^ 1, i0 + q)[s] = A(l, i0 + q)[s] > 0.0f ? c[q] : 0.0f;
        }
        __syncthreads();
        cur ^= 1;
        load_matrix(l - 1);
        __syncthreads();
    }
}

__global__ void __launch_bounds__(AQ_NRC_TRAIN_THREADS)
aq_k_nrc_train_chunk(const float* __restrict__ W, const float* __restrict__ x, const float4* __restrict__ y,
                     uint32_t it, uint32_t batch, float inv_norm, float* __restrict__ g,
                     float* __restrict__ loss_chunk) {
    aq_nrc_train_chunk_dev<false>(W, x, y, it, batch, inv_norm, g, loss_chunk, blockIdx.x);
}

/* grid-wide barrier of the persistent training kernel (cooperative launch: all CTAs are resident);
 * `bar` counts arrivals monotonically, a lost arrival traps instead of hanging the GPU */
__device__ __forceinline__ void aq_nrc_grid_barrier(unsigned int* bar, unsigned int target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        unsigned int spins = 0;
        while (*reinterpret_cast<volatile unsigned int*>(bar) < target)
            if (++spins > (1u << 28)) __trap();
        __threadfence();
    }
    __syncthreads();
}

/* the whole descent in ONE launch (round 2): CTA c = chunk c of every iteration; per iteration
 * forward/loss/backward of the chunk -> barrier -> each CTA sums the chunk gradients (ascending chunk
 * order) and takes the Adam step for its share of the weights -> barrier.  Same arithmetic per record
 * and per weight as aq_k_nrc_train_chunk + aq_k_nrc_adam, so weights and loss curve are unchanged; what
 * goes away are 2 launches per iteration (4096 at the reference's settings). */
__global__ void __launch_bounds__(AQ_NRC_TRAIN_THREADS)
aq_k_nrc_train_persistent(float* W, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ x,
                          const float4* __restrict__ y, uint32_t iters, uint32_t batch, float inv_norm, float lr,
                          const float2* __restrict__ bias, float* __restrict__ g, float* __restrict__ loss_chunk,
                          float* __restrict__ loss, unsigned int* __restrict__ bar) {
    const uint32_t chunk = blockIdx.x, n_chunks = gridDim.x;
    unsigned int target = 0;
    for (uint32_t it = 0; it < iters; ++it) {
        aq_nrc_train_chunk_dev<true>(W, x, y, it, batch, inv_norm, g, loss_chunk, chunk);
        aq_nrc_grid_barrier(bar, target += n_chunks);
        const float2 bc = bias[it];
        for (uint32_t k = chunk * AQ_NRC_TRAIN_THREADS + threadIdx.x; k < AQ_NRC_N_WEIGHTS; k += n_chunks * AQ_NRC_TRAIN_THREADS) {
            float gs = 0.0f;
            for (uint32_t c = 0; c < n_chunks; ++c) gs = gs + __ldcg(g + (size_t)c * AQ_NRC_N_WEIGHTS + k);
            float wk = __ldcg(W + k), mk = m[k], vk = v[k]; /* m, v: only this thread ever touches entry k */
            aq_nrc_adam(gs, lr, bc.x, bc.y, &wk, &mk, &vk);
            W[k] = wk;
            m[k] = mk;
            v[k] = vk;
        }
        if (chunk == 0 && threadIdx.x == 0) {
            float l = 0.0f;
            for (uint32_t c = 0; c < n_chunks; ++c) l += __ldcg(loss_chunk + c);
            loss[it] = l;
        }
        aq_nrc_grid_barrier(bar, target += n_chunks);
    }
}

/* sum the chunk gradients in ascending chunk order, Adam step; thread 0 also folds the loss */
__global__ void aq_k_nrc_adam(float* __restrict__ W, float* __restrict__ m, float* __restrict__ v,
                              const float* __restrict__ g, uint32_t n_chunks, float lr, float bc1, float bc2,
                              const float* __restrict__ loss_chunk, float* __restrict__ loss_it) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0) {
        float l = 0.0f;
        for (uint32_t c = 0; c < n_chunks; ++c) l += loss_chunk[c];
        *loss_it = l;
    }
    if (k >= AQ_NRC_N_WEIGHTS) return;
    float gs = 0.0f;
    for (uint32_t c = 0; c < n_chunks; ++c) gs = gs + g[(size_t)c * AQ_NRC_N_WEIGHTS + k];
    aq_nrc_adam(gs, lr, bc1, bc2, &W[k], &m[k], &v[k]);
}

/* ------------------------------------------------------------------ render: query the cache
 * One thread per entry of the ray queue at the query depth (after its closest-hit pass):
 * L[slot] += emission(vertex) + beta * fac * max(MLP(encode(vertex)), 0).  Activations live in
 * shared memory as [feature][thread]; the weights of one layer at a time next to them. */
template <bool AREA, bool FULL>
__global__ void __launch_bounds__(AQ_NRC_QUERY_THREADS)
aq_k_nrc_query(aq_scene_view sv, aq_nrc_bounds bb, aq_wave_params wp, int depth, aq_queue cur,
               const uint4* __restrict__ hits, const float* __restrict__ W, float4* __restrict__ L,
               const uint32_t* __restrict__ ctrl, const uint32_t* __restrict__ qcnt,
               unsigned long long* __restrict__ stats) {
    extern __shared__ float sm[];
    float* a0 = sm;                                             /* [64][T] */
    float* a1 = a0 + AQ_NRC_WIDTH * AQ_NRC_QUERY_THREADS;        /* [64][T] */
    float* Wl = a1 + AQ_NRC_WIDTH * AQ_NRC_QUERY_THREADS;        /* [64][cols] */
    const uint32_t tid = threadIdx.x;
    const uint32_t n_blocks = ctrl[aqc_blocks_ray(depth)];
    const uint32_t n = n_blocks * AQ_QBLK; /* the queue's capacity walk: entries past a block's count are skipped */
    uint32_t my_hits = 0;
    for (uint32_t base = blockIdx.x * AQ_NRC_QUERY_THREADS; base < n; base += gridDim.x * AQ_NRC_QUERY_THREADS) {
        const uint32_t i = base + tid;
        bool live = false;
        uint32_t slot = 0;
        aq_v3 beta = aq_mk(0.f, 0.f, 0.f), fac = beta, em = beta;
        if (aq_queue_live(qcnt, n_blocks, i)) {
            const uint4 h = AQ_QLD(&hits[i]);
            if (h.x != AQ_MISS_ID) {
                const float4 rdv = AQ_QLD(&cur.d_tmax[i]), bi = AQ_QLD(&cur.beta_id[i]);
                slot = __float_as_uint(bi.w);
                beta = aq_mk(bi.x, bi.y, bi.z);
                aq_vertex_in vi;
                aq_fetch_vertex<FULL>(sv, h.x, __uint_as_float(h.z), __uint_as_float(h.w),
                                      aq_mk(rdv.x, rdv.y, rdv.z), &vi);
                vi.t_hit = __uint_as_float(h.y);
                vi.prev_pdf = AREA ? cur.o_tmin[i].w : 0.0f;
                em = aq_vertex_emitted<AREA>(vi, beta, wp.mis_mode);
                aq_nrc_encode(vi, bb, a0 + tid, AQ_NRC_QUERY_THREADS, &fac);
                live = true;
                ++my_hits;
            }
        }
        if (!live) /* idle lanes still run the products: keep their inputs finite */
            for (int k = 0; k < AQ_NRC_IN; ++k) a0[k * AQ_NRC_QUERY_THREADS + tid] = 0.0f;
        float* in = a0;
        float* out = a1;
        for (int l = 0; l < AQ_NRC_HIDDEN_LAYERS; ++l) {
            __syncthreads(); /* previous users of Wl are done */
            for (int k = tid; k < AQ_NRC_WIDTH * AQ_NRC_WIDTH; k += AQ_NRC_QUERY_THREADS) Wl[k] = W[AQ_NRC_MAT_OFF(l) + k];
            __syncthreads();
            /* 16 neurons per pass share the load of in[k][tid]; every accumulator is the ascending
             * fmaf chain of aq_nrc_dot */
            for (int j0 = 0; j0 < AQ_NRC_WIDTH; j0 += 16) {
                float acc[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) acc[q] = 0.0f;
#pragma unroll 4
                for (int k = 0; k < AQ_NRC_WIDTH; ++k) {
                    const float a = in[k * AQ_NRC_QUERY_THREADS + tid];
                    const float4* wr = reinterpret_cast<const float4*>(Wl + k * AQ_NRC_WIDTH + j0);
                    const float4 w0 = wr[0], w1 = wr[1], w2 = wr[2], w3 = wr[3];
                    acc[0] = fmaf(a, w0.x, acc[0]);   acc[1] = fmaf(a, w0.y, acc[1]);
                    acc[2] = fmaf(a, w0.z, acc[2]);   acc[3] = fmaf(a, w0.w, acc[3]);
                    acc[4] = fmaf(a, w1.x, acc[4]);   acc[5] = fmaf(a, w1.y, acc[5]);
                    acc[6] = fmaf(a, w1.z, acc[6]);   acc[7] = fmaf(a, w1.w, acc[7]);
                    acc[8] = fmaf(a, w2.x, acc[8]);   acc[9] = fmaf(a, w2.y, acc[9]);
                    acc[10] = fmaf(a, w2.z, acc[10]); acc[11] = fmaf(a, w2.w, acc[11]);
                    acc[12] = fmaf(a, w3.x, acc[12]); acc[13] = fmaf(a, w3.y, acc[13]);
                    acc[14] = fmaf(a, w3.z, acc[14]); acc[15] = fmaf(a, w3.w, acc[15]);
                }
#pragma unroll
                for (int q = 0; q < 16; ++q) out[(j0 + q) * AQ_NRC_QUERY_THREADS + tid] = aq_nrc_relu(acc[q]);
            }
            float* t = in;
            in = out;
            out = t;
        }
        __syncthreads();
        for (int k = tid; k < AQ_NRC_WIDTH * AQ_NRC_OUT_PAD; k += AQ_NRC_QUERY_THREADS)
            Wl[k] = W[AQ_NRC_MAT_OFF(AQ_NRC_HIDDEN_LAYERS) + k];
        __syncthreads();
        if (live) {
            aq_v3 yr = aq_mk(aq_nrc_relu(aq_nrc_dot(in + tid, AQ_NRC_QUERY_THREADS, Wl + 0, AQ_NRC_OUT_PAD, AQ_NRC_WIDTH)),
                             aq_nrc_relu(aq_nrc_dot(in + tid, AQ_NRC_QUERY_THREADS, Wl + 1, AQ_NRC_OUT_PAD, AQ_NRC_WIDTH)),
                             aq_nrc_relu(aq_nrc_dot(in + tid, AQ_NRC_QUERY_THREADS, Wl + 2, AQ_NRC_OUT_PAD, AQ_NRC_WIDTH)));
            const float4 l0 = L[slot];
            aq_v3 Ls = aq_add(aq_add(aq_mk(l0.x, l0.y, l0.z), em), aq_mul(beta, aq_mul(fac, yr)));
            L[slot] = make_float4(Ls.x, Ls.y, Ls.z, l0.w);
        }
    }
    uint32_t wsum = my_hits;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wsum += __shfl_xor_sync(0xFFFFFFFFu, wsum, o);
    if ((tid & 31u) == 0u && wsum) atomicAdd(&stats[AQS_BOUNCES], (unsigned long long)wsum);
}

/* ------------------------------------------------------------------ render: query on the tensor cores
 * OPT-IN (AQ_RENDER_NRC_TENSOR): the same lookup as aq_k_nrc_query with the MLP on the
 * 5th-generation tensor cores.  One CTA of 128 threads owns a tile of 128 queue entries; thread t
 * encodes entry t, rounds the 64 features to bf16 and writes row t of the A operand (K-major
 * core-matrix layout, no swizzle) into shared memory; per layer ONE thread issues four
 * tcgen05.mma.cta_group::1.kind::f16 (M 128, N 64 / 16, K 16; accumulator: 64 TMEM columns) and
 * commits them to an mbarrier; the four warps read their 32 TMEM lanes back (tcgen05.ld
 * 32x32b), apply the ReLU, round to bf16 and write the next layer's A row.  The five weight
 * matrices are resident in shared memory in operand layout (aq_k_nrc_pack_weights).
 * Operands are bf16 and the accumulation order inside the MMA is unspecified: this path is
 * compared with the exact one under a tolerance, it is not bit-identical to the oracle.
 * The MMA / TMEM / descriptor part is the kernel validated stand-alone in
 * tools/experimental/nrc_tcgen05_query.cu (B200: correct to bf16 rounding, 65 TFLOP/s). */
#define AQ_NRC_TC_ROWS 128
#define AQ_NRC_TC_NOUT 16 /* N of the output MMA (3 columns used) */
#define AQ_NRC_TC_WT_BYTES (AQ_NRC_HIDDEN_LAYERS * AQ_NRC_WIDTH * AQ_NRC_WIDTH * 2 + AQ_NRC_TC_NOUT * AQ_NRC_WIDTH * 2)
#define AQ_NRC_TC_SMEM_BYTES (AQ_NRC_TC_ROWS * AQ_NRC_WIDTH * 2 + AQ_NRC_TC_WT_BYTES)

/* byte offset of element (r, k) of a bf16 operand tile with R rows and 64 columns: core matrices
 * of 8 rows x 16 B, the 8-row groups contiguous (stride-dimension offset 128 B), the core
 * matrices along K (R/8)*128 B apart (leading-dimension offset) */
__host__ __device__ __forceinline__ uint32_t aq_nrc_tc_off(uint32_t r, uint32_t k, uint32_t R) {
    return (r >> 3) * 128u + (k >> 3) * (R >> 3) * 128u + (r & 7u) * 16u + (k & 7u) * 2u;
}
__device__ __forceinline__ uint32_t aq_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
/* shared-memory matrix descriptor: SWIZZLE_NONE, K-major, descriptor version 1 (sm_100) */
__device__ __forceinline__ uint64_t aq_umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46);
}
/* instruction descriptor: D fp32, A and B bf16, both K-major, dense */
__host__ __device__ constexpr uint32_t aq_umma_idesc(uint32_t M, uint32_t N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ void aq_umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
/* bounded wait: a lost arrive traps instead of hanging the GPU */
__device__ __forceinline__ void aq_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(aq_smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) return;
    }
    asm volatile("trap;\n");
}

/* fp32 weights (aq_nrc.h layout, input-major) -> bf16 operand tiles: hidden layer l at byte
 * l*8192 as [N = 64 rows][K = 64] with B(n, k) = W_l[k][n]; output layer at 4*8192 as [16][64],
 * rows 3..15 zero */
__global__ void aq_k_nrc_pack_weights(const float* __restrict__ W, uint8_t* __restrict__ wt) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t hid = AQ_NRC_HIDDEN_LAYERS * AQ_NRC_WIDTH * AQ_NRC_WIDTH;
    if (t < hid) {
        const uint32_t l = t / (AQ_NRC_WIDTH * AQ_NRC_WIDTH), k = (t / AQ_NRC_WIDTH) % AQ_NRC_WIDTH, j = t % AQ_NRC_WIDTH;
        *reinterpret_cast<__nv_bfloat16*>(wt + (size_t)l * AQ_NRC_WIDTH * AQ_NRC_WIDTH * 2 + aq_nrc_tc_off(j, k, AQ_NRC_WIDTH)) =
            __float2bfloat16_rn(W[t]);
    } else if (t < hid + AQ_NRC_TC_NOUT * AQ_NRC_WIDTH) {
        const uint32_t u = t - hid, c = u / AQ_NRC_WIDTH, k = u % AQ_NRC_WIDTH;
        const float v = c < AQ_NRC_OUT ? W[hid + k * AQ_NRC_OUT_PAD + c] : 0.0f;
        *reinterpret_cast<__nv_bfloat16*>(wt + (size_t)hid * 2 + aq_nrc_tc_off(c, k, AQ_NRC_TC_NOUT)) = __float2bfloat16_rn(v);
    }
}

template <bool AREA, bool FULL>
__global__ void __launch_bounds__(AQ_NRC_TC_ROWS)
aq_k_nrc_query_tc(aq_scene_view sv, aq_nrc_bounds bb, aq_wave_params wp, int depth, aq_queue cur,
                  const uint4* __restrict__ hits, const uint8_t* __restrict__ wt, float4* __restrict__ L,
                  const uint32_t* __restrict__ ctrl, const uint32_t* __restrict__ qcnt,
                  unsigned long long* __restrict__ stats) {
    extern __shared__ __align__(1024) uint8_t smem_tc[];
    uint8_t* sA = smem_tc;                                       /* 128 x 64 bf16 */
    uint8_t* sW = smem_tc + AQ_NRC_TC_ROWS * AQ_NRC_WIDTH * 2;   /* the five weight tiles */
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    const uint32_t n_blocks = ctrl[aqc_blocks_ray(depth)];
    const uint32_t n = n_blocks * AQ_QBLK; /* capacity walk of the block-structured queue */

    for (uint32_t k = tid; k < AQ_NRC_TC_WT_BYTES / 16; k += AQ_NRC_TC_ROWS)
        reinterpret_cast<uint4*>(sW)[k] = reinterpret_cast<const uint4*>(wt)[k];
    if (tid == 0)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(aq_smem_u32(&bar)), "r"(1u) : "memory");
    if (warp == 0) { /* one warp allocates 64 TMEM columns and gives the permit back */
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;\n" ::"r"(aq_smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    uint32_t phase = 0, my_hits = 0;

    for (uint32_t base = blockIdx.x * AQ_NRC_TC_ROWS; base < n; base += gridDim.x * AQ_NRC_TC_ROWS) {
        /* ---- this thread's queue entry: vertex, emission, features -> bf16 row tid of the A operand */
        const uint32_t i = base + tid;
        bool live = false;
        uint32_t slot = 0;
        aq_v3 bf = aq_mk(0.f, 0.f, 0.f), em = bf; /* bf = beta * fac */
        float xr[AQ_NRC_IN];
#pragma unroll
        for (int k = 0; k < AQ_NRC_IN; ++k) xr[k] = 0.0f;
        if (aq_queue_live(qcnt, n_blocks, i)) {
            const uint4 h = AQ_QLD(&hits[i]);
            if (h.x != AQ_MISS_ID) {
                const float4 rdv = AQ_QLD(&cur.d_tmax[i]), bi = AQ_QLD(&cur.beta_id[i]);
                slot = __float_as_uint(bi.w);
                const aq_v3 beta = aq_mk(bi.x, bi.y, bi.z);
                aq_vertex_in vi;
                aq_fetch_vertex<FULL>(sv, h.x, __uint_as_float(h.z), __uint_as_float(h.w),
                                      aq_mk(rdv.x, rdv.y, rdv.z), &vi);
                vi.t_hit = __uint_as_float(h.y);
                vi.prev_pdf = AREA ? cur.o_tmin[i].w : 0.0f;
                em = aq_vertex_emitted<AREA>(vi, beta, wp.mis_mode);
                aq_v3 fac;
                aq_nrc_encode(vi, bb, xr, 1, &fac);
                bf = aq_mul(beta, fac);
                live = true;
                ++my_hits;
            }
        }
#pragma unroll
        for (uint32_t kc = 0; kc < AQ_NRC_WIDTH / 8; ++kc) {
            __nv_bfloat162 v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = __floats2bfloat162_rn(xr[kc * 8 + 2 * e], xr[kc * 8 + 2 * e + 1]);
            *reinterpret_cast<uint4*>(sA + aq_nrc_tc_off(tid, kc * 8, AQ_NRC_TC_ROWS)) = *reinterpret_cast<uint4*>(v);
        }
        for (uint32_t l = 0; l <= AQ_NRC_HIDDEN_LAYERS; ++l) {
            /* generic-proxy writes of sA -> visible to the tensor core (async proxy); all rows written */
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            __syncthreads();
            const uint32_t N = l < AQ_NRC_HIDDEN_LAYERS ? AQ_NRC_WIDTH : AQ_NRC_TC_NOUT;
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const uint32_t a0 = aq_smem_u32(sA), b0 = aq_smem_u32(sW + l * AQ_NRC_WIDTH * AQ_NRC_WIDTH * 2);
                const uint32_t lbo_a = (AQ_NRC_TC_ROWS / 8) * 128, lbo_b = (N / 8) * 128;
                const uint32_t idesc = aq_umma_idesc(AQ_NRC_TC_ROWS, N);
                for (uint32_t s = 0; s < AQ_NRC_WIDTH / 16; ++s)
                    aq_umma_bf16(tmem, aq_umma_desc(a0 + s * 2 * lbo_a, lbo_a, 128), aq_umma_desc(b0 + s * 2 * lbo_b, lbo_b, 128),
                                 idesc, s > 0);
                /* arrives on the mbarrier when the MMAs above have completed */
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(aq_smem_u32(&bar)) : "memory");
            }
            aq_mbar_wait(&bar, phase);
            phase ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const uint32_t taddr = tmem + ((warp * 32u) << 16); /* accumulator row = TMEM lane 32*warp + lane */
            if (l < AQ_NRC_HIDDEN_LAYERS) {
                uint32_t r[64];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
                    "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
                    "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];\n"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                      "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                      "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]),
                      "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
                      "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]),
                      "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]),
                      "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
                    : "r"(taddr)
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
                for (uint32_t kc = 0; kc < AQ_NRC_WIDTH / 8; ++kc) { /* ReLU, round to bf16: the next layer's A row */
                    __nv_bfloat162 v[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        v[e] = __floats2bfloat162_rn(fmaxf(__uint_as_float(r[kc * 8 + 2 * e]), 0.f),
                                                     fmaxf(__uint_as_float(r[kc * 8 + 2 * e + 1]), 0.f));
                    *reinterpret_cast<uint4*>(sA + aq_nrc_tc_off(tid, kc * 8, AQ_NRC_TC_ROWS)) = *reinterpret_cast<uint4*>(v);
                }
            } else {
                uint32_t r[4];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                             : "r"(taddr)
                             : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
                if (live) {
                    const aq_v3 yr = aq_mk(aq_nrc_relu(__uint_as_float(r[0])), aq_nrc_relu(__uint_as_float(r[1])),
                                           aq_nrc_relu(__uint_as_float(r[2])));
                    const float4 l0 = L[slot];
                    const aq_v3 Ls = aq_add(aq_add(aq_mk(l0.x, l0.y, l0.z), em), aq_mul(bf, yr));
                    L[slot] = make_float4(Ls.x, Ls.y, Ls.z, l0.w);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;\n" ::"r"(tmem) : "memory");
    uint32_t wsum = my_hits;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wsum += __shfl_xor_sync(0xFFFFFFFFu, wsum, o);
    if ((tid & 31u) == 0u && wsum) atomicAdd(&stats[AQS_BOUNCES], (unsigned long long)wsum);
}

#endif /* AQ_NRC_CUH */
