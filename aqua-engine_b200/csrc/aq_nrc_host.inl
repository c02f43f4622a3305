/*
 * aq_nrc_host.inl — host side of the `nrc` integrator entry points (include/aqua_cuda.h,
 * aq_nrc_*); included by aq_cuda.cu, whose handles and helpers it uses.  Kernels: aq_nrc.cuh;
 * semantics: aq_nrc.h (scenes/integrator.json:1-8).
 */
namespace {

typedef wave_launcher nrc_waves; /* aq_cuda.cu */

__global__ void aq_k_nrc_count_valid(const float4* __restrict__ y, uint32_t n, unsigned int* __restrict__ out) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int v = (k < n && y[k].w != 0.0f) ? 1u : 0u;
    v = __reduce_add_sync(0xFFFFFFFFu, v);
    if ((threadIdx.x & 31u) == 0u && v) atomicAdd(out, v);
}

template <class T>
int nrc_realloc(aq_ctx* c, T** p, size_t count) {
    if (*p) cudaFreeAsync(*p, c->stream);
    *p = nullptr;
    AQ_CK(c, cudaMallocAsync((void**)p, (count ? count : 1) * sizeof(T), c->stream));
    return AQ_OK;
}

int nrc_check_cfg(aq_ctx* c, const aq_integrator_cfg* cfg, const aq_nrc_cfg* nrc, const char* who) {
    if (cfg->max_depth == 0 || cfg->max_depth > 64) return set_err(c, AQ_ERR_BAD_ARG, "%s: max_depth must be in 1..64", who);
    if (nrc->batch_size == 0 || nrc->training_iters == 0)
        return set_err(c, AQ_ERR_BAD_ARG, "%s: batch_size and training_iters must be positive", who);
    if ((uint64_t)nrc->batch_size * nrc->training_iters > (1ull << 30))
        return set_err(c, AQ_ERR_BAD_ARG, "%s: more than 2^30 training records", who);
    if (!(nrc->learning_rate > 0.0f)) return set_err(c, AQ_ERR_BAD_ARG, "%s: learning_rate must be positive", who);
    return AQ_OK;
}

}  // namespace

extern "C" {

int aq_nrc_train(aq_scene* s, const aq_integrator_cfg* cfg, const aq_nrc_cfg* nrc, aq_nrc_info* info) {
    if (!s || !cfg || !nrc) return set_err(s ? s->ctx : nullptr, AQ_ERR_BAD_ARG, "aq_nrc_train: null argument");
    aq_ctx* c = s->ctx;
    if (!s->built) return set_err(c, AQ_ERR_STATE, "aq_nrc_train: call aq_accel_build first");
    int rc = nrc_check_cfg(c, cfg, nrc, "aq_nrc_train");
    if (rc != AQ_OK) return rc;
    const uint32_t W = cfg->width ? cfg->width : s->camera.res[0];
    const uint32_t H = cfg->height ? cfg->height : s->camera.res[1];
    if (W == 0 || H == 0 || (uint64_t)W * H > 0x7FFFFFFFull)
        return set_err(c, AQ_ERR_BAD_ARG, "aq_nrc_train: bad resolution %ux%u", W, H);
    AQ_CK(c, cudaSetDevice(c->device));
    uint32_t pool = cfg->pool_paths ? cfg->pool_paths : AQ_DEFAULT_POOL;
    if (pool < 1024) pool = 1024;
    rc = ensure_pool(s, pool);
    if (rc != AQ_OK) return rc;
    cudaStream_t st = c->stream;
    const uint32_t B = nrc->batch_size, iters = nrc->training_iters;
    const uint64_t R = (uint64_t)B * iters;
    const uint32_t n_chunks = (B + AQ_NRC_CHUNK - 1) / AQ_NRC_CHUNK;

    s->nrc_trained = false;
    if (!s->d_nrc_w) {
        if ((rc = nrc_realloc(c, &s->d_nrc_w, AQ_NRC_N_WEIGHTS)) != AQ_OK) return rc;
        if ((rc = nrc_realloc(c, &s->d_nrc_m, AQ_NRC_N_WEIGHTS)) != AQ_OK) return rc;
        if ((rc = nrc_realloc(c, &s->d_nrc_v, AQ_NRC_N_WEIGHTS)) != AQ_OK) return rc;
    }
    if ((rc = nrc_realloc(c, &s->d_nrc_x, (size_t)R * AQ_NRC_IN)) != AQ_OK) return rc;
    if ((rc = nrc_realloc(c, &s->d_nrc_y, (size_t)R)) != AQ_OK) return rc;
    if ((rc = nrc_realloc(c, &s->d_nrc_g, (size_t)n_chunks * AQ_NRC_N_WEIGHTS)) != AQ_OK) return rc;
    if ((rc = nrc_realloc(c, &s->d_nrc_loss, (size_t)iters)) != AQ_OK) return rc;
    if ((rc = nrc_realloc(c, &s->d_nrc_loss_chunk, (size_t)n_chunks)) != AQ_OK) return rc;
    s->nrc_records = R;
    s->nrc_iters = iters;
    AQ_CK(c, cudaMemsetAsync(s->d_nrc_x, 0, (size_t)R * AQ_NRC_IN * sizeof(float), st));
    AQ_CK(c, cudaMemsetAsync(s->d_nrc_y, 0, (size_t)R * sizeof(float4), st));
    AQ_CK(c, cudaMemsetAsync(s->d_stats, 0, AQS_WORDS * sizeof(unsigned long long), st));
    aq_k_nrc_init<<<(AQ_NRC_N_WEIGHTS + 255) / 256, 256, 0, st>>>(cfg->seed, s->d_nrc_w, s->d_nrc_m, s->d_nrc_v);

    /* three timing events, destroyed on every way out of this function */
    struct events {
        cudaEvent_t e[3] = {nullptr, nullptr, nullptr};
        ~events() {
            for (cudaEvent_t x : e)
                if (x) cudaEventDestroy(x);
        }
    } ev;
    for (cudaEvent_t& x : ev.e) AQ_CK(c, cudaEventCreate(&x));
    cudaEvent_t e0 = ev.e[0], e1 = ev.e[1], e2 = ev.e[2];
    AQ_CK(c, cudaEventRecord(e0, st));

    /* ---- records: one wavefront pass per record depth (r even: first hit, r odd: second hit) */
    nrc_waves wv(s, cfg->flags);
    aq_wave_params wp;
    wp.cam = aq_cam_derive(s->camera.translate, s->camera.rotate, s->camera.fov, s->camera.lens_radius,
                           s->camera.focal, W, H);
    wp.seed = cfg->seed;
    wp.max_depth = cfg->max_depth;
    wp.spp_begin = 0;
    wp.s0 = 0;
    wp.ns = 1;
    wp.tile_base = 0;
    wp.npix = (uint64_t)W * H;
    wp.mis_mode = (cfg->flags & AQ_RENDER_MIS_NEE_ONLY)    ? AQ_MIS_NEE_ONLY
                  : (cfg->flags & AQ_RENDER_MIS_BSDF_ONLY) ? AQ_MIS_BSDF_ONLY
                                                           : AQ_MIS_BOTH;
    for (uint32_t D = 0; D < 2; ++D) {
        const uint64_t nD = R > D ? (R - D + 1) / 2 : 0;
        for (uint64_t k0 = 0; k0 < nD; k0 += pool) {
            const uint32_t n = (uint32_t)(nD - k0 < pool ? nD - k0 : pool);
            const uint32_t rec_first = (uint32_t)(2 * k0 + D);
            wp.n_paths = n;
            wp.tile_pixels = n;
            wp.skip_emit_depth = D;
            aq_k_nrc_raygen<<<wv.ggrid, AQ_GEN_THREADS, 0, st>>>(wp, rec_first, c->q[0], c->d_L, s->d_ctrl, c->qc.ray[0], s->d_stats);
            wv.closest(0);
            if (D == 1) {
                wv.shade(wp, 0); /* its NEE and emission are discarded: the estimate restarts at the record vertex */
                wv.closest(1);
            }
            if (wv.full)
                aq_k_nrc_record<true><<<wv.ggrid, AQ_SHADE_THREADS, 0, st>>>(wv.sv, s->nrc_bb, (int)D, rec_first, c->q[D & 1],
                                                                             c->d_hits, c->d_L, s->d_ctrl, c->qc.ray[D & 1], s->d_nrc_x, s->d_nrc_y);
            else
                aq_k_nrc_record<false><<<wv.ggrid, AQ_SHADE_THREADS, 0, st>>>(wv.sv, s->nrc_bb, (int)D, rec_first, c->q[D & 1],
                                                                              c->d_hits, c->d_L, s->d_ctrl, c->qc.ray[D & 1], s->d_nrc_x, s->d_nrc_y);
            for (uint32_t depth = D; depth < cfg->max_depth; ++depth) {
                wv.shade(wp, depth);
                wv.shadow(depth);
                if (depth + 1 < cfg->max_depth) wv.closest(depth + 1);
            }
            aq_k_nrc_targets<<<wv.ggrid, AQ_GEN_THREADS, 0, st>>>(n, rec_first, c->d_L, s->d_nrc_y);
        }
    }
    AQ_CK(c, cudaGetLastError());
    AQ_CK(c, cudaEventRecord(e1, st));

    /* ---- training: iteration i descends on records [i*B, (i+1)*B) */
    const size_t smem = AQ_NRC_TRAIN_SMEM_FLOATS * sizeof(float);
    AQ_CK(c, cudaFuncSetAttribute(aq_k_nrc_train_chunk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    AQ_CK(c, cudaFuncSetAttribute(aq_k_nrc_train_persistent, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const float inv_norm = 1.0f / (3.0f * (float)B);
    /* one cooperative launch for the whole descent when every chunk's CTA is resident at once (32 CTAs at
     * the reference's batch of 512); otherwise two launches per iteration (AQUA_NRC_TRAIN=launches forces that) */
    int per_sm = 0, coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, aq_k_nrc_train_persistent, AQ_NRC_TRAIN_THREADS, smem) != cudaSuccess)
        per_sm = 0;
    const char* tm = std::getenv("AQUA_NRC_TRAIN");
    bool persistent = coop && (uint64_t)per_sm * (uint64_t)c->sm_count >= n_chunks && !(tm && !std::strcmp(tm, "launches"));
    if (persistent) {
        std::vector<float2> bias(iters);
        for (uint32_t it = 0; it < iters; ++it) aq_nrc_adam_bias(it + 1, &bias[it].x, &bias[it].y);
        float2* d_bias = nullptr;
        unsigned int* d_bar = nullptr;
        AQ_CK(c, cudaMallocAsync((void**)&d_bias, iters * sizeof(float2), st));
        AQ_CK(c, cudaMallocAsync((void**)&d_bar, sizeof(unsigned int), st));
        AQ_CK(c, cudaMemcpyAsync(d_bias, bias.data(), iters * sizeof(float2), cudaMemcpyHostToDevice, st));
        AQ_CK(c, cudaMemsetAsync(d_bar, 0, sizeof(unsigned int), st));
        AQ_CK(c, cudaStreamSynchronize(st)); /* `bias` (pageable) has been staged */
        float *pW = s->d_nrc_w, *pm = s->d_nrc_m, *pv = s->d_nrc_v, *px = s->d_nrc_x, *pg = s->d_nrc_g, *plc = s->d_nrc_loss_chunk,
              *pl = s->d_nrc_loss;
        float4* py = s->d_nrc_y;
        uint32_t a_iters = iters, a_B = B;
        float a_inv = inv_norm, a_lr = nrc->learning_rate;
        void* args[] = {&pW, &pm, &pv, &px, &py, &a_iters, &a_B, &a_inv, &a_lr, &d_bias, &pg, &plc, &pl, &d_bar};
        cudaError_t le = cudaLaunchCooperativeKernel((const void*)aq_k_nrc_train_persistent, dim3(n_chunks), dim3(AQ_NRC_TRAIN_THREADS),
                                                     args, smem, st);
        cudaFreeAsync(d_bias, st);
        cudaFreeAsync(d_bar, st);
        if (le != cudaSuccess) {
            (void)cudaGetLastError();
            persistent = false; /* fall through to the per-iteration launches */
        }
        if (std::getenv("AQ_BUILD_VERBOSE"))
            std::fprintf(stderr, "[aq nrc] persistent training launch: %s (%d CTAs/SM possible, %u chunks)\n", cudaGetErrorString(le), per_sm, n_chunks);
    } else if (std::getenv("AQ_BUILD_VERBOSE")) {
        std::fprintf(stderr, "[aq nrc] per-iteration training launches (coop %d, %d CTAs/SM, %u chunks)\n", coop, per_sm, n_chunks);
    }
    if (!persistent)
        for (uint32_t it = 0; it < iters; ++it) {
            float bc1, bc2;
            aq_nrc_adam_bias(it + 1, &bc1, &bc2);
            aq_k_nrc_train_chunk<<<n_chunks, AQ_NRC_TRAIN_THREADS, smem, st>>>(s->d_nrc_w, s->d_nrc_x, s->d_nrc_y, it, B,
                                                                                inv_norm, s->d_nrc_g, s->d_nrc_loss_chunk);
            aq_k_nrc_adam<<<(AQ_NRC_N_WEIGHTS + 255) / 256, 256, 0, st>>>(s->d_nrc_w, s->d_nrc_m, s->d_nrc_v, s->d_nrc_g,
                                                                           n_chunks, nrc->learning_rate, bc1, bc2,
                                                                           s->d_nrc_loss_chunk, s->d_nrc_loss + it);
        }
    AQ_CK(c, cudaGetLastError());
    AQ_CK(c, cudaEventRecord(e2, st));
    unsigned int* d_cnt = reinterpret_cast<unsigned int*>(s->d_ctrl + AQC_SPARE); /* spare control word */
    AQ_CK(c, cudaMemsetAsync(d_cnt, 0, sizeof(unsigned int), st));
    aq_k_nrc_count_valid<<<(unsigned)((R + 255) / 256), 256, 0, st>>>(s->d_nrc_y, (uint32_t)R, d_cnt);
    AQ_CK(c, cudaStreamSynchronize(st));
    s->nrc_trained = true;
    if (info) {
        std::memset(info, 0, sizeof *info);
        info->n_weights = AQ_NRC_N_WEIGHTS;
        info->n_records = (uint32_t)R;
        AQ_CK(c, cudaMemcpy(&info->n_valid, d_cnt, sizeof(unsigned int), cudaMemcpyDeviceToHost));
        AQ_CK(c, cudaMemcpy(&info->loss_first, s->d_nrc_loss, sizeof(float), cudaMemcpyDeviceToHost));
        AQ_CK(c, cudaMemcpy(&info->loss_last, s->d_nrc_loss + (iters - 1), sizeof(float), cudaMemcpyDeviceToHost));
        cudaEventElapsedTime(&info->ms_records, e0, e1);
        cudaEventElapsedTime(&info->ms_train, e1, e2);
    }
    return AQ_OK;
}

int aq_nrc_render_device_async(aq_scene* s, const aq_integrator_cfg* cfg, const aq_nrc_cfg* nrc, void* d_film_ext) {
    if (!s || !cfg || !nrc)
        return set_err(s ? s->ctx : nullptr, AQ_ERR_BAD_ARG, "aq_nrc_render: null argument");
    aq_ctx* c = s->ctx;
    if (!s->built) return set_err(c, AQ_ERR_STATE, "aq_nrc_render: call aq_accel_build first");
    if (!s->nrc_trained) return set_err(c, AQ_ERR_STATE, "aq_nrc_render: no trained cache (call aq_nrc_train)");
    const uint32_t W = cfg->width ? cfg->width : s->camera.res[0];
    const uint32_t H = cfg->height ? cfg->height : s->camera.res[1];
    if (W == 0 || H == 0 || (uint64_t)W * H > 0x7FFFFFFFull)
        return set_err(c, AQ_ERR_BAD_ARG, "aq_nrc_render: bad resolution %ux%u", W, H);
    if (cfg->spp_end < cfg->spp_begin) return set_err(c, AQ_ERR_BAD_ARG, "aq_nrc_render: spp_end < spp_begin");
    if (cfg->max_depth == 0 || cfg->max_depth > 64) return set_err(c, AQ_ERR_BAD_ARG, "aq_nrc_render: max_depth must be in 1..64");
    AQ_CK(c, cudaSetDevice(c->device));
    const uint64_t npix = (uint64_t)W * H;
    uint32_t pool = cfg->pool_paths ? cfg->pool_paths : AQ_DEFAULT_POOL;
    if (pool < 1024) pool = 1024;
    int rc = ensure_pool(s, pool);
    if (rc != AQ_OK) return rc;
    float4* film = (float4*)d_film_ext;
    if (!film) {
        if (c->film_pixels < npix) {
            if (c->d_film) cudaFree(c->d_film);
            c->d_film = nullptr;
            c->film_pixels = 0;
            AQ_CK(c, cudaMalloc((void**)&c->d_film, npix * sizeof(float4)));
            c->film_pixels = npix;
        }
        film = c->d_film;
    }
    const uint32_t nspp = cfg->spp_end - cfg->spp_begin;
    float4* samples = nullptr;
    if (cfg->flags & AQ_RENDER_DUMP_SAMPLES) {
        size_t need = (size_t)nspp * npix;
        if (c->samples_count < need) {
            if (c->d_samples) cudaFree(c->d_samples);
            c->d_samples = nullptr;
            c->samples_count = 0;
            AQ_CK(c, cudaMalloc((void**)&c->d_samples, (need ? need : 1) * sizeof(float4)));
            c->samples_count = need;
        }
        samples = c->d_samples;
    }
    cudaStream_t st = c->stream;
    const size_t bytes = npix * sizeof(float4);
    AQ_CK(c, cudaEventRecord(s->ev0, st));
    if (!(cfg->flags & AQ_RENDER_ACCUMULATE)) AQ_CK(c, cudaMemsetAsync(film, 0, bytes, st));
    AQ_CK(c, cudaMemsetAsync(s->d_stats, 0, AQS_WORDS * sizeof(unsigned long long), st));

    nrc_waves wv(s, cfg->flags);
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
    /* the query kernel: as many CTAs as are co-resident with its dynamic shared memory */
    auto query_fn = wv.area ? (wv.full ? aq_k_nrc_query<true, true> : aq_k_nrc_query<true, false>)
                            : (wv.full ? aq_k_nrc_query<false, true> : aq_k_nrc_query<false, false>);
    const size_t qsmem = AQ_NRC_QUERY_SMEM_FLOATS * sizeof(float);
    AQ_CK(c, cudaFuncSetAttribute(query_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)qsmem));
    int q_per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q_per_sm, query_fn, AQ_NRC_QUERY_THREADS, qsmem) != cudaSuccess ||
        q_per_sm < 1)
        q_per_sm = 1;
    const int qgrid = q_per_sm * c->sm_count;
    /* opt-in tensor-core lookup: bf16 operand tiles of the current weights, its own resident grid */
    const bool tensor = (cfg->flags & AQ_RENDER_NRC_TENSOR) != 0;
    auto query_tc_fn = wv.area ? (wv.full ? aq_k_nrc_query_tc<true, true> : aq_k_nrc_query_tc<true, false>)
                               : (wv.full ? aq_k_nrc_query_tc<false, true> : aq_k_nrc_query_tc<false, false>);
    int tcgrid = c->sm_count;
    if (tensor) {
        if (!s->d_nrc_wt) AQ_CK(c, cudaMallocAsync((void**)&s->d_nrc_wt, AQ_NRC_TC_WT_BYTES, st));
        aq_k_nrc_pack_weights<<<(AQ_NRC_HIDDEN_LAYERS * AQ_NRC_WIDTH * AQ_NRC_WIDTH + AQ_NRC_TC_NOUT * AQ_NRC_WIDTH + 255) / 256, 256, 0, st>>>(
            s->d_nrc_w, s->d_nrc_wt);
        AQ_CK(c, cudaFuncSetAttribute(query_tc_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AQ_NRC_TC_SMEM_BYTES));
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, query_tc_fn, AQ_NRC_TC_ROWS, AQ_NRC_TC_SMEM_BYTES) != cudaSuccess ||
            per_sm < 1)
            per_sm = 1;
        tcgrid = per_sm * c->sm_count;
    }
    const uint32_t Dq = nrc->visualize_cache ? 0u : 1u; /* depth index of the vertex that is looked up */
    const uint32_t n_shaded = Dq < cfg->max_depth ? Dq : cfg->max_depth;
    const uint32_t tile_pixels = (uint32_t)(npix < pool ? npix : pool);
    uint32_t S = pool / tile_pixels;
    if (S < 1) S = 1;
    uint32_t waves = 0;
    for (uint64_t tb = 0; tb < npix; tb += tile_pixels) {
        uint32_t tp = (uint32_t)((npix - tb) < tile_pixels ? (npix - tb) : tile_pixels);
        for (uint32_t s0 = cfg->spp_begin; s0 < cfg->spp_end; s0 += S) {
            uint32_t ns = cfg->spp_end - s0 < S ? cfg->spp_end - s0 : S;
            wp.tile_base = (uint32_t)tb;
            wp.tile_pixels = tp;
            wp.s0 = s0;
            wp.ns = ns;
            wp.n_paths = tp * ns;
            wv.raygen(wp);
            for (uint32_t depth = 0; depth < n_shaded; ++depth) {
                wv.closest(depth);
                wv.shade(wp, depth);
                wv.shadow(depth);
            }
            if (Dq < cfg->max_depth) {
                wv.closest(Dq);
                if (tensor)
                    query_tc_fn<<<tcgrid, AQ_NRC_TC_ROWS, AQ_NRC_TC_SMEM_BYTES, st>>>(wv.sv, s->nrc_bb, wp, (int)Dq, c->q[Dq & 1],
                                                                                      c->d_hits, s->d_nrc_wt, c->d_L, s->d_ctrl,
                                                                                      c->qc.ray[Dq & 1], s->d_stats);
                else
                    query_fn<<<qgrid, AQ_NRC_QUERY_THREADS, qsmem, st>>>(wv.sv, s->nrc_bb, wp, (int)Dq, c->q[Dq & 1], c->d_hits,
                                                                         s->d_nrc_w, c->d_L, s->d_ctrl, c->qc.ray[Dq & 1], s->d_stats);
                ++wv.launches;
            }
            wv.film(wp, film, samples);
            ++waves;
        }
    }
    AQ_CK(c, cudaGetLastError());
    AQ_CK(c, cudaEventRecord(s->ev1, st));
    s->last_cfg = *cfg;
    s->last_cfg.width = W;
    s->last_cfg.height = H;
    s->last_launches = wv.launches;
    s->last_waves = waves;
    s->prof_n = 0;
    s->prof_waves = 0;
    s->render_pending = true;
    return AQ_OK;
}

int aq_nrc_render(aq_scene* s, const aq_integrator_cfg* cfg, const aq_nrc_cfg* nrc, float* film_out, aq_stats* stats) {
    if (!s || !cfg || !nrc || !film_out)
        return set_err(s ? s->ctx : nullptr, AQ_ERR_BAD_ARG, "aq_nrc_render: null argument");
    aq_ctx* c = s->ctx;
    const uint32_t W = cfg->width ? cfg->width : s->camera.res[0];
    const uint32_t H = cfg->height ? cfg->height : s->camera.res[1];
    const size_t bytes = (size_t)W * H * sizeof(float4);
    AQ_CK(c, cudaSetDevice(c->device));
    if (cfg->flags & AQ_RENDER_ACCUMULATE) { /* the host film is the accumulator: push it first */
        if (c->film_pixels < (size_t)W * H) {
            if (c->d_film) cudaFree(c->d_film);
            c->d_film = nullptr;
            c->film_pixels = 0;
            AQ_CK(c, cudaMalloc((void**)&c->d_film, bytes));
            c->film_pixels = (size_t)W * H;
        }
        AQ_CK(c, cudaMemcpyAsync(c->d_film, film_out, bytes, cudaMemcpyHostToDevice, c->stream));
    }
    int rc = aq_nrc_render_device_async(s, cfg, nrc, nullptr);
    if (rc != AQ_OK) return rc;
    AQ_CK(c, cudaMemcpyAsync(film_out, c->d_film, bytes, cudaMemcpyDeviceToHost, c->stream));
    return aq_render_finish(s, stats);
}

int aq_nrc_get_weights(aq_scene* s, float* out, size_t n) {
    if (!s || !out) return set_err(s ? s->ctx : nullptr, AQ_ERR_BAD_ARG, "aq_nrc_get_weights: null argument");
    aq_ctx* c = s->ctx;
    if (!s->d_nrc_w) return set_err(c, AQ_ERR_STATE, "aq_nrc_get_weights: no cache");
    if (n < AQ_NRC_N_WEIGHTS) return set_err(c, AQ_ERR_BAD_ARG, "aq_nrc_get_weights: buffer too small");
    AQ_CK(c, cudaSetDevice(c->device));
    AQ_CK(c, cudaStreamSynchronize(c->stream));
    AQ_CK(c, cudaMemcpy(out, s->d_nrc_w, AQ_NRC_N_WEIGHTS * sizeof(float), cudaMemcpyDeviceToHost));
    return AQ_OK;
}

int aq_nrc_set_weights(aq_scene* s, const float* w, size_t n) {
    if (!s || !w) return set_err(s ? s->ctx : nullptr, AQ_ERR_BAD_ARG, "aq_nrc_set_weights: null argument");
    aq_ctx* c = s->ctx;
    if (n != AQ_NRC_N_WEIGHTS) return set_err(c, AQ_ERR_BAD_ARG, "aq_nrc_set_weights: expected %d weights", AQ_NRC_N_WEIGHTS);
    AQ_CK(c, cudaSetDevice(c->device));
    int rc;
    if (!s->d_nrc_w) {
        if ((rc = nrc_realloc(c, &s->d_nrc_w, AQ_NRC_N_WEIGHTS)) != AQ_OK) return rc;
        if ((rc = nrc_realloc(c, &s->d_nrc_m, AQ_NRC_N_WEIGHTS)) != AQ_OK) return rc;
        if ((rc = nrc_realloc(c, &s->d_nrc_v, AQ_NRC_N_WEIGHTS)) != AQ_OK) return rc;
    }
    AQ_CK(c, cudaStreamSynchronize(c->stream));
    AQ_CK(c, cudaMemcpy(s->d_nrc_w, w, AQ_NRC_N_WEIGHTS * sizeof(float), cudaMemcpyHostToDevice));
    s->nrc_trained = true;
    return AQ_OK;
}

int aq_nrc_get_loss(aq_scene* s, float* out, size_t n_iters) {
    if (!s || !out) return set_err(s ? s->ctx : nullptr, AQ_ERR_BAD_ARG, "aq_nrc_get_loss: null argument");
    aq_ctx* c = s->ctx;
    if (!s->d_nrc_loss || n_iters < s->nrc_iters) return set_err(c, AQ_ERR_STATE, "aq_nrc_get_loss: no training run / buffer too small");
    AQ_CK(c, cudaSetDevice(c->device));
    AQ_CK(c, cudaStreamSynchronize(c->stream));
    AQ_CK(c, cudaMemcpy(out, s->d_nrc_loss, s->nrc_iters * sizeof(float), cudaMemcpyDeviceToHost));
    return AQ_OK;
}

int aq_nrc_get_records(aq_scene* s, float* x_out, float* y_out, size_t n_records) {
    if (!s) return set_err(nullptr, AQ_ERR_BAD_ARG, "aq_nrc_get_records: null argument");
    aq_ctx* c = s->ctx;
    if (!s->d_nrc_x || n_records < s->nrc_records) return set_err(c, AQ_ERR_STATE, "aq_nrc_get_records: no training run / buffer too small");
    AQ_CK(c, cudaSetDevice(c->device));
    AQ_CK(c, cudaStreamSynchronize(c->stream));
    if (x_out) AQ_CK(c, cudaMemcpy(x_out, s->d_nrc_x, s->nrc_records * AQ_NRC_IN * sizeof(float), cudaMemcpyDeviceToHost));
    if (y_out) AQ_CK(c, cudaMemcpy(y_out, s->d_nrc_y, s->nrc_records * sizeof(float4), cudaMemcpyDeviceToHost));
    return AQ_OK;
}

}  // extern "C"
