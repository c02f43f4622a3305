/*
 * aq_multi.cu — aq_render_multi: single-process multi-GPU render.
 *
 * Scene + BVH8 replicated on every device (the BVH is built once on the host and cloned
 * device-to-device), the sample range [spp_begin, spp_end) split contiguously per device,
 * every device renders a full-resolution float4 film, and ONE ncclReduce(sum, root = first
 * device) over NVLink combines them.  NCCL is bound at run time with dlopen so that
 * libaqua_cuda.so itself has no link-time dependency on it.
 */
#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "aq_internal.h"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef int ncclResult_t;
struct Nccl {
    void* h = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, int, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load(std::string* err) {
        if (h) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
        if (!h) {
            *err = std::string("cannot dlopen libnccl.so.2: ") + dlerror();
            return false;
        }
#define SYM(field, name)                                       \
    field = reinterpret_cast<decltype(field)>(dlsym(h, name)); \
    if (!field) {                                              \
        *err = std::string("libnccl lacks ") + name;           \
        return false;                                          \
    }
        SYM(CommInitAll, "ncclCommInitAll");
        SYM(CommDestroy, "ncclCommDestroy");
        SYM(GroupStart, "ncclGroupStart");
        SYM(GroupEnd, "ncclGroupEnd");
        SYM(Reduce, "ncclReduce");
        SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
        return true;
    }
};
Nccl g_nccl;
const int kNcclFloat = 7, kNcclSum = 0;

}  // namespace

extern "C" int aq_render_multi(const aq_scene_desc* desc, const aq_integrator_cfg* cfg, int n_gpus,
                               const int* devices, float* film_out, aq_stats* stats) {
    if (!desc || !cfg || !film_out || n_gpus < 1)
        return aq_internal_set_error(nullptr, AQ_ERR_BAD_ARG, "aq_render_multi: bad argument");
    if (cfg->spp_end < cfg->spp_begin)
        return aq_internal_set_error(nullptr, AQ_ERR_BAD_ARG, "aq_render_multi: spp_end < spp_begin");
    std::vector<aq_ctx*> ctx(n_gpus, nullptr);
    std::vector<aq_scene*> sc(n_gpus, nullptr);
    std::vector<ncclComm_t> comms(n_gpus, nullptr);
    std::vector<int> devs(n_gpus);
    for (int i = 0; i < n_gpus; ++i) devs[i] = devices ? devices[i] : i;
    int rc = AQ_OK;
    std::string err;
    auto cleanup = [&]() {
        for (int i = 0; i < n_gpus; ++i) {
            if (comms[i]) g_nccl.CommDestroy(comms[i]);
            if (sc[i]) aq_scene_destroy(sc[i]);
            if (ctx[i]) aq_destroy(ctx[i]);
        }
    };
    for (int i = 0; i < n_gpus && rc == AQ_OK; ++i) {
        rc = aq_init(devs[i], &ctx[i]);
        if (rc == AQ_OK) rc = aq_scene_create(ctx[i], desc, &sc[i]);
        if (rc == AQ_OK) rc = (i == 0) ? aq_accel_build(sc[0], nullptr) : aq_internal_clone_accel(sc[i], sc[0]);
    }
    if (rc != AQ_OK) {
        cleanup();
        return rc; /* message already set by the failing call */
    }
    if (n_gpus > 1) {
        if (!g_nccl.load(&err)) {
            cleanup();
            return aq_internal_set_error(nullptr, AQ_ERR_NCCL, err.c_str());
        }
        ncclResult_t nr = g_nccl.CommInitAll(comms.data(), n_gpus, devs.data());
        if (nr != 0) {
            std::string m = std::string("ncclCommInitAll: ") + g_nccl.GetErrorString(nr);
            for (auto& c : comms) c = nullptr;
            cleanup();
            return aq_internal_set_error(nullptr, AQ_ERR_NCCL, m.c_str());
        }
    }
    const uint32_t W = cfg->width ? cfg->width : desc->camera.res[0];
    const uint32_t H = cfg->height ? cfg->height : desc->camera.res[1];
    const size_t count = (size_t)W * H * 4;
    const uint32_t n = cfg->spp_end - cfg->spp_begin;
    /* every device renders its contiguous share (asynchronously, one stream per device) */
    for (int i = 0; i < n_gpus && rc == AQ_OK; ++i) {
        aq_integrator_cfg c = *cfg;
        c.spp_begin = cfg->spp_begin + (uint32_t)(((uint64_t)n * i) / n_gpus);
        c.spp_end = cfg->spp_begin + (uint32_t)(((uint64_t)n * (i + 1)) / n_gpus);
        c.flags &= ~AQ_RENDER_ACCUMULATE;
        rc = aq_render_device_async(sc[i], &c, nullptr);
    }
    if (rc == AQ_OK && n_gpus > 1) {
        g_nccl.GroupStart();
        for (int i = 0; i < n_gpus; ++i) {
            cudaSetDevice(devs[i]);
            void* f = aq_internal_film(sc[i]);
            ncclResult_t nr = g_nccl.Reduce(f, f, count, kNcclFloat, kNcclSum, 0, comms[i], aq_internal_stream(sc[i]));
            if (nr != 0 && rc == AQ_OK) {
                std::string m = std::string("ncclReduce: ") + g_nccl.GetErrorString(nr);
                rc = aq_internal_set_error(nullptr, AQ_ERR_NCCL, m.c_str());
            }
        }
        g_nccl.GroupEnd();
    }
    aq_stats total;
    std::memset(&total, 0, sizeof total);
    for (int i = 0; i < n_gpus; ++i) {
        aq_stats st;
        std::memset(&st, 0, sizeof st); /* aq_render_finish may fail before it fills the block */
        int r2 = aq_render_finish(sc[i], &st);
        if (r2 != AQ_OK && rc == AQ_OK) rc = r2;
        total.samples += st.samples;
        total.sample_bounces += st.sample_bounces;
        total.rays_closest += st.rays_closest;
        total.rays_shadow += st.rays_shadow;
        total.n_launches += st.n_launches;
        total.n_waves += st.n_waves;
        if (st.ms_total > total.ms_total) total.ms_total = st.ms_total;
    }
    if (rc == AQ_OK) {
        cudaSetDevice(devs[0]);
        cudaError_t e = cudaMemcpy(film_out, aq_internal_film(sc[0]), count * sizeof(float), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = aq_internal_set_error(nullptr, AQ_ERR_CUDA, cudaGetErrorString(e));
    }
    if (stats) *stats = total;
    cleanup();
    return rc;
}
