/*
 * aq_resolve.cu — output stage on the device (§8e: "fuse weight-normalise
 * + tonemap into the post-reduce resolve kernel on root"): float4 film (sum r,g,b, count) ->
 * exposure -> clamp -> sRGB OETF -> RGBA8.  The reference wrote images through the `image`
 * crate (Cargo.toml:13); this replaces that step for the film this library produces.
 *
 * The OETF is evaluated as a search over 255 thresholds computed once in double precision, so
 * the 8-bit result is exact and identical to the same search on the host (no powf ulps).
 */
#include <cuda_runtime.h>

#include <cmath>

#include "aq_internal.h"

namespace {

__constant__ float c_thresh[256];

__device__ __forceinline__ uint32_t to_srgb8(float v) {
    v = v < 0.0f ? 0.0f : v; /* NaN -> 0 as well */
    if (!(v == v)) v = 0.0f;
    uint32_t lo = 0u, hi = 255u; /* number of thresholds <= v */
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (v >= c_thresh[mid])
            lo = mid + 1u;
        else
            hi = mid;
    }
    return lo;
}

__global__ void aq_k_resolve(const float4* __restrict__ film, uint32_t n, float exposure, uchar4* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 f = film[i];
    float w = f.w > 0.0f ? exposure / f.w : 0.0f;
    out[i] = make_uchar4((unsigned char)to_srgb8(f.x * w), (unsigned char)to_srgb8(f.y * w),
                         (unsigned char)to_srgb8(f.z * w), 255);
}

}  // namespace

extern "C" int aq_resolve(aq_ctx* ctx, const void* d_film, const float* h_film, uint32_t width, uint32_t height,
                          float exposure, uint8_t* rgba8_out) {
    if (!ctx || (!d_film && !h_film) || !rgba8_out || !width || !height)
        return aq_internal_set_error(ctx, AQ_ERR_BAD_ARG, "aq_resolve: bad argument");
    cudaStream_t st = aq_internal_ctx_stream(ctx);
    cudaError_t e = cudaSetDevice(aq_internal_ctx_device(ctx));
    static bool thresholds_ready[64] = {false};
    int dev = aq_internal_ctx_device(ctx);
    if (e == cudaSuccess && dev < 64 && !thresholds_ready[dev]) {
        float t[256];
        for (int k = 0; k < 255; ++k) {
            double s = ((double)k + 0.5) / 255.0;
            t[k] = (float)(s <= 0.04045 ? s / 12.92 : std::pow((s + 0.055) / 1.055, 2.4));
        }
        t[255] = 3.0e38f;
        e = cudaMemcpyToSymbolAsync(c_thresh, t, sizeof t, 0, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e == cudaSuccess) thresholds_ready[dev] = true;
    }
    const size_t n = (size_t)width * height;
    float4* tmp_film = nullptr;
    uchar4* d_out = nullptr;
    if (e == cudaSuccess && !d_film) {
        e = cudaMalloc((void**)&tmp_film, n * sizeof(float4));
        if (e == cudaSuccess) e = cudaMemcpyAsync(tmp_film, h_film, n * sizeof(float4), cudaMemcpyHostToDevice, st);
        d_film = tmp_film;
    }
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_out, n * sizeof(uchar4));
    if (e == cudaSuccess) {
        aq_k_resolve<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const float4*)d_film, (uint32_t)n, exposure, d_out);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(rgba8_out, d_out, n * sizeof(uchar4), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (tmp_film) cudaFree(tmp_film);
    if (d_out) cudaFree(d_out);
    if (e != cudaSuccess) return aq_internal_set_error(ctx, AQ_ERR_CUDA, cudaGetErrorString(e));
    return AQ_OK;
}
