/*
 * aq_bvh_build_gpu.cu — device-side acceleration-structure construction (SURVEY §8 row a5,
 * "consider device LBVH for C4"): the host SAH builder needs 6.6 s for the 10 M-triangle soup,
 * 120x the time the traversal kernel then takes for 2^26 rays.
 *
 *   triangles -> padded boxes + 63-bit Morton codes of the centroids -> radix sort (CUB)
 *   -> binary tree over that order, chosen per input (tree_mode):
 *        PLOC (aq_bvh_ploc.h: nearest partner within 16 neighbours, mutual pairs merge, CUB scan compaction,
 *        passes until one cluster is left, then a top-down pass that makes every subtree's leaves contiguous)
 *        for surface-like input — SAH-builder quality on room.json — or
 *        the Karras 2012 radix tree (one thread per internal node) + bottom-up box fit (second-arrival rule)
 *        for a soup and as the 3 ms first tree of the hybrid build
 *   -> cost-optimal collapse tables (aq_dp8, filled while the tree is fitted / merged)
 *   -> level-synchronous collapse to the 8-wide, quantised node format (aq_bvh_emit.h — the
 *      same per-node code the host builder runs)
 *
 * The result obeys the same conservativeness contract as the host builder (padded triangle
 * boxes, outward quantisation), so hit ids stay bit-exact against the brute-force oracle; only
 * the tree quality differs.
 */
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "aq_bvh_build.h"
#include "aq_bvh_emit.h"
#include "aq_bvh_ploc.h"

namespace {

constexpr float kBoxPad = 1.0e-5f; /* same padding rule as aq_bvh_build.cpp */
constexpr uint32_t kNone = 0xFFFFFFFFu;
#ifndef AQ_PLOC_MAX_AREA_RATIO
#define AQ_PLOC_MAX_AREA_RATIO 16.0f /* sum of triangle-box half areas / scene-box half area below which PLOC is used */
#endif

/* order-preserving float <-> uint so atomicMin/Max work on floats */
__device__ __forceinline__ uint32_t f2ord(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord2f(uint32_t u) {
    uint32_t v = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
#if defined(__CUDA_ARCH__)
    return __uint_as_float(v);
#else
    float f;
    memcpy(&f, &v, 4);
    return f;
#endif
}

struct Bounds { /* [0..2] min, [3..5] max, ordered-uint encoded */
    uint32_t v[6];
};

__global__ void k_scene_bounds(const float* __restrict__ pos, const uint32_t* __restrict__ idx, uint32_t n_tris,
                               Bounds* b) {
    float lo[3] = {AQ_INF, AQ_INF, AQ_INF}, hi[3] = {-AQ_INF, -AQ_INF, -AQ_INF};
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n_tris; t += gridDim.x * blockDim.x)
        for (int k = 0; k < 3; ++k) {
            const float* p = pos + 3 * (size_t)idx[3 * (size_t)t + k];
            for (int a = 0; a < 3; ++a) {
                lo[a] = fminf(lo[a], p[a]);
                hi[a] = fmaxf(hi[a], p[a]);
            }
        }
    for (int a = 0; a < 3; ++a) {
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xFFFFFFFFu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xFFFFFFFFu, hi[a], o));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(&b->v[a], f2ord(lo[a]));
            atomicMax(&b->v[3 + a], f2ord(hi[a]));
        }
    }
}

__device__ __forceinline__ unsigned long long expand21(unsigned long long v) {
    v &= 0x1FFFFFull;
    v = (v | v << 32) & 0x1F00000000FFFFull;
    v = (v | v << 16) & 0x1F0000FF0000FFull;
    v = (v | v << 8) & 0x100F00F00F00F00Full;
    v = (v | v << 4) & 0x10C30C30C30C30C3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}

/* padded triangle box into its leaf node slot is written later (after sorting); here only keys */
__global__ void k_morton(const float* __restrict__ pos, const uint32_t* __restrict__ idx, uint32_t n_tris,
                         float3 lo, float3 inv_ext, unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tris) return;
    float c[3];
    for (int a = 0; a < 3; ++a) {
        float mn = AQ_INF, mx = -AQ_INF;
        for (int k = 0; k < 3; ++k) {
            float v = pos[3 * (size_t)idx[3 * (size_t)t + k] + a];
            mn = fminf(mn, v);
            mx = fmaxf(mx, v);
        }
        c[a] = 0.5f * (mn + mx);
    }
    float nx = fminf(fmaxf((c[0] - lo.x) * inv_ext.x, 0.0f), 1.0f);
    float ny = fminf(fmaxf((c[1] - lo.y) * inv_ext.y, 0.0f), 1.0f);
    float nz = fminf(fmaxf((c[2] - lo.z) * inv_ext.z, 0.0f), 1.0f);
    unsigned long long qx = (unsigned long long)fminf(nx * 2097152.0f, 2097151.0f);
    unsigned long long qy = (unsigned long long)fminf(ny * 2097152.0f, 2097151.0f);
    unsigned long long qz = (unsigned long long)fminf(nz * 2097152.0f, 2097151.0f);
    keys[t] = (expand21(qx) << 2) | (expand21(qy) << 1) | expand21(qz);
    vals[t] = t;
}

/* ---- cost-optimal collapse: tables aq_dp8 (aq_bvh_ploc.h).  On the radix-tree path they are filled by the
 * bottom-up pass that fits the boxes (k_fit: the second thread to arrive at a node has both children's
 * tables), on the PLOC path when a parent is created; the decisions are read back top-down by k_emit_level. */
typedef aq_dp8 DPd;
__device__ __forceinline__ void dp_load(const DPd* p, DPd& d) { /* written by another thread of this kernel: read past L1 */
    const uint32_t* w = reinterpret_cast<const uint32_t*>(p);
#pragma unroll
    for (int i = 0; i < 8; ++i) d.c[i] = __uint_as_float(__ldcg(w + i));
    d.split = __ldcg(w + 8);
    d.flags = __ldcg(w + 9);
}

/* leaf k = sorted position k, node index (n-1)+k; area_sum accumulates the half areas of the (unpadded) triangle
 * boxes: their sum over the scene box's half area tells a surface (room.json: 9.4) from a volume-filling soup
 * (C4 at 10^6 triangles: 24.5, at 10^7: 245) */
__global__ void k_leaves(const float* __restrict__ pos, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ order,
                         uint32_t n, float pad, aq_bvh2_node* __restrict__ N, DPd* __restrict__ dp, float Ct, double* area_sum) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    float raw_area = 0.0f;
    if (k < n) {
        uint32_t prim = order[k];
        aq_bvh2_node L;
        float rl[3], rh[3];
        for (int a = 0; a < 3; ++a) {
            float mn = AQ_INF, mx = -AQ_INF;
            for (int v = 0; v < 3; ++v) {
                float x = pos[3 * (size_t)idx[3 * (size_t)prim + v] + a];
                mn = fminf(mn, x);
                mx = fmaxf(mx, x);
            }
            rl[a] = mn;
            rh[a] = mx;
            L.lo[a] = mn - pad;
            L.hi[a] = mx + pad;
        }
        raw_area = aq_box_half_area(rl, rh);
        L.left = L.right = AQ_BVH2_LEAF;
        L.first = k;
        L.count = 1;
        N[(size_t)(n - 1) + k] = L;
        if (dp) {
            DPd D;
            aq_dp8_leaf(D, aq_box_half_area(L.lo, L.hi), 1u, Ct);
            dp[(size_t)(n - 1) + k] = D;
        }
    }
    for (int o = 16; o > 0; o >>= 1) raw_area += __shfl_xor_sync(0xFFFFFFFFu, raw_area, o);
    if ((threadIdx.x & 31) == 0 && raw_area > 0.0f) atomicAdd(area_sum, (double)raw_area);
}

__device__ __forceinline__ int delta(const unsigned long long* __restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    unsigned long long a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz(i ^ j); /* duplicate codes: fall back to the index */
    return __clzll((long long)(a ^ b));
}

/* Karras 2012, "Maximizing Parallelism in the Construction of BVHs, Octrees, and k-d Trees" */
__global__ void k_hierarchy(const unsigned long long* __restrict__ keys, uint32_t n, aq_bvh2_node* __restrict__ N,
                            uint32_t* __restrict__ parent) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int nn = (int)n;
    if (i >= nn - 1) return;
    int d = (delta(keys, nn, i, i + 1) - delta(keys, nn, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = delta(keys, nn, i, i - d);
    int lmax = 2;
    while (delta(keys, nn, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, nn, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = delta(keys, nn, i, j);
    int s = 0, t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(keys, nn, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + (d < 0 ? -1 : 0);
    int first = i < j ? i : j, last = i < j ? j : i;
    uint32_t left = (first == gamma) ? (uint32_t)(nn - 1 + gamma) : (uint32_t)gamma;
    uint32_t right = (last == gamma + 1) ? (uint32_t)(nn - 1 + gamma + 1) : (uint32_t)(gamma + 1);
    N[i].left = left;
    N[i].right = right;
    N[i].first = (uint32_t)first;
    N[i].count = (uint32_t)(last - first + 1);
    parent[left] = (uint32_t)i;
    parent[right] = (uint32_t)i;
    if (i == 0) parent[0] = kNone;
}

/* bottom-up: the second thread to arrive at a node owns it */
__global__ void k_fit(uint32_t n, aq_bvh2_node* N, const uint32_t* __restrict__ parent, uint32_t* flags, DPd* dp, float Ct) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t cur = parent[(size_t)(n - 1) + k];
    while (cur != kNone) {
        if (atomicAdd(&flags[cur], 1u) == 0u) return;
        __threadfence();
        volatile aq_bvh2_node* P = &N[cur];
        const uint32_t l = P->left, r = P->right;
        volatile aq_bvh2_node* A = &N[l];
        volatile aq_bvh2_node* B = &N[r];
        for (int a = 0; a < 3; ++a) {
            P->lo[a] = fminf(A->lo[a], B->lo[a]);
            P->hi[a] = fmaxf(A->hi[a], B->hi[a]);
        }
        /* two small leaf groups merge into one when the SAH says testing all their triangles
         * together is not dearer than descending (same rule as the host builder); merging
         * unconditionally made every leaf hit cost 3 triangle tests (19 instead of 7 per ray on
         * the 10 M-triangle soup) */
        if (dp) { /* the DP decides about leaf groups too (its c[1] leaf alternative, applied by dp_collect) */
            float lo[3], hi[3];
            for (int a = 0; a < 3; ++a) {
                lo[a] = P->lo[a];
                hi[a] = P->hi[a];
            }
            DPd DL, DR, D;
            dp_load(dp + l, DL);
            dp_load(dp + r, DR);
            aq_dp8_inner(D, DL, DR, aq_box_half_area(lo, hi), P->count, Ct);
            dp[cur] = D;
        } else if (P->count <= AQ_LEAF_MAX && A->left == AQ_BVH2_LEAF && B->left == AQ_BVH2_LEAF) {
            float lo[3], hi[3], al[3], ah[3], bl[3], bh[3];
            for (int a = 0; a < 3; ++a) {
                lo[a] = P->lo[a]; hi[a] = P->hi[a];
                al[a] = A->lo[a]; ah[a] = A->hi[a];
                bl[a] = B->lo[a]; bh[a] = B->hi[a];
            }
            float ap = aq_box_half_area(lo, hi);
            float leaf_cost = ap * (float)P->count;
            float split_cost = aq_box_half_area(al, ah) * (float)A->count + aq_box_half_area(bl, bh) * (float)B->count;
            if (leaf_cost <= split_cost + 0.5f * ap) {
                P->left = AQ_BVH2_LEAF;
                P->right = AQ_BVH2_LEAF;
            }
        }
        __threadfence();
        cur = parent[cur];
    }
}

/* ---- PLOC stage (aq_bvh_ploc.h): thin wrappers, one thread per cluster */
__global__ void k_ploc_init(uint32_t n, const aq_bvh2_node* __restrict__ N, uint32_t* __restrict__ cid, aq_box6* __restrict__ cbox) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const aq_bvh2_node& L = N[(size_t)(n - 1) + p];
    aq_box6 b;
    for (int a = 0; a < 3; ++a) {
        b.lo[a] = L.lo[a];
        b.hi[a] = L.hi[a];
    }
    cid[p] = n - 1 + p;
    cbox[p] = b;
}
__global__ void k_ploc_nn(const aq_box6* __restrict__ cbox, uint32_t m, uint32_t R, uint32_t* __restrict__ nn) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) nn[i] = aq_ploc_nearest(cbox, m, i, R);
}
/* counters: [0] internal nodes allocated so far */
__global__ void k_ploc_merge(aq_bvh2_node* N, DPd* dp, const uint32_t* __restrict__ cid, const aq_box6* __restrict__ cbox,
                             const uint32_t* __restrict__ nn, uint32_t m, uint32_t* counters, uint32_t* __restrict__ out_id,
                             aq_box6* __restrict__ out_box, uint32_t* __restrict__ keep, float Ct) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int fate = aq_ploc_fate(nn, i);
    keep[i] = fate != 0 ? 1u : 0u;
    if (fate == 2) {
        const uint32_t j = nn[i], k = atomicAdd(&counters[0], 1u);
        aq_ploc_make_parent(N, dp, k, cid[i], cid[j], cbox[i], cbox[j], Ct, &out_box[i]);
        out_id[i] = k;
    } else if (fate == 1) {
        out_id[i] = cid[i];
        out_box[i] = cbox[i];
    }
}
__global__ void k_ploc_scatter(const uint32_t* __restrict__ out_id, const aq_box6* __restrict__ out_box,
                               const uint32_t* __restrict__ keep, const uint32_t* __restrict__ pos, uint32_t m,
                               uint32_t* __restrict__ cid2, aq_box6* __restrict__ cbox2, uint32_t* counters) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    if (keep[i]) {
        cid2[pos[i]] = out_id[i];
        cbox2[pos[i]] = out_box[i];
    }
    if (i == m - 1) counters[1] = pos[i] + keep[i]; /* size of the next cluster array */
}
__global__ void k_ploc_first(aq_bvh2_node* N, uint32_t begin, uint32_t end) {
    uint32_t k = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (k < end) aq_ploc_assign_first(N, k);
}
__global__ void k_ploc_root(aq_bvh2_node* N, const uint32_t* __restrict__ cid) { N[cid[0]].first = 0u; }
__global__ void k_ploc_order(const aq_bvh2_node* __restrict__ N, uint32_t n, const uint32_t* __restrict__ order_old,
                             uint32_t* __restrict__ order_new) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) order_new[N[(size_t)(n - 1) + p].first] = order_old[p];
}

struct Item {
    uint32_t n2, out;
};

/* one wide node per thread; children of this level are appended to q_out */
__global__ void k_emit_level(aq_bvh2_node* N, const DPd* __restrict__ dp, const uint32_t* __restrict__ order,
                             const float* __restrict__ pos, const uint32_t* __restrict__ idx,
                             const Item* __restrict__ q_in, uint32_t n_in, Item* __restrict__ q_out,
                             uint32_t* counters /* [0] nodes, [1] tris, [2] q_out size */, uint32_t node_cap,
                             aq_u4* __restrict__ nodes, aq_f4* __restrict__ tris) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_in) return;
    Item it = q_in[i];
    aq_node8_plan plan;
    if (dp && N[it.n2].left != AQ_BVH2_LEAF) {
        uint32_t ch[8];
        const int nc = aq_dp8_collect(N, dp, it.n2, ch);
        aq_node8_plan_from(N, ch, nc, &plan);
    } else {
        aq_node8_plan_children(N, it.n2, &plan);
    }
    uint32_t child_base = plan.n_inner ? atomicAdd(&counters[0], plan.n_inner) : 0u;
    uint32_t tri_base = plan.n_tris ? atomicAdd(&counters[1], plan.n_tris) : 0u;
    if (child_base + plan.n_inner > node_cap) { /* reported by the host after the level */
        return;
    }
    uint32_t inner[8];
    aq_node8_write(N, plan, order, pos, idx, child_base, tri_base, nodes + (size_t)it.out * AQ_NODE_WORDS, tris, inner);
    if (plan.n_inner) {
        uint32_t q = atomicAdd(&counters[2], plan.n_inner);
        for (uint32_t k = 0; k < plan.n_inner; ++k) q_out[q + k] = Item{inner[k], child_base + k};
    }
}

#define CK(call)                                 \
    do {                                         \
        cudaError_t e_ = (call);                 \
        if (e_ != cudaSuccess) {                 \
            *err = e_;                           \
            cleanup();                           \
            return -2;                           \
        }                                        \
    } while (0)

}  // namespace

/* Builds the BVH8 of (d_pos, d_idx) on the device.  On success *d_nodes / *d_tris are fresh
 * cudaMalloc'ed buffers owned by the caller.  Returns 0, -1 (tree too deep / node buffer
 * overflow) or -2 (CUDA error in *err). */
int aq_build_bvh8_device(cudaStream_t st, const float* d_pos, const uint32_t* d_idx, uint32_t n_tris,
                         aq_u4** d_nodes, size_t* n_node_words, aq_f4** d_tris, uint32_t* max_depth,
                         cudaError_t* err, int tree_mode) {
    *err = cudaSuccess;
    *d_nodes = nullptr;
    *d_tris = nullptr;
    const uint32_t n = n_tris;
    Bounds* d_bounds = nullptr;
    unsigned long long *d_keys = nullptr, *d_keys2 = nullptr;
    uint32_t *d_vals = nullptr, *d_vals2 = nullptr, *d_parent = nullptr, *d_flags = nullptr, *d_counters = nullptr;
    void* d_tmp = nullptr;
    aq_bvh2_node* d_n2 = nullptr;
    DPd* d_dp = nullptr;
    uint32_t *d_cid = nullptr, *d_cid2 = nullptr, *d_nn = nullptr, *d_oid = nullptr, *d_keep = nullptr, *d_ppos = nullptr, *d_pcnt = nullptr;
    aq_box6 *d_cbox = nullptr, *d_cbox2 = nullptr, *d_obox = nullptr;
    void* d_scan_tmp = nullptr;
    double* d_area = nullptr;
    Item *d_qa = nullptr, *d_qb = nullptr;
    aq_u4* d_nodes_tmp = nullptr;
    aq_f4* d_tris_out = nullptr;
    auto cleanup = [&]() {
        void* ps[] = {d_bounds, d_keys, d_keys2, d_vals, d_vals2, d_parent, d_flags, d_counters, d_tmp, d_n2, d_dp, d_qa, d_qb, d_nodes_tmp,
                      d_cid, d_cid2, d_nn, d_oid, d_keep, d_ppos, d_pcnt, d_cbox, d_cbox2, d_obox, d_scan_tmp, d_area};
        for (void* p : ps)
            if (p) cudaFreeAsync(p, st);
        if (*err != cudaSuccess && d_tris_out) cudaFreeAsync(d_tris_out, st);
    };
    const int T = 256;
    const unsigned G = (n + T - 1) / T;

    /* ---- scene bounds, padding scale */
    CK(cudaMallocAsync((void**)&d_bounds, sizeof(Bounds), st));
    Bounds hb;
    for (int a = 0; a < 3; ++a) {
        hb.v[a] = 0xFFFFFFFFu;
        hb.v[3 + a] = 0u;
    }
    CK(cudaMemcpyAsync(d_bounds, &hb, sizeof hb, cudaMemcpyHostToDevice, st));
    k_scene_bounds<<<1184, T, 0, st>>>(d_pos, d_idx, n, d_bounds);
    CK(cudaMemcpyAsync(&hb, d_bounds, sizeof hb, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    float lo[3], hi[3], scale = 0.f;
    for (int a = 0; a < 3; ++a) {
        lo[a] = ord2f(hb.v[a]);
        hi[a] = ord2f(hb.v[3 + a]);
        scale = fmaxf(scale, hi[a] - lo[a]);
        scale = fmaxf(scale, fabsf(lo[a]));
        scale = fmaxf(scale, fabsf(hi[a]));
    }
    const float pad = kBoxPad * scale;

    /* ---- Morton codes + sort */
    CK(cudaMallocAsync((void**)&d_keys, (size_t)n * 8, st));
    CK(cudaMallocAsync((void**)&d_keys2, (size_t)n * 8, st));
    CK(cudaMallocAsync((void**)&d_vals, (size_t)n * 4, st));
    CK(cudaMallocAsync((void**)&d_vals2, (size_t)n * 4, st));
    float3 flo = make_float3(lo[0], lo[1], lo[2]);
    float3 inv = make_float3(hi[0] > lo[0] ? 1.0f / (hi[0] - lo[0]) : 0.f, hi[1] > lo[1] ? 1.0f / (hi[1] - lo[1]) : 0.f,
                             hi[2] > lo[2] ? 1.0f / (hi[2] - lo[2]) : 0.f);
    k_morton<<<G, T, 0, st>>>(d_pos, d_idx, n, flo, inv, d_keys, d_vals);
    size_t tmp_bytes = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys, d_keys2, d_vals, d_vals2, (int)n, 0, 63, st));
    CK(cudaMallocAsync(&d_tmp, tmp_bytes ? tmp_bytes : 1, st));
    CK(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_keys, d_keys2, d_vals, d_vals2, (int)n, 0, 63, st));
    const unsigned long long* keys = d_keys2;
    const uint32_t* order = d_vals2;

    /* ---- binary radix tree + box fit */
    CK(cudaMallocAsync((void**)&d_n2, (size_t)(2 * (size_t)n) * sizeof(aq_bvh2_node), st));
    CK(cudaMallocAsync((void**)&d_parent, (size_t)(2 * (size_t)n) * 4, st));
    CK(cudaMallocAsync((void**)&d_flags, (size_t)n * 4, st));
    CK(cudaMemsetAsync(d_flags, 0, (size_t)n * 4, st));
    CK(cudaMemsetAsync(d_parent, 0xFF, (size_t)(2 * (size_t)n) * 4, st));
    /* cost-optimal collapse tables (AQUA_COLLAPSE=greedy: the round-1 surface-area collapse).  Triangle / node
     * cost ratio AQUA_BVH_CT, default 1.0 here: on the overlapping boxes of an LBVH tree over the 10 M-triangle
     * soup the host builder's 0.3 merges too many leaves (7.1 -> 12.0 triangle tests per ray, -6 %), 1.0 gives
     * +1.8 % closest / +3.8 % any-hit over the greedy collapse; room.json on the device tree: -5.1 % render time
     * at any ratio (profiles/r02b_ab_device_dp_collapse.log) */
    const char* cm = std::getenv("AQUA_COLLAPSE");
    const bool use_dp = !(cm && !std::strcmp(cm, "greedy"));
    float Ct = 1.0f;
    if (const char* e = std::getenv("AQUA_BVH_CT")) {
        const float c = (float)std::atof(e);
        if (c > 0.f) Ct = c;
    }
    if (use_dp) CK(cudaMallocAsync((void**)&d_dp, (size_t)(2 * (size_t)n) * sizeof(DPd), st));
    CK(cudaMallocAsync((void**)&d_area, sizeof(double), st));
    CK(cudaMemsetAsync(d_area, 0, sizeof(double), st));
    k_leaves<<<G, T, 0, st>>>(d_pos, d_idx, order, n, pad, d_n2, d_dp, Ct, d_area);
    /* binary tree over the Morton order: PLOC or the Karras radix tree.
     *   room.json on the device-built tree: radix 23.48 ms, PLOC 19.28 ms per 1080p x 8 spp render — level with the
     *   host SAH tree (19.35) — for 37 ms instead of 3 ms of build;
     *   C4 soup (10^7 triangles): PLOC -32 % closest-hit (39 instead of 27 node visits per ray), any-hit equal, 193
     *   instead of 22 ms of build (profiles/r02c_ab_device_ploc.log).
     * tree_mode 0 (auto) therefore takes PLOC for surface-like input (box-area ratio, k_leaves) and the radix tree
     * for a soup; 1 / 2 force radix / PLOC (the hybrid build's first tree is always the 3 ms one);
     * AQUA_DEVICE_TREE=lbvh|ploc overrides, AQUA_PLOC_RADIUS sets the search radius.  PLOC gives up after 256
     * passes (an input on which only a few pairs merge per pass) and the radix tree is built instead. */
    bool use_ploc = tree_mode == 2;
    if (tree_mode == 0 && n > 1) {
        double h_area = 0.0;
        CK(cudaMemcpyAsync(&h_area, d_area, sizeof h_area, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const float root_area = aq_box_half_area(lo, hi);
        use_ploc = root_area > 0.0f && h_area / (double)root_area < (double)AQ_PLOC_MAX_AREA_RATIO;
    }
    const char* tree_env = std::getenv("AQUA_DEVICE_TREE");
    if (tree_env) use_ploc = !std::strcmp(tree_env, "ploc");
    uint32_t root_node = 0u; /* the radix tree's root is internal node 0 (or the only leaf, index n-1 = 0) */
    if (use_ploc && n > 1) {
        uint32_t R = 16;
        if (const char* e = std::getenv("AQUA_PLOC_RADIUS")) {
            const int r = std::atoi(e);
            if (r >= 1 && r <= 64) R = (uint32_t)r;
        }
        CK(cudaMallocAsync((void**)&d_cid, (size_t)n * 4, st));
        CK(cudaMallocAsync((void**)&d_cid2, (size_t)n * 4, st));
        CK(cudaMallocAsync((void**)&d_nn, (size_t)n * 4, st));
        CK(cudaMallocAsync((void**)&d_oid, (size_t)n * 4, st));
        CK(cudaMallocAsync((void**)&d_keep, (size_t)n * 4, st));
        CK(cudaMallocAsync((void**)&d_ppos, (size_t)n * 4, st));
        CK(cudaMallocAsync((void**)&d_cbox, (size_t)n * sizeof(aq_box6), st));
        CK(cudaMallocAsync((void**)&d_cbox2, (size_t)n * sizeof(aq_box6), st));
        CK(cudaMallocAsync((void**)&d_obox, (size_t)n * sizeof(aq_box6), st));
        CK(cudaMallocAsync((void**)&d_pcnt, 2 * 4, st));
        CK(cudaMemsetAsync(d_pcnt, 0, 2 * 4, st));
        size_t scan_bytes = 0;
        CK(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, d_keep, d_ppos, (int)n, st));
        CK(cudaMallocAsync(&d_scan_tmp, scan_bytes ? scan_bytes : 1, st));
        k_ploc_init<<<G, T, 0, st>>>(n, d_n2, d_cid, d_cbox);
        std::vector<uint32_t> pass_begin; /* first internal node created by each pass */
        uint32_t m = n, made = 0;
        bool gave_up = false;
        while (m > 1) {
            if (pass_begin.size() >= 256) {
                gave_up = true;
                break;
            }
            pass_begin.push_back(made);
            const unsigned Gm = (m + T - 1) / T;
            k_ploc_nn<<<Gm, T, 0, st>>>(d_cbox, m, R, d_nn);
            k_ploc_merge<<<Gm, T, 0, st>>>(d_n2, d_dp, d_cid, d_cbox, d_nn, m, d_pcnt, d_oid, d_obox, d_keep, Ct);
            CK(cub::DeviceScan::ExclusiveSum(d_scan_tmp, scan_bytes, d_keep, d_ppos, (int)m, st));
            k_ploc_scatter<<<Gm, T, 0, st>>>(d_oid, d_obox, d_keep, d_ppos, m, d_cid2, d_cbox2, d_pcnt);
            uint32_t hp[2];
            CK(cudaMemcpyAsync(hp, d_pcnt, sizeof hp, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            made = hp[0];
            if (hp[1] >= m) { /* no pair merged: cannot happen (the closest pair of a pass is mutual); do not spin */
                gave_up = true;
                break;
            }
            m = hp[1];
            std::swap(d_cid, d_cid2);
            std::swap(d_cbox, d_cbox2);
        }
        if (!gave_up) {
            pass_begin.push_back(made);
            CK(cudaMemcpyAsync(&root_node, d_cid, 4, cudaMemcpyDeviceToHost, st));
            k_ploc_root<<<1, 1, 0, st>>>(d_n2, d_cid);
            for (size_t t = pass_begin.size() - 1; t-- > 0;) { /* parents first: the passes in reverse */
                const uint32_t b0 = pass_begin[t], b1 = pass_begin[t + 1];
                if (b1 > b0) k_ploc_first<<<(b1 - b0 + T - 1) / T, T, 0, st>>>(d_n2, b0, b1);
            }
            k_ploc_order<<<G, T, 0, st>>>(d_n2, n, order, d_vals); /* d_vals (the unsorted ids) is free since the sort */
            CK(cudaStreamSynchronize(st));
            order = d_vals;
        } else {
            use_ploc = false; /* rebuild the leaves (their tables are untouched) and take the radix tree */
        }
    }
    if (!(use_ploc && n > 1) && n > 1) {
        k_hierarchy<<<(n - 1 + T - 1) / T, T, 0, st>>>(keys, n, d_n2, d_parent);
        k_fit<<<G, T, 0, st>>>(n, d_n2, d_parent, d_flags, d_dp, Ct);
    }
    CK(cudaGetLastError());

    /* ---- collapse to 8-wide, one level per launch */
    const uint32_t node_cap = n / 2 + 1024;
    CK(cudaMallocAsync((void**)&d_nodes_tmp, (size_t)node_cap * AQ_NODE_WORDS * sizeof(aq_u4), st));
    CK(cudaMallocAsync((void**)&d_tris_out, (size_t)(n ? n : 1) * AQ_TRI_WORDS * sizeof(aq_f4), st));
    CK(cudaMallocAsync((void**)&d_qa, (size_t)node_cap * sizeof(Item), st));
    CK(cudaMallocAsync((void**)&d_qb, (size_t)node_cap * sizeof(Item), st));
    CK(cudaMallocAsync((void**)&d_counters, 3 * 4, st));
    uint32_t hc[3] = {1u, 0u, 0u}; /* node 0 = root is allocated */
    CK(cudaMemcpyAsync(d_counters, hc, sizeof hc, cudaMemcpyHostToDevice, st));
    Item root{root_node, 0u}; /* BVH2 root: internal node 0 of the radix tree (or the only leaf when n == 1, index n-1 = 0), PLOC's last cluster */
    CK(cudaMemcpyAsync(d_qa, &root, sizeof root, cudaMemcpyHostToDevice, st));
    uint32_t n_in = 1, depth = 0;
    Item *qi = d_qa, *qo = d_qb;
    while (n_in > 0) {
        ++depth;
        k_emit_level<<<(n_in + 127) / 128, 128, 0, st>>>(d_n2, d_dp, order, d_pos, d_idx, qi, n_in, qo, d_counters, node_cap,
                                                          d_nodes_tmp, d_tris_out);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(hc, d_counters, sizeof hc, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (hc[0] > node_cap || depth >= AQ_STACK_MAX) {
            cleanup();
            cudaFreeAsync(d_tris_out, st);
            return -1;
        }
        n_in = hc[2];
        uint32_t zero = 0;
        CK(cudaMemcpyAsync(d_counters + 2, &zero, 4, cudaMemcpyHostToDevice, st));
        Item* t = qi;
        qi = qo;
        qo = t;
    }
    /* ---- exact-size node buffer */
    const size_t words = (size_t)hc[0] * AQ_NODE_WORDS;
    aq_u4* d_final = nullptr;
    CK(cudaMallocAsync((void**)&d_final, words * sizeof(aq_u4), st));
    cudaError_t ce = cudaMemcpyAsync(d_final, d_nodes_tmp, words * sizeof(aq_u4), cudaMemcpyDeviceToDevice, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    if (ce != cudaSuccess) {
        *err = ce;
        cudaFreeAsync(d_final, st);
        cleanup();
        return -2;
    }
    *d_nodes = d_final;
    *n_node_words = words;
    *d_tris = d_tris_out;
    *max_depth = depth;
    cleanup();
    return 0;
}
