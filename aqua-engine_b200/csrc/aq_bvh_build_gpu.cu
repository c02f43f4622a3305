/*
 * aq_bvh_build_gpu.cu — device-side acceleration-structure construction (SURVEY §8 row a5,
 * "consider device LBVH for C4"): the host SAH builder needs 6.6 s for the 10 M-triangle soup,
 * 120x the time the traversal kernel then takes for 2^26 rays.
 *
 *   triangles -> padded boxes + 63-bit Morton codes of the centroids -> radix sort (CUB)
 *   -> Karras 2012 binary radix tree (one thread per internal node)
 *   -> bottom-up box fit (second-arrival rule); subtrees of <= 3 triangles become leaf groups
 *   -> level-synchronous collapse to the 8-wide, quantised node format (aq_bvh_emit.h — the
 *      same per-node code the host builder runs)
 *
 * The result obeys the same conservativeness contract as the host builder (padded triangle
 * boxes, outward quantisation), so hit ids stay bit-exact against the brute-force oracle; only
 * the tree quality differs (LBVH instead of binned SAH).
 */
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>

#include <cub/device/device_radix_sort.cuh>

#include "aq_bvh_build.h"
#include "aq_bvh_emit.h"

namespace {

constexpr float kBoxPad = 1.0e-5f; /* same padding rule as aq_bvh_build.cpp */
constexpr uint32_t kNone = 0xFFFFFFFFu;

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

/* ---- cost-optimal collapse (Ylitie, Karras, Laine 2017, section 3.1), the device form of the dynamic
 * programme of aq_bvh_build.cpp: c[i] = cheapest way to represent a BVH2 subtree with at most i sibling
 * entries of a wide node, each entry either a leaf group (<= AQ_LEAF_MAX triangles, cost A * T * Ct) or an
 * 8-wide node (cost A * Cn + the best split of 8 entries over the two children).  The tables are filled by
 * the bottom-up pass that fits the boxes (k_fit: the second thread to arrive at a node has both children's
 * tables), the decisions are read back top-down by k_emit_level. */
struct DPd {
    float c[8];     /* c[1..7] */
    uint32_t split; /* 3 bits per j = 2..8 (at bit 3*(j-2)): entries given to the left child */
    uint32_t flags; /* bits 2..7: c[i] == c[i-1] ("fewer"); bit 8: c[1] is the leaf alternative */
};
__device__ __forceinline__ void dp_load(const DPd* p, DPd& d) { /* written by another thread of this kernel: read past L1 */
    const uint32_t* w = reinterpret_cast<const uint32_t*>(p);
#pragma unroll
    for (int i = 0; i < 8; ++i) d.c[i] = __uint_as_float(__ldcg(w + i));
    d.split = __ldcg(w + 8);
    d.flags = __ldcg(w + 9);
}
__device__ __forceinline__ void dp_make_leaf(DPd& D, float A, uint32_t count, float Ct) {
    for (int i = 0; i < 8; ++i) D.c[i] = A * (float)count * Ct;
    D.split = 0u;
    D.flags = 0xFCu | 0x100u;
}
__device__ __forceinline__ void dp_make_inner(DPd& D, const DPd& L, const DPd& R, float A, uint32_t count, float Ct) {
    float dist[9];
    D.split = 0u;
    for (int j = 2; j <= 8; ++j) {
        float best = AQ_INF;
        int bk = 1;
        for (int kk = 1; kk < j; ++kk) {
            float v = L.c[kk > 7 ? 7 : kk] + R.c[(j - kk) > 7 ? 7 : (j - kk)];
            if (v < best) {
                best = v;
                bk = kk;
            }
        }
        dist[j] = best;
        D.split |= (uint32_t)bk << (3 * (j - 2));
    }
    const float c_leaf = count <= AQ_LEAF_MAX ? A * (float)count * Ct : AQ_INF;
    const float c_int = dist[8] + A; /* Cn = 1 */
    const bool leaf = c_leaf <= c_int;
    D.c[0] = 0.0f;
    D.c[1] = leaf ? c_leaf : c_int;
    D.flags = leaf ? 0x100u : 0u;
    for (int i = 2; i < 8; ++i) {
        if (D.c[i - 1] <= dist[i]) {
            D.c[i] = D.c[i - 1];
            D.flags |= 1u << i;
        } else {
            D.c[i] = dist[i];
        }
    }
}
/* the <= 8 children of the wide node rooted at BVH2 node `root` according to the DP decisions; a subtree
 * whose cheapest single entry is a leaf group is turned into one (each BVH2 node is reached by exactly one
 * wide node, so the write is not contended) */
__device__ int dp_collect(aq_bvh2_node* N, const DPd* dp, uint32_t root, uint32_t* ch) {
    uint32_t sn[16];
    int sj[16], sp = 0, nc = 0;
    sn[sp] = root;
    sj[sp++] = 8;
    while (sp > 0) {
        const uint32_t n = sn[--sp];
        const int j = sj[sp];
        if (N[n].left == AQ_BVH2_LEAF) {
            ch[nc++] = n;
            continue;
        }
        const uint32_t fl = dp[n].flags, spl = dp[n].split;
        if (j == 1) {
            if (fl & 0x100u) N[n].left = N[n].right = AQ_BVH2_LEAF;
            ch[nc++] = n;
            continue;
        }
        if (j < 8 && (fl & (1u << j))) {
            sn[sp] = n;
            sj[sp++] = j - 1;
            continue;
        }
        const int k = (int)((spl >> (3 * (j - 2))) & 7u);
        sn[sp] = N[n].right;
        sj[sp++] = j - k;
        sn[sp] = N[n].left;
        sj[sp++] = k;
    }
    return nc;
}

/* leaf k = sorted position k, node index (n-1)+k */
__global__ void k_leaves(const float* __restrict__ pos, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ order,
                         uint32_t n, float pad, aq_bvh2_node* __restrict__ N, DPd* __restrict__ dp, float Ct) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t prim = order[k];
    aq_bvh2_node L;
    for (int a = 0; a < 3; ++a) {
        float mn = AQ_INF, mx = -AQ_INF;
        for (int v = 0; v < 3; ++v) {
            float x = pos[3 * (size_t)idx[3 * (size_t)prim + v] + a];
            mn = fminf(mn, x);
            mx = fmaxf(mx, x);
        }
        L.lo[a] = mn - pad;
        L.hi[a] = mx + pad;
    }
    L.left = L.right = AQ_BVH2_LEAF;
    L.first = k;
    L.count = 1;
    N[(size_t)(n - 1) + k] = L;
    if (dp) {
        DPd D;
        dp_make_leaf(D, aq_box_half_area(L.lo, L.hi), 1u, Ct);
        dp[(size_t)(n - 1) + k] = D;
    }
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
            dp_make_inner(D, DL, DR, aq_box_half_area(lo, hi), P->count, Ct);
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
        const int nc = dp_collect(N, dp, it.n2, ch);
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
                         cudaError_t* err) {
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
    Item *d_qa = nullptr, *d_qb = nullptr;
    aq_u4* d_nodes_tmp = nullptr;
    aq_f4* d_tris_out = nullptr;
    auto cleanup = [&]() {
        void* ps[] = {d_bounds, d_keys, d_keys2, d_vals, d_vals2, d_parent, d_flags, d_counters, d_tmp, d_n2, d_dp, d_qa, d_qb, d_nodes_tmp};
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
    k_leaves<<<G, T, 0, st>>>(d_pos, d_idx, order, n, pad, d_n2, d_dp, Ct);
    if (n > 1) {
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
    Item root{0u, 0u}; /* BVH2 root: internal node 0, or the only leaf when n == 1 (index n-1 = 0) */
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
