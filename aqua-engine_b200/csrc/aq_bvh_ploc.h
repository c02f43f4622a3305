/*
 * aq_bvh_ploc.h — per-element steps of the device builder's binary-tree stage, written as host/device
 * functions so that the kernels of aq_bvh_build_gpu.cu are thin wrappers and the same code can be replayed
 * on the CPU (tools/experimental/ploc_emulate.cpp).  SURVEY §8 row a5.
 *
 *  - aq_dp8_*: the cost-optimal BVH2 -> BVH8 collapse tables (Ylitie, Karras, Laine 2017, section 3.1; the
 *    host form is in aq_bvh_build.cpp): c[i] = cheapest way to represent a subtree with at most i sibling
 *    entries of a wide node, each entry a leaf group (<= AQ_LEAF_MAX triangles, cost A * T * Ct) or an
 *    8-wide node (cost A + the best split of 8 entries over the two children).
 *  - aq_ploc_*: parallel locally-ordered clustering (Meister & Bittner 2018) over the Morton order: every
 *    cluster looks R neighbours to each side for the partner with the smallest merged surface area, mutual
 *    choices merge, the cluster array is compacted, until one cluster is left.  On real geometry the tree
 *    costs a quarter fewer node visits per ray than the radix tree over the same order
 *    (profiles/r02c_ploc_prototype_cpu.log).
 */
#ifndef AQ_BVH_PLOC_H
#define AQ_BVH_PLOC_H

#include "aq_bvh_emit.h"

struct aq_dp8 {
    float c[8];     /* c[1..7] */
    uint32_t split; /* 3 bits per j = 2..8 (at bit 3*(j-2)): entries given to the left child */
    uint32_t flags; /* bits 2..7: c[i] == c[i-1] ("fewer"); bit 8: c[1] is the leaf alternative */
};

AQ_HD void aq_dp8_leaf(aq_dp8& D, float A, uint32_t count, float Ct) {
    for (int i = 0; i < 8; ++i) D.c[i] = A * (float)count * Ct;
    D.split = 0u;
    D.flags = 0xFCu | 0x100u;
}
AQ_HD void aq_dp8_inner(aq_dp8& D, const aq_dp8& L, const aq_dp8& R, float A, uint32_t count, float Ct) {
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
AQ_HD int aq_dp8_collect(aq_bvh2_node* N, const aq_dp8* dp, uint32_t root, uint32_t* ch) {
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

/* ---- PLOC.  Cluster i of the current array is BVH2 node cid[i] with box cbox[i] (6 floats: lo, hi). */
struct aq_box6 {
    float lo[3], hi[3];
};

/* the partner of cluster i: index j != i in [i-R, i+R] with the smallest half area of the merged box
 * (ties: the smaller j) */
AQ_HD uint32_t aq_ploc_nearest(const aq_box6* cbox, uint32_t m, uint32_t i, uint32_t R) {
    const aq_box6 a = cbox[i];
    const uint32_t j0 = i > R ? i - R : 0u, j1 = i + R < m ? i + R : m - 1u;
    float best = AQ_INF;
    uint32_t bj = i;
    for (uint32_t j = j0; j <= j1; ++j) {
        if (j == i) continue;
        const aq_box6 b = cbox[j];
        float lo[3], hi[3];
        for (int k = 0; k < 3; ++k) {
            lo[k] = fminf(a.lo[k], b.lo[k]);
            hi[k] = fmaxf(a.hi[k], b.hi[k]);
        }
        const float ar = aq_box_half_area(lo, hi);
        if (ar < best) {
            best = ar;
            bj = j;
        }
    }
    return bj;
}

/* what becomes of cluster i: 0 = it is the right half of a merge (dropped), 1 = kept as it is,
 * 2 = it is the left half of a merge (replaced by the new parent) */
AQ_HD int aq_ploc_fate(const uint32_t* nn, uint32_t i) {
    const uint32_t j = nn[i];
    if (j == i || nn[j] != i) return 1;
    return i < j ? 2 : 0;
}

/* parent node k of the BVH2 nodes a and b (boxes ba, bb), with its collapse table when dp != null */
AQ_HD void aq_ploc_make_parent(aq_bvh2_node* N, aq_dp8* dp, uint32_t k, uint32_t a, uint32_t b, const aq_box6& ba,
                               const aq_box6& bb, float Ct, aq_box6* out_box) {
    aq_bvh2_node P;
    for (int t = 0; t < 3; ++t) {
        P.lo[t] = fminf(ba.lo[t], bb.lo[t]);
        P.hi[t] = fmaxf(ba.hi[t], bb.hi[t]);
        out_box->lo[t] = P.lo[t];
        out_box->hi[t] = P.hi[t];
    }
    P.left = a;
    P.right = b;
    P.first = 0u; /* assigned top-down once the tree is complete (aq_ploc_assign_first) */
    P.count = N[a].count + N[b].count;
    N[k] = P;
    if (dp) {
        aq_dp8 D;
        aq_dp8_inner(D, dp[a], dp[b], aq_box_half_area(P.lo, P.hi), P.count, Ct);
        dp[k] = D;
    }
}

/* top-down: the leaves of every subtree get a contiguous range of the NEW primitive order (the emit code
 * addresses leaf groups as [first, first + count)); run over the internal nodes parents first */
AQ_HD void aq_ploc_assign_first(aq_bvh2_node* N, uint32_t k) {
    const uint32_t l = N[k].left, r = N[k].right;
    N[l].first = N[k].first;
    N[r].first = N[k].first + N[l].count;
}

#endif /* AQ_BVH_PLOC_H */
