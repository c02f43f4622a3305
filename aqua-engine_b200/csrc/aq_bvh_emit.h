/*
 * aq_bvh_emit.h — BVH2 -> BVH8 node emission, shared by the host builder
 * (aq_bvh_build.cpp) and the device builder (aq_bvh_build_gpu.cu): greedy surface-area
 * collapse of a binary subtree into <= 8 children, octant-gain slot assignment, outward 8-bit
 * quantisation, triangle-record emission.  Layout: aq_bvh.h.  SURVEY §8 row a5.
 */
#ifndef AQ_BVH_EMIT_H
#define AQ_BVH_EMIT_H

#include "aq_bvh.h"

#define AQ_BVH2_LEAF 0xFFFFFFFFu
#define AQ_LEAF_MAX 3u

/* binary node; a leaf group (left == AQ_BVH2_LEAF) covers sorted primitives
 * [first, first+count), count <= AQ_LEAF_MAX.  Boxes are already padded. */
struct aq_bvh2_node {
    float lo[3], hi[3];
    uint32_t left, right;
    uint32_t first, count;
};

AQ_HD float aq_box_half_area(const float* lo, const float* hi) {
    float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    if (!(dx >= 0.0f)) return 0.0f;
    return dx * dy + dy * dz + dz * dx;
}

struct aq_node8_plan {
    uint32_t ch[8];          /* BVH2 node per slot, AQ_BVH2_LEAF = empty slot */
    uint32_t n_inner, n_tris;
    float lo[3], hi[3];      /* node box */
};

/* step 1a (greedy variant): choose the <= 8 children of the wide node rooted at BVH2 node `root`
 * by repeatedly opening the inner child with the largest surface area */
AQ_HD int aq_node8_collect_greedy(const aq_bvh2_node* N, uint32_t root, uint32_t* ch) {
    int nc = 0;
    if (N[root].left == AQ_BVH2_LEAF) {
        ch[nc++] = root; /* the whole tree is one leaf group */
        return nc;
    }
    ch[nc++] = N[root].left;
    ch[nc++] = N[root].right;
    while (nc < 8) {
        int best = -1;
        float ba = -1.0f;
        for (int i = 0; i < nc; ++i) {
            const aq_bvh2_node& C = N[ch[i]];
            if (C.left != AQ_BVH2_LEAF) {
                float a = aq_box_half_area(C.lo, C.hi);
                if (a > ba) {
                    ba = a;
                    best = i;
                }
            }
        }
        if (best < 0) break;
        const aq_bvh2_node& C = N[ch[best]];
        ch[best] = C.left;
        ch[nc++] = C.right;
    }
    return nc;
}

/* step 1b: box, counts and octant slots for a chosen child list */
AQ_HD void aq_node8_plan_from(const aq_bvh2_node* N, const uint32_t* ch, int nc, aq_node8_plan* P) {
    for (int a = 0; a < 3; ++a) {
        P->lo[a] = AQ_INF;
        P->hi[a] = -AQ_INF;
    }
    P->n_inner = 0;
    P->n_tris = 0;
    for (int i = 0; i < nc; ++i) {
        const aq_bvh2_node& C = N[ch[i]];
        for (int a = 0; a < 3; ++a) {
            P->lo[a] = fminf(P->lo[a], C.lo[a]);
            P->hi[a] = fmaxf(P->hi[a], C.hi[a]);
        }
        if (C.left == AQ_BVH2_LEAF)
            P->n_tris += C.count;
        else
            P->n_inner++;
    }
    /* slot assignment: gain(c,s) = dot(centroid_c - centroid_node, sign_s), greedy max, so that
     * slot ^ flip orders children roughly front to back for every ray octant */
    int slot_of[8];
    for (int i = 0; i < 8; ++i) {
        slot_of[i] = -1;
        P->ch[i] = AQ_BVH2_LEAF;
    }
    float c3[8][3];
    for (int i = 0; i < nc; ++i) {
        const aq_bvh2_node& C = N[ch[i]];
        for (int a = 0; a < 3; ++a) c3[i][a] = 0.5f * (C.lo[a] + C.hi[a]) - 0.5f * (P->lo[a] + P->hi[a]);
    }
    uint32_t slot_used = 0;
    for (int k = 0; k < nc; ++k) {
        int bi = -1, bs = -1;
        float bg = -AQ_INF;
        for (int i = 0; i < nc; ++i) {
            if (slot_of[i] >= 0) continue;
            for (int s = 0; s < 8; ++s) {
                if (slot_used & (1u << s)) continue;
                float g = 0.0f;
                for (int a = 0; a < 3; ++a) g += ((s >> a) & 1) ? c3[i][a] : -c3[i][a];
                if (bi < 0 || g > bg) {
                    bg = g;
                    bi = i;
                    bs = s;
                }
            }
        }
        slot_of[bi] = bs;
        slot_used |= 1u << bs;
        P->ch[bs] = ch[bi];
    }
}

AQ_HD void aq_put_byte(uint32_t& w, int i, uint32_t v) { w |= (v & 0xFFu) << (8 * i); }

/* one 48-byte triangle record: v0, e1 = v1-v0, e2 = v2-v0 (plain float subtractions), prim id */
AQ_HD void aq_write_tri_record(aq_f4* rec, const float* positions, const uint32_t* indices, uint32_t prim) {
    const float* v0 = positions + 3 * (size_t)indices[3 * (size_t)prim + 0];
    const float* v1 = positions + 3 * (size_t)indices[3 * (size_t)prim + 1];
    const float* v2 = positions + 3 * (size_t)indices[3 * (size_t)prim + 2];
    rec[0].x = v0[0]; rec[0].y = v0[1]; rec[0].z = v0[2]; rec[0].w = v1[0] - v0[0];
    rec[1].x = v1[1] - v0[1]; rec[1].y = v1[2] - v0[2]; rec[1].z = v2[0] - v0[0]; rec[1].w = v2[1] - v0[1];
    rec[2].x = v2[2] - v0[2];
    rec[2].y = aq_u2f(prim);
    rec[2].z = 0.0f;
    rec[2].w = 0.0f;
}

/* step 2: with the output slots known (child_base = index of the first inner child's node,
 * tri_base = index of the first triangle record), write the 80-byte node and its triangle
 * records.  inner_out[k] receives the BVH2 node of the k-th inner child (node index
 * child_base + k).  `order` maps sorted position -> primitive id. */
AQ_HD void aq_node8_write(const aq_bvh2_node* N, const aq_node8_plan& P, const uint32_t* order,
                          const float* positions, const uint32_t* indices, uint32_t child_base,
                          uint32_t tri_base, aq_u4* node_out, aq_f4* tris_out, uint32_t* inner_out) {
    float p[3], sc[3];
    uint32_t eb[3];
    for (int a = 0; a < 3; ++a) {
        p[a] = P.lo[a];
        float ext = P.hi[a] - P.lo[a];
        int e = -100;
        if (ext > 0.0f) {
            e = (int)ceilf(log2f(ext / 255.0f)) - 1; /* the loop below fixes the estimate */
            if (e < -100) e = -100;
        }
        for (;;) { /* smallest power-of-two step whose 255th multiple still covers the node */
            float s = ldexpf(1.0f, e);
            float qh = ceilf((P.hi[a] - p[a]) / s);
            if (qh <= 255.0f && fmaf(255.0f, s, p[a]) >= P.hi[a]) break;
            ++e;
        }
        eb[a] = (uint32_t)(e + 127);
        sc[a] = ldexpf(1.0f, e);
    }
    aq_u4 w0, w1, w2, w3, w4;
    w0.x = aq_f2u(p[0]); w0.y = aq_f2u(p[1]); w0.z = aq_f2u(p[2]);
    w1.x = child_base; w1.y = tri_base; w1.z = 0u; w1.w = 0u;
    w2.x = w2.y = w2.z = w2.w = 0u;
    w3 = w2;
    w4 = w2;
    uint32_t imask = 0u, inner_i = 0u, tri_off = 0u;
    for (int s = 0; s < 8; ++s) {
        uint32_t c = P.ch[s];
        if (c == AQ_BVH2_LEAF) continue;
        const aq_bvh2_node& C = N[c];
        uint32_t meta;
        if (C.left != AQ_BVH2_LEAF) {
            imask |= 1u << s;
            meta = 0x20u | (24u + (uint32_t)s);
            inner_out[inner_i++] = c;
        } else {
            meta = (((1u << C.count) - 1u) << 5) | tri_off;
            for (uint32_t k = 0; k < C.count; ++k)
                aq_write_tri_record(tris_out + (size_t)(tri_base + tri_off + k) * AQ_TRI_WORDS, positions, indices,
                                    order[C.first + k]);
            tri_off += C.count;
        }
        aq_put_byte(s < 4 ? w1.z : w1.w, s & 3, meta);
        uint32_t q[6];
        for (int a = 0; a < 3; ++a) {
            /* outward rounding, verified in the same float expression a decoder would use */
            int ql = (int)floorf((C.lo[a] - p[a]) / sc[a]), qh = (int)ceilf((C.hi[a] - p[a]) / sc[a]);
            ql = ql < 0 ? 0 : (ql > 255 ? 255 : ql);
            qh = qh < 0 ? 0 : (qh > 255 ? 255 : qh);
            while (ql > 0 && fmaf((float)ql, sc[a], p[a]) > C.lo[a]) --ql;
            while (qh < 255 && fmaf((float)qh, sc[a], p[a]) < C.hi[a]) ++qh;
            q[a] = (uint32_t)ql;
            q[3 + a] = (uint32_t)qh;
        }
        /* n2 = qlo_x | qlo_y, n3 = qlo_z | qhi_x, n4 = qhi_y | qhi_z; two words per plane set */
        const int h = s >> 2, b = s & 3;
        aq_put_byte(h ? w2.y : w2.x, b, q[0]);
        aq_put_byte(h ? w2.w : w2.z, b, q[1]);
        aq_put_byte(h ? w3.y : w3.x, b, q[2]);
        aq_put_byte(h ? w3.w : w3.z, b, q[3]);
        aq_put_byte(h ? w4.y : w4.x, b, q[4]);
        aq_put_byte(h ? w4.w : w4.z, b, q[5]);
    }
    w0.w = eb[0] | (eb[1] << 8) | (eb[2] << 16) | (imask << 24);
    node_out[0] = w0;
    node_out[1] = w1;
    node_out[2] = w2;
    node_out[3] = w3;
    node_out[4] = w4;
}

/* step 1 (greedy collapse): collect + plan */
AQ_HD void aq_node8_plan_children(const aq_bvh2_node* N, uint32_t root, aq_node8_plan* P) {
    uint32_t ch[8];
    int nc = aq_node8_collect_greedy(N, root, ch);
    aq_node8_plan_from(N, ch, nc, P);
}

#endif /* AQ_BVH_EMIT_H */
