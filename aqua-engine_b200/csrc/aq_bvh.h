/*
 * aq_bvh.h — compressed 8-wide BVH (80-byte nodes, 48-byte triangle records) and the
 * per-ray traversal loop.
 *
 * Geometry source: TriangleMesh  scenes/ *.mesh; stage rows a5/a6/a7 of
 * SURVEY §8.  Node layout follows Ylitie, Karras, Laine, "Efficient Incoherent Ray
 * Traversal on GPUs Through Compressed Wide BVHs" (HPG 2017): child boxes are 8-bit
 * offsets on a per-node power-of-two grid, rounded OUTWARD, so the box test can only
 * produce false positives, never false negatives.
 *
 *   word  bytes  content
 *   n0    16     origin p.xyz (f32) | ex | ey | ez | imask      (e* = biased f32 exponents)
 *   n1    16     child_base (u32) | tri_base (u32) | meta[0..7]
 *   n2    16     qlo_x[0..7] | qlo_y[0..7]
 *   n3    16     qlo_z[0..7] | qhi_x[0..7]
 *   n4    16     qhi_y[0..7] | qhi_z[0..7]
 *
 *   meta[i] = 0                      empty slot
 *           = 0x20 | (24 + i)        inner child in slot i; node index =
 *                                    child_base + popc(imask & ((1<<i)-1))
 *           = (unary n) << 5 | off   leaf child: n in 1..3 triangles (0b001,0b011,0b111)
 *                                    at records tri_base + off .. +n-1   (off < 24)
 *
 *   triangle record (3 x 16 B): v0.xyz e1.x | e1.yz e2.xy | e2.z prim pad pad
 *
 * The traversal result is independent of visiting order: closest hit = lexicographic min
 * of (t, prim) (aq_hit_closer), nodes are culled with tmin_box <= t_best (inclusive).
 */
#ifndef AQ_BVH_H
#define AQ_BVH_H

#include "aq_core.h"

#define AQ_NODE_WORDS 5 /* 16-byte words per node */
#define AQ_TRI_WORDS 3
#define AQ_STACK_MAX 64 /* builder guarantees depth < AQ_STACK_MAX */
/* entries a traversal stack must hold: one node group per level, plus (interleaved step,
 * aq_trav_step2) one postponed triangle group per level */
#define AQ_STACK_CAP (2 * AQ_STACK_MAX)

#if defined(__CUDA_ARCH__)
#define AQ_LDG_U4(p) aq_ldg_u4(p)
#define AQ_LDG_F4(p) aq_ldg_f4(p)
__device__ __forceinline__ aq_u4 aq_ldg_u4(const aq_u4* p) {
    uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    aq_u4 r;
    r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w;
    return r;
}
__device__ __forceinline__ aq_f4 aq_ldg_f4(const aq_f4* p) {
    float4 v = __ldg(reinterpret_cast<const float4*>(p));
    aq_f4 r;
    r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w;
    return r;
}
AQ_HD uint32_t aq_msb(uint32_t x) { return 31u - (uint32_t)__clz((int)x); }
AQ_HD uint32_t aq_popc(uint32_t x) { return (uint32_t)__popc(x); }
AQ_HD float aq_u2f(uint32_t x) { return __uint_as_float(x); }
AQ_HD uint32_t aq_f2u(float x) { return __float_as_uint(x); }
#else
#define AQ_LDG_U4(p) (*(p))
#define AQ_LDG_F4(p) (*(p))
AQ_HD uint32_t aq_msb(uint32_t x) { return 31u - (uint32_t)__builtin_clz(x); }
AQ_HD uint32_t aq_popc(uint32_t x) { return (uint32_t)__builtin_popcount(x); }
AQ_HD float aq_u2f(uint32_t x) {
    union { uint32_t u; float f; } c;
    c.u = x;
    return c.f;
}
AQ_HD uint32_t aq_f2u(float x) {
    union { uint32_t u; float f; } c;
    c.f = x;
    return c.u;
}
#endif

/* simple array stack used by the host instantiation and as the reference policy */
struct aq_local_stack {
    uint32_t sx[AQ_STACK_CAP], sy[AQ_STACK_CAP];
    int n;
    AQ_HD void reset() { n = 0; }
    AQ_HD bool empty() const { return n == 0; }
    AQ_HD uint32_t top_y() const { return sy[n - 1]; }
    AQ_HD void push(uint32_t x, uint32_t y) {
        sx[n] = x;
        sy[n] = y;
        ++n;
    }
    AQ_HD void pop(uint32_t& x, uint32_t& y) {
        --n;
        x = sx[n];
        y = sy[n];
    }
};

struct aq_trav_counters {
    uint32_t nodes, tris;
};

AQ_HD float aq_safe_rcp_dir(float d) {
    /* avoid 0*inf in the slab test: |d| < 1e-20 is treated as +-1e-20 */
    float a = fabsf(d) > 1.0e-20f ? d : (d < 0.0f ? -1.0e-20f : 1.0e-20f);
#if defined(__CUDA_ARCH__)
    /* the slab test is conservative (padded boxes): a 1-ulp reciprocal is good enough and is
     * one MUFU instead of the IEEE division sequence (the triangle test keeps IEEE division).
     * 1e-20 <= |a| <= ~1, so neither the operand nor the result is subnormal: the bare rcp.approx.ftz
     * (what __fdividef(1, a) multiplies by 1 after six instructions of range handling) is the same value */
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
#else
    return 1.0f / a;
#endif
}

/* byte i of w as a float (I2F.U8 with a byte selector on the device; the PRMT + 2^23 magic
 * alternative measured 3 % slower on B200) */
AQ_HD float aq_byte_f(uint32_t w, int i) { return (float)((w >> (8 * i)) & 0xFFu); }

/* byte i of w, zero-extended (one PRMT on the device) */
AQ_HD uint32_t aq_byte_u(uint32_t w, int i) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(w, 0u, 0x4440u + (uint32_t)i);
#else
    return (w >> (8 * i)) & 0xFFu;
#endif
}
/* b << (s mod 32): the funnel shift's wrap mode ignores the upper bits of s */
AQ_HD uint32_t aq_shl_wrap(uint32_t b, uint32_t s) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(0u, b, s);
#else
    return b << (s & 31u);
#endif
}

/* test the 4 children held in one 32-bit lane group; returns bits into hitmask.
 * The meta bytes are decoded four at a time (round 2: per child the hit-mask update was 8 ALU
 * instructions — extract, classify, permute, shift, or — now 2 PRMT + SHF + LOP3):
 *   inner4  0x01 in the bytes of inner children          (meta & 0x18) == 0x18
 *   sh4     per byte, low 5 bits = target bit in the hit mask: 24 + (slot ^ flip) for an inner
 *           child (slot = meta & 7), the triangle offset for a leaf child
 *   bits4   per byte, the 3-bit field that goes there: 1 / the unary triangle count / 0 (empty) */
AQ_HD uint32_t aq_node_half(uint32_t meta4, uint32_t nx, uint32_t ny, uint32_t nz, uint32_t fx,
                            uint32_t fy, uint32_t fz, aq_v3 adj, aq_v3 org, float tmin, float tmax,
                            uint32_t flip4) {
    const uint32_t inner4 = (meta4 >> 3) & (meta4 >> 4) & 0x01010101u;
    const uint32_t sh4 = meta4 ^ (flip4 & (inner4 * 7u));
    const uint32_t bits4 = (meta4 >> 5) & 0x07070707u;
    uint32_t hm = 0u;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float tnx = fmaf(aq_byte_f(nx, i), adj.x, org.x);
        float tny = fmaf(aq_byte_f(ny, i), adj.y, org.y);
        float tnz = fmaf(aq_byte_f(nz, i), adj.z, org.z);
        float tfx = fmaf(aq_byte_f(fx, i), adj.x, org.x);
        float tfy = fmaf(aq_byte_f(fy, i), adj.y, org.y);
        float tfz = fmaf(aq_byte_f(fz, i), adj.z, org.z);
        float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tmin));
        float tf = fminf(fminf(tfx, tfy), fminf(tfz, tmax));
        const uint32_t c = aq_shl_wrap(aq_byte_u(bits4, i), aq_byte_u(sh4, i));
        hm |= (tn <= tf) ? c : 0u;
    }
    return hm;
}

/*
 * Per-ray traversal state.  The walk is written as init + step so that the GPU kernel can
 * interleave work fetching with traversal (a lane whose ray is finished takes a new ray while
 * its neighbours continue), and the host instantiation simply loops step() to completion.
 * One step = open one child node (5 x 16 B), test its 8 children, test the hit triangles.
 */
struct aq_trav {
    aq_v3 o, d, idir;
    float tmin, tmax;
    uint32_t flip;
    uint32_t ng_x, ng_y; /* current node group: child base index | hit bits (31..24) + imask */
    uint32_t tg_x, tg_y; /* aq_trav_step2 only: pending triangle group: record base | bits (23..0) */
    uint32_t best_prim;
    float best_t, bu, bv;
};

template <class Stack>
AQ_HD void aq_trav_init(aq_trav& T, aq_v3 o, aq_v3 d, float tmin, float tmax, Stack& st) {
    T.o = o;
    T.d = d;
    T.idir = aq_mk(aq_safe_rcp_dir(d.x), aq_safe_rcp_dir(d.y), aq_safe_rcp_dir(d.z));
    /* flip bit i set <=> direction component i is non-negative (near side = low side) */
    T.flip = (d.x >= 0.0f ? 1u : 0u) | (d.y >= 0.0f ? 2u : 0u) | (d.z >= 0.0f ? 4u : 0u);
    T.tmin = tmin;
    T.tmax = tmax;
    T.best_prim = AQ_MISS_ID;
    T.best_t = tmax;
    T.bu = 0.0f;
    T.bv = 0.0f;
    T.ng_x = 0u;
    T.ng_y = 0x80000000u; /* root: base 0, one pending hit, imask 0 */
    T.tg_x = 0u;
    T.tg_y = 0u;
    st.reset();
}

/*
 * One traversal step.  Returns true when the ray is finished.
 * Closest (ANY=false): best_prim/best_t/bu/bv hold the (t,prim)-minimal hit with
 * tmin < t <= tmax, or best_prim == AQ_MISS_ID.  Any (ANY=true): finishes with best_prim = 0
 * as soon as a triangle with tmin < t < tmax is found.
 */
/* open the next child of the current node group (T.ng): fetch its node, test the 8 child boxes
 * against [tmin, best_t]; the hit inner children become the new node group (the rest of the old
 * one goes to the stack), the hit leaf triangles are returned as (tg_x, tg_y) */
template <bool COUNT, class Stack>
AQ_HD void aq_trav_open_node(const aq_u4* __restrict__ nodes, aq_trav& T, Stack& st, aq_trav_counters* cnt,
                             uint32_t& tg_x, uint32_t& tg_y) {
    uint32_t b = aq_msb(T.ng_y);
    uint32_t imask = T.ng_y & 0xFFu;
    uint32_t base = T.ng_x;
    T.ng_y &= ~(1u << b);
    if (T.ng_y > 0x00FFFFFFu) st.push(T.ng_x, T.ng_y);
    uint32_t slot = (b - 24u) ^ T.flip;
    uint32_t rel = aq_popc(imask & ((1u << slot) - 1u));
    const aq_u4* np = nodes + (size_t)(base + rel) * AQ_NODE_WORDS;
    aq_u4 n0 = AQ_LDG_U4(np + 0), n1 = AQ_LDG_U4(np + 1), n2 = AQ_LDG_U4(np + 2),
          n3 = AQ_LDG_U4(np + 3), n4 = AQ_LDG_U4(np + 4);
    if (COUNT) cnt->nodes++;
    aq_v3 adj = aq_mk(aq_u2f((n0.w & 0xFFu) << 23) * T.idir.x,
                      aq_u2f(((n0.w >> 8) & 0xFFu) << 23) * T.idir.y,
                      aq_u2f(((n0.w >> 16) & 0xFFu) << 23) * T.idir.z);
    aq_v3 org = aq_mk((aq_u2f(n0.x) - T.o.x) * T.idir.x, (aq_u2f(n0.y) - T.o.y) * T.idir.y,
                      (aq_u2f(n0.z) - T.o.z) * T.idir.z);
    /* near/far plane words per axis */
    bool px = (T.flip & 1u) != 0u, py = (T.flip & 2u) != 0u, pz = (T.flip & 4u) != 0u;
    uint32_t nxl = px ? n2.x : n3.z, nxh = px ? n2.y : n3.w; /* qlo_x : qhi_x */
    uint32_t fxl = px ? n3.z : n2.x, fxh = px ? n3.w : n2.y;
    uint32_t nyl = py ? n2.z : n4.x, nyh = py ? n2.w : n4.y; /* qlo_y : qhi_y */
    uint32_t fyl = py ? n4.x : n2.z, fyh = py ? n4.y : n2.w;
    uint32_t nzl = pz ? n3.x : n4.z, nzh = pz ? n3.y : n4.w; /* qlo_z : qhi_z */
    uint32_t fzl = pz ? n4.z : n3.x, fzh = pz ? n4.w : n3.y;
    const uint32_t flip4 = T.flip * 0x01010101u;
    uint32_t hm = aq_node_half(n1.z, nxl, nyl, nzl, fxl, fyl, fzl, adj, org, T.tmin, T.best_t, flip4) |
                  aq_node_half(n1.w, nxh, nyh, nzh, fxh, fyh, fzh, adj, org, T.tmin, T.best_t, flip4);
    T.ng_x = n1.x;
    T.ng_y = (hm & 0xFF000000u) | (n0.w >> 24);
    tg_x = n1.y;
    tg_y = hm & 0x00FFFFFFu;
}

/* test triangle record `rec`; returns true when the ray is finished (any-hit found) */
template <bool ANY, bool COUNT>
AQ_HD bool aq_trav_test_tri(const aq_f4* __restrict__ tris, aq_trav& T, uint32_t rec, aq_trav_counters* cnt) {
    const aq_f4* tp = tris + (size_t)rec * AQ_TRI_WORDS;
    aq_f4 t0 = AQ_LDG_F4(tp + 0), t1 = AQ_LDG_F4(tp + 1), t2 = AQ_LDG_F4(tp + 2);
    if (COUNT) cnt->tris++;
    float t, u, v;
    if (aq_tri_test(T.o, T.d, T.tmin, aq_mk(t0.x, t0.y, t0.z), aq_mk(t0.w, t1.x, t1.y),
                    aq_mk(t1.z, t1.w, t2.x), &t, &u, &v)) {
        uint32_t prim = aq_f2u(t2.y);
        if (ANY) {
            if (t < T.tmax) {
                T.best_prim = 0u;
                return true;
            }
        } else if (t <= T.tmax && aq_hit_closer(t, prim, T.best_t, T.best_prim)) {
            T.best_t = t;
            T.best_prim = prim;
            T.bu = u;
            T.bv = v;
        }
    }
    return false;
}

template <bool ANY, bool COUNT, class Stack>
AQ_HD bool aq_trav_step(const aq_u4* __restrict__ nodes, const aq_f4* __restrict__ tris, aq_trav& T,
                        Stack& st, aq_trav_counters* cnt) {
    uint32_t tg_x, tg_y;
    /* ---- pop one child of the current node group and open it */
    aq_trav_open_node<COUNT>(nodes, T, st, cnt, tg_x, tg_y);

    /* ---- triangles of this node */
    while (tg_y) {
        uint32_t i = aq_msb(tg_y);
        tg_y &= ~(1u << i);
        if (aq_trav_test_tri<ANY, COUNT>(tris, T, tg_x + i, cnt)) return true;
    }

    /* ---- next node group */
    if (T.ng_y <= 0x00FFFFFFu) {
        if (st.empty()) return true;
        st.pop(T.ng_x, T.ng_y);
    }
    return false;
}

/*
 * Interleaved step: one node visit AND at most AQ_TRIS_PER_STEP triangle tests per call, for the
 * same ray.  In aq_trav_step a lane that found k leaf triangles runs k test iterations while the
 * lanes of its warp that found none wait (ncu, room.json: the triangle loop is 30 % of the
 * traversal kernel's warp instructions at 9 of 32 threads).  Here the triangles a node visit
 * produced are tested one per step, next to the following node visits of the same ray, so the
 * triangle phase of a step runs once for every lane that has any triangle pending.  A triangle
 * group that arrives while another is still pending goes to the traversal stack (entries with
 * y <= 0x00FFFFFF are triangle groups).  best_t is updated up to a few steps later than in
 * aq_trav_step, so slightly more nodes are opened; the result is the same (t, prim) minimum.
 */
#ifndef AQ_TRIS_PER_STEP
#define AQ_TRIS_PER_STEP 1
#endif
template <bool ANY, bool COUNT, class Stack>
AQ_HD bool aq_trav_step2(const aq_u4* __restrict__ nodes, const aq_f4* __restrict__ tris, aq_trav& T,
                         Stack& st, aq_trav_counters* cnt) {
    /* ---- refill the two work registers from the stack */
    if (T.tg_y == 0u && !st.empty() && st.top_y() <= 0x00FFFFFFu) st.pop(T.tg_x, T.tg_y);
    if (T.ng_y <= 0x00FFFFFFu && !st.empty() && st.top_y() > 0x00FFFFFFu) st.pop(T.ng_x, T.ng_y);
    /* ---- node phase */
    if (T.ng_y > 0x00FFFFFFu) {
        uint32_t nx, ny;
        aq_trav_open_node<COUNT>(nodes, T, st, cnt, nx, ny);
        if (ny) {
            if (T.tg_y) {
                st.push(nx, ny);
            } else {
                T.tg_x = nx;
                T.tg_y = ny;
            }
        }
    }
    /* ---- triangle phase */
#pragma unroll
    for (int k = 0; k < AQ_TRIS_PER_STEP; ++k) {
        if (T.tg_y) {
            uint32_t i = aq_msb(T.tg_y);
            T.tg_y &= ~(1u << i);
            if (aq_trav_test_tri<ANY, COUNT>(tris, T, T.tg_x + i, cnt)) return true;
        }
    }
    return T.ng_y <= 0x00FFFFFFu && T.tg_y == 0u && st.empty();
}

/* which step the product uses (A/B switch; both give the same hits) */
#if defined(AQ_TRAV_INTERLEAVED) && AQ_TRAV_INTERLEAVED
#define AQ_TRAV_STEP2 true
#else
#define AQ_TRAV_STEP2 false
#endif

/* whole-ray convenience wrapper (host walk, simple kernels) */
template <bool ANY, bool COUNT, bool STEP2 = AQ_TRAV_STEP2, class Stack>
AQ_HD bool aq_bvh8_trace(const aq_u4* __restrict__ nodes, const aq_f4* __restrict__ tris, aq_v3 o,
                         aq_v3 d, float tmin, float tmax, Stack& st, uint32_t& best_prim,
                         float& best_t, float& bu, float& bv, aq_trav_counters* cnt) {
    aq_trav T;
    aq_trav_init(T, o, d, tmin, tmax, st);
    if (STEP2) {
        while (!aq_trav_step2<ANY, COUNT>(nodes, tris, T, st, cnt)) {
        }
    } else {
        while (!aq_trav_step<ANY, COUNT>(nodes, tris, T, st, cnt)) {
        }
    }
    best_prim = T.best_prim;
    best_t = T.best_t;
    bu = T.bu;
    bv = T.bv;
    return best_prim != AQ_MISS_ID;
}

#endif /* AQ_BVH_H */
