/*
 * aq_bvh_build.cpp — host-side acceleration-structure construction:
 *   triangles -> binned-SAH BVH2 (multi-threaded) -> greedy collapse to 8-wide ->
 *   octant-aware slot assignment -> outward-rounded 8-bit quantisation into the 80 B node /
 *   48 B triangle-record layout of aq_bvh.h.
 *
 * Input geometry: TriangleMesh{vertices,indices}  scenes/ *.mesh flattened by
 * the host into one indexed list.  The reference only hints at its builder through the
 * `ordered-float` dependency (Cargo.toml:19); there is no reference builder to follow.
 *
 * Conservativeness contract (needed for bit-exact hit ids vs a brute-force loop): every
 * triangle box is padded by AQ_BOX_PAD * scene_scale before building and child boxes are
 * rounded outward onto the node grid, so a ray that passes aq_tri_test() for a triangle
 * always passes the slab test of all of that triangle's ancestors.
 */
#include "aq_bvh_build.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

namespace {

constexpr int kBins = 16;
constexpr float kBoxPad = 1.0e-5f;
constexpr uint32_t kLeafMax = 3;

struct Box {
    float lo[3], hi[3];
    void reset() {
        for (int a = 0; a < 3; ++a) {
            lo[a] = INFINITY;
            hi[a] = -INFINITY;
        }
    }
    void grow(const Box& b) {
        for (int a = 0; a < 3; ++a) {
            lo[a] = std::min(lo[a], b.lo[a]);
            hi[a] = std::max(hi[a], b.hi[a]);
        }
    }
    void grow(const float* p) {
        for (int a = 0; a < 3; ++a) {
            lo[a] = std::min(lo[a], p[a]);
            hi[a] = std::max(hi[a], p[a]);
        }
    }
    float half_area() const {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (!(dx >= 0.f)) return 0.f;
        return dx * dy + dy * dz + dz * dx;
    }
};

struct Node2 {
    Box box;
    uint32_t left, right; /* children (inner) */
    uint32_t first, count; /* leaf: range in prim index array; count==0 => inner */
};

struct Builder {
    const float* pos;
    const uint32_t* idx;
    uint32_t n;
    std::vector<Box> pbox;
    std::vector<float> cent; /* 3 per prim */
    std::vector<uint32_t> order;
    std::vector<Node2> nodes;
    std::atomic<uint32_t> n_nodes{0};
    std::atomic<int> free_threads{0};

    uint32_t alloc() { return n_nodes.fetch_add(1); }

    void build_range(uint32_t node, uint32_t b, uint32_t e, int depth) {
        Node2& N = nodes[node];
        Box bb, cb;
        bb.reset();
        cb.reset();
        for (uint32_t i = b; i < e; ++i) {
            uint32_t p = order[i];
            bb.grow(pbox[p]);
            cb.grow(&cent[3 * (size_t)p]);
        }
        N.box = bb;
        uint32_t cnt = e - b;
        if (cnt == 1) {
            N.first = b;
            N.count = 1;
            return;
        }
        /* binned SAH over the centroid bounds */
        float best_cost = INFINITY;
        int best_axis = -1, best_bin = -1;
        for (int a = 0; a < 3; ++a) {
            float lo = cb.lo[a], ext = cb.hi[a] - cb.lo[a];
            if (!(ext > 0.f)) continue;
            float k = (float)kBins / ext;
            Box bbx[kBins];
            uint32_t bc[kBins];
            for (int i = 0; i < kBins; ++i) {
                bbx[i].reset();
                bc[i] = 0;
            }
            for (uint32_t i = b; i < e; ++i) {
                uint32_t p = order[i];
                int bi = (int)((cent[3 * (size_t)p + a] - lo) * k);
                bi = bi < 0 ? 0 : (bi >= kBins ? kBins - 1 : bi);
                bbx[bi].grow(pbox[p]);
                bc[bi]++;
            }
            float ra[kBins];
            uint32_t rc[kBins];
            Box acc;
            acc.reset();
            uint32_t c = 0;
            for (int i = kBins - 1; i > 0; --i) {
                acc.grow(bbx[i]);
                c += bc[i];
                ra[i] = acc.half_area();
                rc[i] = c;
            }
            acc.reset();
            c = 0;
            for (int i = 0; i < kBins - 1; ++i) {
                acc.grow(bbx[i]);
                c += bc[i];
                if (c == 0 || rc[i + 1] == 0) continue;
                float cost = acc.half_area() * (float)c + ra[i + 1] * (float)rc[i + 1];
                if (cost < best_cost) {
                    best_cost = cost;
                    best_axis = a;
                    best_bin = i;
                }
            }
        }
        float leaf_cost = bb.half_area() * (float)cnt;
        if (cnt <= kLeafMax && (best_axis < 0 || leaf_cost <= best_cost + 0.5f * bb.half_area())) {
            N.first = b;
            N.count = cnt;
            return;
        }
        uint32_t mid;
        if (best_axis >= 0) {
            float lo = cb.lo[best_axis], k = (float)kBins / (cb.hi[best_axis] - cb.lo[best_axis]);
            int a = best_axis, bbn = best_bin;
            auto it = std::partition(order.begin() + b, order.begin() + e, [&](uint32_t p) {
                int bi = (int)((cent[3 * (size_t)p + a] - lo) * k);
                bi = bi < 0 ? 0 : (bi >= kBins ? kBins - 1 : bi);
                return bi <= bbn;
            });
            mid = (uint32_t)(it - order.begin());
        } else {
            mid = b + cnt / 2; /* all centroids coincide */
        }
        if (mid == b || mid == e) mid = b + cnt / 2;
        uint32_t l = alloc(), r = alloc();
        N.left = l;
        N.right = r;
        N.count = 0;
        bool spawn = false;
        if (cnt > 32768 && free_threads.load(std::memory_order_relaxed) > 0) {
            if (free_threads.fetch_sub(1) > 0)
                spawn = true;
            else
                free_threads.fetch_add(1);
        }
        if (spawn) {
            std::thread t([=]() {
                build_range(l, b, mid, depth + 1);
                free_threads.fetch_add(1);
            });
            build_range(r, mid, e, depth + 1);
            t.join();
        } else {
            build_range(l, b, mid, depth + 1);
            build_range(r, mid, e, depth + 1);
        }
    }
};

inline void put_byte(uint32_t& w, int i, uint32_t v) { w |= (v & 0xFFu) << (8 * i); }

}  // namespace

int aq_build_bvh8(const float* positions, const uint32_t* indices, uint32_t n_tris, int n_threads,
                  aq_bvh8* out) {
    out->nodes.clear();
    out->tris.clear();
    out->max_depth = 0;
    out->sah_cost = 0.f;
    out->n_bvh2_nodes = 0;

    if (n_threads <= 0) {
        n_threads = (int)std::thread::hardware_concurrency();
        if (n_threads <= 0) n_threads = 1;
    }

    if (n_tris == 0) { /* one empty node so traversal has a root */
        out->nodes.resize(AQ_NODE_WORDS);
        std::memset(out->nodes.data(), 0, sizeof(aq_u4) * AQ_NODE_WORDS);
        out->nodes[0].w = 127u | (127u << 8) | (127u << 16);
        return 0;
    }

    Builder B;
    B.pos = positions;
    B.idx = indices;
    B.n = n_tris;
    B.pbox.resize(n_tris);
    B.cent.resize(3 * (size_t)n_tris);
    B.order.resize(n_tris);
    /* scene scale for the conservative padding */
    float scale = 0.f;
    {
        Box sb;
        sb.reset();
        for (uint32_t t = 0; t < n_tris; ++t)
            for (int k = 0; k < 3; ++k) sb.grow(positions + 3 * (size_t)indices[3 * (size_t)t + k]);
        for (int a = 0; a < 3; ++a) {
            scale = std::max(scale, sb.hi[a] - sb.lo[a]);
            scale = std::max(scale, std::fabs(sb.lo[a]));
            scale = std::max(scale, std::fabs(sb.hi[a]));
        }
    }
    const float pad = kBoxPad * scale;
    for (uint32_t t = 0; t < n_tris; ++t) {
        Box b;
        b.reset();
        for (int k = 0; k < 3; ++k) b.grow(positions + 3 * (size_t)indices[3 * (size_t)t + k]);
        for (int a = 0; a < 3; ++a) {
            B.cent[3 * (size_t)t + a] = 0.5f * (b.lo[a] + b.hi[a]);
            b.lo[a] -= pad;
            b.hi[a] += pad;
        }
        B.pbox[t] = b;
        B.order[t] = t;
    }
    B.nodes.resize(2 * (size_t)n_tris);
    B.free_threads = n_threads - 1;
    uint32_t root = B.alloc();
    B.build_range(root, 0, n_tris, 0);
    out->n_bvh2_nodes = B.n_nodes.load();

    /* ---- collapse + emit, breadth first */
    struct Item {
        uint32_t n2;    /* BVH2 node */
        uint32_t out;   /* BVH8 node index */
        uint32_t depth;
    };
    std::vector<Item> queue;
    queue.reserve(n_tris / 2 + 16);
    out->nodes.reserve((size_t)(n_tris / 3 + 16) * AQ_NODE_WORDS);
    out->tris.resize((size_t)n_tris * AQ_TRI_WORDS);
    size_t tri_cursor = 0;
    out->nodes.resize(AQ_NODE_WORDS);
    queue.push_back({root, 0u, 1u});
    double sah = 0.0;
    const float root_area = std::max(B.nodes[root].box.half_area(), 1e-30f);

    for (size_t qi = 0; qi < queue.size(); ++qi) {
        Item it = queue[qi];
        out->max_depth = std::max(out->max_depth, it.depth);
        const Node2& P = B.nodes[it.n2];
        uint32_t ch[8];
        int nc = 0;
        if (P.count > 0) {
            ch[nc++] = it.n2; /* root is itself a leaf */
        } else {
            ch[nc++] = P.left;
            ch[nc++] = P.right;
            while (nc < 8) {
                int best = -1;
                float ba = -1.f;
                for (int i = 0; i < nc; ++i) {
                    const Node2& C = B.nodes[ch[i]];
                    if (C.count == 0) {
                        float a = C.box.half_area();
                        if (a > ba) {
                            ba = a;
                            best = i;
                        }
                    }
                }
                if (best < 0) break;
                const Node2& C = B.nodes[ch[best]];
                ch[best] = C.left;
                ch[nc++] = C.right;
            }
        }
        /* node box */
        Box nb;
        nb.reset();
        for (int i = 0; i < nc; ++i) nb.grow(B.nodes[ch[i]].box);
        sah += (double)nb.half_area() / root_area;
        /* slot assignment: gain(c,s) = dot(centroid_c - centroid_node, sign_s), greedy max */
        int slot_of[8], child_in_slot[8];
        for (int i = 0; i < 8; ++i) {
            slot_of[i] = -1;
            child_in_slot[i] = -1;
        }
        float gain[8][8];
        float nc3[3] = {0.5f * (nb.lo[0] + nb.hi[0]), 0.5f * (nb.lo[1] + nb.hi[1]),
                        0.5f * (nb.lo[2] + nb.hi[2])};
        for (int i = 0; i < nc; ++i) {
            const Box& cb = B.nodes[ch[i]].box;
            float c3[3];
            for (int a = 0; a < 3; ++a) c3[a] = 0.5f * (cb.lo[a] + cb.hi[a]) - nc3[a];
            for (int s = 0; s < 8; ++s) {
                float g = 0.f;
                for (int a = 0; a < 3; ++a) g += ((s >> a) & 1) ? c3[a] : -c3[a];
                gain[i][s] = g;
            }
        }
        for (int k = 0; k < nc; ++k) {
            int bi = -1, bs = -1;
            float bg = -INFINITY;
            for (int i = 0; i < nc; ++i) {
                if (slot_of[i] >= 0) continue;
                for (int s = 0; s < 8; ++s) {
                    if (child_in_slot[s] >= 0) continue;
                    if (gain[i][s] > bg) {
                        bg = gain[i][s];
                        bi = i;
                        bs = s;
                    }
                }
            }
            slot_of[bi] = bs;
            child_in_slot[bs] = bi;
        }
        /* quantisation grid */
        float p[3];
        uint32_t eb[3];
        float sc[3];
        for (int a = 0; a < 3; ++a) {
            p[a] = nb.lo[a];
            float ext = nb.hi[a] - nb.lo[a];
            int e = -100;
            if (ext > 0.f) {
                e = (int)std::ceil(std::log2((double)ext / 255.0));
                if (e < -100) e = -100;
            }
            /* make sure the largest offset fits in 8 bits in float arithmetic */
            for (;;) {
                float s = std::ldexp(1.0f, e);
                float qh = std::ceil((nb.hi[a] - p[a]) / s);
                if (qh <= 255.f && std::fmaf(255.f, s, p[a]) >= nb.hi[a]) break;
                ++e;
            }
            eb[a] = (uint32_t)(e + 127);
            sc[a] = std::ldexp(1.0f, e);
        }
        aq_u4 w0, w1, w2, w3, w4;
        std::memset(&w0, 0, sizeof w0);
        std::memset(&w1, 0, sizeof w1);
        std::memset(&w2, 0, sizeof w2);
        std::memset(&w3, 0, sizeof w3);
        std::memset(&w4, 0, sizeof w4);
        std::memcpy(&w0.x, &p[0], 4);
        std::memcpy(&w0.y, &p[1], 4);
        std::memcpy(&w0.z, &p[2], 4);
        uint32_t imask = 0;
        uint32_t n_inner = 0;
        for (int s = 0; s < 8; ++s)
            if (child_in_slot[s] >= 0 && B.nodes[ch[child_in_slot[s]]].count == 0) {
                imask |= 1u << s;
                ++n_inner;
            }
        w0.w = eb[0] | (eb[1] << 8) | (eb[2] << 16) | (imask << 24);
        uint32_t child_base = (uint32_t)(out->nodes.size() / AQ_NODE_WORDS);
        out->nodes.resize(out->nodes.size() + (size_t)n_inner * AQ_NODE_WORDS);
        uint32_t tri_base = (uint32_t)tri_cursor;
        w1.x = child_base;
        w1.y = tri_base;
        uint32_t inner_i = 0, tri_off = 0;
        uint32_t* qw[6] = {&w2.x, &w2.z, &w3.x, &w3.z, &w4.x, &w4.z}; /* lox loy loz hix hiy hiz */
        for (int s = 0; s < 8; ++s) {
            int ci = child_in_slot[s];
            if (ci < 0) continue;
            const Node2& C = B.nodes[ch[ci]];
            uint32_t meta;
            if (C.count == 0) {
                meta = 0x20u | (24u + (uint32_t)s);
                queue.push_back({ch[ci], child_base + inner_i, it.depth + 1});
                ++inner_i;
            } else {
                uint32_t unary = (1u << C.count) - 1u;
                meta = (unary << 5) | tri_off;
                for (uint32_t k = 0; k < C.count; ++k) {
                    uint32_t prim = B.order[C.first + k];
                    const float* v0 = positions + 3 * (size_t)indices[3 * (size_t)prim + 0];
                    const float* v1 = positions + 3 * (size_t)indices[3 * (size_t)prim + 1];
                    const float* v2 = positions + 3 * (size_t)indices[3 * (size_t)prim + 2];
                    aq_f4* rec = &out->tris[(tri_cursor + k) * AQ_TRI_WORDS];
                    float e1[3] = {v1[0] - v0[0], v1[1] - v0[1], v1[2] - v0[2]};
                    float e2[3] = {v2[0] - v0[0], v2[1] - v0[1], v2[2] - v0[2]};
                    rec[0].x = v0[0]; rec[0].y = v0[1]; rec[0].z = v0[2]; rec[0].w = e1[0];
                    rec[1].x = e1[1]; rec[1].y = e1[2]; rec[1].z = e2[0]; rec[1].w = e2[1];
                    rec[2].x = e2[2];
                    std::memcpy(&rec[2].y, &prim, 4);
                    rec[2].z = 0.f;
                    rec[2].w = 0.f;
                }
                tri_cursor += C.count;
                tri_off += C.count;
            }
            uint32_t* mw = s < 4 ? &w1.z : &w1.w;
            put_byte(*mw, s & 3, meta);
            for (int a = 0; a < 3; ++a) {
                double dl = ((double)C.box.lo[a] - (double)p[a]) / (double)sc[a];
                double dh = ((double)C.box.hi[a] - (double)p[a]) / (double)sc[a];
                int ql = (int)std::floor(dl), qh = (int)std::ceil(dh);
                ql = std::max(0, std::min(255, ql));
                qh = std::max(0, std::min(255, qh));
                while (ql > 0 && std::fmaf((float)ql, sc[a], p[a]) > C.box.lo[a]) --ql;
                while (qh < 255 && std::fmaf((float)qh, sc[a], p[a]) < C.box.hi[a]) ++qh;
                put_byte(qw[a][s >> 2], s & 3, (uint32_t)ql);
                put_byte(qw[3 + a][s >> 2], s & 3, (uint32_t)qh);
            }
        }
        aq_u4* dst = &out->nodes[(size_t)it.out * AQ_NODE_WORDS];
        dst[0] = w0;
        dst[1] = w1;
        dst[2] = w2;
        dst[3] = w3;
        dst[4] = w4;
    }
    out->sah_cost = (float)sah;
    if (out->max_depth >= AQ_STACK_MAX) return -1;
    return 0;
}
