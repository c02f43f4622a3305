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
#include "aq_bvh_emit.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace {

constexpr int kBins = 32; /* 16 -> 32 bins: 10-15 % fewer node visits per ray on room.json; 64/128 add nothing */
constexpr float kBoxPad = 1.0e-5f;
/* leaf size and the triangle/node cost ratio of the collapse: tunable for A/B runs
 * (AQUA_BVH_LEAF = 1..3, AQUA_BVH_CT = float); defaults below */
inline uint32_t leaf_max() {
    static const uint32_t v = [] {
        const char* e = std::getenv("AQUA_BVH_LEAF");
        int k = e ? std::atoi(e) : (int)AQ_LEAF_MAX;
        return (uint32_t)(k < 1 ? 1 : (k > (int)AQ_LEAF_MAX ? (int)AQ_LEAF_MAX : k));
    }();
    return v;
}
inline float tri_cost() {
    static const float v = [] {
        const char* e = std::getenv("AQUA_BVH_CT");
        float c = e ? (float)std::atof(e) : 0.3f;
        return c > 0.f ? c : 0.3f;
    }();
    return v;
}
#define kLeafMax leaf_max()

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

typedef aq_bvh2_node Node2;
inline void set_box(Node2& n, const Box& b) {
    for (int a = 0; a < 3; ++a) {
        n.lo[a] = b.lo[a];
        n.hi[a] = b.hi[a];
    }
}

/* one primitive of the build: everything a node's passes touch, kept IN the array that is partitioned, so that
 * binning and partitioning stream through memory instead of chasing an index permutation into two other arrays
 * (round 2: the build of room.json's 394k triangles was bound by those cache misses) */
struct Prim {
    Box box;     /* padded triangle box */
    float c[3];  /* centroid */
    uint32_t id; /* primitive id */
};

/* spin barrier for the few threads that share one big node (phases of microseconds: no futex round trip) */
struct SpinBarrier {
    std::atomic<int> count{0}, phase{0};
    int n = 1;
    void wait() {
        const int ph = phase.load(std::memory_order_acquire);
        if (count.fetch_add(1, std::memory_order_acq_rel) == n - 1) {
            count.store(0, std::memory_order_relaxed);
            phase.store(ph + 1, std::memory_order_release);
        } else {
            while (phase.load(std::memory_order_acquire) == ph) std::this_thread::yield();
        }
    }
};

constexpr uint32_t kParallelNode = 32768; /* nodes of at least this many primitives are binned and partitioned by several threads */

struct Builder {
    uint32_t n;
    std::vector<Prim> tmp; /* scatter target of the parallel (stable) partition */
    std::vector<Prim> prims;
    std::vector<Node2> nodes;
    std::atomic<uint32_t> n_nodes{0};
    std::atomic<int> free_threads{0};

    uint32_t alloc() { return n_nodes.fetch_add(1); }

    int take_threads(int want) {
        int got = 0;
        int f = free_threads.load(std::memory_order_relaxed);
        while (got < want && f > 0) {
            if (free_threads.compare_exchange_weak(f, f - 1)) ++got;
        }
        return got;
    }

    struct Bins {
        Box bbx[3][kBins];
        uint32_t bc[3][kBins];
        void reset() {
            for (int a = 0; a < 3; ++a)
                for (int i = 0; i < kBins; ++i) {
                    bbx[a][i].reset();
                    bc[a][i] = 0;
                }
        }
    };
    void bin_range(uint32_t b, uint32_t e, const bool* axis_ok, const float* bin_lo, const float* bin_k, Bins* B) const {
        for (uint32_t i = b; i < e; ++i) {
            const Prim& P = prims[i];
            for (int a = 0; a < 3; ++a) {
                if (!axis_ok[a]) continue;
                int bi = (int)((P.c[a] - bin_lo[a]) * bin_k[a]);
                bi = bi < 0 ? 0 : (bi >= kBins ? kBins - 1 : bi);
                B->bbx[a][bi].grow(P.box);
                B->bc[a][bi]++;
            }
        }
    }

    /* bounds of prims [b, e): boxes and centroids */
    void bounds(uint32_t b, uint32_t e, Box* bb, Box* cb) const {
        bb->reset();
        cb->reset();
        for (uint32_t i = b; i < e; ++i) {
            bb->grow(prims[i].box);
            cb->grow(prims[i].c);
        }
    }

    /* bb / cb: the bounds of the range, handed down by the parent's partition pass (one pass per node less) */
    void build_range(uint32_t node, uint32_t b, uint32_t e, int depth, const Box& bb, const Box& cb) {
        Node2& N = nodes[node];
        set_box(N, bb);
        uint32_t cnt = e - b;
        N.first = b;
        N.count = cnt;
        if (cnt == 1) {
            N.left = N.right = AQ_BVH2_LEAF;
            return;
        }
        /* binned SAH over the centroid bounds */
        float best_cost = INFINITY;
        int best_axis = -1, best_bin = -1;
        /* one pass over the primitives fills the bins of all three axes; a big node (the serial spine of an
         * unbalanced SAH tree: 30 % of room.json's build work sits in 26 nodes) is shared with idle threads */
        Bins bins;
        Box (&bbx)[3][kBins] = bins.bbx;
        uint32_t (&bc)[3][kBins] = bins.bc;
        float bin_lo[3], bin_k[3];
        bool axis_ok[3];
        for (int a = 0; a < 3; ++a) {
            const float ext = cb.hi[a] - cb.lo[a];
            axis_ok[a] = ext > 0.f;
            bin_lo[a] = cb.lo[a];
            bin_k[a] = axis_ok[a] ? (float)kBins / ext : 0.f;
        }
        bins.reset();
        const int helpers = cnt >= kParallelNode ? take_threads(7) : 0;
        if (helpers > 0) {
            const int T = helpers + 1;
            std::vector<Bins> part((size_t)helpers);
            std::vector<std::thread> th;
            for (int t = 1; t < T; ++t)
                th.emplace_back([&, t]() {
                    part[t - 1].reset();
                    bin_range(b + (uint32_t)((uint64_t)cnt * t / T), b + (uint32_t)((uint64_t)cnt * (t + 1) / T), axis_ok, bin_lo, bin_k, &part[t - 1]);
                });
            bin_range(b, b + (uint32_t)((uint64_t)cnt / T), axis_ok, bin_lo, bin_k, &bins);
            for (auto& x : th) x.join();
            for (const Bins& Pb : part) /* min / max / counts: exact, the result does not depend on the chunking */
                for (int a = 0; a < 3; ++a)
                    for (int i = 0; i < kBins; ++i) {
                        bbx[a][i].grow(Pb.bbx[a][i]);
                        bc[a][i] += Pb.bc[a][i];
                    }
        } else {
            bin_range(b, e, axis_ok, bin_lo, bin_k, &bins);
        }
        for (int a = 0; a < 3; ++a) {
            if (!axis_ok[a]) continue;
            float ra[kBins];
            uint32_t rc[kBins];
            Box acc;
            acc.reset();
            uint32_t c = 0;
            for (int i = kBins - 1; i > 0; --i) {
                acc.grow(bbx[a][i]);
                c += bc[a][i];
                ra[i] = acc.half_area();
                rc[i] = c;
            }
            acc.reset();
            c = 0;
            for (int i = 0; i < kBins - 1; ++i) {
                acc.grow(bbx[a][i]);
                c += bc[a][i];
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
            N.left = N.right = AQ_BVH2_LEAF;
            return;
        }
        uint32_t mid;
        Box lbb, lcb, rbb, rcb; /* children's bounds, gathered by the partition pass */
        lbb.reset();
        lcb.reset();
        rbb.reset();
        rcb.reset();
        bool have_child_bounds = false;
        if (best_axis >= 0 && cnt >= kParallelNode) {
            /* big node: stable partition, by several threads when some are idle (alone otherwise: the output must
             * not depend on who was free) — count and gather the children's bounds per chunk,
             * prefix, scatter into `tmp`, copy back.  (A stable partition orders the primitives inside the
             * children differently from the in-place one below; the split itself, and therefore every box and
             * every SAH decision further down, depends on the SETS only.) */
            const float lo = cb.lo[best_axis], k = (float)kBins / (cb.hi[best_axis] - cb.lo[best_axis]);
            const int a = best_axis, bbn = best_bin, T = helpers + 1;
            auto side = [&](const Prim& P) {
                int bi = (int)((P.c[a] - lo) * k);
                bi = bi < 0 ? 0 : (bi >= kBins ? kBins - 1 : bi);
                return bi <= bbn;
            };
            struct Part {
                uint32_t nl = 0, loff = 0, roff = 0;
                Box lbb, lcb, rbb, rcb;
            };
            std::vector<Part> pr((size_t)T);
            SpinBarrier bar;
            bar.n = T;
            uint32_t total_left = 0;
            auto work = [&](int t) {
                const uint32_t cb0 = b + (uint32_t)((uint64_t)cnt * t / T), ce0 = b + (uint32_t)((uint64_t)cnt * (t + 1) / T);
                Part& q = pr[(size_t)t];
                q.lbb.reset();
                q.lcb.reset();
                q.rbb.reset();
                q.rcb.reset();
                for (uint32_t i = cb0; i < ce0; ++i) {
                    const Prim& P = prims[i];
                    const bool left = side(P);
                    (left ? q.lbb : q.rbb).grow(P.box);
                    (left ? q.lcb : q.rcb).grow(P.c);
                    q.nl += left ? 1u : 0u;
                }
                bar.wait();
                if (t == 0) {
                    uint32_t l = 0, r = 0;
                    for (int u = 0; u < T; ++u) {
                        const uint32_t c0 = b + (uint32_t)((uint64_t)cnt * u / T), c1 = b + (uint32_t)((uint64_t)cnt * (u + 1) / T);
                        pr[(size_t)u].loff = l;
                        pr[(size_t)u].roff = r;
                        l += pr[(size_t)u].nl;
                        r += (c1 - c0) - pr[(size_t)u].nl;
                    }
                    total_left = l;
                }
                bar.wait();
                uint32_t lo_ = b + q.loff, ro_ = b + total_left + q.roff;
                for (uint32_t i = cb0; i < ce0; ++i) {
                    if (side(prims[i]))
                        tmp[lo_++] = prims[i];
                    else
                        tmp[ro_++] = prims[i];
                }
                bar.wait();
                for (uint32_t i = cb0; i < ce0; ++i) prims[i] = tmp[i];
            };
            std::vector<std::thread> th;
            for (int t = 1; t < T; ++t) th.emplace_back(work, t);
            work(0);
            for (auto& x : th) x.join();
            for (const Part& q : pr) {
                lbb.grow(q.lbb);
                lcb.grow(q.lcb);
                rbb.grow(q.rbb);
                rcb.grow(q.rcb);
            }
            mid = b + total_left;
            have_child_bounds = true;
        } else if (best_axis >= 0) {
            const float lo = cb.lo[best_axis], k = (float)kBins / (cb.hi[best_axis] - cb.lo[best_axis]);
            const int a = best_axis, bbn = best_bin;
            auto goes_left = [&](const Prim& P) { /* called exactly once per primitive */
                int bi = (int)((P.c[a] - lo) * k);
                bi = bi < 0 ? 0 : (bi >= kBins ? kBins - 1 : bi);
                const bool left = bi <= bbn;
                (left ? lbb : rbb).grow(P.box);
                (left ? lcb : rcb).grow(P.c);
                return left;
            };
            /* the bidirectional partition of libstdc++ written out (same permutation as std::partition over the
             * index array this replaced: the tree, and with it every output byte, is unchanged) */
            Prim *first = prims.data() + b, *last = prims.data() + e;
            for (;;) {
                for (;;) {
                    if (first == last) goto done;
                    if (goes_left(*first)) ++first; else break;
                }
                --last;
                for (;;) {
                    if (first == last) goto done;
                    if (!goes_left(*last)) --last; else break;
                }
                std::swap(*first, *last);
                ++first;
            }
        done:
            mid = (uint32_t)(first - prims.data());
            have_child_bounds = true;
        } else {
            mid = b + cnt / 2; /* all centroids coincide */
        }
        if (helpers > 0) free_threads.fetch_add(helpers); /* back to the pool before the children are spawned */
        if (mid == b || mid == e) {
            mid = b + cnt / 2;
            have_child_bounds = false;
        }
        if (!have_child_bounds) {
            bounds(b, mid, &lbb, &lcb);
            bounds(mid, e, &rbb, &rcb);
        }
        uint32_t l = alloc(), r = alloc();
        N.left = l;
        N.right = r;
        bool spawn = false;
        if (cnt > 8192 && free_threads.load(std::memory_order_relaxed) > 0) {
            if (free_threads.fetch_sub(1) > 0)
                spawn = true;
            else
                free_threads.fetch_add(1);
        }
        if (spawn) {
            std::thread t([=]() {
                build_range(l, b, mid, depth + 1, lbb, lcb);
                free_threads.fetch_add(1);
            });
            build_range(r, mid, e, depth + 1, rbb, rcb);
            t.join();
        } else {
            build_range(l, b, mid, depth + 1, lbb, lcb);
            build_range(r, mid, e, depth + 1, rbb, rcb);
        }
    }
};

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
    B.n = n_tris;
    B.prims.resize(n_tris);
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
        Prim& P = B.prims[t];
        for (int a = 0; a < 3; ++a) {
            P.c[a] = 0.5f * (b.lo[a] + b.hi[a]);
            b.lo[a] -= pad;
            b.hi[a] += pad;
        }
        P.box = b;
        P.id = t;
    }
    const bool verbose = std::getenv("AQ_BUILD_VERBOSE") != nullptr;
    auto tp0 = std::chrono::steady_clock::now();
    B.nodes.resize(2 * (size_t)n_tris);
    B.free_threads = n_threads - 1;
    if (n_tris >= kParallelNode) B.tmp.resize(n_tris);
    uint32_t root = B.alloc();
    {
        Box rbb, rcb;
        B.bounds(0, n_tris, &rbb, &rcb);
        B.build_range(root, 0, n_tris, 0, rbb, rcb);
    }
    /* the primitive order the emit code reads (leaf groups are ranges of it) */
    std::vector<uint32_t> order_of(n_tris);
    for (uint32_t i = 0; i < n_tris; ++i) order_of[i] = B.prims[i].id;
    auto tp1 = std::chrono::steady_clock::now();
    out->n_bvh2_nodes = B.n_nodes.load();

    /* ---- collapse + emit, breadth first (aq_bvh_emit.h does the per-node work) */
    struct Item {
        uint32_t n2;    /* BVH2 node */
        uint32_t out;   /* BVH8 node index */
        uint32_t depth;
    };
    std::vector<Item> queue;
    queue.reserve(n_tris / 2 + 16);
    out->nodes.reserve((size_t)(n_tris / 3 + 16) * AQ_NODE_WORDS);
    out->tris.resize((size_t)n_tris * AQ_TRI_WORDS);
    uint32_t tri_cursor = 0;
    out->nodes.resize(AQ_NODE_WORDS);
    queue.push_back({root, 0u, 1u});
    double sah = 0.0;
    const float root_area = std::max(aq_box_half_area(B.nodes[root].lo, B.nodes[root].hi), 1e-30f);
    /* ---- optional: cost-optimal collapse (Ylitie, Karras, Laine 2017, section 3.1).  Dynamic
     * programme over the BVH2: c(n,i) = cheapest way to represent subtree n with at most i
     * sibling entries of a wide node, each entry either a leaf (<= 3 triangles, cost A*T*Ct) or
     * an 8-wide node (cost A*Cn + best split of 8 entries over the two children).  Children
     * have larger indices than their parent (alloc order), so a reverse sweep is a post-order. */
    struct DP {
        float c[8];       /* c[1..7] */
        uint8_t split[9]; /* j = 2..8: entries given to the left child */
        uint8_t fewer;    /* bit i (2..7): c[i] == c[i-1] */
        uint8_t leaf;     /* c[1] is the leaf alternative */
    };
    const char* cm = std::getenv("AQUA_COLLAPSE");
    const bool use_dp = cm ? !std::strcmp(cm, "dp") : n_tris <= 4000000u;
    std::vector<DP> dp;
    if (use_dp) {
        const float Cn = 1.0f, Ct = tri_cost();
        const uint32_t nn = B.n_nodes.load();
        dp.resize(nn);
        for (uint32_t k = nn; k-- > 0;) {
            Node2& N = B.nodes[k];
            DP& D = dp[k];
            const float A = aq_box_half_area(N.lo, N.hi);
            if (N.left == AQ_BVH2_LEAF) {
                for (int i = 1; i < 8; ++i) D.c[i] = A * (float)N.count * Ct;
                D.leaf = 1;
                D.fewer = 0xFF;
                continue;
            }
            const DP &L = dp[N.left], &R = dp[N.right];
            float dist[9];
            for (int j = 2; j <= 8; ++j) {
                float best = INFINITY;
                int bk = 1;
                for (int kk = 1; kk < j; ++kk) {
                    float v = L.c[kk > 7 ? 7 : kk] + R.c[(j - kk) > 7 ? 7 : (j - kk)];
                    if (v < best) {
                        best = v;
                        bk = kk;
                    }
                }
                dist[j] = best;
                D.split[j] = (uint8_t)bk;
            }
            const float c_leaf = N.count <= kLeafMax ? A * (float)N.count * Ct : INFINITY;
            const float c_int = dist[8] + A * Cn;
            D.leaf = c_leaf <= c_int ? 1 : 0;
            D.c[1] = D.leaf ? c_leaf : c_int;
            D.fewer = 0;
            for (int i = 2; i < 8; ++i) {
                if (D.c[i - 1] <= dist[i]) {
                    D.c[i] = D.c[i - 1];
                    D.fewer |= (uint8_t)(1u << i);
                } else {
                    D.c[i] = dist[i];
                }
            }
        }
    }
    /* children of the wide node rooted at BVH2 node n according to the DP decisions */
    struct Collect {
        Builder& B;
        std::vector<DP>& dp;
        uint32_t ch[8];
        int nc = 0;
        void go(uint32_t n, int j) {
            Node2& N = B.nodes[n];
            if (N.left == AQ_BVH2_LEAF) {
                ch[nc++] = n;
                return;
            }
            if (j == 1) {
                if (dp[n].leaf) N.left = N.right = AQ_BVH2_LEAF; /* merged into one leaf group */
                ch[nc++] = n;
                return;
            }
            if (j < 8 && (dp[n].fewer & (1u << j))) {
                go(n, j - 1);
                return;
            }
            int k = dp[n].split[j];
            go(N.left, k);
            go(N.right, j - k);
        }
    };

    /* level-synchronous emit: the plans of one level are computed in parallel (a plan only touches the
     * BVH2 subtree of its item), a serial prefix assigns node and record slots in queue order — so the
     * output is byte-identical to a serial breadth-first walk — and the nodes are written in parallel */
    auto parallel_for = [&](size_t n, auto&& fn) {
        int T = (int)std::min<size_t>((size_t)n_threads, (n + 511) / 512);
        if (T <= 1) {
            for (size_t i = 0; i < n; ++i) fn(i);
            return;
        }
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
            th.emplace_back([&, t]() {
                for (size_t i = n * t / T, e = n * (t + 1) / T; i < e; ++i) fn(i);
            });
        for (auto& x : th) x.join();
    };
    std::vector<Item> next;
    std::vector<aq_node8_plan> plans;
    std::vector<uint32_t> cbase, tbase;
    while (!queue.empty()) {
        const size_t L = queue.size();
        plans.resize(L);
        parallel_for(L, [&](size_t i) {
            const Item& it = queue[i];
            if (use_dp && B.nodes[it.n2].left != AQ_BVH2_LEAF) {
                Collect col{B, dp};
                col.go(it.n2, 8);
                aq_node8_plan_from(B.nodes.data(), col.ch, col.nc, &plans[i]);
            } else {
                aq_node8_plan_children(B.nodes.data(), it.n2, &plans[i]);
            }
        });
        cbase.resize(L);
        tbase.resize(L);
        const uint32_t cb0 = (uint32_t)(out->nodes.size() / AQ_NODE_WORDS);
        uint32_t cb = cb0;
        for (size_t i = 0; i < L; ++i) {
            out->max_depth = std::max(out->max_depth, queue[i].depth);
            sah += (double)aq_box_half_area(plans[i].lo, plans[i].hi) / root_area;
            cbase[i] = cb;
            tbase[i] = tri_cursor;
            cb += plans[i].n_inner;
            tri_cursor += plans[i].n_tris;
        }
        out->nodes.resize((size_t)cb * AQ_NODE_WORDS);
        next.resize(cb - cb0);
        parallel_for(L, [&](size_t i) {
            const Item& it = queue[i];
            uint32_t inner[8];
            aq_node8_write(B.nodes.data(), plans[i], order_of.data(), positions, indices, cbase[i], tbase[i],
                           &out->nodes[(size_t)it.out * AQ_NODE_WORDS], out->tris.data(), inner);
            for (uint32_t k = 0; k < plans[i].n_inner; ++k) next[cbase[i] - cb0 + k] = {inner[k], cbase[i] + k, it.depth + 1};
        });
        queue.swap(next);
    }
    out->sah_cost = (float)sah;
    if (verbose) {
        auto tp2 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[aq_bvh_build] %u tris, %d threads: BVH2 %.1f ms, collapse+emit %.1f ms\n", n_tris, n_threads,
                     std::chrono::duration<double, std::milli>(tp1 - tp0).count(),
                     std::chrono::duration<double, std::milli>(tp2 - tp1).count());
    }
    if (out->max_depth >= AQ_STACK_MAX) return -1;
    return 0;
}
