// CPU replay of the device builder's PLOC stage (development aid): runs the per-element functions of
// aq_bvh_ploc.h in loops exactly as the kernels of aq_bvh_build_gpu.cu call them, then checks the tree
// (every leaf reachable once, boxes enclose children, counts add up, the leaves of every subtree are a
// contiguous range of the new order) and walks the DP collapse as the emit pass does.
//   ploc_emulate pos.bin idx.bin [radius]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <algorithm>
#include <numeric>
#include "aq_bvh_ploc.h"

template <class T> static std::vector<T> load(const char* p) {
    FILE* f = fopen(p, "rb"); if (!f) { perror(p); exit(1); }
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<T> v(n / sizeof(T)); if (fread(v.data(), sizeof(T), v.size(), f) != v.size()) exit(1); fclose(f); return v;
}
static inline unsigned long long expand21(unsigned long long v) {
    v &= 0x1FFFFFull; v = (v | v << 32) & 0x1F00000000FFFFull; v = (v | v << 16) & 0x1F0000FF0000FFull;
    v = (v | v << 8) & 0x100F00F00F00F00Full; v = (v | v << 4) & 0x10C30C30C30C30C3ull; v = (v | v << 2) & 0x1249249249249249ull; return v;
}
int main(int argc, char** argv) {
    auto P = load<float>(argv[1]); auto I = load<uint32_t>(argv[2]); uint32_t R = argc > 3 ? atoi(argv[3]) : 16; const float Ct = 1.0f;
    const uint32_t n = (uint32_t)I.size() / 3;
    std::vector<aq_box6> tb(n); float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t t = 0; t < n; ++t) { for (int k = 0; k < 3; ++k) { tb[t].lo[k] = INFINITY; tb[t].hi[k] = -INFINITY; }
        for (int v = 0; v < 3; ++v) for (int k = 0; k < 3; ++k) { float x = P[3 * (size_t)I[3 * (size_t)t + v] + k]; tb[t].lo[k] = std::min(tb[t].lo[k], x); tb[t].hi[k] = std::max(tb[t].hi[k], x); }
        for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], tb[t].lo[k]); hi[k] = std::max(hi[k], tb[t].hi[k]); } }
    std::vector<unsigned long long> k0(n); std::vector<uint32_t> order(n); std::iota(order.begin(), order.end(), 0u);
    for (uint32_t t = 0; t < n; ++t) { unsigned long long q[3]; for (int k = 0; k < 3; ++k) { float c = 0.5f * (tb[t].lo[k] + tb[t].hi[k]); float x = hi[k] > lo[k] ? (c - lo[k]) / (hi[k] - lo[k]) : 0.f; q[k] = (unsigned long long)std::min(x * 2097152.0f, 2097151.0f); }
        k0[t] = (expand21(q[0]) << 2) | (expand21(q[1]) << 1) | expand21(q[2]); }
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return k0[a] < k0[b] || (k0[a] == k0[b] && a < b); });
    // device layout: internal nodes [0, n-1), leaf of sorted position p at (n-1)+p
    std::vector<aq_bvh2_node> N(2 * (size_t)n); std::vector<aq_dp8> dp(2 * (size_t)n);
    std::vector<uint32_t> cid(n), cid2(n), nn(n), out_id(n), keep(n), pos(n); std::vector<aq_box6> cbox(n), cbox2(n), out_box(n);
    for (uint32_t p = 0; p < n; ++p) { aq_bvh2_node L; for (int k = 0; k < 3; ++k) { L.lo[k] = tb[order[p]].lo[k]; L.hi[k] = tb[order[p]].hi[k]; }
        L.left = L.right = AQ_BVH2_LEAF; L.first = p; L.count = 1; N[(size_t)(n - 1) + p] = L; aq_dp8_leaf(dp[(size_t)(n - 1) + p], aq_box_half_area(L.lo, L.hi), 1u, Ct);
        cid[p] = n - 1 + p; cbox[p] = tb[order[p]]; }
    uint32_t m = n, counter = 0, iters = 0; std::vector<uint32_t> range_start;
    while (m > 1) {
        range_start.push_back(counter);
        for (uint32_t i = 0; i < m; ++i) nn[i] = aq_ploc_nearest(cbox.data(), m, i, R);              // k_ploc_nn
        for (uint32_t i = 0; i < m; ++i) {                                                          // k_ploc_merge
            int fate = aq_ploc_fate(nn.data(), i); keep[i] = fate != 0;
            if (fate == 2) { uint32_t k = counter++; aq_ploc_make_parent(N.data(), dp.data(), k, cid[i], cid[nn[i]], cbox[i], cbox[nn[i]], Ct, &out_box[i]); out_id[i] = k; }
            else if (fate == 1) { out_id[i] = cid[i]; out_box[i] = cbox[i]; }
        }
        uint32_t s = 0; for (uint32_t i = 0; i < m; ++i) { pos[i] = s; s += keep[i]; }                 // scan
        for (uint32_t i = 0; i < m; ++i) if (keep[i]) { cid2[pos[i]] = out_id[i]; cbox2[pos[i]] = out_box[i]; } // k_ploc_scatter
        cid.swap(cid2); cbox.swap(cbox2); m = s; ++iters;
    }
    range_start.push_back(counter);
    const uint32_t root = cid[0];
    printf("n=%u iterations=%u internal=%u root=%u\n", n, iters, counter, root);
    if (counter != n - 1) { printf("FAIL: internal node count\n"); return 1; }
    N[root].first = 0;
    for (size_t t = range_start.size() - 1; t-- > 0;)                                                // k_ploc_first, last iteration first
        for (uint32_t k = range_start[t]; k < range_start[t + 1]; ++k) aq_ploc_assign_first(N.data(), k);
    std::vector<uint32_t> order_new(n, 0xFFFFFFFFu);
    for (uint32_t p = 0; p < n; ++p) { uint32_t f = N[(size_t)(n - 1) + p].first; if (f >= n || order_new[f] != 0xFFFFFFFFu) { printf("FAIL: leaf slot %u\n", f); return 1; } order_new[f] = order[p]; }
    // invariants
    double sah = 0; const float ra = aq_box_half_area(N[root].lo, N[root].hi); size_t seen = 0;
    std::vector<uint32_t> st{root};
    while (!st.empty()) { uint32_t k = st.back(); st.pop_back(); const aq_bvh2_node& nd = N[k]; sah += aq_box_half_area(nd.lo, nd.hi) / ra;
        if (nd.left == AQ_BVH2_LEAF) { ++seen; continue; }
        const aq_bvh2_node &a = N[nd.left], &b = N[nd.right];
        if (a.count + b.count != nd.count || a.first != nd.first || b.first != nd.first + a.count) { printf("FAIL: ranges at node %u\n", k); return 1; }
        for (int t = 0; t < 3; ++t) if (nd.lo[t] > std::min(a.lo[t], b.lo[t]) || nd.hi[t] < std::max(a.hi[t], b.hi[t])) { printf("FAIL: box at node %u\n", k); return 1; }
        st.push_back(nd.left); st.push_back(nd.right); }
    if (seen != n) { printf("FAIL: %zu leaves reachable\n", seen); return 1; }
    printf("tree ok: sah=%.2f (cost 1 per node, leaves of one triangle)\n", sah);
    // the emit pass: breadth first over wide nodes through the DP decisions
    std::vector<uint32_t> q{root}, covered(n, 0); size_t wide = 0, groups = 0, tris = 0; uint32_t depth = 0;
    while (!q.empty()) { std::vector<uint32_t> next; ++depth;
        for (uint32_t it : q) { ++wide; uint32_t ch[8]; int nc;
            if (N[it].left != AQ_BVH2_LEAF) nc = aq_dp8_collect(N.data(), dp.data(), it, ch); else { ch[0] = it; nc = 1; }
            if (nc < 1 || nc > 8) { printf("FAIL: %d children\n", nc); return 1; }
            aq_node8_plan plan; aq_node8_plan_from(N.data(), ch, nc, &plan);
            for (int c = 0; c < nc; ++c) { const aq_bvh2_node& C = N[ch[c]];
                if (C.left == AQ_BVH2_LEAF) { if (C.count < 1 || C.count > AQ_LEAF_MAX) { printf("FAIL: leaf group of %u\n", C.count); return 1; }
                    ++groups; tris += C.count; for (uint32_t x = C.first; x < C.first + C.count; ++x) covered[x]++; }
                else next.push_back(ch[c]); } }
        q.swap(next); }
    for (uint32_t x = 0; x < n; ++x) if (covered[x] != 1) { printf("FAIL: slot %u covered %u times\n", x, covered[x]); return 1; }
    printf("emit ok: %zu wide nodes, %zu leaf groups, %zu triangles, depth %u\n", wide, groups, tris, depth);
    return 0;
}
