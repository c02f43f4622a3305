// Tree-quality prototype (CPU only, development aid): LBVH (highest differing Morton bit) vs PLOC
// (parallel locally-ordered clustering, Meister & Bittner 2018) on the same Morton order; reports the
// SAH cost of the binary tree and the BVH2 node visits / triangle tests per closest-hit ray.
//   ploc_proto pos.bin idx.bin rays.bin [radius]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <algorithm>
#include <numeric>
#include "aq_core.h"

struct Ray { float o[3], tmin, d[3], tmax; };
struct Node { float lo[3], hi[3]; int left, right; int first, count; };
template <class T> static std::vector<T> load(const char* p) {
    FILE* f = fopen(p, "rb"); if (!f) { perror(p); exit(1); }
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<T> v(n / sizeof(T)); if (fread(v.data(), sizeof(T), v.size(), f) != v.size()) exit(1); fclose(f); return v;
}
static inline float harea(const float* lo, const float* hi) { float x = hi[0]-lo[0], y = hi[1]-lo[1], z = hi[2]-lo[2]; return x*y + y*z + z*x; }
static inline unsigned long long expand21(unsigned long long v) {
    v &= 0x1FFFFFull; v = (v | v << 32) & 0x1F00000000FFFFull; v = (v | v << 16) & 0x1F0000FF0000FFull;
    v = (v | v << 8) & 0x100F00F00F00F00Full; v = (v | v << 4) & 0x10C30C30C30C30C3ull; v = (v | v << 2) & 0x1249249249249249ull; return v;
}
static std::vector<float> P; static std::vector<uint32_t> I; static std::vector<Node> N; static std::vector<uint32_t> order;
static std::vector<unsigned long long> keys;

static int lbvh(int first, int last) { // [first,last] sorted positions; returns node index
    if (first == last) return first; // leaves occupy N[0..n)
    unsigned long long a = keys[first], b = keys[last];
    int split;
    if (a == b) split = (first + last) >> 1;
    else {
        int prefix = __builtin_clzll(a ^ b);
        split = first; int step = last - first;
        do { step = (step + 1) >> 1; int ns = split + step;
             if (ns < last) { unsigned long long c = keys[ns]; if (c == a || __builtin_clzll(a ^ c) > prefix) split = ns; } } while (step > 1);
    }
    int l = lbvh(first, split), r = lbvh(split + 1, last);
    Node m; for (int k = 0; k < 3; ++k) { m.lo[k] = std::min(N[l].lo[k], N[r].lo[k]); m.hi[k] = std::max(N[l].hi[k], N[r].hi[k]); }
    m.left = l; m.right = r; m.first = first; m.count = last - first + 1; N.push_back(m); return (int)N.size() - 1;
}
static int ploc(int n, int R) {
    std::vector<int> C(n), C2, nn(n); std::iota(C.begin(), C.end(), 0);
    int iters = 0;
    while (C.size() > 1) {
        int m = (int)C.size(); nn.resize(m);
        for (int i = 0; i < m; ++i) {
            float best = INFINITY; int bj = -1; const Node& a = N[C[i]];
            for (int j = std::max(0, i - R); j <= std::min(m - 1, i + R); ++j) if (j != i) {
                const Node& b = N[C[j]]; float lo[3], hi[3];
                for (int k = 0; k < 3; ++k) { lo[k] = std::min(a.lo[k], b.lo[k]); hi[k] = std::max(a.hi[k], b.hi[k]); }
                float ar = harea(lo, hi); if (ar < best) { best = ar; bj = j; }
            }
            nn[i] = bj;
        }
        C2.clear();
        for (int i = 0; i < m; ++i) {
            int j = nn[i];
            if (nn[j] == i) { if (i < j) { Node mnode; const Node &a = N[C[i]], &b = N[C[j]];
                    for (int k = 0; k < 3; ++k) { mnode.lo[k] = std::min(a.lo[k], b.lo[k]); mnode.hi[k] = std::max(a.hi[k], b.hi[k]); }
                    mnode.left = C[i]; mnode.right = C[j]; mnode.first = -1; mnode.count = a.count + b.count; N.push_back(mnode); C2.push_back((int)N.size() - 1); } }
            else C2.push_back(C[i]);
        }
        C.swap(C2); ++iters;
    }
    fprintf(stderr, "ploc: %d iterations\n", iters);
    return C[0];
}
static double sah(int root) { double c = 0; float ra = harea(N[root].lo, N[root].hi); std::vector<int> st{root};
    while (!st.empty()) { int k = st.back(); st.pop_back(); const Node& nd = N[k]; float a = harea(nd.lo, nd.hi) / ra;
        if (nd.left < 0) c += a * 1.0; else { c += a * 1.0; st.push_back(nd.left); st.push_back(nd.right); } } return c; }
static void trace(int root, const std::vector<Ray>& rays, const std::vector<int>& leaf_prim, double& nodes, double& tris) {
    nodes = tris = 0;
    for (const Ray& r : rays) {
        aq_v3 o = aq_mk(r.o[0], r.o[1], r.o[2]), d = aq_mk(r.d[0], r.d[1], r.d[2]);
        float id[3] = {1.0f / (fabsf(r.d[0]) > 1e-20f ? r.d[0] : 1e-20f), 1.0f / (fabsf(r.d[1]) > 1e-20f ? r.d[1] : 1e-20f), 1.0f / (fabsf(r.d[2]) > 1e-20f ? r.d[2] : 1e-20f)};
        float best = r.tmax; int st[128], sp = 0; st[sp++] = root;
        while (sp) { int k = st[--sp]; const Node& nd = N[k]; nodes += 1;
            float t0 = r.tmin, t1 = best; bool hit = true;
            for (int a = 0; a < 3; ++a) { float ta = (nd.lo[a] - r.o[a]) * id[a], tb = (nd.hi[a] - r.o[a]) * id[a]; if (ta > tb) std::swap(ta, tb); t0 = std::max(t0, ta); t1 = std::min(t1, tb); }
            if (!(t0 <= t1)) continue;
            if (nd.left < 0) { int prim = leaf_prim[k]; tris += 1;
                const float* v0 = &P[3 * I[3 * prim]]; const float* v1 = &P[3 * I[3 * prim + 1]]; const float* v2 = &P[3 * I[3 * prim + 2]];
                float t, u, v; if (aq_tri_test(o, d, r.tmin, aq_mk(v0[0], v0[1], v0[2]), aq_mk(v1[0]-v0[0], v1[1]-v0[1], v1[2]-v0[2]), aq_mk(v2[0]-v0[0], v2[1]-v0[1], v2[2]-v0[2]), &t, &u, &v) && t < best) best = t;
            } else { // near child first (by box centre along the dominant axis: cheap proxy)
                st[sp++] = nd.right; st[sp++] = nd.left; }
        }
    }
    nodes /= rays.size(); tris /= rays.size();
}
int main(int argc, char** argv) {
    P = load<float>(argv[1]); I = load<uint32_t>(argv[2]); auto rays = load<Ray>(argv[3]); int R = argc > 4 ? atoi(argv[4]) : 16;
    if (rays.size() > 20000) rays.resize(20000);
    int n = (int)I.size() / 3; float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    std::vector<Node> leaves(n);
    for (int t = 0; t < n; ++t) { Node& L = leaves[t]; for (int k = 0; k < 3; ++k) { L.lo[k] = INFINITY; L.hi[k] = -INFINITY; }
        for (int v = 0; v < 3; ++v) for (int k = 0; k < 3; ++k) { float x = P[3 * I[3 * t + v] + k]; L.lo[k] = std::min(L.lo[k], x); L.hi[k] = std::max(L.hi[k], x); }
        for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], L.lo[k]); hi[k] = std::max(hi[k], L.hi[k]); } L.left = L.right = -1; L.count = 1; }
    keys.resize(n); order.resize(n); std::iota(order.begin(), order.end(), 0u); std::vector<unsigned long long> k0(n);
    for (int t = 0; t < n; ++t) { unsigned long long q[3]; for (int k = 0; k < 3; ++k) { float c = 0.5f * (leaves[t].lo[k] + leaves[t].hi[k]); float x = hi[k] > lo[k] ? (c - lo[k]) / (hi[k] - lo[k]) : 0.f; q[k] = (unsigned long long)std::min(x * 2097152.0f, 2097151.0f); }
        k0[t] = (expand21(q[0]) << 2) | (expand21(q[1]) << 1) | expand21(q[2]); }
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return k0[a] < k0[b] || (k0[a] == k0[b] && a < b); });
    for (int i = 0; i < n; ++i) keys[i] = k0[order[i]];
    for (int variant = 0; variant < 2; ++variant) {
        N.clear(); N.reserve(2 * (size_t)n); std::vector<int> leaf_prim(2 * (size_t)n, -1);
        for (int i = 0; i < n; ++i) { Node L = leaves[order[i]]; L.first = i; N.push_back(L); leaf_prim[i] = (int)order[i]; }
        int root = variant == 0 ? lbvh(0, n - 1) : ploc(n, R);
        double nd, tr; trace(root, rays, leaf_prim, nd, tr);
        printf("%s: nodes=%zu sah=%.2f bvh2 node visits/ray=%.1f tri tests/ray=%.2f\n", variant == 0 ? "lbvh" : "ploc", N.size(), sah(root), nd, tr);
    }
    return 0;
}
