// Warp-level cost model of the BVH8 traversal loop (development aid, CPU only): replays the
// per-lane-refill scheduling of aq_k_trace on a ray set with the product's own traversal
// template (aq_bvh.h) and counts warp instructions under different step policies.
//   warp_sim nodes.bin tris.bin rays.bin policy K [any]
// policy: 0 = product (open node, then all triangles), 1 = capped (at most K triangle tests per
// step; a lane with triangles left does not open a node in the next step), 2 = vote (triangle phase
// only when >= K lanes have triangles pending or no lane has node work; pending groups on the stack)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include "aq_bvh.h"

struct Ray { float o[3], tmin, d[3], tmax; };
// instruction costs read off the SASS of aq_k_trace<3,false> (round 2, after the node-visit diet)
static const int C_LOOP = 28, C_REFILL = 55, C_NODE = 258, C_FINISH = 12;
static const int C_TRI[5] = {0, 26, 38, 52, 80}; // stage reached: 1 det, 2 U, 3 V, 4 full

static int tri_stage(const aq_trav& T, const aq_f4* tp) {
    aq_f4 t0 = tp[0], t1 = tp[1], t2 = tp[2];
    aq_v3 v0 = aq_mk(t0.x, t0.y, t0.z), e1 = aq_mk(t0.w, t1.x, t1.y), e2 = aq_mk(t1.z, t1.w, t2.x);
    aq_v3 pvec = aq_cross(T.d, e2);
    float det = aq_dot(e1, pvec), adet = fabsf(det);
    if (!(adet > 0.0f)) return 1;
    float sg = det < 0.0f ? -1.0f : 1.0f;
    aq_v3 tvec = aq_sub(T.o, v0);
    float U = aq_dot(tvec, pvec) * sg;
    if (U < 0.0f || U > adet) return 2;
    aq_v3 qvec = aq_cross(tvec, e1);
    float V = aq_dot(T.d, qvec) * sg;
    if (V < 0.0f || U + V > adet) return 3;
    return 4;
}

template <class T> static std::vector<T> load(const char* p) {
    FILE* f = fopen(p, "rb"); if (!f) { perror(p); exit(1); }
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<T> v(n / sizeof(T)); if (fread(v.data(), sizeof(T), v.size(), f) != v.size()) exit(1); fclose(f); return v;
}

struct Lane { aq_trav T; aq_local_stack st; bool active = false; uint32_t tg_x = 0, tg_y = 0; };

int main(int argc, char** argv) {
    if (argc < 6) { fprintf(stderr, "usage\n"); return 1; }
    auto nodes = load<aq_u4>(argv[1]); auto tris = load<aq_f4>(argv[2]); auto rays = load<Ray>(argv[3]);
    int policy = atoi(argv[4]), K = atoi(argv[5]); bool any = argc > 6 && atoi(argv[6]); int K2 = argc > 7 ? atoi(argv[7]) : 0; int R = argc > 8 ? atoi(argv[8]) : 1; /* refill only when >= R lanes are idle (or none is active) */
    const size_t SEG = 32 * 64; // rays one simulated warp works through
    double instr = 0, thr = 0, node_instr = 0, tri_instr = 0, node_thr = 0, tri_thr = 0, steps = 0, node_visits = 0, tri_tests = 0;
    uint64_t checksum = 0;
    for (size_t seg = 0; seg < rays.size(); seg += SEG) {
        size_t next = seg, end = std::min(rays.size(), seg + SEG);
        std::vector<Lane> L(32);
        for (;;) {
            // refill
            int refilled = 0, n_idle = 0, n_busy = 0;
            for (auto& l : L) { n_idle += !l.active; n_busy += l.active; }
            if (n_idle >= R || n_busy == 0)
            for (auto& l : L) if (!l.active && next < end) {
                const Ray& r = rays[next++];
                aq_trav_init(l.T, aq_mk(r.o[0], r.o[1], r.o[2]), aq_mk(r.d[0], r.d[1], r.d[2]), r.tmin, r.tmax, l.st);
                l.tg_y = 0; l.active = true; ++refilled;
            }
            int nact = 0; for (auto& l : L) nact += l.active;
            if (!nact) break;
            instr += C_LOOP; thr += C_LOOP * 32.0; steps += 1;
            if (refilled) { instr += C_REFILL; thr += C_REFILL * (double)refilled; }
            // ---- node phase
            auto wants_node = [&](Lane& l) {
                if (!l.active) return false;
                if (policy == 1 || policy == 3 || policy == 5) return l.tg_y == 0 && l.T.ng_y > 0x00FFFFFFu;
                return l.T.ng_y > 0x00FFFFFFu;
            };
            if (policy == 2) // refill work registers from the stack
                for (auto& l : L) if (l.active) {
                    if (l.tg_y == 0 && !l.st.empty() && l.st.top_y() <= 0x00FFFFFFu) l.st.pop(l.tg_x, l.tg_y);
                    if (l.T.ng_y <= 0x00FFFFFFu && !l.st.empty() && l.st.top_y() > 0x00FFFFFFu) l.st.pop(l.T.ng_x, l.T.ng_y);
                }
            int nn = 0, ntp = 0;
            for (auto& l : L) { nn += wants_node(l); ntp += l.active && l.tg_y != 0; }
            bool do_node = nn > 0, do_tri = true;
            if (policy == 3 && K2 > 0 && nn < K2 && ntp > 0) do_node = false; // too few lanes want a node: triangles first
            if (policy == 2) { do_tri = ntp >= K || nn == 0; if (do_tri && ntp >= K) do_node = do_node && true; }
            if (do_node) {
                for (auto& l : L) if (wants_node(l)) {
                    uint32_t nx, ny;
                    aq_trav_open_node<false>(nodes.data(), l.T, l.st, nullptr, nx, ny);
                    node_visits += 1;
                    if (policy == 2 && l.tg_y != 0) { if (ny) l.st.push(nx, ny); }
                    else { l.tg_x = nx; l.tg_y = ny; }
                }
                instr += C_NODE; thr += C_NODE * (double)nn; node_instr += C_NODE; node_thr += C_NODE * (double)nn;
            }
            // ---- triangle phase
            if (do_tri) {
                for (int it = 0; (policy == 0 || it < K || policy == 2 || policy == 3 || policy == 5) ; ++it) {
                    if (policy == 2 && it >= 1) break; // vote policy: one triangle per lane per phase
                    if (policy == 5) { // shipped: first iteration always, then while >= K lanes pending; one vote per iteration
                        int np = 0;
                        for (auto& l : L) np += l.active && l.tg_y != 0;
                        if (np == 0 || (it > 0 && np < K)) break;
                        instr += 6; thr += 6 * 32.0;
                    }
                    if (policy == 3) { // dynamic: iterate while >= K lanes have triangles pending (or nobody can open a node)
                        int np = 0, nw = 0;
                        for (auto& l : L) { np += l.active && l.tg_y != 0; nw += l.active && l.tg_y == 0 && (l.T.ng_y > 0x00FFFFFFu || !l.st.empty()); }
                        if (np == 0) break;
                        if (np < K && nw > 0 && it > 0) break;
                    }
                    int na = 0, mx = 0;
                    for (auto& l : L) if (l.active && l.tg_y) {
                        uint32_t i = aq_msb(l.tg_y); l.tg_y &= ~(1u << i);
                        const aq_f4* tp = tris.data() + (size_t)(l.tg_x + i) * AQ_TRI_WORDS;
                        int sgt = tri_stage(l.T, tp);
                        bool fin = any ? aq_trav_test_tri<true, false>(tris.data(), l.T, l.tg_x + i, nullptr)
                                       : aq_trav_test_tri<false, false>(tris.data(), l.T, l.tg_x + i, nullptr);
                        tri_tests += 1; ++na; mx = std::max(mx, sgt);
                        if (fin) { l.active = false; l.tg_y = 0; checksum += 1; instr += 0; }
                    }
                    if (!na) break;
                    instr += C_TRI[mx]; thr += C_TRI[mx] * (double)na; tri_instr += C_TRI[mx]; tri_thr += C_TRI[mx] * (double)na;
                }
            }
            // ---- next node group / finish
            int nfin = 0;
            for (auto& l : L) if (l.active) {
                if (l.tg_y) continue; // triangles left: stay
                if (l.T.ng_y <= 0x00FFFFFFu) {
                    if (l.st.empty()) { l.active = false; ++nfin; checksum += l.T.best_prim * 2654435761u + (uint64_t)aq_f2u(l.T.best_t); }
                    else if (policy != 2) l.st.pop(l.T.ng_x, l.T.ng_y);
                }
            }
            if (nfin) { instr += C_FINISH; thr += C_FINISH * (double)nfin; }
        }
    }
    double n = (double)rays.size();
    printf("policy=%d K=%d any=%d rays=%zu | warp-instr/ray=%.1f thr/inst=%.2f | node: %.1f%% @%.1f  tri: %.1f%% @%.1f | steps/ray=%.2f nodes/ray=%.2f tris/ray=%.2f checksum=%llx\n",
           policy, K, (int)any, rays.size(), instr / n * 32.0 / 32.0, thr / instr, 100 * node_instr / instr, node_thr / node_instr,
           100 * tri_instr / instr, tri_thr / std::max(1.0, tri_instr), steps / n * 32, node_visits / n, tri_tests / n, (unsigned long long)checksum);
    return 0;
}
