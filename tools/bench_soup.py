"""BASELINE config C4: N random triangles (seed 12345, centres U[0,1]^3, half-extent 0.005),
R uniform random rays; incoherent-ray intersection throughput of the BVH8 traversal kernel,
its roofline from device-counted node/triangle fetches, and a hit-id bit-exact check against
the CPU oracle on subsets.  Prints one JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import aqua_engine_b200 as aq


def soup(n, seed=12345, r=0.005):
    g = np.random.default_rng(seed)
    c = g.uniform(0, 1, (n, 1, 3)).astype(np.float32)
    v = c + g.uniform(-r, r, (n, 3, 3)).astype(np.float32)
    return v.reshape(-1, 3), np.arange(3 * n, dtype=np.uint32).reshape(-1, 3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tris", type=int, default=10_000_000)
    ap.add_argument("--rays", type=int, default=1 << 26)
    ap.add_argument("--check-brute", type=int, default=1 << 12)
    ap.add_argument("--check-bvh", type=int, default=1 << 20)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--brief", action="store_true", help="print only the throughput numbers")
    a = ap.parse_args()
    import torch
    t0 = time.time()
    pos, idx = soup(a.tris)
    t_gen = time.time() - t0
    sc = aq.Scene.from_arrays(pos, idx)
    r = aq.Renderer(0)
    r.set_stream(torch.cuda.current_stream().cuda_stream)
    t0 = time.time()
    ds = r.upload(sc)
    t_build = time.time() - t0
    g = torch.Generator(device="cuda").manual_seed(7)
    rays = torch.empty(a.rays, 8, device="cuda")
    rays[:, 0:3] = torch.rand(a.rays, 3, device="cuda", generator=g)
    d = torch.randn(a.rays, 3, device="cuda", generator=g)
    rays[:, 4:7] = d / d.norm(dim=1, keepdim=True)
    rays[:, 3] = 0.0
    rays[:, 7] = 3.0e38
    hits = torch.empty(a.rays, 4, device="cuda", dtype=torch.int32)
    out = {}
    for any_hit in (False, True):
        best = None
        for _ in range(a.reps):
            ds.trace_counters(reset=True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ds.intersect_device(rays.data_ptr(), a.rays, hits.data_ptr(), any_hit)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            nn, nt = ds.trace_counters(reset=True)
            if best is None or ms < best[0]:
                best = (ms, nn, nt)
        ms, nn, nt = best
        out_b = 4 if any_hit else 16
        alg = a.rays * (32 + out_b) + 80 * nn + 48 * nt
        out["any" if any_hit else "closest"] = {
            "ms": ms, "mrays_s": a.rays / ms / 1e3, "nodes_per_ray": nn / a.rays, "tris_per_ray": nt / a.rays,
            "algorithmic_bytes_per_ray": alg / a.rays, "achieved_gbs": alg / ms / 1e6,
            "frac_of_measured_hbm": alg / ms / 1e6 / 6555.8}
        if not any_hit:
            h_closest = hits.cpu().numpy().view(aq.HIT_DTYPE).reshape(-1).copy()
    # ---- parity on subsets
    import aq_oracle as ao
    o = ao.OracleScene(sc, build_bvh=True)
    rh = rays.cpu().numpy().view(aq.RAY_DTYPE).reshape(-1)
    chk = {}
    for name, cnt, mode in (("brute_force", a.check_brute, 0), ("oracle_bvh2", a.check_bvh, 1)):
        if cnt <= 0:
            continue
        t0 = time.time()
        oh = o.intersect(rh[:cnt], mode=mode)
        gh = h_closest[:cnt]
        same = bool(np.array_equal(oh["prim"], gh["prim"]) and np.array_equal(oh["t"], gh["t"])
                    and np.array_equal(oh["u"], gh["u"]) and np.array_equal(oh["v"], gh["v"]))
        chk[name] = {"rays": cnt, "bit_exact": same, "mismatches": int((oh["prim"] != gh["prim"]).sum()),
                     "cpu_s": time.time() - t0, "cpu_mrays_s": cnt / (time.time() - t0) / 1e6}
    line = {"config": f"C4: {a.tris} random triangles, {a.rays} uniform random rays", "n_nodes": ds.accel.n_nodes,
            "builder": "device LBVH" if ds.accel.builder else "host binned SAH",
            "bvh_depth": ds.accel.max_depth, "build_ms": ds.accel.build_ms, "upload_build_s": t_build,
            "gen_s": t_gen, "hit_fraction": float((h_closest["prim"] != aq.AQ_MISS).mean()), **out, "parity": chk,
            "cpu_threads": ao.threads()}
    if a.brief:
        print(os.environ.get("AQUA_CUDA_LIB", "base"), {k: (round(v["mrays_s"]), round(v["ms"], 2), round(v["nodes_per_ray"], 2), round(v["tris_per_ray"], 2)) for k, v in out.items()},
              "nodes", ds.accel.n_nodes, "build_ms", round(ds.accel.build_ms, 1))
    else:
        print(json.dumps(line))


if __name__ == "__main__":
    main()
