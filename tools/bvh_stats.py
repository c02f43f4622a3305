"""Nodes / triangle records fetched per ray for the host-built BVH8 of a scene (CPU walk of the
same traversal template as the kernel) — development aid for builder A/B runs
(AQUA_BVH_LEAF, AQUA_BVH_CT, AQUA_COLLAPSE are read by the builder at first use)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import aq_oracle as ao
import aqua_engine_b200 as aq

scene = aq.Scene.load(os.path.join(aq.scenes_dir(), (sys.argv[1] if len(sys.argv) > 1 else "room") + ".json"))
pos, idx, *_ = scene.arrays()
nodes, tris, info = aq.build_accel_host(pos, idx)
o = ao.OracleScene(scene, build_bvh=True)
cfg = aq.Integrator(spp=1).cfg(width=320, height=180)
cam = o.camera_rays(cfg, 0)
h = o.intersect(cam, mode=1)
# bounce rays: cosine-ish random directions from the camera hit points (incoherent, like depth >= 1)
g = np.random.default_rng(1)
ok = h["prim"] != aq.AQ_MISS
P = cam["o"][ok] + h["t"][ok, None] * cam["d"][ok]
d = g.normal(size=P.shape)
d /= np.linalg.norm(d, axis=1, keepdims=True)
b = np.zeros(len(P), aq.RAY_DTYPE)
b["o"], b["d"], b["tmin"], b["tmax"] = P + 1e-3 * d, d, 0.0, 3e38
out = [f"leaf={os.environ.get('AQUA_BVH_LEAF','-')} ct={os.environ.get('AQUA_BVH_CT','-')} nodes={info.n_nodes} recs={info.n_tri_records} depth={info.max_depth}"]
for name, r in (("camera", cam), ("bounce", b)):
    for anyhit in (False, True):
        if anyhit:
            r = r.copy()
            r["tmax"] = 2.0
        hh, nn, nt = ao.bvh8_intersect(nodes, tris, r, any_hit=anyhit)
        out.append(f"{name}{'-any' if anyhit else ''}: n/ray={nn/len(r):.2f} t/ray={nt/len(r):.2f} cost(296n+178t)={(296*nn+178*nt)/len(r):.0f}")
print(" | ".join(out))
