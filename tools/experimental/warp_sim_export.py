"""Export BVH8 + a depth-1-like ray set (pixel order: coherent origins, random directions) for warp_sim."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import aq_oracle as ao
import aqua_engine_b200 as aq
name, W, H, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
scene = aq.Scene.load(os.path.join(aq.scenes_dir(), name + ".json"))
pos, idx, *_ = scene.arrays()
nodes, tris, info = aq.build_accel_host(pos, idx)
o = ao.OracleScene(scene, build_bvh=True)
cam = o.camera_rays(aq.Integrator(spp=1).cfg(width=W, height=H), 0)
h = o.intersect(cam, mode=1)
g = np.random.default_rng(1)
ok = h["prim"] != aq.AQ_MISS
P = cam["o"][ok] + h["t"][ok, None] * cam["d"][ok]
d = g.normal(size=P.shape); d /= np.linalg.norm(d, axis=1, keepdims=True)
b = np.zeros(len(P), aq.RAY_DTYPE)
b["o"], b["d"], b["tmin"], b["tmax"] = P + 1e-3 * d, d, 0.0, 3e38
os.makedirs(out, exist_ok=True)
np.ascontiguousarray(nodes).tofile(os.path.join(out, "nodes.bin")); np.ascontiguousarray(tris).tofile(os.path.join(out, "tris.bin"))
b.tofile(os.path.join(out, "rays.bin")); cam.tofile(os.path.join(out, "cam.bin"))
sh = b.copy(); sh["tmax"] = 2.0; sh.tofile(os.path.join(out, "shadow.bin"))
print(name, len(b), "rays", info.n_nodes, "nodes", nodes.dtype, nodes.shape, tris.shape, b.dtype.itemsize)
