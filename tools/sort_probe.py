"""How much would ray ORDER buy the traversal kernels?  (sort / bin continuation and
shadow rays by direction octant + origin cell.)  Builds the depth-1 ray sets of a render (bounce rays
and point-light shadow rays from the camera hit points), reorders them in several ways and times the
same traversal kernel (aq_intersect_device_async) on each order.  Development probe; one JSON line."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aqua_engine_b200 as aq

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="cbox")
ap.add_argument("--res", type=int, nargs=2, default=[2048, 2048])
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
scene = aq.Scene.load(os.path.join(aq.scenes_dir(), a.scene + ".json"))
r = aq.Renderer(0)
r.set_stream(torch.cuda.current_stream().cuda_stream)
ds = r.upload(scene)
ds.accel_wait()
cfg = aq.Integrator(spp=1).cfg(width=a.res[0], height=a.res[1])
cam = ds.camera_rays(cfg, 0)
h = ds.intersect(cam)
ok = h["prim"] != aq.AQ_MISS
P = torch.from_numpy(cam["o"][ok] + h["t"][ok, None] * cam["d"][ok]).cuda()
n = P.shape[0]
g = torch.Generator(device="cuda").manual_seed(1)
d = torch.randn(n, 3, device="cuda", generator=g)
d = d / d.norm(dim=1, keepdim=True)
light = torch.tensor(scene.light_positions()[0] if hasattr(scene, "light_positions") else ([0, 1.7, 0.1] if a.scene == "cbox" else [0, 1.2, 2.0]),
                     device="cuda", dtype=torch.float32)


def pack(o, d, tmax):
    rays = torch.empty(o.shape[0], 8, device="cuda")
    rays[:, 0:3] = o
    rays[:, 3] = 0.0
    rays[:, 4:7] = d
    rays[:, 7] = tmax
    return rays


bounce = pack(P + 1e-3 * d, d, 3.0e38)
to_l = light[None, :] - P
dist = to_l.norm(dim=1)
shadow = pack(P + 1e-3 * to_l / dist[:, None], to_l / dist[:, None], dist * (1 - 1e-3))


def octant(rays):
    dd = rays[:, 4:7]
    return ((dd[:, 0] >= 0).long() | ((dd[:, 1] >= 0).long() << 1) | ((dd[:, 2] >= 0).long() << 2))


def morton(rays, bits=5):
    o = rays[:, 0:3]
    lo, hi = o.min(0).values, o.max(0).values
    q = ((o - lo) / (hi - lo + 1e-9) * ((1 << bits) - 1)).long().clamp(0, (1 << bits) - 1)
    code = torch.zeros(o.shape[0], dtype=torch.long, device="cuda")
    for b in range(bits):
        for ax in range(3):
            code |= ((q[:, ax] >> b) & 1) << (3 * b + ax)
    return code


def local_sort(key, group):
    idx = torch.arange(key.shape[0], device="cuda")
    return torch.argsort((idx // group) * (int(key.max()) + 1) + key, stable=True)


def time_it(rays, any_hit):
    hits = torch.empty(rays.shape[0], 4, device="cuda", dtype=torch.int32)
    best = 1e30
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ds.intersect_device(rays.data_ptr(), rays.shape[0], hits.data_ptr(), any_hit)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return round(rays.shape[0] / best / 1e3, 1)


out = {"scene": a.scene, "rays": n, "unit": "Mrays/s"}
for name, rays, any_hit in (("bounce_closest", bounce, False), ("shadow_any", shadow, True)):
    oc, mo = octant(rays), morton(rays)
    orders = {
        "pixel_order": torch.arange(n, device="cuda"),
        "shuffled": torch.randperm(n, device="cuda", generator=g),
        "octant_within_512": local_sort(oc, 512),
        "octant_within_4096": local_sort(oc, 4096),
        "octant_global": torch.argsort(oc, stable=True),
        "octant_morton_within_4096": local_sort(oc * (1 << 15) + mo, 4096),
        "octant_morton_global": torch.argsort(oc * (1 << 15) + mo, stable=True),
        "morton_octant_global": torch.argsort(mo * 8 + oc, stable=True),
    }
    out[name] = {k: time_it(rays[v].contiguous(), any_hit) for k, v in orders.items()}
print(json.dumps(out))
