"""Does render time depend on WHERE the queue counters live?  (with one contended
atomic per 32 queue entries the same cbox render took 13.4 or 14.0-15.0 ms depending only on the
address of the control block.)  Each scene is created with its control block at another offset of
its allocation (AQUA_DEBUG_CTRL_OFFSET) and the same render is timed; prints min / max / spread."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aqua_engine_b200 as aq

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="cbox")
ap.add_argument("--res", type=int, nargs=2, default=[1024, 1024])
ap.add_argument("--spp", type=int, default=32)
ap.add_argument("--offsets", type=int, default=16)
ap.add_argument("--reps", type=int, default=4)
a = ap.parse_args()
scene = aq.Scene.load(os.path.join(aq.scenes_dir(), a.scene + ".json"))
r = aq.Renderer(0)
cfg = aq.Integrator(spp=a.spp, max_depth=5).cfg(width=a.res[0], height=a.res[1])
rows = []
for k in range(a.offsets):
    off = k * 4096 + (k % 4) * 1024
    os.environ["AQUA_DEBUG_CTRL_OFFSET"] = str(off)
    ds = r.upload(scene)
    ds.accel_wait()
    best = 1e30
    for _ in range(a.reps):
        ds.render_device_async(cfg)
        best = min(best, ds.finish()["ms_total"])
    rows.append((off, best))
    ds.close()
ms = [m for _, m in rows]
print(json.dumps({"lib": os.environ.get("AQUA_CUDA_LIB", "libaqua_cuda.so"), "scene": a.scene, "spp": a.spp,
                  "offsets": len(rows), "ms_min": round(min(ms), 3), "ms_max": round(max(ms), 3),
                  "spread_pct": round(100 * (max(ms) - min(ms)) / min(ms), 2), "ms": [round(m, 3) for m in ms]}))
