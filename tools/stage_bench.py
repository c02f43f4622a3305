"""Per-stage device times of one render (AQ_RENDER_PROFILE) — development aid."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aqua_engine_b200 as aq

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="cbox")
ap.add_argument("--res", type=int, nargs=2, default=[1024, 1024])
ap.add_argument("--spp", type=int, default=64)
ap.add_argument("--pool", type=int, default=0)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--full-bsdf", action="store_true", help="force the full-Principled shade instantiation (same image)")
a = ap.parse_args()
scene = aq.Scene.load(os.path.join(aq.scenes_dir(), a.scene + ".json"))
ds = aq.Renderer(0).upload(scene)
ds.accel_wait()  # hybrid build: time the final (SAH) tree
for prof in (0, aq.AQ_RENDER_PROFILE):
    cfg = aq.Integrator(spp=a.spp, max_depth=5).cfg(width=a.res[0], height=a.res[1], pool_paths=a.pool,
                                                    flags=prof | (aq.AQ_RENDER_FORCE_FULL_BSDF if a.full_bsdf else 0))
    best = None
    for _ in range(a.reps):
        ds.render_device_async(cfg)
        st = ds.finish()
        if best is None or st["ms_total"] < best["ms_total"]:
            best = st
    st = best
    s = st["ms_total"] * 1e-3
    print(f"{os.environ.get('AQUA_CUDA_LIB', 'base')} {a.scene} pool={a.pool} full={int(a.full_bsdf)} prof={prof}: {st['ms_total']:.2f} ms  sb/s={st['sample_bounces']/s:.4g} Mrays/s={(st['rays_closest']+st['rays_shadow'])/s/1e6:.0f} "
          f"| raygen {st['ms_raygen']:.2f} closest {st['ms_trace']:.2f} shade {st['ms_shade']:.2f} shadow {st['ms_shadow']:.2f} film {st['ms_film']:.2f}")
