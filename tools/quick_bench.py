"""Ad-hoc device-side timing of one scene (development aid; bench.py is the contract)."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aqua_engine_b200 as aq


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="cbox")
    ap.add_argument("--res", type=int, nargs=2, default=[1024, 1024])
    ap.add_argument("--spp", type=int, default=64)
    ap.add_argument("--depth", type=int, default=5)
    ap.add_argument("--pool", type=int, default=0)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    scene = aq.Scene.load(os.path.join(aq.scenes_dir(), a.scene + ".json"))
    r = aq.Renderer(0)
    t = time.time()
    ds = r.upload(scene)
    ds.accel_wait()
    print(f"upload+build {time.time()-t:.3f}s  nodes={ds.accel.n_nodes} depth={ds.accel.max_depth} build_ms={ds.accel.build_ms:.1f}")
    cfg = aq.Integrator(spp=a.spp, max_depth=a.depth).cfg(width=a.res[0], height=a.res[1], pool_paths=a.pool)
    for i in range(a.reps):
        ds.render_device_async(cfg)
        st = ds.finish()
        s = st["ms_total"] * 1e-3
        print(f"rep {i}: {st['ms_total']:.2f} ms  samples/s={st['samples']/s:.4g}  sample-bounces/s={st['sample_bounces']/s:.4g} "
              f"Mrays/s={(st['rays_closest']+st['rays_shadow'])/s/1e6:.1f}  launches={st['n_launches']} waves={st['n_waves']} "
              f"bounces/sample={st['sample_bounces']/st['samples']:.3f}")
    if a.out:
        film, st = ds.render(cfg)
        from PIL import Image
        Image.fromarray(aq.tonemap(film)).save(a.out)


if __name__ == "__main__":
    main()
