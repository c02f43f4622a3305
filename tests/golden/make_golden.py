"""Regenerates tests/golden/*.npz from the CPU oracle (run in the dev container).

The reference holds NO golden vectors for this path (src/lib.rs:0 is empty), so
these fixtures pin the oracle's own semantics against accidental drift, and give the GPU
tests committed vectors to compare with.  They are not reference outputs."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import aq_oracle as ao  # noqa: E402
import aqua_engine_b200 as aq  # noqa: E402


def main():
    out = os.path.dirname(os.path.abspath(__file__))
    sc = aq.Scene.load(os.path.join(aq.scenes_dir(), "cbox.json"))
    o = ao.OracleScene(sc)
    integ = aq.Integrator(spp=4, max_depth=5, seed=0)
    cfg = integ.cfg(width=64, height=64)
    rays = o.camera_rays(cfg, 0)
    hits = o.intersect(rays)
    cfg2 = integ.cfg(width=32, height=32)
    film, samples, st = o.render(cfg2, want_samples=True)
    np.savez_compressed(os.path.join(out, "cbox_golden.npz"), rays=rays.view(np.float32).reshape(-1, 8),
                        hit_prim=hits["prim"], hit_t=hits["t"], film=film, samples=samples,
                        sample_bounces=np.uint64(st["sample_bounces"]), rays_shadow=np.uint64(st["rays_shadow"]))
    room = aq.Scene.load(os.path.join(aq.scenes_dir(), "room.json"))
    orm = ao.OracleScene(room, build_bvh=True)
    cfgr = aq.Integrator(spp=2, max_depth=5, seed=3).cfg(width=48, height=27)
    rr = orm.camera_rays(cfgr, 0)
    hr = orm.intersect(rr, mode=0)  # brute force over all 394,269 triangles
    filmr, _, str_ = orm.render(cfgr, mode=1)
    np.savez_compressed(os.path.join(out, "room_golden.npz"), hit_prim=hr["prim"], hit_t=hr["t"], film=filmr,
                        sample_bounces=np.uint64(str_["sample_bounces"]))
    print("wrote golden fixtures:", st, str_)


if __name__ == "__main__":
    main()
