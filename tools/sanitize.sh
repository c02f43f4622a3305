#!/bin/bash
# compute-sanitizer passes over a small render + intersect (run under gpurun): memcheck, racecheck,
# initcheck, synccheck.  Output: gpurun_out/sanitize_*.log
mkdir -p gpurun_out
cat > /tmp/san_job.py <<'PY'
import os, sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np
import aqua_engine_b200 as aq
for name, (w, h, spp) in {"cbox": (96, 96, 2), "room": (64, 36, 1)}.items():
    sc = aq.Scene.load(os.path.join(aq.scenes_dir(), name + ".json"))
    for builder in ("host", "device"):
        os.environ["AQUA_ACCEL_BUILDER"] = builder
        ds = aq.Renderer(0).upload(sc)
        cfg = aq.Integrator(spp=spp, max_depth=4, seed=1).cfg(width=w, height=h, pool_paths=4096, flags=aq.AQ_RENDER_DUMP_SAMPLES)
        film, st = ds.render(cfg)
        rays = ds.camera_rays(cfg, 0)
        hits = ds.intersect(rays)
        occ = ds.intersect(rays, any_hit=True)
        print(name, builder, st["sample_bounces"], int((hits["prim"] != aq.AQ_MISS).sum()), float(film[..., :3].sum()))
    if name == "cbox":  # the full-Principled shade instantiation and the nrc integrator (tiny sizes)
        cfgf = aq.Integrator(spp=1, max_depth=4, seed=1).cfg(width=48, height=48, pool_paths=4096, flags=aq.AQ_RENDER_FORCE_FULL_BSDF)
        film, st = ds.render(cfgf)
        integ = aq.Integrator(spp=1, max_depth=3, seed=2, type="nrc", batch_size=96, training_iters=3)
        cfgn, nrc = integ.cfg(width=40, height=40, pool_paths=2048), integ.nrc_cfg()
        info = ds.nrc_train(cfgn, nrc)
        film, st = ds.nrc_render(cfgn, nrc)
        print("full + nrc", info["n_valid"], st["sample_bounces"], float(film[..., :3].sum()))
PY
export AQUA_CTRL_PLACEMENT=off  # (the slot ranking would run 33 calibration renders under the sanitizer)
for tool in memcheck racecheck initcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 77 python /tmp/san_job.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool exit=$?" | tee -a gpurun_out/sanitize_summary.log
  tail -3 gpurun_out/sanitize_$tool.log
done
