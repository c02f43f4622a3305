#!/bin/bash
# compute-sanitizer passes over a small render + intersect (run under gpurun): memcheck, racecheck,
# initcheck, synccheck.  Output: gpurun_out/${TAG:-r02b}_sanitize_*.txt
tag=${TAG:-r02b}
mkdir -p gpurun_out
cat > /tmp/san_job.py <<'PY'
import os, sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np
import aqua_engine_b200 as aq
for name, (w, h, spp) in {"cbox": (96, 96, 2), "room": (64, 36, 1)}.items():
    sc = aq.Scene.load(os.path.join(aq.scenes_dir(), name + ".json"))
    for builder in ("host", "device", "hybrid"):
        os.environ["AQUA_ACCEL_BUILDER"] = builder
        ds = aq.Renderer(0).upload(sc)
        cfg = aq.Integrator(spp=spp, max_depth=4, seed=1).cfg(width=w, height=h, pool_paths=4096, flags=aq.AQ_RENDER_DUMP_SAMPLES)
        film, st = ds.render(cfg)
        if builder == "hybrid":  # once more on the tree the background thread swapped in
            ds.accel_wait()
            film2, _ = ds.render(cfg)
            assert np.array_equal(film, film2)
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
        cfgt = integ.cfg(width=40, height=40, pool_paths=2048, flags=aq.AQ_RENDER_NRC_TENSOR)
        filmt, _ = ds.nrc_render(cfgt, nrc)  # the tcgen05 lookup
        print("full + nrc", info["n_valid"], st["sample_bounces"], float(film[..., :3].sum()), float(filmt[..., :3].sum()))
PY
for tool in ${TOOLS:-memcheck racecheck initcheck synccheck}; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 77 python /tmp/san_job.py > gpurun_out/${tag}_sanitize_$tool.txt 2>&1
  echo "$tool exit=$?" | tee -a gpurun_out/${tag}_sanitize_summary.txt
  tail -3 gpurun_out/${tag}_sanitize_$tool.txt
done
