import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import aqua_engine_b200 as aq
scene = aq.Scene.load(os.path.join(aq.scenes_dir(), "cbox.json"))
r = aq.Renderer(0)
r.set_stream(torch.cuda.current_stream().cuda_stream)
integ = aq.Integrator(spp=1024, max_depth=5)
cfg = integ.cfg(width=1024, height=1024)
pinned = torch.zeros(1024, 1024, 4).pin_memory()
ds = r.upload(scene)
for i in range(3):
    t0 = time.perf_counter(); ds.render_device_async(cfg); st = ds.finish(); t1 = time.perf_counter()
    print("device path", round((t1 - t0) * 1e3, 1), "ms wall", round(st["ms_total"], 1), "ms events")
for i in range(3):
    t0 = time.perf_counter(); ds2 = r.upload(scene); t1 = time.perf_counter()
    _, st = ds2.render(cfg, film=pinned.numpy()); t2 = time.perf_counter()
    ds2.close(); t3 = time.perf_counter()
    print("e2e: upload", round((t1 - t0) * 1e3, 2), "render", round((t2 - t1) * 1e3, 1), "(events", round(st["ms_total"], 1), ") close", round((t3 - t2) * 1e3, 2))
film = np.zeros((1024, 1024, 4), np.float32)
for i in range(2):
    t0 = time.perf_counter(); _, st = ds.render(cfg, film=film); t1 = time.perf_counter()
    print("host-film render on resident scene", round((t1 - t0) * 1e3, 1), "events", round(st["ms_total"], 1))
