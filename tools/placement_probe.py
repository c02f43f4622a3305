"""Does the ADDRESS of the control block (queue counters: the target of every claim / compaction
atomic) change the render time?  bench.py's two arms differed by 10 % on identical work; this
prints, for every slot of the ctx slab, the calibration probe's time next to the measured render
time, and what the calibrated allocator picks.  Run on a GPU box."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aqua_engine_b200 as aq

scene_name = sys.argv[1] if len(sys.argv) > 1 else "cbox"
scene = aq.Scene.load(os.path.join(aq.scenes_dir(), scene_name + ".json"))
r = aq.Renderer(0)
cfg = (aq.Integrator(spp=32, max_depth=5).cfg(width=1024, height=1024) if scene_name == "cbox"
       else aq.Integrator(spp=4, max_depth=5).cfg(width=1920, height=1080))


def t(ds, reps=3):
    best = 1e9
    for _ in range(reps):
        ds.render_device_async(cfg)
        best = min(best, ds.finish()["ms_total"])
    return best


keep = [r.upload(scene) for _ in range(6)]
print("calibrated allocator, 6 scenes in a row (ms):", " ".join(f"{t(ds):.2f}" for ds in keep))
os.environ["AQUA_CTRL_PLACEMENT"] = "off"
keep2 = [r.upload(scene) for _ in range(6)]
print("plain cudaMalloc,      6 scenes in a row (ms):", " ".join(f"{t(ds):.2f}" for ds in keep2))
ds = keep2[0]
L = r.lib
L.aq_debug_ctrl_slot.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float)]
probe, rend = [], []
for k in range(64):
    p = C.c_float()
    assert L.aq_debug_ctrl_slot(ds.handle, k, C.byref(p)) == 0
    probe.append(p.value)  # calibration render time of the ranked slots (0 = not ranked)
    rend.append(t(ds, 2))
probe, rend = np.array(probe), np.array(rend)
print("slot probe_ms render_ms")
for k in range(64):
    print(k, f"{probe[k]:.4f} {rend[k]:.2f}")
rk = probe > 0
print("ranked slots:", int(rk.sum()), "corr(calibration, render) =", float(np.corrcoef(probe[rk], rend[rk])[0, 1]))
order = np.nonzero(rk)[0][np.argsort(probe[rk], kind="stable")]
print("render ms of the ranked slots, calibration order:", rend[order].round(2))
print("render ms min/median/max over all slots:", rend.min().round(2), np.median(rend).round(2), rend.max().round(2))
