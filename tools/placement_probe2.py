"""Follow-up of placement_probe.py: (a) which stage slows down on a slow control-block slot,
(b) is a slot's speed a property of the slot alone (same pattern for a second scene object and for
room.json)?"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aqua_engine_b200 as aq

r = aq.Renderer(0)
L = r.lib
L.aq_debug_ctrl_slot.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float)]


def sweep(ds, cfg, slots, prof=False):
    out = []
    for k in slots:
        assert L.aq_debug_ctrl_slot(ds.handle, int(k), None) == 0
        best = None
        for _ in range(2):
            ds.render_device_async(cfg)
            st = ds.finish()
            if best is None or st["ms_total"] < best["ms_total"]:
                best = st
        out.append(best)
    return out


cbox = aq.Scene.load(os.path.join(aq.scenes_dir(), "cbox.json"))
room = aq.Scene.load(os.path.join(aq.scenes_dir(), "room.json"))
cfgc = aq.Integrator(spp=32, max_depth=5).cfg(width=1024, height=1024)
cfgp = aq.Integrator(spp=32, max_depth=5).cfg(width=1024, height=1024, flags=aq.AQ_RENDER_PROFILE)
cfgr = aq.Integrator(spp=4, max_depth=5).cfg(width=1920, height=1080)
slots = list(range(24))
a, b = r.upload(cbox), r.upload(cbox)
ta = np.array([s["ms_total"] for s in sweep(a, cfgc, slots)])
tb = np.array([s["ms_total"] for s in sweep(b, cfgc, slots)])
rm = r.upload(room)
tr = np.array([s["ms_total"] for s in sweep(rm, cfgr, slots)])
print("cbox scene A:", ta.round(2))
print("cbox scene B:", tb.round(2))
print("room        :", tr.round(2))
print("corr A/B", float(np.corrcoef(ta, tb)[0, 1]), "corr A/room", float(np.corrcoef(ta, tr)[0, 1]))
fast, slow = int(np.argmin(ta)), int(np.argmax(ta))
for name, k in (("fast", fast), ("slow", slow)):
    st = sweep(a, cfgp, [k])[0]
    print(name, "slot", k, {x: round(st[x], 2) for x in ("ms_total", "ms_raygen", "ms_trace", "ms_shade", "ms_shadow", "ms_film")})
