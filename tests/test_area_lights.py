"""Row (f)-1 of SURVEY §8: mesh area lights (Bsdf::Principled.emission > 0, scenes/cbox.json:61-63;
the source MTL has `Ke 17 12 4` on the cbox light, CornellBox-Original.mtl:88) with next-event
estimation + multiple importance sampling.  CPU oracle checks here; GPU parity at the end."""
import copy
import ctypes as C

import numpy as np
import pytest

from test_oracle import bsdf_f64


def emissive_cbox(aq, cbox, emission=(17.0, 12.0, 4.0), keep_point_light=False):
    """cbox.json with the `light` material's emission restored from the MTL."""
    pos, idx, nrm, uv, tm = cbox.arrays()
    mats = []
    for k, name in enumerate(cbox.material_names()):
        m = type(cbox.desc.materials[k])()
        C.memmove(C.byref(m), C.byref(cbox.desc.materials[k]), C.sizeof(m))
        if name == "light":
            m.emission[:] = emission
        mats.append(m)
    lights = [aq.point_light((0.0, 1.7, 0.1), (1.0, 1.0, 1.0))] if keep_point_light else []
    cam = type(cbox.desc.camera)()
    C.memmove(C.byref(cam), C.byref(cbox.desc.camera), C.sizeof(cam))
    return aq.Scene.from_arrays(pos.copy(), idx.copy(), normals=nrm.copy(), tri_material=tm.copy(),
                                materials=mats, lights=lights, camera=cam)


def mean_rgb(film):
    return (film[..., :3] / film[..., 3:]).reshape(-1, 3).mean(axis=0)


def lamp_room(aq, s=0.6):
    """Floor + back wall + a square emitter: smooth direct and indirect light, no 1/d^2 hot spots
    (in cbox.json the light quad hangs 1 cm under the ceiling, which makes NEE-only very noisy)."""
    pos = np.array([[-6, 0, -2], [6, 0, -2], [6, 0, 6], [-6, 0, 6],          # floor
                    [-6, 0, -2], [6, 0, -2], [6, 5, -2], [-6, 5, -2],        # back wall
                    [-s, 2, 1 - s], [s, 2, 1 - s], [s, 2, 1 + s], [-s, 2, 1 + s]], np.float32)
    idx = np.array([[0, 2, 1], [0, 3, 2], [4, 5, 6], [4, 6, 7], [8, 9, 10], [8, 10, 11]], np.uint32)
    floor = aq.default_material(color=(0.7, 0.6, 0.5), roughness=0.5)
    wall = aq.default_material(color=(0.3, 0.5, 0.8), roughness=0.3, metallic=0.5)
    lamp = aq.default_material(color=(0.0, 0.0, 0.0), roughness=0.5)
    lamp.emission[:] = (6.0, 5.0, 4.0)
    cam = aq.default_camera(res=(32, 32), fov=50.0, translate=(0.0, 1.2, 5.5))
    cam.rotate[:] = (-0.15, 0.0, 0.0)
    return aq.Scene.from_arrays(pos, idx, materials=[floor, wall, lamp], tri_material=[0, 0, 1, 1, 2, 2], camera=cam)


def test_three_estimators_agree(aq, ao):
    sc = lamp_room(aq)
    o = ao.OracleScene(sc)
    res = {}
    # NEE at vertex k reaches the light with k+1 segments, a BSDF-sampled hit needs the light at
    # vertex k+1: BSDF-only with max_depth d+1 covers the same path set as NEE-only / MIS with d
    for name, fl, md in (("mis", 0, 3), ("nee", aq.AQ_RENDER_MIS_NEE_ONLY, 3), ("bsdf", aq.AQ_RENDER_MIS_BSDF_ONLY, 4)):
        integ = aq.Integrator(spp=256, max_depth=md, seed=1)
        film, _, st = o.render(integ.cfg(width=32, height=32, flags=fl))
        img = film[..., :3] / film[..., 3:]
        res[name] = (mean_rgb(film), img)
        assert np.isfinite(film).all() and img.min() >= 0
    assert np.allclose(res["mis"][0], res["nee"][0], rtol=0.02), (res["mis"][0], res["nee"][0])
    assert np.allclose(res["mis"][0], res["bsdf"][0], rtol=0.04), (res["mis"][0], res["bsdf"][0])
    assert res["mis"][0][0] > 0.01
    # MIS beats the worse single strategy (per-pixel error against the average of all three)
    ref = (res["mis"][1] + res["nee"][1] + res["bsdf"][1]) / 3
    err = {k: float(np.mean((v[1] - ref) ** 2)) for k, v in res.items()}
    assert err["mis"] < max(err["nee"], err["bsdf"])


def test_emissive_cbox_renders(aq, ao, cbox):
    """cbox.json with the MTL's emitter restored: finite, positive, orange-ish, NEE active."""
    sc = emissive_cbox(aq, cbox)
    o = ao.OracleScene(sc)
    film, _, st = o.render(aq.Integrator(spp=16, max_depth=4, seed=1).cfg(width=48, height=48))
    m = mean_rgb(film)
    assert np.isfinite(film).all() and m[0] > m[1] > m[2] > 0 and st["rays_shadow"] > 0


def test_directly_visible_emitter_and_light_table(aq, ao, cbox):
    sc = emissive_cbox(aq, cbox)
    o = ao.OracleScene(sc)
    cfg = aq.Integrator(spp=1, max_depth=1, seed=0).cfg(width=64, height=64)
    rays = o.camera_rays(cfg, 0)
    hits = o.intersect(rays)
    film, samples, st = o.render(cfg, want_samples=True)
    on_light = np.isin(hits["prim"], [34, 35])
    if on_light.any():  # the light quad faces down; seen from the camera only at grazing angles
        assert np.allclose(samples[0].reshape(-1, 4)[on_light, :3], [17, 12, 4])
    # max_depth 1: floor pixels outside the boxes' shadows receive NEE from the quad (no point light here)
    floor = np.isin(hits["prim"], [0, 1])
    assert floor.any() and (samples[0].reshape(-1, 4)[floor, 0] > 0).mean() > 0.5
    assert st["rays_shadow"] > 0


def test_closed_form_square_light_over_floor(aq, ao):
    """max_depth 1, square emitter over a floor: L = integral over the light of
    f(wo,wi) cos_i * Le * cos_l / d^2 dA, integrated numerically in float64."""
    s = 0.5
    pos = np.array([[-20, 0, -20], [20, 0, -20], [20, 0, 20], [-20, 0, 20],
                    [-s, 2, -s], [s, 2, -s], [s, 2, s], [-s, 2, s]], np.float32)
    idx = np.array([[0, 2, 1], [0, 3, 2], [4, 5, 6], [4, 6, 7]], np.uint32)
    rho, rough, Le = 0.7, 0.6, 5.0
    floor = aq.default_material(color=(rho, rho, rho), roughness=rough)
    lamp = aq.default_material(color=(0.0, 0.0, 0.0), roughness=0.5)
    lamp.emission[:] = (Le, Le, Le)
    cam = aq.default_camera(res=(4, 4), fov=2.0, translate=(0.3, 1.0, 3.0))
    cam.rotate[:] = (-0.5, 0.0, 0.0)
    sc = aq.Scene.from_arrays(pos, idx, materials=[floor, lamp], tri_material=[0, 0, 1, 1], camera=cam)
    o = ao.OracleScene(sc)
    cfg = aq.Integrator(spp=4096, max_depth=1, seed=3).cfg(width=4, height=4)
    film, _, st = o.render(cfg)
    got = mean_rgb(film)[0]
    # reference: quadrature at the mean hit point / direction of the (tiny) field of view
    rays = o.camera_rays(aq.Integrator(spp=1).cfg(width=4, height=4), 0)
    hits = o.intersect(rays)
    assert (hits["prim"] < 2).all()
    P = (rays["o"].astype(float) + hits["t"][:, None] * rays["d"].astype(float)).mean(axis=0)
    wo = -rays["d"].astype(float).mean(axis=0)
    wo /= np.linalg.norm(wo)
    n = 400
    xs = (np.arange(n) + 0.5) / n * 2 * s - s
    X, Z = np.meshgrid(xs, xs)
    Y = np.stack([X.ravel(), np.full(n * n, 2.0), Z.ravel()], 1)
    dl = Y - P
    d2 = (dl ** 2).sum(1)
    wi = dl / np.sqrt(d2)[:, None]
    cos_l = np.abs(wi[:, 1])
    to_local = lambda w: np.array([w[0], w[2], w[1]])
    params = [rho, rho, rho, 0.0, rough, 0.0, 0.0, 0.0, 0.5, 0.0]
    acc = 0.0
    step = 16  # subsample the BSDF evaluation grid (smooth integrand)
    sub = np.arange(0, n * n, step)
    for k in sub:
        fcos, _ = bsdf_f64(params, to_local(wo), to_local(wi[k]))
        acc += fcos[0] * Le * cos_l[k] / d2[k]
    want = acc / len(sub) * (2 * s) ** 2
    assert abs(got / want - 1) < 0.02, (got, want)


def lamp_room_with(aq, lights, emission):
    sc = lamp_room(aq)
    pos, idx, nrm, uv, tm = sc.arrays()
    mats = []
    for k in range(3):
        m = copy.copy(sc.desc.materials[k])
        if k == 2:
            m.emission[:] = emission
        mats.append(m)
    return aq.Scene.from_arrays(pos.copy(), idx.copy(), tri_material=tm.copy(), materials=mats, lights=lights,
                                camera=sc.desc.camera)


def test_power_proportional_light_pick_is_unbiased(aq, ao):
    """A bright and a dim point light plus the emissive quad, picked by power: the render equals
    the sum of the three single-light renders (linearity of light transport)."""
    la = aq.point_light((1.5, 1.5, 2.0), (4.0, 4.0, 4.0))
    lb = aq.point_light((-2.0, 0.6, 1.0), (0.2, 0.4, 0.2))
    em = (6.0, 5.0, 4.0)
    cfg = aq.Integrator(spp=256, max_depth=2, seed=5).cfg(width=32, height=32)
    r = lambda L, e: mean_rgb(ao.OracleScene(lamp_room_with(aq, L, e)).render(cfg)[0])
    full = r([la, lb], em)
    parts = r([la], (0, 0, 0)) + r([lb], (0, 0, 0)) + r([], em)
    assert np.allclose(full, parts, rtol=0.03), (full, parts)
    assert np.allclose(r([la, lb], (0, 0, 0)), r([la], (0, 0, 0)) + r([lb], (0, 0, 0)), rtol=0.02)


@pytest.mark.gpu
def test_gpu_parity_with_area_lights(aq, ao, cbox, renderer):
    for keep in (False, True):
        sc = emissive_cbox(aq, cbox, keep_point_light=keep)
        ds, o = renderer.upload(sc), ao.OracleScene(sc)
        for fl in (0, aq.AQ_RENDER_MIS_NEE_ONLY, aq.AQ_RENDER_MIS_BSDF_ONLY):
            cfg = aq.Integrator(spp=6, max_depth=5, seed=2).cfg(width=96, height=96, flags=fl | aq.AQ_RENDER_DUMP_SAMPLES)
            film, st = ds.render(cfg)
            samples = ds.samples(cfg)
            ofilm, osamples, ost = o.render(cfg, want_samples=True)
            err = np.abs(samples - osamples).max() / max(1e-20, np.abs(osamples).max())
            assert err <= 1e-4, err
            assert np.array_equal(samples, osamples) and np.array_equal(film, ofilm)
            assert st["rays_shadow"] == ost["rays_shadow"] and st["sample_bounces"] == ost["sample_bounces"]
