"""The CPU oracle against (i) independent float64 numpy restatements of each definitional
function, (ii) analytic properties: furnace, pdf normalisation, sampling
weight consistency, reciprocity, closed-form one-bounce radiance, (iii) its own brute-force
intersector, (iv) the committed golden fixtures.  No GPU."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import hits_equal, random_rays, triangle_soup

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


# ---------------------------------------------------------------- scalar definitions
def test_sincos_polynomial(ao):
    L = ao.lib()
    us = np.concatenate([np.linspace(0, 1, 4001, endpoint=False), [0.125, 0.25, 0.5, 0.75, 0.999999]]).astype(np.float32)
    s, c = C.c_float(), C.c_float()
    err = 0.0
    for u in us:
        L.aqo_sincos_2pi(float(u), C.byref(s), C.byref(c))
        err = max(err, abs(s.value - np.sin(2 * np.pi * float(u))), abs(c.value - np.cos(2 * np.pi * float(u))))
    assert err < 4e-7


def test_rng_is_a_pure_function_and_uniform(ao):
    L = ao.lib()
    k = L.aqo_rng_key(0, 12345, 7)
    assert k == L.aqo_rng_key(0, 12345, 7) and k != L.aqo_rng_key(1, 12345, 7) and k != L.aqo_rng_key(0, 12346, 7)

    # independent restatement of the PCG hash
    def pcg(v):
        s = (v * 747796405 + 2891336453) & 0xFFFFFFFF
        w = (((s >> ((s >> 28) + 4)) ^ s) * 277803737) & 0xFFFFFFFF
        return (w >> 22) ^ w
    assert k == pcg((7 + pcg((12345 + pcg(0)) & 0xFFFFFFFF)) & 0xFFFFFFFF)
    xs = np.array([L.aqo_rng(L.aqo_rng_key(0, p, 0), d) for p in range(2000) for d in range(8)])
    assert xs.min() >= 0.0 and xs.max() < 1.0
    assert abs(xs.mean() - 0.5) < 0.01 and abs(xs.var() - 1 / 12) < 0.005
    assert abs(L.aqo_rng(k, 3) - (pcg((k + 3) & 0xFFFFFFFF) >> 8) / 16777216.0) == 0


def mt_f64(o, d, v0, v1, v2):
    e1, e2 = v1 - v0, v2 - v0
    p = np.cross(d, e2)
    det = e1 @ p
    if det == 0:
        return None
    t_ = o - v0
    u = (t_ @ p) / det
    q = np.cross(t_, e1)
    v = (d @ q) / det
    t = (e2 @ q) / det
    return t, u, v


def test_triangle_test_against_float64(ao):
    L = ao.lib()
    g = np.random.default_rng(5)
    n_hit = 0
    for _ in range(3000):
        v = g.uniform(-1, 1, (3, 3)).astype(np.float32)
        o = g.uniform(-2, 2, 3).astype(np.float32)
        tgt = (v[0] * 0.3 + v[1] * 0.3 + v[2] * 0.4 + g.normal(0, 0.6, 3)).astype(np.float32)
        d = tgt - o
        d = (d / np.linalg.norm(d)).astype(np.float32)
        out = np.zeros(3, np.float32)
        hit = L.aqo_tri_test(fptr(o), fptr(d), 0.0, fptr(v[0]), fptr(v[1]), fptr(v[2]), fptr(out))
        ref = mt_f64(o.astype(np.float64), d.astype(np.float64), *v.astype(np.float64))
        inside = ref is not None and ref[0] > 0 and ref[1] >= 0 and ref[2] >= 0 and ref[1] + ref[2] <= 1
        margin = 1e-4 if ref is None else min(abs(ref[1]), abs(ref[2]), abs(1 - ref[1] - ref[2]), abs(ref[0]))
        if margin > 1e-4:  # away from edges the decision must agree
            assert bool(hit) == bool(inside)
        if hit and inside:
            n_hit += 1
            assert np.allclose(out, ref, rtol=2e-4, atol=2e-5)
    assert n_hit > 300


# ---------------------------------------------------------------- Principled BSDF
def bsdf_f64(p, wo, wi):
    """Independent float64 restatement of DESIGN.md's Principled definition -> (f*cos, pdf)."""
    base, metallic, rough, spec, stint, sheen, shtint, trans = np.array(p[:3], float), *[float(x) for x in p[3:10]]
    if wi[2] <= 0 or wo[2] <= 0:
        return None
    h = wo + wi
    h = h / np.linalg.norm(h)
    ldh = wi @ h
    lum = 0.2126 * base[0] + 0.7152 * base[1] + 0.0722 * base[2]
    tint = base / lum if lum > 0 else np.ones(3)
    f0 = (1 - metallic) * (0.08 * spec * ((1 - stint) + stint * tint)) + metallic * base
    f0 = 0.08 * spec * (1 + stint * (tint - 1))
    f0 = f0 + metallic * (base - f0)
    dw = (1 - metallic) * (1 - trans)
    alpha = max(rough * rough, 1e-4)
    Fo = f0 + (1 - f0) * (1 - min(max(wo[2], 0), 1)) ** 5
    ws, wd = Fo.max(), dw * base.max()
    ps = ws / (ws + wd) if ws + wd > 0 else 0.0
    f = np.zeros(3)
    pdf = 0.0
    if dw > 0:
        fd90 = 0.5 + 2 * rough * ldh * ldh
        fd = (1 + (fd90 - 1) * (1 - wi[2]) ** 5) * (1 + (fd90 - 1) * (1 - wo[2]) ** 5)
        sh = sheen * (1 + shtint * (tint - 1)) * (1 - ldh) ** 5
        f += dw * (base / np.pi * fd + sh)
        pdf += (1 - ps) * wi[2] / np.pi
    if ps > 0:
        a2 = alpha * alpha
        D = a2 / (np.pi * (h[2] ** 2 * (a2 - 1) + 1) ** 2)
        lam = lambda c: 0.5 * (np.sqrt(1 + a2 * (1 - c * c) / (c * c)) - 1)
        G = 1 / (1 + lam(wo[2]) + lam(wi[2]))
        F = f0 + (1 - f0) * (1 - ldh) ** 5
        f += F * D * G / (4 * wo[2] * wi[2])
        pdf += ps * D / (1 + lam(wo[2])) / (4 * wo[2])
    return f * wi[2], pdf


MATS = {
    "cbox_diffuse": [0.73, 0.71, 0.68, 0.0, 0.4, 0.0, 0.0, 0.0, 0.5, 0.0],
    "cbox_tallbox_metal": [0.73, 0.71, 0.68, 1.0, 0.1, 0.0, 0.0, 0.0, 0.5, 0.0],
    "room_floor": [0.4, 0.2, 0.1, 0.1, 0.0535, 0.0, 0.0, 0.0, 0.5, 0.0],
    "mixed_spec_sheen": [0.8, 0.3, 0.2, 0.5, 0.35, 0.5, 0.5, 0.3, 0.5, 0.0],
    "rough_metal": [0.9, 0.9, 0.9, 1.0, 0.8, 0.0, 0.0, 0.0, 0.5, 0.0],
}


def dirs(n, seed):
    g = np.random.default_rng(seed)
    z = g.uniform(0.02, 1, n)
    ph = g.uniform(0, 2 * np.pi, n)
    r = np.sqrt(1 - z * z)
    return np.stack([r * np.cos(ph), r * np.sin(ph), z], 1).astype(np.float32)


@pytest.mark.parametrize("name", list(MATS))
def test_bsdf_eval_against_float64(ao, name):
    L = ao.lib()
    p = np.array(MATS[name], np.float32)
    for wo, wi in zip(dirs(400, 1), dirs(400, 2)):
        f = np.zeros(3, np.float32)
        pdf = C.c_float()
        ok = L.aqo_bsdf_eval(fptr(p), fptr(wo), fptr(wi), fptr(f), C.byref(pdf))
        ref = bsdf_f64(p.astype(np.float64), wo.astype(np.float64), wi.astype(np.float64))
        assert ok and ref is not None
        assert np.allclose(f, ref[0], rtol=2e-3, atol=1e-6), (f, ref[0])
        assert np.isclose(pdf.value, ref[1], rtol=2e-3, atol=1e-6)


@pytest.mark.parametrize("name", list(MATS))
def test_bsdf_sampling_consistency_furnace_and_pdf_normalisation(ao, name):
    L = ao.lib()
    p = np.array(MATS[name], np.float32)
    g = np.random.default_rng(11)
    for wo in dirs(3, 7):
        wsum = np.zeros(3)
        n = 4000
        n_ok = 0
        for _ in range(n):
            u = g.uniform(0, 1, 3).astype(np.float32)
            wi, w = np.zeros(3, np.float32), np.zeros(3, np.float32)
            pdf = C.c_float()
            if not L.aqo_bsdf_sample(fptr(p), fptr(wo), fptr(u), fptr(wi), fptr(w), C.byref(pdf)):
                continue
            n_ok += 1
            assert wi[2] > 0 and abs(np.linalg.norm(wi) - 1) < 1e-4
            f = np.zeros(3, np.float32)
            pdf2 = C.c_float()
            assert L.aqo_bsdf_eval(fptr(p), fptr(wo), fptr(wi), fptr(f), C.byref(pdf2))
            assert np.isclose(pdf.value, pdf2.value, rtol=1e-5)           # sample() reports eval()'s pdf
            assert np.allclose(w, f / pdf2.value, rtol=1e-5, atol=1e-7)   # weight = f cos / pdf
            wsum += w
        albedo = wsum / n
        assert (albedo <= 1.05).all(), albedo  # white furnace: no energy gain
        # pdf integrates to <= 1 over the hemisphere (VNDF mass below the horizon is lost)
        m = 40000
        z = g.uniform(0, 1, m)
        ph = g.uniform(0, 2 * np.pi, m)
        r = np.sqrt(1 - z * z)
        wis = np.stack([r * np.cos(ph), r * np.sin(ph), z], 1).astype(np.float32)
        acc = 0.0
        for wi in wis[:: 8 if p[4] > 0.3 else 1][:5000]:
            f = np.zeros(3, np.float32)
            pdf = C.c_float()
            if L.aqo_bsdf_eval(fptr(p), fptr(wo), fptr(wi), fptr(f), C.byref(pdf)):
                acc += pdf.value
        if p[4] > 0.3:  # smooth lobes need importance sampling to integrate; check rough ones only
            est = acc / 5000 * 2 * np.pi
            # the pdf integrates to the probability that a sample lands above the horizon
            # (VNDF mass reflected below it is lost, more so at grazing wo)
            assert abs(est - n_ok / n) < 0.05 and est < 1.05, (est, n_ok / n, wo)


def test_bsdf_reciprocity(ao):
    L = ao.lib()
    p = np.array(MATS["rough_metal"], np.float32)
    for wo, wi in zip(dirs(200, 3), dirs(200, 4)):
        fa, fb = np.zeros(3, np.float32), np.zeros(3, np.float32)
        pa, pb = C.c_float(), C.c_float()
        assert L.aqo_bsdf_eval(fptr(p), fptr(wo), fptr(wi), fptr(fa), C.byref(pa))
        assert L.aqo_bsdf_eval(fptr(p), fptr(wi), fptr(wo), fptr(fb), C.byref(pb))
        assert np.allclose(fa / wi[2], fb / wo[2], rtol=1e-4)  # f(wo,wi) == f(wi,wo)


# ---------------------------------------------------------------- whole-path checks
def test_closed_form_one_bounce_radiance(aq, ao):
    """Point light over a floor, max_depth 1: L = f(wo,wi) cos(theta_i) * I / d^2, with f from
    the independent float64 restatement above (Burley diffuse + the Schlick/GGX grazing term)."""
    pos = np.array([[-50, 0, -50], [50, 0, -50], [50, 0, 50], [-50, 0, 50]], np.float32)
    idx = np.array([[0, 2, 1], [0, 3, 2]], np.uint32)
    rho, rough, I, hgt = 0.6, 0.5, 10.0, 2.0
    mat = aq.default_material(color=(rho, rho, rho), roughness=rough)
    params = [rho, rho, rho, 0.0, rough, 0.0, 0.0, 0.0, 0.5, 0.0]
    cam = aq.default_camera(res=(8, 8), fov=20.0, translate=(0, 1.0, 3))
    cam.rotate[:] = (-0.6, 0.0, 0.0)  # pitch down towards the floor
    sc = aq.Scene.from_arrays(pos, idx, materials=[mat], lights=[aq.point_light((0, hgt, 0), (I, I, I))], camera=cam)
    o = ao.OracleScene(sc)
    cfg = aq.Integrator(spp=1, max_depth=1).cfg(width=8, height=8)
    rays = o.camera_rays(cfg, 0)
    hits = o.intersect(rays)
    film, samples, st = o.render(cfg, want_samples=True)
    assert (hits["prim"] != aq.AQ_MISS).all() and st["rays_shadow"] == 64
    for k in range(len(rays)):
        P = rays["o"][k].astype(float) + hits["t"][k] * rays["d"][k].astype(float)
        assert abs(P[1]) < 1e-5
        wo = -rays["d"][k].astype(float)
        Lv = np.array([0, hgt, 0.0]) - P
        d2 = Lv @ Lv
        wi = Lv / np.sqrt(d2)
        to_local = lambda w: np.array([w[0], w[2], w[1]])  # floor normal +y -> local +z
        fcos, _ = bsdf_f64(params, to_local(wo), to_local(wi))
        want = fcos * I / d2
        got = samples[0].reshape(-1, 4)[k, :3]
        assert np.allclose(got, want, rtol=5e-4), (k, got, want)
        lambert = rho / np.pi * wi[1] * I / d2
        assert abs(got[0] / lambert - 1) < 0.1  # and it is within 10% of plain Lambert


def test_camera_convention(aq, ao, cbox):
    """Looks down -z, +y up, pixel (0,0) top-left, fov = full angle across the larger side."""
    o = ao.OracleScene(cbox)
    cfg = aq.Integrator(spp=1).cfg(width=64, height=32)
    r = o.camera_rays(cfg, 0).reshape(32, 64)
    assert np.allclose(r["o"], [0, 1, 9])
    assert (r["d"][..., 2] < -0.98).all()
    assert (r["d"][:, 0, 0] < 0).all() and (r["d"][:, -1, 0] > 0).all()      # x grows to the right
    assert (r["d"][0, :, 1] > 0).all() and (r["d"][-1, :, 1] < 0).all()      # row 0 is the top
    half = np.degrees(np.arctan2(np.abs(r["d"][16, 0, 0]), -r["d"][16, 0, 2]))
    assert 7.0 < half < 7.5                                                   # 15 degrees across the width
    halfv = np.degrees(np.arctan2(np.abs(r["d"][0, 32, 1]), -r["d"][0, 32, 2]))
    assert 3.4 < halfv < 3.8


def test_oracle_bvh2_and_product_bvh8_equal_brute_force(aq, ao, cbox):
    o = ao.OracleScene(cbox, build_bvh=True)
    rays = random_rays(aq, 50000, [-1.2, -0.2, -1.2], [1.2, 2.2, 1.2], seed=3)
    hb = o.intersect(rays, mode=0)
    assert hits_equal(o.intersect(rays, mode=1), hb)
    pos, idx, *_ = cbox.arrays()
    nodes, tris, info = aq.build_accel_host(pos, idx)
    assert info.n_tri_records == 36 and nodes.shape[1] * 4 == 80 and tris.shape[1] * 4 == 48
    h8, nn, nt = ao.bvh8_intersect(nodes, tris, rays)
    assert hits_equal(h8, hb)
    # duplicated triangles: the smaller id of each coincident pair must win
    dup_hi = {20, 21, 32, 33}
    assert not (set(np.unique(hb["prim"]).tolist()) & dup_hi)
    assert {16, 17, 30, 31} & set(np.unique(hb["prim"]).tolist())
    # any-hit
    rays["tmax"] = np.where(np.arange(len(rays)) % 2 == 0, 0.7, 3e38).astype(np.float32)
    ab = o.intersect(rays, any_hit=True, mode=0)
    a8, _, _ = ao.bvh8_intersect(nodes, tris, rays, any_hit=True)
    assert np.array_equal(ab["prim"], a8["prim"]) and np.array_equal(ab["prim"], o.intersect(rays, any_hit=True, mode=1)["prim"])


def test_bvh8_on_triangle_soup_equals_brute_force(aq, ao):
    pos, idx = triangle_soup(60000, r=0.02)
    sc = aq.Scene.from_arrays(pos, idx)
    o = ao.OracleScene(sc, build_bvh=True)
    nodes, tris, info = aq.build_accel_host(pos, idx)
    assert info.max_depth < 64
    # every triangle appears exactly once in the records
    prims = tris.view(np.uint32)[:, 9]
    assert np.array_equal(np.sort(prims), np.arange(60000, dtype=np.uint32))
    rays = random_rays(aq, 4096, 0.0, 1.0, seed=9)
    hb = o.intersect(rays, mode=0)
    h8, nn, nt = ao.bvh8_intersect(nodes, tris, rays)
    assert hits_equal(h8, hb) and hits_equal(o.intersect(rays, mode=1), hb)
    assert (hb["prim"] != aq.AQ_MISS).mean() > 0.5
    assert nn / len(rays) < 60 and nt / len(rays) < 40  # the tree actually culls


def test_degenerate_inputs(aq, ao):
    """Empty scene, zero-area triangles, axis-parallel rays, rays starting on a surface."""
    nodes, tris, info = aq.build_accel_host(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint32))
    assert info.n_nodes == 1 and info.n_tri_records == 0
    rays = random_rays(aq, 64, -1, 1)
    h, _, _ = ao.bvh8_intersect(nodes, tris, rays)
    assert (h["prim"] == aq.AQ_MISS).all()
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0.5, 0.5, 0], [2, 2, 0], [2, 2, 0]], np.float32)
    idx = np.array([[0, 1, 2], [0, 3, 3], [4, 5, 4], [0, 1, 2]], np.uint32)  # 2 degenerate + 1 duplicate
    sc = aq.Scene.from_arrays(pos, idx)
    o = ao.OracleScene(sc)
    nodes, tris, info = aq.build_accel_host(pos, idx)
    rays = np.zeros(5, aq.RAY_DTYPE)
    rays["o"] = [[0.2, 0.2, 1], [0.2, 0.2, 1], [0.5, 0.0, 1], [0.2, 0.2, 0.0], [5, 5, 1]]
    rays["d"] = [[0, 0, -1], [0, 0, 1], [0, 0, -1], [0, 0, -1], [0, 0, -1]]
    rays["tmax"] = 3e38
    hb = o.intersect(rays)
    h8, _, _ = ao.bvh8_intersect(nodes, tris, rays)
    assert hits_equal(h8, hb)
    assert hb["prim"].tolist() == [0, aq.AQ_MISS, 0, aq.AQ_MISS, aq.AQ_MISS]  # edge hit counts, t=0 does not


# ---------------------------------------------------------------- golden fixtures
def test_cbox_against_golden(aq, ao, cbox):
    g = np.load(os.path.join(GOLD, "cbox_golden.npz"))
    o = ao.OracleScene(cbox)
    integ = aq.Integrator(spp=4, max_depth=5, seed=0)
    cfg = integ.cfg(width=64, height=64)
    rays = o.camera_rays(cfg, 0)
    assert np.array_equal(rays.view(np.float32).reshape(-1, 8), g["rays"])
    h = o.intersect(rays)
    assert np.array_equal(h["prim"], g["hit_prim"]) and np.array_equal(h["t"], g["hit_t"])
    film, samples, st = o.render(integ.cfg(width=32, height=32), want_samples=True)
    assert np.array_equal(film, g["film"]) and np.array_equal(samples, g["samples"])
    assert st["sample_bounces"] == int(g["sample_bounces"]) and st["rays_shadow"] == int(g["rays_shadow"])
    # single thread == all threads (film accumulates per pixel in sample order)
    film1, _, _ = o.render(integ.cfg(width=32, height=32), n_threads=1)
    assert np.array_equal(film1, film)


def test_room_against_golden(aq, ao, room):
    g = np.load(os.path.join(GOLD, "room_golden.npz"))
    o = ao.OracleScene(room, build_bvh=True)
    cfg = aq.Integrator(spp=2, max_depth=5, seed=3).cfg(width=48, height=27)
    rays = o.camera_rays(cfg, 0)
    h = o.intersect(rays, mode=1)  # BVH2 here; the fixture was made by brute force over 394,269 triangles
    assert np.array_equal(h["prim"], g["hit_prim"]) and np.array_equal(h["t"], g["hit_t"])
    film, _, st = o.render(cfg, mode=1)
    assert np.array_equal(film, g["film"]) and st["sample_bounces"] == int(g["sample_bounces"])
    assert np.isfinite(film).all() and film[..., :3].min() >= 0


def test_interleaved_traversal_step_gives_the_same_hits(aq, ao, cbox):
    """aq_trav_step2 (one node visit + a bounded number of triangle tests per step, postponed
    triangle groups on the stack; an A/B option of the kernel) against aq_trav_step."""
    pos, idx = triangle_soup(8000, r=0.03)
    for p, i, lo, hi in ((pos, idx, 0.0, 1.0), (*cbox.arrays()[:2], -1.2, 2.2)):
        nodes, tris, info = aq.build_accel_host(p, i)
        rays = random_rays(aq, 6000, lo, hi, seed=4)
        h0, n0, t0 = ao.bvh8_intersect(nodes, tris, rays, step=0)
        h1, n1, t1 = ao.bvh8_intersect(nodes, tris, rays, step=1)
        assert hits_equal(h0, h1) and n1 >= n0 and t1 >= t0 and n1 < 1.2 * n0  # later culling: a little more work
        rays["tmax"] = 0.3
        a0, _, _ = ao.bvh8_intersect(nodes, tris, rays, any_hit=True, step=0)
        a1, _, _ = ao.bvh8_intersect(nodes, tris, rays, any_hit=True, step=1)
        assert np.array_equal(a0["prim"], a1["prim"])
