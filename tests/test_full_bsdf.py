"""Row (f)-4 of SURVEY §8: the Principled lobes that are zero in every shipped material
(scenes/cbox.json:15-17,48-58 — clearcoat, clearcoat_roughness, ior, transmission, subsurface,
subsurface_color): the FULL instantiation of the vertex code (aq_core.h, aq_bsdf_*_full,
aq_shade_vertex<AREA, true>).

CPU checks of the oracle: an independent float64 restatement, the dielectric Fresnel term,
bit-identity with the fast path when the extra inputs are zero (function level and whole renders,
against the committed golden films), sampling density == eval density, energy conservation of
glass, the generalised reciprocity of refraction, closed-form transmittance of a glass slab,
closed-form one-bounce radiance through a transmissive floor.  GPU parity at the end."""
import ctypes as C
import os

import numpy as np
import pytest

from test_oracle import GOLD, MATS, dirs, fptr

# params: base.rgb metallic roughness specular specular_tint sheen sheen_tint transmission
#         clearcoat clearcoat_roughness ior subsurface subsurface_color.rgb
FULL_MATS = {
    "clearcoat_paint": [0.7, 0.1, 0.1, 0.0, 0.5, 0.5, 0.0, 0.0, 0.5, 0.0, 1.0, 0.2, 1.45, 0.0, 0, 0, 0],
    "clearcoat_metal": [0.9, 0.7, 0.3, 1.0, 0.35, 0.0, 0.0, 0.0, 0.5, 0.0, 0.6, 0.3, 1.45, 0.0, 0, 0, 0],
    "rough_glass": [0.95, 0.97, 1.0, 0.0, 0.4, 0.5, 0.0, 0.0, 0.5, 1.0, 0.0, 0.03, 1.5, 0.0, 0, 0, 0],
    "half_glass": [0.8, 0.6, 0.4, 0.0, 0.5, 0.5, 0.3, 0.2, 0.5, 0.5, 0.3, 0.4, 1.33, 0.0, 0, 0, 0],
    "skin_like": [0.8, 0.5, 0.4, 0.0, 0.6, 0.3, 0.0, 0.0, 0.5, 0.0, 0.0, 0.03, 1.45, 0.7, 0.9, 0.3, 0.2],
    "everything": [0.6, 0.7, 0.8, 0.2, 0.45, 0.6, 0.4, 0.3, 0.6, 0.4, 0.8, 0.35, 1.6, 0.5, 0.9, 0.4, 0.3],
}


def fresnel_f64(c, eta):
    c = min(max(float(c), 0.0), 1.0)
    s2 = (1 - c * c) / (eta * eta)
    if s2 >= 1:
        return 1.0
    ct = np.sqrt(1 - s2)
    rp = (eta * c - ct) / (eta * c + ct)
    rs = (c - eta * ct) / (c + eta * ct)
    return 0.5 * (rp * rp + rs * rs)


def bsdf_full_f64(p, eta, wo, wi):
    """Independent float64 restatement of DESIGN.md's full Principled definition -> (f*|cos|, pdf)."""
    p = [float(x) for x in p]
    base = np.array(p[:3])
    metallic, rough, spec, stint, sheen, shtint, trans, cc, ccr, _ior, ss = p[3:14]
    sscol = np.array(p[14:17])
    wo, wi = np.asarray(wo, float), np.asarray(wi, float)
    if wo[2] <= 0 or wi[2] == 0:
        return None
    tw = (1 - metallic) * trans
    dw = (1 - metallic) * (1 - trans)
    lum = 0.2126 * base[0] + 0.7152 * base[1] + 0.0722 * base[2]
    tint = base / lum if lum > 0 else np.ones(3)
    f0 = 0.08 * spec * (1 + stint * (tint - 1))
    f0 = f0 + metallic * (base - f0)
    base_d = base + ss * (sscol - base)
    al, alc = max(rough * rough, 1e-4), max(ccr * ccr, 1e-4)
    lam = lambda a, c: 0.5 * (np.sqrt(1 + a * a * (1 - c * c) / (c * c)) - 1)
    D = lambda a, h: a * a / (np.pi * (h[2] ** 2 * (a * a - 1) + 1) ** 2)
    fo5 = (1 - min(wo[2], 1.0)) ** 5
    Fo = f0 + (1 - f0) * fo5
    fdo = fresnel_f64(wo[2], eta)
    if tw > 0:
        Fo = (1 - tw) * Fo + tw * fdo
    w_s, w_d = Fo.max(), dw * base_d.max()
    w_t = tw * max(1 - fdo, min(al, 0.5)) * base.max()
    w_c = 0.25 * cc * (0.04 + 0.96 * fo5)
    tot = w_s + w_d + w_t + w_c
    if tot <= 0:
        return None
    ps, pd, pt, pc = w_s / tot, w_d / tot, w_t / tot, w_c / tot
    if wi[2] > 0:
        h = wo + wi
        h /= np.linalg.norm(h)
        ldh = wi @ h
        f = np.zeros(3)
        pdf = 0.0
        if dw > 0:
            fl, fv = (1 - wi[2]) ** 5, (1 - wo[2]) ** 5
            fd90 = 0.5 + 2 * rough * ldh * ldh
            fd = (1 + (fd90 - 1) * fl) * (1 + (fd90 - 1) * fv)
            fss90 = ldh * ldh * rough
            fss = (1 + (fss90 - 1) * fl) * (1 + (fss90 - 1) * fv)
            sst = 1.25 * (fss * (1 / (wi[2] + wo[2]) - 0.5) + 0.5)
            shape = (1 - ss) * fd + ss * sst
            sh = sheen * (1 + shtint * (tint - 1)) * (1 - ldh) ** 5
            f += dw * (base_d / np.pi * shape + sh)
            pdf += pd * wi[2] / np.pi
        if ps > 0:
            F = f0 + (1 - f0) * (1 - ldh) ** 5
            if tw > 0:
                F = (1 - tw) * F + tw * fresnel_f64(ldh, eta)
            G = 1 / (1 + lam(al, wo[2]) + lam(al, wi[2]))
            f += F * D(al, h) * G / (4 * wo[2] * wi[2])
            pdf += ps * D(al, h) / (1 + lam(al, wo[2])) / (4 * wo[2])
        if pc > 0:
            Fc = 0.04 + 0.96 * (1 - ldh) ** 5
            G = 1 / (1 + lam(alc, wo[2]) + lam(alc, wi[2]))
            f += 0.25 * cc * Fc * D(alc, h) * G / (4 * wo[2] * wi[2])
            pdf += pc * D(alc, h) / (1 + lam(alc, wo[2])) / (4 * wo[2])
        return f * wi[2], pdf
    if tw <= 0 or pt <= 0:
        return None
    h = wo + eta * wi
    h /= np.linalg.norm(h)
    if h[2] < 0:
        h = -h
    odh, idh = wo @ h, wi @ h
    if odh <= 0 or idh >= 0:
        return None
    F = fresnel_f64(odh, eta)
    den = odh + eta * idh
    G2 = 1 / (1 + lam(al, wo[2]) + lam(al, -wi[2]))
    G1 = 1 / (1 + lam(al, wo[2]))
    ft = tw * base * (1 - F) * D(al, h) * G2 * abs(idh) * odh / (wo[2] * abs(wi[2]) * den * den)  # radiance form
    pdf = pt * D(al, h) * G1 * odh / wo[2] * eta * eta * abs(idh) / (den * den)
    return ft * abs(wi[2]), pdf


def sphere_dirs(n, seed):
    g = np.random.default_rng(seed)
    z = g.uniform(-1, 1, n)
    z = np.where(np.abs(z) < 0.02, 0.5, z)
    ph = g.uniform(0, 2 * np.pi, n)
    r = np.sqrt(1 - z * z)
    return np.stack([r * np.cos(ph), r * np.sin(ph), z], 1).astype(np.float32)


# ---------------------------------------------------------------- scalar pieces
def test_fresnel_dielectric(ao):
    L = ao.lib()
    for eta in (1.5, 1 / 1.5, 1.33, 2.4, 1.0001):
        assert np.isclose(L.aqo_fresnel_dielectric(1.0, eta), ((eta - 1) / (eta + 1)) ** 2, rtol=1e-4, atol=1e-9)
        for c in np.linspace(0.0, 1.0, 41):
            assert np.isclose(L.aqo_fresnel_dielectric(float(c), eta), fresnel_f64(c, eta), rtol=2e-4, atol=1e-7)
    # total internal reflection past the critical angle, continuity just before it
    crit = np.sqrt(1 - (1 / 1.5) ** 2)  # cos of the critical angle leaving ior 1.5
    assert L.aqo_fresnel_dielectric(float(crit) - 1e-3, 1 / 1.5) == 1.0
    assert 0.5 < L.aqo_fresnel_dielectric(float(crit) + 1e-4, 1 / 1.5) < 1.0
    # the same interface seen from the other side reflects the same fraction
    for c in (0.2, 0.5, 0.9):
        ct = np.sqrt(1 - (1 - c * c) / 1.5 ** 2)
        assert np.isclose(L.aqo_fresnel_dielectric(c, 1.5), L.aqo_fresnel_dielectric(float(ct), 1 / 1.5), rtol=1e-4)


@pytest.mark.parametrize("name", list(FULL_MATS))
def test_full_eval_against_float64(ao, name):
    p = np.array(FULL_MATS[name], np.float32)
    n_refr = 0
    for eta in (p[12], 1.0 / p[12]):
        wos = dirs(24, 21)
        for wo in wos:
            wis = sphere_dirs(60, int(1000 * wo[2]))
            f, pdf, ok = ao.bsdf_eval_full(p, eta, wo, wis)
            for k, wi in enumerate(wis):
                ref = bsdf_full_f64(p.astype(np.float64), float(eta), wo.astype(np.float64), wi.astype(np.float64))
                if ref is None or ref[1] <= 0:
                    assert not ok[k] or pdf[k] < 1e-6
                    continue
                assert ok[k], (wo, wi, ref)
                n_refr += wi[2] < 0
                assert np.allclose(f[k], ref[0], rtol=3e-3, atol=1e-6), (wo, wi, f[k], ref[0])
                assert np.isclose(pdf[k], ref[1], rtol=3e-3, atol=1e-6)
    if p[9] > 0:
        assert n_refr > 100  # the refraction branch was exercised (most far-side directions lie outside the refraction cone)


def test_full_with_zero_extras_is_bit_identical_to_fast_functions(ao):
    """clearcoat = transmission = subsurface = 0: the FULL functions run the same operations in
    the same order as the fast ones (eval, pdf, sampled direction, weight: same bits)."""
    L = ao.lib()
    g = np.random.default_rng(2)
    for name, m in MATS.items():
        if m[9] != 0.0:
            continue
        p10 = np.array(m, np.float32)
        p17 = np.array(list(m) + [0.0, 0.03, 1.45, 0.0, 0.3, 0.6, 0.9], np.float32)
        for wo in dirs(40, 5):
            wis = dirs(50, 6)
            f, pdf, ok = ao.bsdf_eval_full(p17, 1.45, wo, wis)
            for k, wi in enumerate(wis):
                fa = np.zeros(3, np.float32)
                pa = C.c_float()
                oka = L.aqo_bsdf_eval(fptr(p10), fptr(wo), fptr(wi), fptr(fa), C.byref(pa))
                assert bool(oka) == bool(ok[k])
                if oka:
                    assert np.array_equal(fa.view(np.uint32), f[k].view(np.uint32))
                    assert np.float32(pa.value).view(np.uint32) == pdf[k].view(np.uint32)
            u3 = g.uniform(0, 1, (200, 3)).astype(np.float32)
            wi_f, w_f, pdf_f, ok_f = ao.bsdf_sample_full(p17, 1.45, wo, u3)
            for k in range(len(u3)):
                wi, w = np.zeros(3, np.float32), np.zeros(3, np.float32)
                pa = C.c_float()
                oka = L.aqo_bsdf_sample(fptr(p10), fptr(wo), fptr(u3[k]), fptr(wi), fptr(w), C.byref(pa))
                assert bool(oka) == bool(ok_f[k])
                if oka:
                    assert np.array_equal(wi.view(np.uint32), wi_f[k].view(np.uint32))
                    assert np.array_equal(w.view(np.uint32), w_f[k].view(np.uint32))


def test_forced_full_render_equals_golden_films(aq, ao, cbox, room):
    """Whole renders of both shipped scenes through aq_shade_vertex<.., FULL=true> reproduce the
    committed golden films bit for bit (their materials have all extra inputs at zero)."""
    g = np.load(os.path.join(GOLD, "cbox_golden.npz"))
    integ = aq.Integrator(spp=4, max_depth=5, seed=0)
    film, samples, st = ao.OracleScene(cbox).render(
        integ.cfg(width=32, height=32, flags=aq.AQ_RENDER_FORCE_FULL_BSDF), want_samples=True)
    assert np.array_equal(film, g["film"]) and np.array_equal(samples, g["samples"])
    assert st["sample_bounces"] == int(g["sample_bounces"]) and st["rays_shadow"] == int(g["rays_shadow"])
    g = np.load(os.path.join(GOLD, "room_golden.npz"))
    cfg = aq.Integrator(spp=2, max_depth=5, seed=3).cfg(width=48, height=27, flags=aq.AQ_RENDER_FORCE_FULL_BSDF)
    film, _, st = ao.OracleScene(room, build_bvh=True).render(cfg, mode=1)
    assert np.array_equal(film, g["film"]) and st["sample_bounces"] == int(g["sample_bounces"])


# ---------------------------------------------------------------- statistical properties
def sphere_grid(nz, nphi):
    """midpoint grid on the sphere, equal solid angle 4 pi / (nz * nphi) per cell"""
    z = (np.arange(nz) + 0.5) / nz * 2 - 1
    ph = (np.arange(nphi) + 0.5) / nphi * 2 * np.pi
    Z, P = np.meshgrid(z, ph, indexing="ij")
    r = np.sqrt(1 - Z * Z)
    return np.stack([r * np.cos(P), r * np.sin(P), Z], -1).reshape(-1, 3).astype(np.float32), 4 * np.pi / (nz * nphi)


@pytest.mark.parametrize("name", ["clearcoat_metal", "rough_glass", "half_glass", "everything"])
def test_full_sampling_density_matches_eval(ao, name):
    """(i) the mean sampling weight equals the quadrature of f*cos over the sphere: sample() draws
    from exactly the density eval() reports, in both hemispheres; (ii) that density integrates to
    the fraction of samples that are not lost (below-horizon reflections, total internal
    reflection)."""
    p = np.array(FULL_MATS[name], np.float32)
    grid, dw = sphere_grid(1200, 600)
    g = np.random.default_rng(17)
    for eta in (p[12], 1.0 / p[12]):
        for wo in dirs(2, 31):
            f, pdf, ok = ao.bsdf_eval_full(p, eta, wo, grid)
            quad = f.astype(np.float64).sum(0) * dw
            mass = pdf.astype(np.float64).sum() * dw
            n = 600000
            wi, w, spdf, sok = ao.bsdf_sample_full(p, eta, wo, g.uniform(0, 1, (n, 3)))
            mean = w.astype(np.float64).sum(0) / n
            se = w.astype(np.float64).std(0) / np.sqrt(n)
            assert (np.abs(mean - quad) <= 4 * se + 0.015 * quad + 1e-4).all(), (name, eta, wo, mean, quad, se)
            assert abs(mass - sok.mean()) < 0.015, (name, eta, wo, mass, sok.mean())
            assert np.allclose(np.linalg.norm(wi[sok], axis=1), 1, atol=1e-4)
            if p[9] > 0 and eta > 1:  # (leaving the medium at a grazing wo everything is totally reflected)
                assert (wi[sok, 2] < 0).mean() > 0.05  # refraction is being sampled
            # upper-hemisphere samples of a refracted direction never appear, and vice versa
            assert mass <= 1.005


def test_glass_conserves_energy(ao):
    """base = 1, transmission = 1: reflected + transmitted POWER <= 1 and close to 1 at moderate
    roughness (single-scattering microfacet models lose the multiply scattered part).  The BTDF is
    in radiance form, so transmitted power = eta^2 * its integral.  Integrated by importance
    sampling (the lobes are too narrow for a grid; the sampling density is checked above)."""
    g = np.random.default_rng(23)
    for rough, lo in ((0.05, 0.99), (0.2, 0.97), (0.5, 0.85)):
        p = np.array([1, 1, 1, 0.0, rough, 0.5, 0.0, 0.0, 0.5, 1.0, 0.0, 0.03, 1.5, 0.0, 0, 0, 0], np.float32)
        for eta in (1.5, 1 / 1.5):
            for wo in ([0.0, 0.0, 1.0], [0.6, 0.0, 0.8]):
                n = 400000
                wi, w, _, ok = ao.bsdf_sample_full(p, eta, np.array(wo, np.float32), g.uniform(0, 1, (n, 3)))
                up = wi[:, 2] > 0
                R = w[ok & up, 0].astype(np.float64).sum() / n
                T = w[ok & ~up, 0].astype(np.float64).sum() / n * eta * eta
                assert R + T <= 1.005 and R + T >= lo, (rough, eta, wo, R, T)
                if wo[2] == 1.0 and rough <= 0.2:
                    assert abs(R - 0.04) < 0.01  # ((1.5-1)/(1.5+1))^2 from either side


def test_refraction_obeys_generalised_reciprocity(ao):
    """Veach: f(wi->wo) / eta_o^2 = f(wo->wi) / eta_i^2 for the radiance BTDF.  Swapping the roles
    of the two directions means evaluating from the other side of the interface (frame flipped,
    relative index inverted)."""
    p = np.array(FULL_MATS["rough_glass"], np.float32)
    eta = 1.5
    n = 0
    for wo in dirs(30, 8):
        wis = sphere_dirs(40, int(wo[2] * 999))
        wis = wis[wis[:, 2] < 0]
        fa, _, oka = ao.bsdf_eval_full(p, eta, wo, wis)
        for k, wi in enumerate(wis):
            wo2 = np.array([wi[0], wi[1], -wi[2]], np.float32)
            wi2 = np.array([wo[0], wo[1], -wo[2]], np.float32)
            fb, _, okb = ao.bsdf_eval_full(p, 1 / eta, wo2, wi2[None])
            hv = wo.astype(np.float64) + eta * wi.astype(np.float64)
            hv /= np.linalg.norm(hv)
            if min(abs(wo @ hv), abs(wi @ hv)) < 0.05:
                continue  # grazing microfacet: 1 - F and the Jacobian are ill-conditioned in f32
            assert bool(oka[k]) == bool(okb[0])
            if oka[k]:
                a = fa[k] / abs(wi[2])      # BTDF value, wo outside (index 1), wi inside (index 1.5)
                b = fb[0] / abs(wi2[2])
                assert np.allclose(a * eta * eta, b, rtol=2e-3, atol=1e-7), (wo, wi, a, b)
                n += 1
    assert n > 80


# ---------------------------------------------------------------- whole-path closed forms
def glass_slab_scene(aq, ior=1.5, Le=3.0):
    """camera -> glass slab (front face z=0 facing the camera, back face z=-0.2 facing away)
    -> emissive wall at z=-1."""
    q = lambda z, flip: (np.array([[-5, -5, z], [5, -5, z], [5, 5, z], [-5, 5, z]], np.float32),
                         [[0, 2, 1], [0, 3, 2]] if flip else [[0, 1, 2], [0, 2, 3]])
    pos, idx = [], []
    for z, flip in ((0.0, False), (-0.2, True), (-1.0, False)):
        P, I = q(z, flip)
        idx += [[a + len(pos) * 4 for a in t] for t in I]
        pos.append(P)
    pos = np.concatenate(pos)
    glass = aq.default_material(color=(1.0, 1.0, 1.0), roughness=0.0)
    glass.transmission, glass.ior, glass.specular = 1.0, ior, 0.5
    lamp = aq.default_material(color=(0.0, 0.0, 0.0))
    lamp.emission[:] = (Le, Le, Le)
    cam = aq.default_camera(res=(4, 4), fov=0.5, translate=(0.0, 0.0, 3.0))
    return aq.Scene.from_arrays(pos, np.array(idx, np.uint32), materials=[glass, lamp],
                                tri_material=[0, 0, 0, 0, 1, 1], camera=cam)


def test_glass_slab_transmittance_closed_form(aq, ao):
    """Normal incidence through a smooth slab: radiance is scaled by 1/eta^2 entering and eta^2
    leaving (they cancel); with all inter-reflections L = Le (1-F)^2 / (1-F^2) = Le (1-F)/(1+F)."""
    ior, Le = 1.5, 3.0
    sc = glass_slab_scene(aq, ior, Le)
    o = ao.OracleScene(sc)
    rays = o.camera_rays(aq.Integrator(spp=1).cfg(width=4, height=4), 0)
    h = o.intersect(rays)
    assert (h["prim"] < 2).all() and np.allclose(h["t"], 3.0, atol=1e-3)
    # front face normal points at the camera (outside), back face away from it
    F = ((ior - 1) / (ior + 1)) ** 2
    cfg = aq.Integrator(spp=8192, max_depth=12, seed=1).cfg(width=4, height=4, flags=aq.AQ_RENDER_MIS_BSDF_ONLY)
    film, _, st = o.render(cfg)
    got = (film[..., :3] / film[..., 3:]).reshape(-1, 3).mean(0)
    want = Le * (1 - F) / (1 + F)
    assert np.allclose(got, want, rtol=0.01), (got, want)
    assert st["rays_shadow"] == 0
    # two bounces only: the directly transmitted term Le (1-F)^2 (emitter found at the third vertex)
    cfg = aq.Integrator(spp=8192, max_depth=3, seed=2).cfg(width=4, height=4, flags=aq.AQ_RENDER_MIS_BSDF_ONLY)
    film, _, _ = o.render(cfg)
    got = (film[..., :3] / film[..., 3:]).reshape(-1, 3).mean(0)
    assert np.allclose(got, Le * (1 - F) ** 2, rtol=0.01), (got, Le * (1 - F) ** 2)


@pytest.mark.parametrize("name,below", [("clearcoat_paint", False), ("half_glass", False), ("half_glass", True),
                                        ("everything", True)])
def test_closed_form_one_bounce_radiance_full(aq, ao, name, below):
    """Point light above (reflection lobes) or BELOW (refraction: the shadow ray starts on the far
    side of the floor) a floor with a full-Principled material, max_depth 1:
    L = f(wo,wi) |cos(theta_i)| I / d^2 with f from the float64 restatement."""
    pos = np.array([[-50, 0, -50], [50, 0, -50], [50, 0, 50], [-50, 0, 50]], np.float32)
    idx = np.array([[0, 2, 1], [0, 3, 2]], np.uint32)  # normal +y: the camera is on the outside
    p = FULL_MATS[name]
    mat = aq.default_material(color=p[:3], metallic=p[3], roughness=p[4])
    (mat.specular, mat.specular_tint, mat.sheen, mat.sheen_tint, mat.transmission, mat.clearcoat,
     mat.clearcoat_roughness, mat.ior, mat.subsurface) = p[5:14]
    mat.subsurface_color[:] = p[14:17]
    I, hgt = 10.0, (-1.5 if below else 2.0)
    cam = aq.default_camera(res=(8, 8), fov=20.0, translate=(0, 1.0, 3))
    cam.rotate[:] = (-0.6, 0.0, 0.0)
    sc = aq.Scene.from_arrays(pos, idx, materials=[mat], lights=[aq.point_light((0.3, hgt, 0.2), (I, I, I))], camera=cam)
    o = ao.OracleScene(sc)
    cfg = aq.Integrator(spp=1, max_depth=1).cfg(width=8, height=8)
    rays = o.camera_rays(cfg, 0)
    hits = o.intersect(rays)
    film, samples, st = o.render(cfg, want_samples=True)
    # (from below, directions outside the refraction cone carry nothing: no shadow ray for them)
    assert (hits["prim"] != aq.AQ_MISS).all() and (st["rays_shadow"] == 64 or (below and st["rays_shadow"] > 32))
    to_local = lambda w: np.array([w[0], -w[2], w[1]])  # floor normal +y -> local +z (right-handed)
    n_lit = 0
    for k in range(len(rays)):
        P = rays["o"][k].astype(float) + hits["t"][k] * rays["d"][k].astype(float)
        wo = -rays["d"][k].astype(float)
        Lv = np.array([0.3, hgt, 0.2]) - P
        d2 = Lv @ Lv
        wi = Lv / np.sqrt(d2)
        ref = bsdf_full_f64(p, p[12], to_local(wo), to_local(wi))
        got = samples[0].reshape(-1, 4)[k, :3]
        if ref is None:
            assert below and got.max() == 0
            continue
        want = ref[0] * I / d2
        n_lit += 1
        assert got.max() > 0
        assert np.allclose(got, want, rtol=3e-3), (k, got, want)
    assert n_lit > 32


def full_cbox(aq, cbox):
    """cbox.json with the lobes switched on: glass short box, clear-coated tall box, subsurface
    floor, half-transmissive back wall, and the MTL's emitter restored next to the point light."""
    pos, idx, nrm, uv, tm = cbox.arrays()
    mats = []
    for k, name in enumerate(cbox.material_names()):
        m = type(cbox.desc.materials[k])()
        C.memmove(C.byref(m), C.byref(cbox.desc.materials[k]), C.sizeof(m))
        if name == "shortBox":
            m.transmission, m.ior, m.roughness, m.specular = 1.0, 1.45, 0.15, 0.5
        elif name == "tallBox":
            m.clearcoat, m.clearcoat_roughness, m.metallic = 1.0, 0.1, 0.0
        elif name == "floor":
            m.subsurface = 0.6
            m.subsurface_color[:] = (0.9, 0.3, 0.2)
        elif name == "backWall":
            m.transmission, m.clearcoat, m.ior = 0.5, 0.5, 1.3
        elif name == "light":
            m.emission[:] = (17.0, 12.0, 4.0)
        mats.append(m)
    cam = type(cbox.desc.camera)()
    C.memmove(C.byref(cam), C.byref(cbox.desc.camera), C.sizeof(cam))
    return aq.Scene.from_arrays(pos.copy(), idx.copy(), normals=nrm.copy(), tri_material=tm.copy(), materials=mats,
                                lights=[aq.point_light((0.0, 1.7, 0.1), (1.0, 1.0, 1.0))], camera=cam)


def test_full_cbox_renders_and_refracts(aq, ao, cbox):
    sc = full_cbox(aq, cbox)
    o = ao.OracleScene(sc)
    film, samples, st = o.render(aq.Integrator(spp=8, max_depth=6, seed=4).cfg(width=48, height=48), want_samples=True)
    assert np.isfinite(film).all() and film[..., :3].min() >= 0 and film[..., :3].mean() > 0.01
    # more path vertices than the opaque scene reaches at the same depth budget: paths continue inside the glass box
    plain = ao.OracleScene(cbox).render(aq.Integrator(spp=8, max_depth=6, seed=4).cfg(width=48, height=48))[2]
    assert st["sample_bounces"] != plain["sample_bounces"]
    film1, _, _ = o.render(aq.Integrator(spp=8, max_depth=6, seed=4).cfg(width=48, height=48), n_threads=1)
    assert np.array_equal(film, film1)


# ---------------------------------------------------------------- GPU parity
@pytest.mark.gpu
def test_gpu_parity_full_principled(aq, ao, cbox, renderer):
    """The FULL instantiations of the shade kernel (with and without emissive triangles) against
    the oracle: per-sample radiance, film and ray counts bit for bit."""
    sc = full_cbox(aq, cbox)
    ds, o = renderer.upload(sc), ao.OracleScene(sc)
    for fl in (0, aq.AQ_RENDER_MIS_BSDF_ONLY):
        cfg = aq.Integrator(spp=6, max_depth=7, seed=2).cfg(width=96, height=96, flags=fl | aq.AQ_RENDER_DUMP_SAMPLES)
        film, st = ds.render(cfg)
        samples = ds.samples(cfg)
        ofilm, osamples, ost = o.render(cfg, want_samples=True)
        err = np.abs(samples - osamples).max() / max(1e-20, np.abs(osamples).max())
        assert err <= 1e-4, err  # the north-star tolerance; the design target is bit equality:
        assert np.array_equal(samples, osamples) and np.array_equal(film, ofilm)
        assert st["rays_shadow"] == ost["rays_shadow"] and st["sample_bounces"] == ost["sample_bounces"]
    # AREA = false, FULL = true: the glass slab (no point light, emission only through BSDF hits is AREA;
    # so use the floor scene with a point light below a transmissive floor)
    sc2 = glass_slab_scene(aq)
    ds2, o2 = renderer.upload(sc2), ao.OracleScene(sc2)
    cfg = aq.Integrator(spp=64, max_depth=8, seed=3).cfg(width=16, height=16, flags=aq.AQ_RENDER_DUMP_SAMPLES)
    film, st = ds2.render(cfg)
    ofilm, osamples, ost = o2.render(cfg, want_samples=True)
    assert np.array_equal(ds2.samples(cfg), osamples) and np.array_equal(film, ofilm)


@pytest.mark.gpu
def test_gpu_forced_full_kernel_equals_fast_kernel(aq, ao, cbox, room, renderer):
    """aq_k_shade<false,true> on the shipped scenes == aq_k_shade<false,false> == oracle, bit for bit."""
    for sc, w, h, spp in ((cbox, 128, 128, 8), (room, 96, 54, 4)):
        ds = renderer.upload(sc)
        integ = aq.Integrator(spp=spp, max_depth=5, seed=7)
        fa, sa = ds.render(integ.cfg(width=w, height=h))
        fb, sb = ds.render(integ.cfg(width=w, height=h, flags=aq.AQ_RENDER_FORCE_FULL_BSDF))
        assert np.array_equal(fa, fb)
        assert sa["sample_bounces"] == sb["sample_bounces"] and sa["rays_shadow"] == sb["rays_shadow"]
    sc3 = None
    # a point light below a half-transmissive floor: AREA=false, FULL=true with two-sided NEE
    pos = np.array([[-50, 0, -50], [50, 0, -50], [50, 0, 50], [-50, 0, 50]], np.float32)
    idx = np.array([[0, 2, 1], [0, 3, 2]], np.uint32)
    p = FULL_MATS["everything"]
    mat = aq.default_material(color=p[:3], metallic=p[3], roughness=p[4])
    (mat.specular, mat.specular_tint, mat.sheen, mat.sheen_tint, mat.transmission, mat.clearcoat,
     mat.clearcoat_roughness, mat.ior, mat.subsurface) = p[5:14]
    mat.subsurface_color[:] = p[14:17]
    cam = aq.default_camera(res=(64, 64), fov=40.0, translate=(0, 1.0, 3))
    cam.rotate[:] = (-0.6, 0.0, 0.0)
    sc3 = aq.Scene.from_arrays(pos, idx, materials=[mat], camera=cam,
                               lights=[aq.point_light((0.3, -1.5, 0.2), (10, 10, 10)), aq.point_light((0.0, 2.0, 0.0), (5, 5, 5))])
    ds3, o3 = renderer.upload(sc3), ao.OracleScene(sc3)
    cfg = aq.Integrator(spp=8, max_depth=3, seed=5).cfg(width=64, height=64, flags=aq.AQ_RENDER_DUMP_SAMPLES)
    film, st = ds3.render(cfg)
    ofilm, osamples, ost = o3.render(cfg, want_samples=True)
    assert np.array_equal(ds3.samples(cfg), osamples) and np.array_equal(film, ofilm)
    assert st["rays_shadow"] == ost["rays_shadow"] > 0
