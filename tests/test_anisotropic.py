"""`anisotropic` / `anisotropic_rotation` of Bsdf::Principled (scenes/cbox.json:29-34; zero in both shipped
scenes, parsed but not evaluated in round 1).  Round 2 evaluates them in the FULL instantiation: GGX with
alpha_x = max(r^2/s, 1e-4), alpha_y = max(r^2 s, 1e-4), s = sqrt(1 - 0.9 a), lobe axes = surface tangent dp/du
rotated by `anisotropic_rotation` turns about the shading normal.  Checked against a float64 numpy
restatement, by the usual sampling identities, and GPU == oracle bit for bit."""
import numpy as np
import pytest

WHITE_METAL = [1.0, 1.0, 1.0, 1.0, 0.35, 0.0, 0.0, 0.0, 0.5, 0.0, 0.0, 0.03, 1.45, 0.0, 1.0, 1.0, 1.0]
COPPER = [0.95, 0.64, 0.54, 1.0, 0.25, 0.0, 0.0, 0.0, 0.5, 0.0, 0.0, 0.03, 1.45, 0.0, 1.0, 1.0, 1.0]
MIXED = [0.8, 0.3, 0.2, 0.3, 0.4, 0.5, 0.2, 0.0, 0.5, 0.0, 0.5, 0.1, 1.45, 0.0, 1.0, 1.0, 1.0]


def hemi(n, seed, zmin=0.02):
    g = np.random.default_rng(seed)
    z = g.uniform(zmin, 1, n)
    ph = g.uniform(0, 2 * np.pi, n)
    r = np.sqrt(1 - z * z)
    return np.stack([r * np.cos(ph), r * np.sin(ph), z], 1)


def metal_f64(base, rough, aniso, rot, tang, wo, wi):
    """float64: F D G2 / (4 wo.z wi.z) * wi.z and the VNDF pdf D G1(wo) / (4 wo.z) of an anisotropic GGX metal."""
    asp = np.sqrt(1 - 0.9 * aniso)
    ax, ay = max(rough * rough / asp, 1e-4), max(rough * rough * asp, 1e-4)
    th = tang + 2 * np.pi * rot
    c, s = np.cos(th), np.sin(th)
    R = np.array([[c, s, 0], [-s, c, 0], [0, 0, 1.0]])          # shading frame -> lobe axes
    o, i = R @ wo, R @ wi
    h = o + i
    h /= np.linalg.norm(h)
    D = 1.0 / (np.pi * ax * ay * ((h[0] / ax) ** 2 + (h[1] / ay) ** 2 + h[2] ** 2) ** 2)
    lam = lambda w: 0.5 * (np.sqrt(1 + ((ax * w[0]) ** 2 + (ay * w[1]) ** 2) / w[2] ** 2) - 1)
    F = np.array(base) + (1 - np.array(base)) * (1 - i @ h) ** 5
    f_cos = F * D / (1 + lam(o) + lam(i)) / (4 * o[2])
    pdf = D / (1 + lam(o)) / (4 * o[2])
    return f_cos, pdf


def test_anisotropic_metal_against_float64(ao):
    wo = np.array([0.5, -0.2, 0.8])
    wo /= np.linalg.norm(wo)
    wis = hemi(400, 3)
    for params, aniso in ((COPPER, (0.8, 0.1, 0.3)), (WHITE_METAL, (0.5, 0.0, 0.0)), (COPPER, (1.0, 0.37, -1.1))):
        f, pdf, ok = ao.bsdf_eval_aniso(params, aniso, 1.45, wo, wis)
        assert ok.all()
        for k in range(len(wis)):
            wf, wp = metal_f64(params[:3], params[4], aniso[0], aniso[1], aniso[2], wo, wis[k])
            assert np.allclose(f[k], wf, rtol=2e-4, atol=1e-7), (k, f[k], wf)
            assert np.isclose(pdf[k], wp, rtol=2e-4), (k, pdf[k], wp)


def test_zero_anisotropy_is_the_isotropic_code_bit_for_bit(ao):
    wo = np.array([0.3, 0.4, 0.8660254])
    wis, u3 = hemi(300, 5), np.random.default_rng(6).random((300, 3))
    for params in (COPPER, MIXED):
        f0, p0, ok0 = ao.bsdf_eval_full(params, 1.45, wo, wis)
        f1, p1, ok1 = ao.bsdf_eval_aniso(params, (0.0, 0.3, 1.0), 1.45, wo, wis)
        assert np.array_equal(f0, f1) and np.array_equal(p0, p1) and np.array_equal(ok0, ok1)
        s0, s1 = ao.bsdf_sample_full(params, 1.45, wo, u3), ao.bsdf_sample_aniso(params, (0.0, 0.3, 1.0), 1.45, wo, u3)
        assert all(np.array_equal(a, b) for a, b in zip(s0, s1))


def test_rotation_parameter_and_tangent_angle_are_the_same_rotation(ao):
    wo = np.array([0.6, 0.1, 0.79])
    wo /= np.linalg.norm(wo)
    wis = hemi(200, 8)
    a = ao.bsdf_eval_aniso(COPPER, (0.7, 0.0, 2 * np.pi * 0.2), 1.45, wo, wis)
    b = ao.bsdf_eval_aniso(COPPER, (0.7, 0.2, 0.0), 1.45, wo, wis)
    assert np.allclose(a[0], b[0], rtol=2e-4, atol=1e-7) and np.allclose(a[1], b[1], rtol=2e-4)
    # a quarter turn swaps the axes: different from no rotation, equal to three quarters + half a turn
    c = ao.bsdf_eval_aniso(COPPER, (0.7, 0.25, 0.0), 1.45, wo, wis)
    d = ao.bsdf_eval_aniso(COPPER, (0.7, 0.75, 0.0), 1.45, wo, wis)
    assert not np.allclose(a[0], c[0], rtol=1e-2) and np.allclose(c[0], d[0], rtol=2e-4, atol=1e-7)


@pytest.mark.parametrize("params,aniso", [(WHITE_METAL, (0.9, 0.15, 0.4)), (MIXED, (0.6, 0.6, -0.7))])
def test_sampling_identities_hold_with_anisotropy(ao, params, aniso):
    """weight * pdf == f*cos at the sampled direction; a white metal never gains energy; the pdf integrates
    to the probability of a valid sample."""
    g = np.random.default_rng(9)
    for wo in (np.array([0.0, 0.0, 1.0]), np.array([0.7, -0.3, 0.648]), np.array([-0.2, 0.95, 0.24])):
        wo = wo / np.linalg.norm(wo)
        u3 = g.random((20000, 3))
        wi, w, pdf, ok = ao.bsdf_sample_aniso(params, aniso, 1.45, wo, u3)
        f, p2, ok2 = ao.bsdf_eval_aniso(params, aniso, 1.45, wo, wi[ok])
        assert ok2.all() and np.allclose(p2, pdf[ok], rtol=1e-4)
        assert np.allclose(w[ok] * pdf[ok, None], f, rtol=2e-4, atol=1e-7)
        if params is WHITE_METAL:
            assert w.max() <= 1.0 + 1e-4 and 0.5 < w.mean() <= 1.0 + 1e-4     # F G2/G1 <= 1
        # integral of the pdf over the hemisphere by uniform sampling = fraction of valid samples
        wu = hemi(200000, 10, zmin=0.0)
        _, pu, oku = ao.bsdf_eval_aniso(params, aniso, 1.45, wo, wu)
        assert abs((pu * oku).mean() * 2 * np.pi - ok.mean()) < 0.03


def _aniso_quad(aq, with_uv, aniso=0.85, rot=0.1):
    pos = np.array([[-1, -1, 0], [1, -1, 0], [1, 1, 0], [-1, 1, 0], [-1, -1, -0.5], [1, -1, -0.5], [1, 1, -0.5]], np.float32)
    uv = np.array([[0, 0], [1, 0.2], [1.1, 1], [0, 1], [0, 0], [1, 0], [1, 1]], np.float32)   # sheared: dp/du is not an edge
    idx = np.array([[0, 1, 2], [0, 2, 3], [4, 5, 6]], np.uint32)
    m = aq.default_material(color=(0.95, 0.64, 0.54), metallic=1.0, roughness=0.3)
    m.anisotropic, m.anisotropic_rotation = aniso, rot
    d = aq.default_material(color=(0.6, 0.6, 0.6), roughness=0.5)
    cam = aq.default_camera(res=(48, 48), fov=38.0, translate=(0.2, -0.4, 3))
    return aq.Scene.from_arrays(pos, idx, normals=np.tile([0, 0, 1.0], (7, 1)), uvs=uv if with_uv else None,
                                tri_material=[0, 0, 1], materials=[m, d],
                                lights=[aq.point_light((0.5, 0.6, 1.0), (3, 3, 3)), aq.point_light((-0.8, -0.2, 0.7), (2, 1, 1))], camera=cam)


def test_oracle_renders_an_anisotropic_highlight_that_follows_the_uv_tangent(aq, ao):
    cfg = aq.Integrator(spp=8, max_depth=2, seed=3).cfg(width=48, height=48)
    iso, _, _ = ao.OracleScene(_aniso_quad(aq, True, aniso=0.0)).render(cfg)
    a0, _, _ = ao.OracleScene(_aniso_quad(aq, True, rot=0.0)).render(cfg)
    a1, _, _ = ao.OracleScene(_aniso_quad(aq, True, rot=0.25)).render(cfg)
    assert np.isfinite(a0).all() and np.isfinite(a1).all()
    assert not np.allclose(iso, a0, rtol=1e-2) and not np.allclose(a0, a1, rtol=1e-2)
    nouv, _, _ = ao.OracleScene(_aniso_quad(aq, False)).render(cfg)          # no uvs: the frame's own tangent
    assert np.isfinite(nouv).all() and nouv[..., :3].sum() > 0


@pytest.mark.gpu
def test_gpu_anisotropic_bit_exact_vs_oracle(aq, ao, renderer):
    for with_uv in (True, False):
        sc = _aniso_quad(aq, with_uv)
        cfg = aq.Integrator(spp=4, max_depth=4, seed=5).cfg(width=48, height=48, flags=aq.AQ_RENDER_DUMP_SAMPLES)
        ds = renderer.upload(sc)
        film, st = ds.render(cfg)
        smp = ds.samples(cfg)
        ofilm, osmp, ost = ao.OracleScene(sc).render(cfg, want_samples=True)
        assert st["sample_bounces"] == ost["sample_bounces"]
        assert np.array_equal(smp.view(np.uint32), osmp.view(np.uint32)) and np.array_equal(film, ofilm)
