"""An INDEPENDENT float64 numpy path tracer (own camera, own intersector, own BSDF restatement,
own sampling strategy: uniform-hemisphere BSDF-only, no NEE, numpy RNG) against the C++ oracle on
a small emissive scene.  Nothing here touches aq_core.h, so agreement of the means checks the
oracle's conventions (camera, two-sided emission, frames, throughput recursion, max_depth
semantics, MIS bookkeeping) rather than restating them."""
import numpy as np

from test_area_lights import lamp_room
from test_oracle import bsdf_f64


def np_intersect(o, d, V0, E1, E2):
    """closest two-sided hit of rays (n,3) against triangles (m,3): returns prim (-1 = miss), t."""
    n, m = len(o), len(V0)
    P = np.cross(d[:, None, :], E2[None, :, :])
    det = (E1[None] * P).sum(-1)
    ok = np.abs(det) > 0
    inv = np.where(ok, 1.0 / np.where(ok, det, 1), 0)
    T = o[:, None, :] - V0[None]
    u = (T * P).sum(-1) * inv
    Q = np.cross(T, E1[None])
    v = (d[:, None, :] * Q).sum(-1) * inv
    t = (E2[None] * Q).sum(-1) * inv
    hit = ok & (u >= 0) & (v >= 0) & (u + v <= 1) & (t > 1e-6)
    t = np.where(hit, t, np.inf)
    prim = t.argmin(1)
    tt = t[np.arange(n), prim]
    return np.where(np.isfinite(tt), prim, -1), tt


def np_render_pixel(scene_np, px, py, W, H, n, max_depth, rng):
    V0, E1, E2, tri_mat, mats, cam = scene_np
    # camera: looks down -z, +y up, fov across the larger side, pixel (0,0) top-left, R = Rx only here
    u0, u1 = rng.random(n), rng.random(n)
    sx = (px + u0) / W * 2 - 1
    sy = 1 - (py + u1) / H * 2
    t = np.tan(np.radians(cam["fov"]) / 2)
    tx, ty = (t, t * H / W) if W >= H else (t * W / H, t)
    dc = np.stack([sx * tx, sy * ty, -np.ones(n)], 1)
    dc /= np.linalg.norm(dc, axis=1, keepdims=True)
    a = cam["rot_x"]
    R = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    d = dc @ R.T
    o = np.tile(np.array(cam["pos"], float), (n, 1))
    beta = np.ones((n, 3))
    L = np.zeros((n, 3))
    alive = np.ones(n, bool)
    for depth in range(max_depth):
        prim, tt = np_intersect(o, d, V0, E1, E2)
        alive &= prim >= 0
        if not alive.any():
            break
        idx = np.nonzero(alive)[0]
        p = prim[idx]
        m = tri_mat[p]
        L[idx] += beta[idx] * mats["emission"][m]
        if depth + 1 >= max_depth:
            break
        ng = np.cross(E1[p], E2[p])
        ng /= np.linalg.norm(ng, axis=1, keepdims=True)
        wo = -d[idx]
        ng = np.where(((ng * wo).sum(1) < 0)[:, None], -ng, ng)
        # uniform hemisphere around ng
        z = rng.random(len(idx))
        ph = rng.random(len(idx)) * 2 * np.pi
        r = np.sqrt(1 - z * z)
        hlp = np.where(np.abs(ng[:, [0]]) > 0.9, np.array([[0, 1.0, 0]]), np.array([[1.0, 0, 0]]))
        tv = np.cross(hlp, ng)
        tv /= np.linalg.norm(tv, axis=1, keepdims=True)
        bv = np.cross(ng, tv)
        wi = r[:, None] * np.cos(ph)[:, None] * tv + r[:, None] * np.sin(ph)[:, None] * bv + z[:, None] * ng
        hitp = o[idx] + tt[idx, None] * d[idx]
        w = np.zeros((len(idx), 3))
        for k in range(len(idx)):  # BSDF in the local frame of ng (isotropic: any tangent frame)
            tl = lambda x: np.array([x @ tv[k], x @ bv[k], x @ ng[k]])
            e = bsdf_f64(mats["params"][m[k]], tl(wo[k]), tl(wi[k]))
            if e is not None:
                w[k] = e[0] * 2 * np.pi  # f cos / (1 / 2pi)
        beta[idx] *= w
        o[idx] = hitp + 1e-5 * ng
        d[idx] = wi
        alive[idx] &= w.max(1) > 0
    return L.mean(0), L.std(0) / np.sqrt(n)


def test_numpy_path_tracer_agrees_with_the_oracle(aq, ao):
    sc = lamp_room(aq, s=1.2)  # a large lamp keeps the BSDF-only estimator's variance manageable
    pos, idx, nrm, uv, tm = sc.arrays()
    V0 = pos[idx[:, 0]].astype(float)
    E1 = pos[idx[:, 1]].astype(float) - V0
    E2 = pos[idx[:, 2]].astype(float) - V0
    mats = {"params": [], "emission": []}
    for k in range(sc.desc.n_materials):
        m = sc.desc.materials[k]
        mats["params"].append([*m.color, m.metallic, m.roughness, m.specular, m.specular_tint, m.sheen, m.sheen_tint, m.transmission])
        mats["emission"].append(list(m.emission))
    mats["emission"] = np.array(mats["emission"], float)
    cam = {"fov": sc.desc.camera.fov, "pos": list(sc.desc.camera.translate), "rot_x": sc.desc.camera.rotate[0]}
    W = H = 8
    o = ao.OracleScene(sc)
    max_depth = 3
    film, _, st = o.render(aq.Integrator(spp=20000, max_depth=max_depth, seed=9).cfg(width=W, height=H))
    img = film[..., :3] / film[..., 3:]
    rng = np.random.default_rng(4)
    checked = 0
    for (px, py) in [(2, 6), (5, 5), (4, 2)]:  # two floor pixels, one wall pixel
        mean, se = np_render_pixel((V0, E1, E2, tm, mats, cam), px, py, W, H, 12000, max_depth, rng)
        got = img[py, px]
        assert got.max() > 1e-3
        tol = 4 * se + 0.03 * mean  # 4 sigma of the numpy estimator + 3 %
        assert (np.abs(got - mean) <= tol).all(), ((px, py), got, mean, se)
        checked += 1
    assert checked == 3


# ---------------------------------------------------------------- cbox itself, point-light NEE
def _cbox_np(cbox):
    pos, idx, nrm, uv, tm = cbox.arrays()
    V = pos[idx].astype(float)                       # [tri, corner, xyz]
    N = nrm[idx].astype(float)
    mats = []
    for k in range(cbox.desc.n_materials):
        m = cbox.desc.materials[k]
        mats.append([*m.color, m.metallic, m.roughness, m.specular, m.specular_tint, m.sheen, m.sheen_tint, m.transmission])
    lt = cbox.desc.lights[0]
    cam = cbox.desc.camera
    return V, N, np.asarray(tm), mats, (np.array(list(lt.pos), float), np.array(list(lt.intensity), float)), \
        (np.array(list(cam.translate), float), float(cam.fov))


def np_cbox_pixel(sc, px, py, W, H, n, max_depth, rng):
    """Expected radiance of one pixel of scenes/cbox.json from an estimator that shares nothing with
    aq_core.h: float64, numpy RNG, uniform-hemisphere continuation, no Russian roulette, NEE towards the
    point light at every vertex (L = sum over vertices of beta * f cos * I / d^2 * visibility).  Follows
    only the WRITTEN conventions of DESIGN.md section 3 (camera, two-sided shading normal = normalised
    barycentric blend flipped to the geometric side, Li = I / d^2, max_depth = vertices)."""
    V, N, tm, mats, (lpos, lint), (cpos, fov) = sc
    V0, E1, E2 = V[:, 0], V[:, 1] - V[:, 0], V[:, 2] - V[:, 0]
    t = np.tan(np.radians(fov) / 2)
    tx, ty = (t, t * H / W) if W >= H else (t * W / H, t)
    sx = (px + rng.random(n)) / W * 2 - 1
    sy = 1 - (py + rng.random(n)) / H * 2
    d = np.stack([sx * tx, sy * ty, -np.ones(n)], 1)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    o = np.tile(cpos, (n, 1))
    beta, L = np.ones((n, 3)), np.zeros((n, 3))
    alive = np.ones(n, bool)
    for depth in range(max_depth):
        prim, tt = np_intersect(o, d, V0, E1, E2)
        alive &= prim >= 0
        idx = np.nonzero(alive)[0]
        if len(idx) == 0:
            break
        p = prim[idx]
        hp = o[idx] + tt[idx, None] * d[idx]
        # barycentrics of the hit -> shading normal
        T = hp - V0[p]
        d00, d01, d11 = (E1[p] * E1[p]).sum(1), (E1[p] * E2[p]).sum(1), (E2[p] * E2[p]).sum(1)
        d20, d21 = (T * E1[p]).sum(1), (T * E2[p]).sum(1)
        den = d00 * d11 - d01 * d01
        bu, bv = (d11 * d20 - d01 * d21) / den, (d00 * d21 - d01 * d20) / den
        ns = (1 - bu - bv)[:, None] * N[p, 0] + bu[:, None] * N[p, 1] + bv[:, None] * N[p, 2]
        ns /= np.linalg.norm(ns, axis=1, keepdims=True)
        ng = np.cross(E1[p], E2[p])
        ng /= np.linalg.norm(ng, axis=1, keepdims=True)
        wo = -d[idx]
        ng = np.where(((ng * wo).sum(1) < 0)[:, None], -ng, ng)
        ns = np.where(((ns * ng).sum(1) < 0)[:, None], -ns, ns)
        hlp = np.where(np.abs(ns[:, [0]]) > 0.9, np.array([[0, 1.0, 0]]), np.array([[1.0, 0, 0]]))
        tv = np.cross(hlp, ns)
        tv /= np.linalg.norm(tv, axis=1, keepdims=True)
        bvv = np.cross(ns, tv)
        loc = lambda k, x: np.array([x @ tv[k], x @ bvv[k], x @ ns[k]])
        # ---- NEE: the light is a point, so this term has no variance of its own
        tol = lpos[None] - hp
        dist = np.linalg.norm(tol, axis=1)
        wl = tol / dist[:, None]
        so = hp + 1e-4 * ng * np.sign((wl * ng).sum(1))[:, None]
        sp, st_ = np_intersect(so, wl, V0, E1, E2)
        vis = (sp < 0) | (st_ > dist * (1 - 1e-3))
        for k in np.nonzero(vis & ((wl * ng).sum(1) > 0))[0]:
            e = bsdf_f64(mats[tm[p[k]]], loc(k, wo[k]), loc(k, wl[k]))
            if e is not None:
                L[idx[k]] += beta[idx[k]] * e[0] * lint / dist[k] ** 2
        if depth + 1 >= max_depth:
            break
        # ---- continuation: uniform hemisphere about the shading normal
        z = rng.random(len(idx))
        ph = rng.random(len(idx)) * 2 * np.pi
        r = np.sqrt(1 - z * z)
        wi = r[:, None] * np.cos(ph)[:, None] * tv + r[:, None] * np.sin(ph)[:, None] * bvv + z[:, None] * ns
        w = np.zeros((len(idx), 3))
        for k in range(len(idx)):
            if (wi[k] @ ng[k]) <= 0:
                continue  # leaves through the back of the geometric surface: no transport (opaque)
            e = bsdf_f64(mats[tm[p[k]]], loc(k, wo[k]), loc(k, wi[k]))
            if e is not None:
                w[k] = e[0] * 2 * np.pi
        beta[idx] *= w
        o[idx] = hp + 1e-4 * ng
        d[idx] = wi
        alive[idx] &= w.max(1) > 0
    return L.mean(0), L.std(0) / np.sqrt(n)


CBOX_PIXELS = [(16, 8), (8, 2), (29, 6), (20, 16)]  # (x, y) at 32x32: back wall, ceiling, right wall, back wall beside the boxes (all directly lit)


def test_numpy_estimator_of_cbox_agrees_with_the_oracle(aq, ao, cbox):
    """scenes/cbox.json at 32x32, depth 3: pixel means of the independent estimator vs the oracle's film
    (same expected value, different sampling: BSDF importance sampling + RR vs uniform hemisphere)."""
    W = H = 32
    max_depth = 3
    film, _, _ = ao.OracleScene(cbox).render(aq.Integrator(spp=2048, max_depth=max_depth, seed=5).cfg(width=W, height=H))
    img = film[..., :3] / film[..., 3:]
    sc = _cbox_np(cbox)
    rng = np.random.default_rng(11)
    for (px, py) in CBOX_PIXELS:
        mean, se = np_cbox_pixel(sc, px, py, W, H, 3000, max_depth, rng)
        got = img[py, px]
        assert got.max() > 1e-3
        assert (np.abs(got - mean) <= 4 * se + 0.03 * mean + 1e-4).all(), ((px, py), got, mean, se)


# ---------------------------------------------------------------- nrc: which path a record is, and its 64 input features
def test_nrc_record_addressing_and_encoding_restated_in_numpy(aq, ao, cbox):
    """First-hit training records (even r) of the `nrc` integrator on scenes/cbox.json, restated without
    aq_nrc.h / aq_core.h: the record's pixel and RNG key from the written hash chain (DESIGN.md section 8,
    integer arithmetic in Python), the jittered camera ray, an own float64 intersection, and the 64 input
    features (normalised position + 8 triangle-wave octaves, quartic one-blobs of wo / shading normal /
    roughness, diffuse albedo, F0, bias) — against the records the oracle generates."""
    SALT, M = 0xA5A5A5A5, 0xFFFFFFFF

    def pcg(v):
        s = (v * 747796405 + 2891336453) & M
        w = (((s >> ((s >> 28) + 4)) ^ s) * 277803737) & M
        return ((w >> 22) ^ w) & M

    def key_of(seed, pixel, sample):
        return pcg((sample + pcg((pixel + pcg(seed)) & M)) & M)

    unit = lambda k, dim: (pcg((k + dim) & M) >> 8) / 16777216.0
    W = H = 48
    from test_nrc import small_nrc
    integ = small_nrc(aq, batch_size=128, training_iters=4)
    cfg, nrc = integ.cfg(width=W, height=H), integ.nrc_cfg()
    x, y = ao.OracleScene(cbox).nrc_records(cfg, nrc)
    V, N, tm, mats, _, (cpos, fov) = _cbox_np(cbox)
    V0, E1, E2 = V[:, 0], V[:, 1] - V[:, 0], V[:, 2] - V[:, 0]
    pos = cbox.arrays()[0].astype(float)
    lo, ext = pos.min(0), pos.max(0) - pos.min(0)
    t = np.tan(np.radians(fov) / 2)

    def blob(v):
        xx = min(max(v, 0.0), 1.0)
        q = 1.0 - ((xx - (np.arange(4) + 0.5) / 4) * 4) ** 2
        return np.where(q > 0, q * q, 0.0)

    seed = cfg.seed
    checked = close = 0
    for r in range(0, len(x), 2):
        if y[r, 3] != 1:
            continue
        pixel = (pcg((key_of(seed ^ SALT, r, 0x4E5243) + 0) & M) * (W * H)) >> 32
        key = key_of(seed ^ SALT, pixel, r)
        px, py = pixel % W, pixel // W
        sx = (px + unit(key, 0)) / W * 2 - 1
        sy = 1 - (py + unit(key, 1)) / H * 2
        d = np.array([sx * t, sy * t, -1.0])
        d /= np.linalg.norm(d)
        prim, tt = np_intersect(cpos[None], d[None], V0, E1, E2)
        if prim[0] < 0:
            continue
        p = prim[0]
        hp = cpos + tt[0] * d
        T = hp - V0[p]
        d00, d01, d11, d20, d21 = E1[p] @ E1[p], E1[p] @ E2[p], E2[p] @ E2[p], T @ E1[p], T @ E2[p]
        den = d00 * d11 - d01 * d01
        bu, bv = (d11 * d20 - d01 * d21) / den, (d00 * d21 - d01 * d20) / den
        ns = (1 - bu - bv) * N[p, 0] + bu * N[p, 1] + bv * N[p, 2]
        ns /= np.linalg.norm(ns)
        ng = np.cross(E1[p], E2[p])
        wo = -d
        if ng @ wo < 0:
            ng = -ng
        if ns @ ng < 0:
            ns = -ns
        base, metallic, rough, spec, spec_tint, _, _, transmission = np.array(mats[tm[p]][:3]), *mats[tm[p]][3:]
        pn = np.clip((hp - lo) / ext, 0, 1)
        f = [pn]
        for k in range(8):
            f.append(np.abs(2 * ((pn * 2 ** k) % 1.0) - 1))
        feat = np.concatenate(f + [blob(v * 0.5 + 0.5) for v in wo] + [blob(v * 0.5 + 0.5) for v in ns] + [blob(rough)])
        lum = 0.2126 * base[0] + 0.7152 * base[1] + 0.0722 * base[2]
        tint = base / lum if lum > 0 else np.ones(3)
        diel = (1 + (tint - 1) * spec_tint) * 0.08 * spec
        f0 = diel + (base - diel) * metallic
        feat = np.concatenate([feat, base * (1 - metallic) * (1 - transmission), f0, [1.0, 0.0, 0.0]])
        assert feat.shape == (64,)
        checked += 1
        # the high triangle-wave octaves amplify the float32 position error by up to 2^8
        tol = np.full(64, 2e-5)
        tol[3:27] = 2e-5 * 2 ** (np.arange(24) // 3 + 1)
        close += bool((np.abs(feat - x[r]) <= tol + 1e-4 * (np.arange(64) >= 27) * (np.arange(64) < 55)).all())
    assert checked >= 150 and close >= 0.98 * checked, (checked, close)


def test_numpy_estimator_of_cbox_at_depth_six_checks_russian_roulette(aq, ao, cbox):
    """The same comparison at max_depth 6: vertices 4..6 of the oracle's paths survive Russian roulette
    (q = min(max beta, 0.95), throughput / q), the independent estimator has none — equal pixel means say
    the roulette is unbiased and the throughput recursion holds over six vertices."""
    W = H = 32
    max_depth = 6
    film, _, st = ao.OracleScene(cbox).render(aq.Integrator(spp=4096, max_depth=max_depth, seed=6).cfg(width=W, height=H))
    assert st["sample_bounces"] > 2.3 * st["samples"]  # (a depth-3 render of this view: 2.17 vertices per sample)
    img = film[..., :3] / film[..., 3:]
    sc = _cbox_np(cbox)
    rng = np.random.default_rng(12)
    for (px, py) in CBOX_PIXELS[:2]:
        mean, se = np_cbox_pixel(sc, px, py, W, H, 4000, max_depth, rng)
        got = img[py, px]
        d3, _ = np_cbox_pixel(sc, px, py, W, H, 2000, 3, rng)
        assert (mean > d3 * 1.02).any()  # the deeper bounces carry visible energy in this pixel
        assert (np.abs(got - mean) <= 4 * se + 0.03 * mean + 1e-4).all(), ((px, py), got, mean, se)
