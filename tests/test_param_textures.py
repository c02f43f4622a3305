"""Texture::Image on Principled parameters other than `color` (the schema allows an image on every
field, scenes/cbox.json:5-63; round 1 rejected it): parameter = constant x first channel of the bilinear,
linearised texel (subsurface_color: constant x rgb).  Builder-defined convention (the reference has no
code), written down in include/aqua_cuda.h and DESIGN.md section 3."""
import json
import os

import numpy as np
import pytest

from test_oracle import bsdf_f64


def _checker(n=64, lo=40, hi=230):
    y, x = np.mgrid[0:n, 0:n]
    t = np.zeros((n, n, 4), np.uint8)
    t[..., :3] = np.where(((x // 8 + y // 8) % 2 == 0)[..., None], hi, lo)
    t[..., 1] //= 2          # channels differ: only the first one may drive a scalar
    t[..., 3] = 255
    return t


def _quad_scene(aq, slot_values, tex, full=False):
    pos = np.array([[-1, -1, 0], [1, -1, 0], [1, 1, 0], [-1, 1, 0]], np.float32)
    uv = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float32)
    idx = np.array([[0, 1, 2], [0, 2, 3]], np.uint32)
    mat = aq.default_material(color=(0.8, 0.7, 0.6), metallic=1.0, roughness=1.0)
    for name, const in slot_values.items():
        setattr(mat, name, const)
        mat.param_tex[aq._abi.PTEX[name]] = 1
    cam = aq.default_camera(res=(32, 32), fov=35.0, translate=(0, 0, 3))
    return aq.Scene.from_arrays(pos, idx, normals=np.tile([0, 0, 1.0], (4, 1)), uvs=uv, materials=[mat],
                                lights=[aq.point_light((0.4, 0.3, 1.2), (3, 3, 3))], camera=cam, textures=[tex])


def test_oracle_roughness_and_metallic_textures_follow_the_written_convention(aq, ao):
    tex = _checker()
    sc = _quad_scene(aq, {"roughness": 0.9, "metallic": 1.0}, tex)
    o = ao.OracleScene(sc)
    cfg = aq.Integrator(spp=1, max_depth=1, seed=3).cfg(width=32, height=32)
    rays = o.camera_rays(cfg, 0)
    hits = o.intersect(rays)
    _, samples, _ = o.render(cfg, want_samples=True)
    got = samples[0].reshape(-1, 4)[:, :3]
    srgb = lambda c: np.where(c <= 0.04045, c / 12.92, ((c + 0.055) / 1.055) ** 2.4)
    lin = srgb(tex[..., 0].astype(float) / 255.0)           # first channel only
    n = tex.shape[0]
    lpos, I = np.array([0.4, 0.3, 1.2]), 3.0
    checked = 0
    for k in range(len(rays)):
        if hits["prim"][k] == aq.AQ_MISS:
            continue
        P = rays["o"][k].astype(float) + float(hits["t"][k]) * rays["d"][k].astype(float)
        u, v = (P[0] + 1) / 2, (P[1] + 1) / 2
        x, y = u * n - 0.5, (1 - v) * n - 0.5
        x0, y0 = int(np.floor(x)), int(np.floor(y))
        fx, fy = x - x0, y - y0
        tx = lambda a, b: lin[b % n, a % n]
        s = (tx(x0, y0) * (1 - fx) + tx(x0 + 1, y0) * fx) * (1 - fy) + (tx(x0, y0 + 1) * (1 - fx) + tx(x0 + 1, y0 + 1) * fx) * fy
        wo = -rays["d"][k].astype(float)
        Lv = lpos - P
        d2 = Lv @ Lv
        fcos, _ = bsdf_f64([0.8, 0.7, 0.6, 1.0 * s, 0.9 * s, 0.0, 0.0, 0.0, 0.5, 0.0], wo, Lv / np.sqrt(d2))
        assert np.allclose(got[k], fcos * I / d2, rtol=3e-3, atol=1e-5), (k, got[k], fcos * I / d2, s)
        checked += 1
    assert checked > 400
    # and it is not the constant-parameter image
    plain = _quad_scene(aq, {}, tex)
    _, s2, _ = ao.OracleScene(plain).render(cfg, want_samples=True)
    assert not np.allclose(s2[0].reshape(-1, 4)[:, :3], got, rtol=1e-2)


def test_loader_accepts_images_on_scalar_parameters(aq, ao, scenes, tmp_path):
    """A copy of cbox.json whose floor takes `roughness` and whose tallBox takes `metallic` from one of the
    room's JPEGs loads, renders, and differs from the plain cbox; an image on `emission` is still rejected."""
    src = json.load(open(os.path.join(scenes, "cbox.json")))
    img = sorted(os.listdir(os.path.join(scenes, "textures")))[0]
    for sh in src["shapes"]:                          # the loader resolves meshes relative to the JSON file
        rel = sh["Mesh"][0].replace("\\", "/")
        os.symlink(os.path.join(scenes, rel), tmp_path / rel)
    os.makedirs(tmp_path / "textures")
    os.symlink(os.path.join(scenes, "textures", img), tmp_path / "textures" / img)
    src["named_bsdfs"]["floor"]["Principled"]["roughness"] = {"Image": "textures\\" + img}
    src["named_bsdfs"]["tallBox"]["Principled"]["metallic"] = {"Image": "textures/" + img}
    p = tmp_path / "cbox_ptex.json"
    p.write_text(json.dumps(src))
    sc = aq.Scene.load(str(p))
    names = sc.material_names()
    mf, mt = sc.desc.materials[names.index("floor")], sc.desc.materials[names.index("tallBox")]
    assert mf.param_tex[aq._abi.PTEX["roughness"]] == 1 and mf.roughness == 1.0
    assert mt.param_tex[aq._abi.PTEX["metallic"]] == 1 and mt.metallic == 1.0 and sc.desc.n_textures == 1
    cfg = aq.Integrator(spp=2, max_depth=3, seed=1).cfg(width=48, height=48)
    film, _, _ = ao.OracleScene(sc).render(cfg)
    ref, _, _ = ao.OracleScene(aq.Scene.load(os.path.join(scenes, "cbox.json"))).render(cfg)
    assert np.isfinite(film).all() and not np.array_equal(film, ref)
    src["named_bsdfs"]["floor"]["Principled"]["emission"] = {"Image": "textures/" + img}
    p.write_text(json.dumps(src))
    with pytest.raises(aq.AquaError) as e:
        aq.Scene.load(str(p))
    assert e.value.code == -4 and "UNSUPPORTED" in str(e.value)


@pytest.mark.gpu
def test_gpu_parameter_textures_bit_exact_vs_oracle(aq, ao, renderer):
    tex = _checker()
    for slots, flags in (({"roughness": 0.9, "metallic": 1.0}, 0),
                         ({"specular": 1.0, "sheen": 0.8, "roughness": 0.7}, 0),
                         ({"clearcoat": 1.0, "clearcoat_roughness": 0.5, "subsurface": 0.9}, 0)):
        sc = _quad_scene(aq, slots, tex)
        cfg = aq.Integrator(spp=4, max_depth=4, seed=5).cfg(width=48, height=48, flags=aq.AQ_RENDER_DUMP_SAMPLES | flags)
        ds = renderer.upload(sc)
        film, st = ds.render(cfg)
        smp = ds.samples(cfg)
        ofilm, osmp, ost = ao.OracleScene(sc).render(cfg, want_samples=True)
        assert st["sample_bounces"] == ost["sample_bounces"]
        assert np.array_equal(smp.view(np.uint32), osmp.view(np.uint32)), slots
        assert np.array_equal(film, ofilm)
        with pytest.raises(aq.AquaError):             # an index past the texture table is refused
            bad = _quad_scene(aq, slots, tex)
            bad.desc.materials[0].param_tex[0] = 9
            renderer.upload(bad)
