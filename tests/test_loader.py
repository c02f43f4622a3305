"""Host-side ingest: the only behaviour the reference's DATA pins.

BSON .mesh parser, scene/integrator JSON, sRGB linearisation, Named-BSDF resolution,
missing-asset tolerance, JPEG decode."""
import glob
import json
import os
import struct

import numpy as np
import pytest

REF = "/root/reference/scenes"
needs_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")


def test_cbox_counts_and_bounds(cbox):
    i = cbox.info
    assert (i.n_shapes, i.n_meshes_loaded, i.n_meshes_missing) == (8, 8, 0)
    assert (i.n_verts, i.n_tris, i.n_materials, i.n_textures, i.n_lights) == (64, 36, 8, 0, 1)
    assert np.allclose(list(i.bounds_min), [-1.02, 0.0, -1.04], atol=1e-6)
    assert np.allclose(list(i.bounds_max), [1.0, 1.99, 0.99], atol=1e-6)
    d = cbox.desc
    assert (d.camera.res[0], d.camera.res[1], d.camera.fov) == (512, 512, 15.0)
    assert list(d.camera.translate) == [0.0, 1.0, 9.0]
    assert np.allclose(list(d.lights[0].pos), [0.0, 1.7, 0.1])
    assert list(d.lights[0].intensity) == [1.0, 1.0, 1.0]


def test_cbox_global_prim_id_prefix_table(cbox):
    """floor 0-1, ceiling 2-3, backWall 4-5, rightWall 6-7, leftWall 8-9,
    shortBox 10-21, tallBox 22-33, light 34-35."""
    names = cbox.material_names()
    want = [("floor", 0, 2), ("ceiling", 2, 2), ("backWall", 4, 2), ("rightWall", 6, 2),
            ("leftWall", 8, 2), ("shortBox", 10, 12), ("tallBox", 22, 12), ("light", 34, 2)]
    for s, (name, first, cnt) in enumerate(want):
        f, c, m = cbox.shape_range(s)
        assert (f, c, names[m]) == (first, cnt, name)


def test_cbox_duplicate_triangles(cbox):
    """4 exactly duplicated triangles (shortBox tri 6,7 == 10,11; tallBox tri 8,9 == 10,11):
    the (t, prim) tie-break is observable on this scene."""
    pos, idx, *_ = cbox.arrays()
    for base, a in ((10, 6), (22, 8)):
        assert np.array_equal(idx[base + a], idx[base + 10])
        assert np.array_equal(idx[base + a + 1], idx[base + 11])


def test_cbox_materials(cbox):
    names = cbox.material_names()
    d = cbox.desc
    for k, n in enumerate(names):
        m = d.materials[k]
        if n == "tallBox":
            assert (m.metallic, m.roughness) == (1.0, np.float32(0.1))
        else:
            assert (m.metallic, m.roughness) == (0.0, np.float32(0.4))
        assert m.color_tex == -1 and list(m.emission) == [0, 0, 0]
        assert abs(m.ior - 1.45) < 1e-6 and abs(m.clearcoat_roughness - 0.03) < 1e-6


@needs_ref
def test_srgb_linearisation_matches_mtl_kd(cbox):
    """Srgb constants are the OETF of the MTL's linear Kd."""
    kd = {}
    cur = None
    for line in open(os.path.join(REF, "CornellBox-Original.mtl")):
        t = line.split()
        if t[:1] == ["newmtl"]:
            cur = t[1]
        elif t[:1] == ["Kd"]:
            kd[cur] = [float(x) for x in t[1:4]]
    names = cbox.material_names()
    for k, n in enumerate(names):
        got = list(cbox.desc.materials[k].color)
        assert np.allclose(got, kd[n], atol=2e-6), (n, got, kd[n])


@needs_ref
def test_cbox_mesh_vertices_equal_obj(aq):
    """.mesh vertices are the OBJ `v` lines as f32; quads are fan-triangulated."""
    obj_v = [[float(x) for x in l.split()[1:4]] for l in open(os.path.join(REF, "CornellBox-Original.obj"))
             if l.startswith("v ")]
    obj_v = np.array(obj_v, np.float32)
    m = aq.load_mesh(os.path.join(REF, "CornellBox-Original_floor_0.mesh"))
    assert m["name"] == "floor"
    assert m["vertices"].shape == (4, 3) and m["indices"].tolist() == [[0, 1, 2], [0, 2, 3]]
    for v in m["vertices"]:
        assert (np.abs(obj_v - v).max(axis=1) == 0).any(), v
    assert m["texcoords"].shape[0] == 0
    lw = aq.load_mesh(os.path.join(REF, "CornellBox-Original_leftWall_4.mesh"))
    assert np.allclose(lw["normals"][0], [0.99991262, 0.01004899, 0.00492531], atol=1e-7)


def test_bson_length_equals_file_size_for_every_mesh(aq, scenes):
    files = sorted(glob.glob(os.path.join(scenes, "*.mesh")))
    assert len(files) == 160
    tot_v = tot_t = 0
    for f in files:
        assert struct.unpack("<i", open(f, "rb").read(4))[0] == os.path.getsize(f)
    for f in files[:40]:  # full parse of a quarter keeps the CPU suite short; the scene test parses all
        m = aq.load_mesh(f)
        nv = len(m["vertices"])
        assert len(m["normals"]) == nv and len(m["texcoords"]) in (0, nv)
        assert m["indices"].max() == nv - 1
        assert np.isfinite(m["vertices"]).all()
        tot_v += nv
        tot_t += len(m["indices"])
    assert tot_v > 0 and tot_t > 0


def test_room_counts_textures_and_missing_mesh(room):
    i = room.info
    assert (i.n_shapes, i.n_meshes_loaded, i.n_meshes_missing) == (153, 152, 1)
    assert (i.n_verts, i.n_tris, i.n_materials, i.n_textures, i.n_lights) == (253238, 394269, 42, 16, 1)
    assert np.allclose(list(i.bounds_min), [-2.582, 0.024, -0.425], atol=1e-3)
    assert np.allclose(list(i.bounds_max), [3.148, 3.171, 8.174], atol=1e-3)
    # the missing mesh keeps an empty range so shape indices stay aligned with the JSON
    assert room.shape_range(43)[1] == 0
    pos, idx, nrm, uv, tm = room.arrays()
    assert uv is not None and uv.shape == (253238, 2)
    assert tm.max() < 42 and len(np.unique(tm)) >= 40
    n_tex_mats = sum(1 for k in range(42) if room.desc.materials[k].color_tex >= 0)
    assert n_tex_mats == 16
    names = room.material_names()
    fl = room.desc.materials[names.index("Floor")]
    assert abs(fl.metallic - 0.1) < 1e-7 and abs(fl.roughness - 0.05345224738121033) < 1e-7


def test_every_named_bsdf_resolves_and_is_used(room, scenes):
    js = json.load(open(os.path.join(scenes, "room.json")))
    used = {s["Mesh"][1]["Named"] for s in js["shapes"]}
    assert used == set(js["named_bsdfs"]) == set(room.material_names())


def test_integrator_json(aq, scenes):
    it = aq.Integrator.load(os.path.join(scenes, "integrator.json"))
    assert (it.type, it.spp, it.max_depth) == ("nrc", 4, 5)
    c = it.cfg(width=256, height=256)
    assert (c.spp_begin, c.spp_end, c.max_depth) == (0, 4, 5)


def test_loader_errors(aq, tmp_path):
    with pytest.raises(aq.AquaError):
        aq.Scene.load(str(tmp_path / "nope.json"))
    p = tmp_path / "bad.json"
    p.write_text('{"named_bsdfs": {"a": {"Glass": {}}}, "shapes": []}')
    with pytest.raises(aq.AquaError) as e:
        aq.Scene.load(str(p))
    assert "Principled" in str(e.value)
    bad = tmp_path / "bad.mesh"
    bad.write_bytes(b"\x10\x00\x00\x00garbage")
    with pytest.raises(aq.AquaError):
        aq.load_mesh(str(bad))
    p2 = tmp_path / "integ.json"
    p2.write_text('{"type": "bdpt", "spp": 1}')
    with pytest.raises(aq.AquaError):
        aq.Integrator.load(str(p2))


def test_jpeg_decoder_against_libjpeg(aq, scenes):
    """Baseline and progressive files; tolerance covers IDCT / chroma-upsampling differences
    between decoders (texel parity with the reference's jpeg-decoder crate is unpinned)."""
    from PIL import Image
    for name, tol_max, tol_mean in [("photo1.jpg", 2, 0.1), ("wood4.jpg", 4, 0.2), ("apple.jpg", 12, 0.4)]:
        f = os.path.join(scenes, "textures", name)
        a = aq.decode_jpeg(f)
        b = np.asarray(Image.open(f).convert("RGB")).astype(int)
        assert a.shape[:2] == b.shape[:2] and (a[..., 3] == 255).all()
        d = np.abs(a[..., :3].astype(int) - b)
        assert d.max() <= tol_max and d.mean() <= tol_mean, (name, d.max(), d.mean())
