"""OBJ/MTL -> scene JSON + BSON .mesh.  Pinned by reference DATA: importing the
reference's own CornellBox-Original.obj must reproduce the reference's .mesh files byte for byte
and the Srgb colours of cbox.json."""
import filecmp
import glob
import json
import os

import numpy as np
import pytest

REF = "/root/reference/scenes"
needs_ref = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "CornellBox-Original.obj")),
                               reason="reference checkout not present")


@needs_ref
def test_import_reproduces_reference_meshes_byte_for_byte(aq, tmp_path):
    jp = aq.import_obj(os.path.join(REF, "CornellBox-Original.obj"), str(tmp_path), "cbox_imported")
    ref_meshes = sorted(glob.glob(os.path.join(REF, "CornellBox-Original_*.mesh")))
    assert len(ref_meshes) == 8
    for f in ref_meshes:
        mine = tmp_path / os.path.basename(f)
        assert mine.exists(), f"importer did not write {mine.name}"
        assert filecmp.cmp(f, str(mine), shallow=False), f"{mine.name} differs from the reference file"
    sc = aq.Scene.load(jp)
    ref = aq.Scene.load(os.path.join(REF, "cbox.json"))
    assert (sc.info.n_verts, sc.info.n_tris, sc.info.n_materials) == (64, 36, 8)
    p0, i0, n0, *_ = sc.arrays()
    p1, i1, n1, *_ = ref.arrays()
    assert np.array_equal(p0, p1) and np.array_equal(i0, i1) and np.array_equal(n0, n1)
    for k, name in enumerate(sc.material_names()):
        j = ref.material_names().index(name)
        assert np.allclose(list(sc.desc.materials[k].color), list(ref.desc.materials[j].color), atol=2e-7)
        assert sc.desc.materials[k].metallic == 0.0  # Ks = 0 everywhere in the MTL
    js = json.load(open(jp))
    assert js["shapes"][5] == {"Mesh": ["CornellBox-Original_shortBox_5.mesh", {"Named": "shortBox"}]}
    assert js["named_bsdfs"]["floor"]["Principled"]["hint"] == "ltc"


def test_import_material_mapping_and_uvs(aq, tmp_path):
    """roughness = sqrt(2/(Ns+2)), metallic = 1/(1+max Kd) when Ks != 0, Ks as colour when Kd = 0,
    map_Kd -> Image with a Windows separator, negative and v/vt indices, quads and pentagons."""
    (tmp_path / "t.mtl").write_text(
        "newmtl Apple\nNs 7\nKd 0.5 0.1 0.1\nKs 0 0 0\n"
        "newmtl Ceramic\nNs 98\nKd 0.3 0.3 0.3\nKs 0.5 0.5 0.5\n"
        "newmtl BlackMarble\nNs 30\nKd 0 0 0\nKs 0.2 0.3 0.4\n"
        "newmtl Mag\nNs 10\nKd 1 1 1\nKs 0 0 0\nmap_Kd textures/magazine.jpg\n")
    (tmp_path / "t.obj").write_text(
        "mtllib t.mtl\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv 0.5 1.5 0\nvt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\n"
        "g a\nusemtl Apple\nf 1 2 3 4\n"
        "g b\nusemtl Ceramic\nf -5 -4 -3 -2 -1\n"
        "g c\nusemtl BlackMarble\nf 1 2 3\n"
        "g d\nusemtl Mag\nf 1/1 2/2 3/3 4/4\n")
    jp = aq.import_obj(str(tmp_path / "t.obj"), str(tmp_path / "out"), "t")
    js = json.load(open(jp))
    b = js["named_bsdfs"]
    assert abs(b["Apple"]["Principled"]["roughness"]["Float"] - np.float32(np.sqrt(2 / 9))) < 1e-7
    assert b["Apple"]["Principled"]["metallic"]["Float"] == 0.0
    assert abs(b["Ceramic"]["Principled"]["metallic"]["Float"] - np.float32(1 / 1.3)) < 1e-7
    want = [aq._abi.host_lib().aq_host_srgb_to_linear(c) for c in b["BlackMarble"]["Principled"]["color"]["Srgb"]]
    assert np.allclose(want, [0.2, 0.3, 0.4], atol=1e-6)                     # Ks used because Kd == 0
    assert b["Mag"]["Principled"]["color"] == {"Image": "textures\\magazine.jpg"}
    ma = aq.load_mesh(str(tmp_path / "out" / "t_a_0.mesh"))
    assert ma["indices"].tolist() == [[0, 1, 2], [0, 2, 3]] and ma["texcoords"].shape[0] == 0
    mb = aq.load_mesh(str(tmp_path / "out" / "t_b_1.mesh"))
    assert mb["indices"].tolist() == [[0, 1, 2], [0, 2, 3], [0, 3, 4]] and len(mb["vertices"]) == 5
    md = aq.load_mesh(str(tmp_path / "out" / "t_d_3.mesh"))
    assert md["texcoords"].tolist() == [[0, 0], [1, 0], [1, 1], [0, 1]]
    assert np.allclose(ma["normals"], [[0, 0, 1]] * 4)
    for f in glob.glob(str(tmp_path / "out" / "*.mesh")):
        assert int.from_bytes(open(f, "rb").read(4), "little") == os.path.getsize(f)


def test_import_errors(aq, tmp_path):
    with pytest.raises(aq.AquaError):
        aq.import_obj(str(tmp_path / "missing.obj"), str(tmp_path), "x")
    (tmp_path / "empty.obj").write_text("v 0 0 0\n")
    with pytest.raises(aq.AquaError):
        aq.import_obj(str(tmp_path / "empty.obj"), str(tmp_path), "x")
