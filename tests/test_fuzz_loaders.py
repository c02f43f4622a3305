"""Robustness of the host ingest against damaged inputs: every failure must surface as an
AquaError (AQ_ERR_IO), never as a crash or a hang."""
import os

import numpy as np
import pytest


def test_truncated_and_corrupted_mesh_files(aq, scenes, tmp_path):
    src = open(os.path.join(scenes, "CornellBox-Original_shortBox_5.mesh"), "rb").read()
    g = np.random.default_rng(0)
    n_err = 0
    for k in range(200):
        b = bytearray(src)
        if k % 2 == 0:
            b = b[: int(g.integers(0, len(b)))]
        else:
            for _ in range(int(g.integers(1, 6))):
                b[int(g.integers(0, len(b)))] = int(g.integers(0, 256))
        p = tmp_path / "m.mesh"
        p.write_bytes(bytes(b))
        try:
            m = aq.load_mesh(str(p))
            assert m["indices"].size == 0 or m["indices"].max() < max(1, len(m["vertices"]))
        except aq.AquaError as e:
            assert e.code == -7
            n_err += 1
    assert n_err > 100


def test_corrupted_jpeg_files(aq, scenes, tmp_path):
    src = open(os.path.join(scenes, "textures", "photo1.jpg"), "rb").read()
    g = np.random.default_rng(1)
    for k in range(60):
        b = bytearray(src)
        if k % 3 == 0:
            b = b[: int(g.integers(2, len(b)))]
        else:
            for _ in range(int(g.integers(1, 20))):
                b[int(g.integers(2, len(b)))] = int(g.integers(0, 256))
        p = tmp_path / "t.jpg"
        p.write_bytes(bytes(b))
        try:
            img = aq.decode_jpeg(str(p))
            assert img.ndim == 3 and img.shape[2] == 4
        except aq.AquaError as e:
            assert e.code == -7


def test_malformed_scene_json(aq, scenes, tmp_path):
    src = open(os.path.join(scenes, "cbox.json")).read()
    cases = [src[: len(src) // 2], src.replace('"Principled"', '"Glass"', 1), src.replace('"Mesh"', '"Sphere"', 1),
             src.replace('"Named": "floor"', '"Named": "nope"', 1), "[]", "", "{\"named_bsdfs\": 3}"]
    for c in cases:
        p = tmp_path / "s.json"
        p.write_text(c)
        with pytest.raises(aq.AquaError):
            aq.Scene.load(str(p))
