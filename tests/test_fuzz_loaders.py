"""Robustness of the host ingest against damaged inputs: every failure must surface as an
AquaError (AQ_ERR_IO), never as a crash or a hang."""
import os
import struct

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


def _bson_doc(elems):
    body = b"".join(elems) + b"\0"
    return struct.pack("<i", len(body) + 4) + body


def _bson_str(key, raw_len, payload):
    return b"\x02" + key + b"\0" + struct.pack("<i", raw_len) + payload


def _bson_arr(key, raw_len, payload=b""):
    return b"\x04" + key + b"\0" + struct.pack("<i", raw_len) + payload


@pytest.mark.timeout(20)
def test_crafted_bson_lengths_neither_abort_nor_hang(aq, tmp_path):
    """a `name` string with length <= 0 made vsize-5 underflow (length_error through the C
    ABI -> abort); a negative array length made the element walker step backwards (endless loop)."""
    empty_arr = struct.pack("<i", 5) + b"\0"
    good_tail = [b"\x04" + k + b"\0" + empty_arr for k in (b"vertices", b"normals", b"texcoords", b"indices")]
    cases = [
        _bson_doc([_bson_str(b"name", 0, b"")] + good_tail),
        _bson_doc([_bson_str(b"name", -1, b"")] + good_tail),
        _bson_doc([_bson_str(b"name", -2**31, b"")] + good_tail),
        _bson_doc([_bson_str(b"name", 3, b"abc")] + good_tail),          # not NUL-terminated
        _bson_doc([_bson_arr(b"vertices", -3)] + good_tail),
        _bson_doc([_bson_arr(b"vertices", 0)] + good_tail),
        _bson_doc([_bson_arr(b"vertices", 4)] + good_tail),
        _bson_doc([_bson_arr(b"vertices", 2**31 - 1)] + good_tail),
        _bson_doc([b"\x04vertices" + b"x" * 40]),                         # element name never ends
    ]
    for k, doc in enumerate(cases):
        p = tmp_path / f"c{k}.mesh"
        p.write_bytes(doc)
        with pytest.raises(aq.AquaError) as e:
            aq.load_mesh(str(p))
        assert e.value.code == -7, k


def test_mesh_without_normals_returns_zero_normals(aq, tmp_path):
    """load_mesh read n_verts*3 floats from a 1-byte allocation when `normals` was absent."""
    def arr(items):
        return _bson_doc([b"\x04" + str(i).encode() + b"\0" + it for i, it in enumerate(items)])
    vec = lambda t, vals: _bson_doc([t + str(i).encode() + b"\0" + v for i, v in enumerate(vals)])
    verts = arr([vec(b"\x01", [struct.pack("<d", x) for x in v]) for v in ((0, 0, 0), (1, 0, 0), (0, 1, 0))])
    idx = arr([vec(b"\x12", [struct.pack("<q", x) for x in (0, 1, 2)])])
    doc = _bson_doc([_bson_str(b"name", 2, b"t\0"), b"\x04vertices\0" + verts, b"\x04indices\0" + idx])
    p = tmp_path / "nonormals.mesh"
    p.write_bytes(doc)
    m = aq.load_mesh(str(p))
    assert m["vertices"].shape == (3, 3) and m["indices"].tolist() == [[0, 1, 2]]
    assert m["normals"].shape == (3, 3) and not m["normals"].any()


def test_json_nesting_depth_is_capped(aq, tmp_path):
    p = tmp_path / "deep.json"
    p.write_text("[" * 100000)
    with pytest.raises(aq.AquaError):
        aq.Scene.load(str(p))
    p.write_text('{"named_bsdfs": ' + "[" * 70 + "]" * 70 + "}")
    with pytest.raises(aq.AquaError):
        aq.Scene.load(str(p))
    p.write_text("1e")  # strtod at the very end of the buffer
    with pytest.raises(aq.AquaError):
        aq.Scene.load(str(p))


def test_jpeg_with_undefined_tables_or_huge_dimensions(aq, tmp_path):
    """SOS never checked that its Huffman tables exist; SOF dimensions were unbounded."""
    soi, eoi = b"\xff\xd8", b"\xff\xd9"
    dqt = b"\xff\xdb" + struct.pack(">H", 67) + b"\0" + bytes([1] * 64)
    def sof(w, h):
        return b"\xff\xc0" + struct.pack(">HBHHB", 11, 8, h, w, 1) + bytes([1, 0x11, 0])
    sos = b"\xff\xda" + struct.pack(">HB", 8, 1) + bytes([1, 0x00, 0, 63, 0]) + b"\x12\x34\x56" * 20
    for k, body in enumerate([dqt + sof(16, 16) + sos, dqt + sof(65535, 65535) + sos]):
        p = tmp_path / f"j{k}.jpg"
        p.write_bytes(soi + body + eoi)
        with pytest.raises(aq.AquaError):
            aq.decode_jpeg(str(p))
