"""output stage — PNG / PFM writers on the host, resolve kernel on the device."""
import os

import numpy as np
import pytest


def test_png_writer_round_trip(aq, tmp_path):
    from PIL import Image
    g = np.random.default_rng(0)
    for (h, w) in [(1, 1), (37, 53), (300, 257)]:   # > 65535 raw bytes => several stored-deflate blocks
        img = g.integers(0, 256, (h, w, 4)).astype(np.uint8)
        p = str(tmp_path / f"t{h}.png")
        aq.write_png(p, img)
        assert np.array_equal(np.asarray(Image.open(p)), img)


def test_srgb_threshold_table_matches_the_oetf(aq):
    t = np.zeros(255, np.float32)
    aq._abi.host_lib().aq_host_srgb_thresholds(t.ctypes.data)
    assert (np.diff(t) > 0).all() and 0 < t[0] < 1e-3 and 0.99 < t[-1] < 1
    v = np.linspace(0, 1, 20001).astype(np.float64)
    oetf = np.where(v <= 0.0031308, 12.92 * v, 1.055 * v ** (1 / 2.4) - 0.055)
    want = np.round(oetf * 255).astype(int)
    got = np.searchsorted(t, v.astype(np.float32), side="right")
    assert np.abs(got - want).max() <= 1 and (got != want).mean() < 2e-3   # only at exact half-way points


def test_pfm_writer(aq, tmp_path):
    film = np.zeros((3, 2, 4), np.float32)
    film[..., :3] = np.arange(18, dtype=np.float32).reshape(3, 2, 3)
    film[..., 3] = 2
    p = str(tmp_path / "f.pfm")
    assert aq._abi.host_lib().aq_host_write_pfm(os.fsencode(p), film.ctypes.data, 2, 3) == 0
    raw = open(p, "rb").read()
    assert raw.startswith(b"PF\n2 3\n-1.0\n")
    data = np.frombuffer(raw[len(b"PF\n2 3\n-1.0\n"):], np.float32).reshape(3, 2, 3)
    assert np.array_equal(data[::-1], film[..., :3] / 2)


@pytest.mark.gpu
def test_resolve_kernel_is_exact(aq, renderer, cbox):
    ds = renderer.upload(cbox)
    film, _ = ds.render(aq.Integrator(spp=8, max_depth=5, seed=1).cfg(width=160, height=96))
    for exposure in (1.0, 2.5):
        got = renderer.resolve(film, exposure=exposure)
        assert np.array_equal(got, aq.srgb8_reference(film, exposure))
    edge = np.zeros((2, 3, 4), np.float32)
    edge[0, 0] = (np.nan, -1.0, 1e30, 1)      # NaN / negative -> 0, huge -> 255
    edge[0, 1] = (0.5, 0.5, 0.5, 0)           # zero weight -> black
    edge[1, 2] = (2.0, 0.2, 0.02, 2)
    out = renderer.resolve(edge)
    assert out[0, 0].tolist() == [0, 0, 255, 255] and out[0, 1].tolist() == [0, 0, 0, 255]
    assert np.array_equal(out[1, 2], aq.srgb8_reference(edge)[1, 2])
    import torch
    d = torch.from_numpy(film).cuda()
    assert np.array_equal(renderer.resolve(film.shape[:2], d_film_ptr=d.data_ptr()), aq.srgb8_reference(film))
