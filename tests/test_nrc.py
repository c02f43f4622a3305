"""Row (f)-4 of SURVEY §8: the `nrc` integrator scenes/integrator.json names (:2 type, :4
batch_size, :6 training_iters, :7 learning_rate, :8 visualize_cache).  The reference holds no
implementation (src/lib.rs:0); the semantics are defined in aq_nrc.h and checked
here: an independent numpy restatement of the encoding / MLP / loss gradient / Adam, properties of
the records, that training reduces the error and that the cached render agrees with the path
tracer — on the CPU oracle; then the GPU against the oracle, bit for bit."""
import os

import numpy as np
import pytest

from test_area_lights import emissive_cbox, lamp_room

NW, HID = 16640, 4


def small_nrc(aq, **kw):
    a = dict(spp=4, max_depth=4, seed=3, type="nrc", batch_size=128, training_iters=24, learning_rate=2e-3)
    a.update(kw)
    return aq.Integrator(**a)


def np_forward(w, x):
    """float64 restatement: 4 hidden ReLU layers of 64, input-major weights, 64x4 output matrix"""
    a = np.asarray(x, np.float64)
    acts = [a]
    for l in range(HID):
        a = np.maximum(a @ w[l * 4096:(l + 1) * 4096].reshape(64, 64).astype(np.float64), 0)
        acts.append(a)
    return a @ w[HID * 4096:].reshape(64, 4).astype(np.float64)[:, :3], acts


def test_integrator_json_carries_the_nrc_keys(aq, scenes):
    it = aq.Integrator.load(os.path.join(scenes, "integrator.json"))
    assert (it.type, it.spp, it.max_depth, it.batch_size, it.training_iters, it.visualize_cache) == ("nrc", 4, 5, 512, 2048, False)
    assert abs(it.learning_rate - 1e-3) < 1e-9
    n = it.nrc_cfg()
    assert (n.batch_size, n.training_iters, n.visualize_cache) == (512, 2048, 0)


def test_initial_weights_and_forward_against_numpy(ao):
    w = ao.nrc_init_weights(7)
    assert w.shape == (NW,) and np.array_equal(w, ao.nrc_init_weights(7)) and not np.array_equal(w, ao.nrc_init_weights(8))
    hid, out = w[:HID * 4096], w[HID * 4096:].reshape(64, 4)
    b = np.sqrt(6 / 128)
    assert hid.min() >= -b and hid.max() <= b and abs(hid.mean()) < 2e-3 and abs(hid.std() - b / np.sqrt(3)) < 2e-3
    assert (out[:, 3] == 0).all() and np.abs(out[:, :3]).max() <= np.sqrt(6 / 67) + 1e-6
    x = np.random.default_rng(1).uniform(0, 1, (200, 64)).astype(np.float32)
    y = ao.nrc_forward(w, x)
    ref, _ = np_forward(w, x)
    assert np.allclose(y, ref, rtol=1e-4, atol=1e-5)


def test_records_describe_path_vertices(aq, ao, cbox):
    integ = small_nrc(aq, batch_size=256, training_iters=8)
    o = ao.OracleScene(cbox)
    cfg, nrc = integ.cfg(width=40, height=40), integ.nrc_cfg()
    x, y = o.nrc_records(cfg, nrc)
    assert x.shape == (2048, 64) and y.shape == (2048, 4)
    v = y[:, 3] == 1
    assert 0.6 < v.mean() <= 1.0 and set(np.unique(y[:, 3])) <= {0.0, 1.0}
    assert (x[~v] == 0).all() and (y[~v] == 0).all()
    xv = x[v]
    assert xv.min() >= 0 and xv.max() <= 1 and (xv[:, 61] == 1).all() and (xv[:, 62:] == 0).all()
    # triangle waves of the normalised position, octave 0 and 3
    assert np.allclose(xv[:, 3:6], np.abs(2 * xv[:, 0:3] - 1), atol=1e-6)
    assert np.allclose(xv[:, 12:15], np.abs(2 * ((xv[:, 0:3] * 8) % 1) - 1), atol=2e-5)
    # each one-blob (quartic kernel, 4 bins) has 1 or 2 non-zero bins
    nz = (xv[:, 27:55].reshape(-1, 7, 4) > 0).sum(-1)
    assert nz.min() >= 1 and nz.max() <= 2
    # targets are non-negative radiance / fac; first-hit records (even r) are valid at least as often as second-hit ones
    assert y[v, :3].min() >= 0 and y[v, :3].mean() > 0.01
    assert y[0::2, 3].mean() >= y[1::2, 3].mean()
    # deterministic, and independent of the thread count
    x1, y1 = o.nrc_records(cfg, nrc, n_threads=1)
    assert np.array_equal(x, x1) and np.array_equal(y, y1)
    assert np.isfinite(y).all()


def test_record_targets_estimate_scattered_radiance(aq, ao):
    """Mean target * fac of the first-hit records == mean radiance of a path-traced image minus
    the emission the camera sees directly, on a scene with an area light (both estimate the same
    integral over the image plane)."""
    sc = lamp_room(aq)
    o = ao.OracleScene(sc)
    integ = small_nrc(aq, batch_size=4096, training_iters=8, max_depth=3)
    cfg, nrc = integ.cfg(width=32, height=32), integ.nrc_cfg()
    x, y = o.nrc_records(cfg, nrc)
    even = np.arange(len(y)) % 2 == 0
    fac = np.maximum(x[:, 55:58] + x[:, 58:61], 0.02)
    L_rec = (y[:, :3] * fac)[even].mean(0)  # invalid records (camera ray missed) count as 0
    film, _, _ = o.render(aq.Integrator(spp=64, max_depth=3, seed=11).cfg(width=32, height=32))
    img = film[..., :3] / film[..., 3:]
    em1, _, _ = o.render(aq.Integrator(spp=64, max_depth=1, seed=11).cfg(width=32, height=32, flags=aq.AQ_RENDER_MIS_BSDF_ONLY))
    # max_depth 1 + BSDF-only: what is left is the emission seen directly by the camera
    direct_emission = (em1[..., :3] / em1[..., 3:]).mean((0, 1))
    want = img.mean((0, 1)) - direct_emission
    assert np.allclose(L_rec, want, rtol=0.06), (L_rec, want)


def test_one_training_step_against_numpy(aq, ao, cbox):
    """Gradient of the relative loss through the MLP and one Adam step, restated in float64."""
    integ = small_nrc(aq, batch_size=128, training_iters=1, learning_rate=1e-2)
    o = ao.OracleScene(cbox)
    cfg, nrc = integ.cfg(width=32, height=32), integ.nrc_cfg()
    w1, loss, x, y = o.nrc_train(cfg, nrc)
    w0 = ao.nrc_init_weights(cfg.seed)
    v = y[:, 3] == 1
    out, acts = np_forward(w0, x)
    t = y[:, :3].astype(np.float64)
    inv = 1.0 / (3 * 128)
    want_loss = ((out - t) ** 2 / (out ** 2 + 0.01) * inv)[v].sum()
    assert np.isclose(loss[0], want_loss, rtol=1e-4)
    dy = np.where(v[:, None], 2 * (out - t) / (out ** 2 + 0.01) * inv, 0)
    d = np.zeros((len(x), 64))
    d[:, :3] = dy
    grads = np.zeros(NW)
    for l in range(HID, -1, -1):
        cols = 4 if l == HID else 64
        Wl = w0[l * 4096:l * 4096 + 64 * cols].reshape(64, cols).astype(np.float64)
        grads[l * 4096:l * 4096 + 64 * cols] = (acts[l].T @ d[:, :cols]).ravel()
        if l > 0:
            d = (d[:, :cols] @ Wl.T) * (acts[l] > 0)
    # first Adam step: m = 0.1 g, v = 0.01 g^2, both bias-corrected back to g and g^2
    want = w0 - 1e-2 * grads / (np.abs(grads) + 1e-8)
    moved = np.abs(grads) > 1e-6  # (below that the sign-like first step is decided by rounding noise)
    assert moved.sum() > 4000
    assert np.allclose(w1[moved], want[moved], rtol=2e-3, atol=2e-5)
    assert (w1[HID * 4096:].reshape(64, 4)[:, 3] == 0).all()


def test_training_reduces_the_error_and_render_agrees_with_the_path_tracer(aq, ao, cbox):
    integ = small_nrc(aq, spp=8, max_depth=5, batch_size=256, training_iters=400, learning_rate=2e-3, seed=1)
    o = ao.OracleScene(cbox)
    cfg, nrc = integ.cfg(width=40, height=40), integ.nrc_cfg()
    w, loss, x, y = o.nrc_train(cfg, nrc)
    assert np.isfinite(w).all() and np.isfinite(loss).all()
    # the per-iteration loss is dominated by the noise of the 1-sample targets; judge the fit on
    # all records at once: relative error of the trained cache against the untrained one
    v = y[:, 3] == 1

    def rel_err(wt):
        out, _ = np_forward(wt, x[v])
        return np.mean((out - y[v, :3]) ** 2 / (out ** 2 + 0.01))
    assert rel_err(w) < 0.5 * rel_err(ao.nrc_init_weights(cfg.seed))
    assert np.mean(loss[-100:]) < np.mean(loss[:20])
    film, _, st = o.nrc_render(cfg, nrc, w)
    img = film[..., :3] / film[..., 3:]
    pt, _, st_pt = o.render(aq.Integrator(spp=64, max_depth=5, seed=2).cfg(width=40, height=40))
    ref = pt[..., :3] / pt[..., 3:]
    assert np.isfinite(img).all() and img.min() >= 0
    assert np.allclose(img.mean((0, 1)), ref.mean((0, 1)), rtol=0.25), (img.mean((0, 1)), ref.mean((0, 1)))
    # the cached render stops every path at its second hit: at most 2 closest-hit rays per sample
    assert st["rays_closest"] <= 2 * st["samples"] and st["rays_closest"] / st["samples"] < st_pt["rays_closest"] / st_pt["samples"]
    nrc.visualize_cache = 1
    vis, _, stv = o.nrc_render(cfg, nrc, w)
    assert stv["rays_closest"] == stv["samples"] and stv["rays_shadow"] == 0
    v = vis[..., :3] / vis[..., 3:]
    assert np.allclose(v.mean((0, 1)), ref.mean((0, 1)), rtol=0.35)
    # untrained weights give a different (wrong) image: the training is what makes it agree
    nrc.visualize_cache = 0
    bad, _, _ = o.nrc_render(cfg, nrc, ao.nrc_init_weights(1))
    b = bad[..., :3] / bad[..., 3:]
    assert np.abs(b.mean((0, 1)) - ref.mean((0, 1))).sum() > np.abs(img.mean((0, 1)) - ref.mean((0, 1))).sum()


def test_cache_with_max_depth_one_is_the_plain_first_vertex(aq, ao, cbox):
    """max_depth 1: no second vertex exists, so nothing is looked up (same image as the path tracer)."""
    integ = small_nrc(aq, max_depth=1)
    o = ao.OracleScene(cbox)
    cfg, nrc = integ.cfg(width=24, height=24), integ.nrc_cfg()
    film, _, _ = o.nrc_render(cfg, nrc, ao.nrc_init_weights(0))
    pt, _, _ = o.render(cfg)
    assert np.array_equal(film, pt)


# ---------------------------------------------------------------- GPU parity
@pytest.mark.gpu
def test_gpu_nrc_records_training_and_render_equal_the_oracle(aq, ao, cbox, renderer):
    for sc, kw in ((cbox, {}), (emissive_cbox(aq, cbox, keep_point_light=True), {"batch_size": 96})):
        integ = small_nrc(aq, **kw)
        cfg, nrc = integ.cfg(width=48, height=48), integ.nrc_cfg()
        ds, o = renderer.upload(sc), ao.OracleScene(sc)
        info = ds.nrc_train(cfg, nrc)
        R = nrc.batch_size * nrc.training_iters
        w, loss, x, y = o.nrc_train(cfg, nrc)
        gx, gy = ds.nrc_records(R)
        assert info["n_records"] == R and info["n_weights"] == NW and info["n_valid"] == int(y[:, 3].sum())
        assert np.array_equal(gx, x) and np.array_equal(gy, y)                 # records: bit for bit
        assert np.array_equal(ds.nrc_loss(nrc.training_iters), loss)            # every iteration's loss
        assert np.array_equal(ds.nrc_weights(), w)                              # the trained weights
        assert info["loss_first"] == loss[0] and info["loss_last"] == loss[-1]
        for vis in (0, 1):
            nrc.visualize_cache = vis
            cfg = integ.cfg(width=48, height=48, flags=aq.AQ_RENDER_DUMP_SAMPLES)
            film, st = ds.nrc_render(cfg, nrc)
            ofilm, osamples, ost = o.nrc_render(cfg, nrc, w, want_samples=True)
            assert np.array_equal(ds.samples(cfg), osamples) and np.array_equal(film, ofilm)
            for k in ("samples", "sample_bounces", "rays_closest", "rays_shadow"):
                assert st[k] == ost[k], k


@pytest.mark.gpu
def test_gpu_nrc_reference_config_on_room(aq, ao, room, renderer):
    """scenes/integrator.json as shipped (512 x 2048 records, lr 1e-3, spp 4) on room.json: trains,
    renders, and the cached image agrees with the path tracer's; set_weights round-trips."""
    it = aq.Integrator.load(os.path.join(aq.scenes_dir(), "integrator.json"))
    cfg, nrc = it.cfg(width=160, height=90), it.nrc_cfg()
    ds = renderer.upload(room)
    info = ds.nrc_train(cfg, nrc)
    assert info["n_records"] == 512 * 2048 and info["n_valid"] > 0.5 * info["n_records"]
    loss = ds.nrc_loss(nrc.training_iters)
    assert np.isfinite(loss).all() and np.median(loss[-100:]) < np.median(loss[:20])
    film, st = ds.nrc_render(cfg, nrc)
    img = film[..., :3] / film[..., 3:]
    pt, _ = ds.render(aq.Integrator(spp=64, max_depth=5, seed=5).cfg(width=160, height=90))
    ref = pt[..., :3] / pt[..., 3:]
    assert np.isfinite(img).all() and np.allclose(img.mean((0, 1)), ref.mean((0, 1)), rtol=0.3)
    w = ds.nrc_weights()
    ds2 = renderer.upload(room)
    with pytest.raises(aq.AquaError):
        ds2.nrc_render(cfg, nrc)  # no cache yet
    ds2.nrc_set_weights(w)
    film2, _ = ds2.nrc_render(cfg, nrc)
    assert np.array_equal(film, film2)
    print("nrc room:", info, {k: st[k] for k in ("ms_total", "samples", "rays_closest")})


@pytest.mark.gpu
def test_gpu_dist_renderer_honours_the_integrator_type(aq, cbox, renderer):
    """DistRenderer (what tools/render.py and bench.py drive) trains and uses the cache when the
    integrator says `nrc`; the film equals the C-ABI calls made by hand."""
    import torch
    from aqua_engine_b200 import dist as aqd
    integ = small_nrc(aq)
    dr = aqd.DistRenderer(cbox, 0)
    film = dr.render_async(integ, 48, 48, nrc_exact=True).clone()
    dr.finish()
    torch.cuda.synchronize()
    ds = renderer.upload(cbox)
    cfg, nrc = integ.cfg(width=48, height=48), integ.nrc_cfg()
    ds.nrc_train(cfg, nrc)
    want, _ = ds.nrc_render(cfg, nrc)
    assert dr.nrc_info["n_records"] == 128 * 24 and np.array_equal(film.cpu().numpy(), want)
    film_t = dr.render_async(integ, 48, 48)           # default: the tcgen05 lookup, same cache
    dr.finish()
    torch.cuda.synchronize()
    want_t, _ = ds.nrc_render(integ.cfg(width=48, height=48, flags=aq.AQ_RENDER_NRC_TENSOR), nrc)
    assert np.array_equal(film_t.cpu().numpy(), want_t) and not np.array_equal(want_t, want)
    integ.type = "pt"
    film = dr.render_async(integ, 48, 48)
    dr.finish()
    torch.cuda.synchronize()
    assert np.array_equal(film.cpu().numpy(), ds.render(cfg)[0])


def degenerate_scenes(aq):
    """(name, scene): nothing to hit at all / a single triangle far from the camera axis."""
    empty = aq.Scene.from_arrays(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint32),
                                 camera=aq.default_camera(res=(16, 16)))
    tri = aq.Scene.from_arrays(np.array([[50, 50, -5], [51, 50, -5], [50, 51, -5]], np.float32),
                               np.array([[0, 1, 2]], np.uint32), camera=aq.default_camera(res=(16, 16)),
                               lights=[aq.point_light((0, 0, 0), (1, 1, 1))])
    return [("empty", empty), ("off-axis triangle", tri)]


def test_nrc_on_scenes_without_valid_records(aq, ao):
    """No camera ray hits anything: every record is invalid, the descent steps are no-ops (zero
    gradient: the weights stay at their initial values, loss 0) and the render is black."""
    for name, sc in degenerate_scenes(aq):
        integ = small_nrc(aq, batch_size=80, training_iters=3)
        o = ao.OracleScene(sc)
        cfg, nrc = integ.cfg(width=16, height=16), integ.nrc_cfg()
        w, loss, x, y = o.nrc_train(cfg, nrc)
        assert (y == 0).all() and (x == 0).all() and (loss == 0).all(), name
        assert np.array_equal(w, ao.nrc_init_weights(cfg.seed)), name
        film, _, st = o.nrc_render(cfg, nrc, w)
        assert (film[..., :3] == 0).all() and (film[..., 3] == integ.spp).all() and st["sample_bounces"] == 0


@pytest.mark.gpu
def test_gpu_nrc_on_scenes_without_valid_records(aq, ao, renderer):
    for name, sc in degenerate_scenes(aq):
        integ = small_nrc(aq, batch_size=80, training_iters=3)
        ds, o = renderer.upload(sc), ao.OracleScene(sc)
        cfg, nrc = integ.cfg(width=16, height=16), integ.nrc_cfg()
        info = ds.nrc_train(cfg, nrc)
        w, loss, x, y = o.nrc_train(cfg, nrc)
        assert info["n_valid"] == 0 and np.array_equal(ds.nrc_weights(), w) and np.array_equal(ds.nrc_loss(3), loss)
        film, st = ds.nrc_render(cfg, nrc)
        ofilm, _, ost = o.nrc_render(cfg, nrc, w)
        assert np.array_equal(film, ofilm) and st["sample_bounces"] == 0


@pytest.mark.gpu
def test_gpu_nrc_argument_errors(aq, cbox, renderer):
    ds = renderer.upload(cbox)
    integ = small_nrc(aq)
    cfg, nrc = integ.cfg(width=16, height=16), integ.nrc_cfg()
    with pytest.raises(aq.AquaError) as e:
        ds.nrc_render(cfg, nrc)                      # not trained yet
    assert e.value.code == -5
    bad = integ.nrc_cfg()
    bad.batch_size = 0
    with pytest.raises(aq.AquaError) as e:
        ds.nrc_train(cfg, bad)
    assert e.value.code == -1
    bad = integ.nrc_cfg()
    bad.learning_rate = 0.0
    with pytest.raises(aq.AquaError):
        ds.nrc_train(cfg, bad)
    with pytest.raises(aq.AquaError):
        ds.nrc_set_weights(np.zeros(10, np.float32))  # wrong size
    ds.nrc_train(cfg, nrc)                            # and a valid call still works afterwards
    film, _ = ds.nrc_render(cfg, nrc)
    assert np.isfinite(film).all()


@pytest.mark.gpu
def test_gpu_nrc_tensor_lookup_within_tolerance(aq, cbox, room, renderer):
    """AQ_RENDER_NRC_TENSOR: the cache's MLP on the tensor cores (bf16 operands, fp32 accumulate) agrees
    with the exact fp32 lookup within the bf16 tolerance; everything else of the render is shared.
    Tolerances: per sample 3 % of the largest sample (a few bf16 ulps through 5 layers), image mean
    5e-3 relative.  First product run on B200: round 2, profiles/r02_nrc_tensor_first_product_run.log
    (room 1080p x 4 spp: 6.3 ms against 15.4 ms exact, max deviation 3.1e-4 of the film maximum)."""
    for sc, w, h in ((cbox, 96, 96), (room, 160, 90)):
        integ = small_nrc(aq, batch_size=256, training_iters=64, max_depth=5)
        ds = renderer.upload(sc)
        cfg, nrc = integ.cfg(width=w, height=h, flags=aq.AQ_RENDER_DUMP_SAMPLES), integ.nrc_cfg()
        ds.nrc_train(cfg, nrc)
        exact, ste = ds.nrc_render(cfg, nrc)
        se = ds.samples(cfg)
        cfgt = integ.cfg(width=w, height=h, flags=aq.AQ_RENDER_DUMP_SAMPLES | aq.AQ_RENDER_NRC_TENSOR)
        tens, stt = ds.nrc_render(cfgt, nrc)
        stn = ds.samples(cfgt)
        for k in ("samples", "sample_bounces", "rays_closest", "rays_shadow"):
            assert ste[k] == stt[k], k
        assert np.isfinite(tens).all() and np.array_equal(tens[..., 3], exact[..., 3])
        scale = np.abs(se[..., :3]).max()
        assert np.abs(stn[..., :3] - se[..., :3]).max() <= 0.03 * scale           # per sample: a few bf16 ulps through 5 layers
        me, mt = exact[..., :3].mean((0, 1)), tens[..., :3].mean((0, 1))
        assert np.allclose(mt, me, rtol=5e-3)                                       # no bias in the mean
        assert not np.array_equal(tens, exact)                                      # (it really is the other kernel)
