"""Parity of the CUDA path (through the C ABI, libaqua_cuda.so) against the CPU oracle.

Bars (BASELINE.json north_star): hit triangle ids bit-exact; per-sample radiance within 1e-4
relative under the same RNG stream (measured: bit-exact, the definitional functions are
single-sourced and compiled without FP contraction on both sides); film deterministic."""
import os

import numpy as np
import pytest

from conftest import hits_equal, random_rays, triangle_soup

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REL_TOL = 1e-4  # north_star's per-sample radiance tolerance


def rel_err(a, b):
    """largest PER-SAMPLE relative error (north_star: 1e-4 per sample under the same RNG stream); samples whose
    reference magnitude is below 1e-6 of the largest one are measured against that floor"""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    floor = 1e-6 * max(1e-20, float(np.abs(b).max()))
    return float((np.abs(a - b) / np.maximum(np.abs(b), floor)).max())


@pytest.fixture(scope="module")
def d_cbox(renderer, cbox):
    return renderer.upload(cbox)


@pytest.fixture(scope="module")
def d_room(renderer, room):
    return renderer.upload(room)


@pytest.fixture(scope="module")
def o_cbox(ao, cbox):
    return ao.OracleScene(cbox, build_bvh=True)


@pytest.fixture(scope="module")
def o_room(ao, room):
    return ao.OracleScene(room, build_bvh=True)


def test_device_is_b200(renderer):
    info = renderer.device_info()
    assert info["cc"][0] == 10 and info["sm_count"] >= 100


# ---------------------------------------------------------------- raygen (row a4)
def test_camera_rays_bit_exact(aq, d_cbox, o_cbox):
    for (w, h, s) in [(256, 256, 0), (64, 32, 5), (33, 17, 2)]:
        cfg = aq.Integrator(spp=8, seed=11).cfg(width=w, height=h)
        g = d_cbox.camera_rays(cfg, s)
        o = o_cbox.camera_rays(cfg, s)
        assert np.array_equal(g.view(np.uint8), o.view(np.uint8))


def test_thin_lens_camera_rays_bit_exact(aq, ao, renderer):
    pos = np.array([[-1, -1, 0], [1, -1, 0], [0, 1, 0]], np.float32)
    cam = aq.default_camera(res=(32, 32), fov=40.0, translate=(0.1, 0.2, 3))
    cam.lens_radius, cam.focal = 0.05, 3.0
    cam.rotate[:] = (0.1, -0.2, 0.05)
    sc = aq.Scene.from_arrays(pos, np.array([[0, 1, 2]], np.uint32), camera=cam)
    cfg = aq.Integrator(spp=1, seed=1).cfg()
    g = renderer.upload(sc).camera_rays(cfg, 3)
    o = ao.OracleScene(sc).camera_rays(cfg, 3)
    assert np.array_equal(g.view(np.uint8), o.view(np.uint8))
    assert np.abs(g["o"] - [0.1, 0.2, 3]).max() <= 0.05 + 1e-6 and np.abs(g["o"] - [0.1, 0.2, 3]).max() > 0


# ---------------------------------------------------------------- traversal (rows a6, a7)
def test_cbox_hit_ids_bit_exact_camera_and_random_rays(aq, d_cbox, o_cbox):
    cfg = aq.Integrator(spp=1).cfg(width=256, height=256)  # BASELINE config C1's ray set
    rays = d_cbox.camera_rays(cfg, 0)
    g = d_cbox.intersect(rays)
    o = o_cbox.intersect(rays, mode=0)  # brute force over all 36 triangles
    assert hits_equal(g, o)
    rr = random_rays(aq, 1 << 18, [-1.2, -0.2, -1.2], [1.2, 2.2, 1.2], seed=2)
    g, o = d_cbox.intersect(rr), o_cbox.intersect(rr, mode=0)
    assert hits_equal(g, o)
    assert not (set(np.unique(g["prim"]).tolist()) & {20, 21, 32, 33})  # duplicates: smaller id wins
    rr["tmax"] = np.where(np.arange(len(rr)) % 3 == 0, 0.6, 3e38).astype(np.float32)
    rr["tmin"] = np.where(np.arange(len(rr)) % 5 == 0, 0.3, 0.0).astype(np.float32)
    assert hits_equal(d_cbox.intersect(rr), o_cbox.intersect(rr, mode=0))
    assert np.array_equal(d_cbox.intersect(rr, any_hit=True)["prim"], o_cbox.intersect(rr, any_hit=True, mode=0)["prim"])


def test_cbox_hit_ids_match_golden_fixture(aq, d_cbox):
    g = np.load(os.path.join(GOLD, "cbox_golden.npz"))
    cfg = aq.Integrator(spp=4, max_depth=5, seed=0).cfg(width=64, height=64)
    rays = d_cbox.camera_rays(cfg, 0)
    assert np.array_equal(rays.view(np.float32).reshape(-1, 8), g["rays"])
    h = d_cbox.intersect(rays)
    assert np.array_equal(h["prim"], g["hit_prim"]) and np.array_equal(h["t"], g["hit_t"])


def test_room_hit_ids_bit_exact(aq, d_room, o_room, room):
    g0 = np.load(os.path.join(GOLD, "room_golden.npz"))
    cfg = aq.Integrator(spp=2, max_depth=5, seed=3).cfg(width=48, height=27)
    rays = d_room.camera_rays(cfg, 0)
    h = d_room.intersect(rays)
    assert np.array_equal(h["prim"], g0["hit_prim"]) and np.array_equal(h["t"], g0["hit_t"])  # brute-force fixture
    lo, hi = np.array(room.info.bounds_min), np.array(room.info.bounds_max)
    rr = random_rays(aq, 4096, lo, hi, seed=4)
    assert hits_equal(d_room.intersect(rr), o_room.intersect(rr, mode=0))       # brute force, 394,269 tris
    rr = random_rays(aq, 1 << 19, lo, hi, seed=5)
    assert hits_equal(d_room.intersect(rr), o_room.intersect(rr, mode=1))       # oracle BVH2
    rr["tmax"] = 1.5
    assert np.array_equal(d_room.intersect(rr, any_hit=True)["prim"], o_room.intersect(rr, any_hit=True, mode=1)["prim"])
    cfg = aq.Integrator(spp=1).cfg(width=640, height=360)
    cr = d_room.camera_rays(cfg, 0)
    assert hits_equal(d_room.intersect(cr), o_room.intersect(cr, mode=1))


def test_triangle_soup_hit_ids_bit_exact(aq, ao, renderer):
    """BASELINE config C4 at 1M triangles (the 10M run lives in bench/tools; same generator)."""
    pos, idx = triangle_soup(1_000_000)
    sc = aq.Scene.from_arrays(pos, idx)
    ds = renderer.upload(sc)
    assert ds.accel.n_tri_records == 1_000_000 and ds.accel.max_depth < 64
    o = ao.OracleScene(sc, build_bvh=True)
    rb = random_rays(aq, 1 << 12, 0.0, 1.0, seed=6)
    assert hits_equal(ds.intersect(rb), o.intersect(rb, mode=0))                # brute force
    rr = random_rays(aq, 1 << 20, 0.0, 1.0, seed=7)
    g = ds.intersect(rr)
    assert hits_equal(g, o.intersect(rr, mode=1))
    assert (g["prim"] != aq.AQ_MISS).mean() > 0.3
    # the downloaded tree walked on the CPU gives the same answer (isolates kernel from builder)
    nodes, tris = ds.download_accel()
    h8, nn, nt = ao.bvh8_intersect(nodes, tris, rr[: 1 << 16])
    assert hits_equal(h8, g[: 1 << 16])


def test_intersect_edge_cases(aq, ao, renderer):
    sc = aq.Scene.from_arrays(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint32))
    ds = renderer.upload(sc)
    rays = random_rays(aq, 1000, -1, 1)
    assert (ds.intersect(rays)["prim"] == aq.AQ_MISS).all()
    assert len(ds.intersect(rays[:0])) == 0                                      # empty ray set
    film, st = ds.render(aq.Integrator(spp=2).cfg(width=16, height=16))
    assert (film[..., :3] == 0).all() and (film[..., 3] == 2).all() and st["sample_bounces"] == 0
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0.5, 0.5, 0], [2, 2, 0]], np.float32)
    idx = np.array([[0, 1, 2], [0, 3, 3], [4, 4, 4], [0, 1, 2]], np.uint32)       # degenerate + duplicate
    sc = aq.Scene.from_arrays(pos, idx)
    ds, o = renderer.upload(sc), ao.OracleScene(sc)
    r = np.zeros(6, aq.RAY_DTYPE)
    r["o"] = [[0.2, 0.2, 1], [0.2, 0.2, 1], [0.5, 0.0, 1], [0.2, 0.2, 0.0], [5, 5, 1], [0.2, 0.2, 1]]
    r["d"] = [[0, 0, -1], [0, 0, 1], [0, 0, -1], [0, 0, -1], [0, 0, -1], [0, 0, -1]]
    r["tmax"] = [3e38, 3e38, 3e38, 3e38, 3e38, 1.0]                                # last: t == tmax is a hit
    assert hits_equal(ds.intersect(r), o.intersect(r))
    assert ds.intersect(r)["prim"].tolist() == [0, aq.AQ_MISS, 0, aq.AQ_MISS, aq.AQ_MISS, 0]


# ---------------------------------------------------------------- full path (rows a8-a13)
def test_cbox_per_sample_radiance_and_film(aq, d_cbox, o_cbox):
    """BASELINE config C1 shape at 128^2 x 8 spp: per-sample radiance, film, counters."""
    integ = aq.Integrator(spp=8, max_depth=5, seed=0)
    cfg = integ.cfg(width=128, height=128, flags=aq.AQ_RENDER_DUMP_SAMPLES)
    film, st = d_cbox.render(cfg)
    samples = d_cbox.samples(cfg)
    ofilm, osamples, ost = o_cbox.render(cfg, mode=0, want_samples=True)
    assert rel_err(samples, osamples) <= REL_TOL
    assert np.array_equal(samples, osamples), "expected bit-exact radiance (same FP sequence)"
    assert np.array_equal(film, ofilm), "film accumulates per pixel in sample order on both sides"
    for k in ("samples", "sample_bounces", "rays_closest", "rays_shadow"):
        assert st[k] == ost[k], (k, st[k], ost[k])
    assert np.isfinite(film).all() and (film[..., 3] == 8).all()


def test_cbox_c1_image_rmse_vs_oracle(aq, d_cbox, o_cbox):
    """Config C1 exactly: 256x256, 16 spp, depth 5.  RMSE threshold: 1e-6 of the mean radiance
    (in practice 0: the images are bit-identical)."""
    cfg = aq.Integrator(spp=16, max_depth=5, seed=0).cfg(width=256, height=256)
    film, st = d_cbox.render(cfg)
    ofilm, _, ost = o_cbox.render(cfg, mode=1)
    img, oimg = film[..., :3] / film[..., 3:], ofilm[..., :3] / ofilm[..., 3:]
    rmse = float(np.sqrt(np.mean((img - oimg) ** 2)))
    assert rmse <= 1e-6 * float(oimg.mean()), rmse
    assert st["samples"] == 256 * 256 * 16 and st["sample_bounces"] == ost["sample_bounces"]


def test_cbox_film_matches_golden_fixture(aq, d_cbox):
    g = np.load(os.path.join(GOLD, "cbox_golden.npz"))
    cfg = aq.Integrator(spp=4, max_depth=5, seed=0).cfg(width=32, height=32, flags=aq.AQ_RENDER_DUMP_SAMPLES)
    film, st = d_cbox.render(cfg)
    assert np.array_equal(film, g["film"]) and np.array_equal(d_cbox.samples(cfg), g["samples"])
    assert st["sample_bounces"] == int(g["sample_bounces"]) and st["rays_shadow"] == int(g["rays_shadow"])


def test_room_per_sample_radiance_textures_and_film(aq, d_room, o_room):
    g = np.load(os.path.join(GOLD, "room_golden.npz"))
    cfg = aq.Integrator(spp=2, max_depth=5, seed=3).cfg(width=48, height=27)
    film, st = d_room.render(cfg)
    assert np.array_equal(film, g["film"]) and st["sample_bounces"] == int(g["sample_bounces"])
    cfg = aq.Integrator(spp=4, max_depth=5, seed=9).cfg(width=160, height=90, flags=aq.AQ_RENDER_DUMP_SAMPLES)
    film, st = d_room.render(cfg)
    samples = d_room.samples(cfg)
    ofilm, osamples, ost = o_room.render(cfg, mode=1, want_samples=True)
    assert rel_err(samples, osamples) <= REL_TOL
    assert np.array_equal(film, ofilm)
    assert st["sample_bounces"] == ost["sample_bounces"] and st["rays_shadow"] == ost["rays_shadow"]


def test_balanced_triangle_phase_is_the_same_walk(aq, d_room, d_cbox, room, monkeypatch):
    """Round 2b: from depth 1 on (and in aq_intersect) the traversal kernels run the triangle phase of a
    step at warp level (aq_k_trace<.., DYN>): a lane with triangles left does not open a node in the next
    step.  Every ray must see the operation order of aq_trav_step: same hits, same t/u/v bits, the same
    number of nodes and triangle records fetched, the same film (AQUA_TRI_DYN=0 selects the plain step)."""
    lo, hi = np.array(room.info.bounds_min), np.array(room.info.bounds_max)
    rr = random_rays(aq, 1 << 18, lo, hi, seed=21)
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("AQUA_TRI_DYN", mode)
        d_room.trace_counters(reset=True)
        h = d_room.intersect(rr)
        c_closest = d_room.trace_counters(reset=True)
        rs = rr.copy()
        rs["tmax"] = 1.5
        a = d_room.intersect(rs, any_hit=True)
        c_any = d_room.trace_counters(reset=True)
        cfg = aq.Integrator(spp=4, max_depth=5, seed=5).cfg(width=160, height=90)
        film_room, st_room = d_room.render(cfg)
        film_cbox, st_cbox = d_cbox.render(aq.Integrator(spp=4, max_depth=5, seed=5).cfg(width=128, height=128))
        out[mode] = (h, a, c_closest, c_any, film_room, film_cbox, st_room["sample_bounces"], st_cbox["rays_shadow"])
    p, q = out["0"], out["1"]
    assert hits_equal(p[0], q[0]) and np.array_equal(p[1]["prim"], q[1]["prim"])
    assert p[2] == q[2] and p[2][0] > 0 and p[2][1] > 0   # closest: nodes, triangle records fetched
    assert p[3] == q[3] and p[3][0] > 0                   # any-hit: the per-ray test order is the same too
    assert np.array_equal(p[4], q[4]) and np.array_equal(p[5], q[5]) and p[6] == q[6] and p[7] == q[7]


# ---------------------------------------------------------------- size-independent properties
def test_spp_ranges_are_additive_and_pool_size_does_not_matter(aq, d_cbox):
    integ = aq.Integrator(spp=12, max_depth=5, seed=5)
    full, st = d_cbox.render(integ.cfg(width=200, height=120))
    # [0,5) then [5,12) accumulated == [0,12): the multi-GPU partition relies on this
    part, _ = d_cbox.render(integ.cfg(width=200, height=120, spp_begin=0, spp_end=5))
    part, _ = d_cbox.render(integ.cfg(width=200, height=120, spp_begin=5, spp_end=12, flags=aq.AQ_RENDER_ACCUMULATE), film=part)
    assert np.array_equal(part, full)
    # pool smaller than the image (tiles), odd pool, pool holding many samples per pixel
    for pool in (1024, 5000, 24000, 1 << 18):
        f, s = d_cbox.render(integ.cfg(width=200, height=120, pool_paths=pool))
        assert np.array_equal(f, full), pool
        assert s["sample_bounces"] == st["sample_bounces"]
    # determinism run to run
    again, _ = d_cbox.render(integ.cfg(width=200, height=120))
    assert np.array_equal(again, full)
    # a different seed gives a different (but statistically equal) image
    other, _ = d_cbox.render(aq.Integrator(spp=12, max_depth=5, seed=6).cfg(width=200, height=120))
    assert not np.array_equal(other, full)
    assert abs(other[..., :3].mean() / full[..., :3].mean() - 1) < 0.05


def test_max_depth_semantics(aq, d_cbox, o_cbox):
    for md in (1, 2, 3):
        cfg = aq.Integrator(spp=2, max_depth=md, seed=1).cfg(width=64, height=64)
        film, st = d_cbox.render(cfg)
        ofilm, _, ost = o_cbox.render(cfg)
        assert np.array_equal(film, ofilm) and st["rays_closest"] == ost["rays_closest"]
        assert st["sample_bounces"] <= st["samples"] * md


def test_full_resolution_properties_c2(aq, d_cbox):
    """BASELINE config C2 geometry (1024x1024, depth 5) at 8 spp: counters are consistent and
    the device-resident path equals the host-buffer path."""
    cfg = aq.Integrator(spp=8, max_depth=5, seed=0).cfg(width=1024, height=1024)
    film, st = d_cbox.render(cfg)
    assert st["samples"] == 8 * 1024 * 1024
    assert st["rays_closest"] >= st["sample_bounces"] >= st["rays_shadow"] > 0
    assert st["rays_closest"] <= st["samples"] + st["sample_bounces"]
    assert (film[..., 3] == 8).all() and np.isfinite(film).all() and film[..., :3].min() >= 0
    import torch
    d_film = torch.zeros(1024, 1024, 4, device="cuda")
    torch.cuda.synchronize()
    d_cbox.render_device_async(cfg, d_film.data_ptr())
    st2 = d_cbox.finish()
    assert np.array_equal(d_film.cpu().numpy(), film) and st2["sample_bounces"] == st["sample_bounces"]
    # the middle of the image sees the back wall: mean radiance well above zero
    assert film[400:600, 400:600, :3].mean() / 8 > 0.02


# ---------------------------------------------------------------- error behaviour
def test_error_paths(aq, renderer, cbox):
    ds = renderer.upload(cbox, build=False)
    cfg = aq.Integrator(spp=1).cfg(width=8, height=8)
    with pytest.raises(aq.AquaError) as e:
        ds.render(cfg)
    assert e.value.code == -5  # AQ_ERR_STATE: accel not built
    with pytest.raises(aq.AquaError):
        ds.intersect(random_rays(aq, 4, -1, 1))
    ds.build()
    bad = aq.Integrator(spp=1, max_depth=0).cfg(width=8, height=8)
    with pytest.raises(aq.AquaError) as e:
        ds.render(bad)
    assert e.value.code == -1
    with pytest.raises(aq.AquaError):
        ds.samples(cfg)  # last render had no AQ_RENDER_DUMP_SAMPLES
    pos, idx, *_ = cbox.arrays()
    broken = aq.Scene.from_arrays(pos, idx.copy() + 1000)
    with pytest.raises(aq.AquaError) as e:
        renderer.upload(broken)
    assert e.value.code == -1 and "out of range" in str(e.value)
    with pytest.raises(aq.AquaError):
        aq.Renderer(99)


# ---------------------------------------------------------------- device BVH builder (row a5)
def _with_builder(kind, fn):
    old = os.environ.get("AQUA_ACCEL_BUILDER")
    os.environ["AQUA_ACCEL_BUILDER"] = kind
    try:
        return fn()
    finally:
        if old is None:
            del os.environ["AQUA_ACCEL_BUILDER"]
        else:
            os.environ["AQUA_ACCEL_BUILDER"] = old


def test_hybrid_build_renders_the_same_film_before_and_after_the_swap(aq, renderer, room, cbox):
    """room.json builds HYBRID by default (device LBVH tree at once, host SAH tree swapped in by the
    render loop when its thread is done): the film must not depend on which tree a wave used."""
    ds = renderer.upload(room)
    assert ds.accel.builder == 2                      # usable before the SAH tree exists
    cfg = aq.Integrator(spp=2, max_depth=5, seed=9).cfg(width=320, height=180, pool_paths=1 << 14)  # many waves
    early, st_e = ds.render(cfg)                      # starts on the LBVH tree; may switch between two waves
    n_lbvh = ds.accel.n_nodes
    info = ds.accel_wait()
    assert info.builder == 0 and info.n_nodes != n_lbvh and info.sah_cost > 0
    late, st_l = ds.render(cfg)
    assert np.array_equal(early, late)
    for k in ("samples", "sample_bounces", "rays_closest", "rays_shadow"):
        assert st_e[k] == st_l[k]
    nodes, tris = ds.download_accel()                  # sized from the final tree
    assert nodes.shape[0] == info.n_nodes
    # forced on a tiny scene; destroying the scene right away joins the builder thread
    dc = _with_builder("hybrid", lambda: renderer.upload(cbox))
    assert dc.accel.builder == 2
    dc.close()
    dc = _with_builder("hybrid", lambda: renderer.upload(cbox))
    f1, _ = dc.render(aq.Integrator(spp=2, seed=1).cfg(width=64, height=64))
    dc.accel_wait()
    f2, _ = dc.render(aq.Integrator(spp=2, seed=1).cfg(width=64, height=64))
    assert np.array_equal(f1, f2)


def test_device_lbvh_builder_hit_ids_bit_exact(aq, ao, renderer, cbox, room, o_cbox, o_room):
    """The Morton/LBVH device builder yields a different tree, never a different answer."""
    ds = _with_builder("device", lambda: renderer.upload(cbox))
    assert ds.accel.builder == 1 and ds.accel.n_tri_records == 36
    rr = random_rays(aq, 1 << 16, [-1.2, -0.2, -1.2], [1.2, 2.2, 1.2], seed=12)
    assert hits_equal(ds.intersect(rr), o_cbox.intersect(rr, mode=0))
    cfg = aq.Integrator(spp=4, max_depth=5, seed=0).cfg(width=64, height=64)
    film, st = ds.render(cfg)
    ofilm, _, ost = o_cbox.render(cfg)
    assert np.array_equal(film, ofilm) and st["sample_bounces"] == ost["sample_bounces"]

    dr = _with_builder("device", lambda: renderer.upload(room))
    assert dr.accel.builder == 1 and dr.accel.max_depth < 64
    lo, hi = np.array(room.info.bounds_min), np.array(room.info.bounds_max)
    rr = random_rays(aq, 4096, lo, hi, seed=13)
    assert hits_equal(dr.intersect(rr), o_room.intersect(rr, mode=0))            # brute force
    rr = random_rays(aq, 1 << 18, lo, hi, seed=14)
    assert hits_equal(dr.intersect(rr), o_room.intersect(rr, mode=1))
    rr["tmax"] = 2.0
    assert np.array_equal(dr.intersect(rr, any_hit=True)["prim"], o_room.intersect(rr, any_hit=True, mode=1)["prim"])
    # every triangle appears exactly once in the device-built records
    nodes, tris = dr.download_accel()
    assert np.array_equal(np.sort(tris.view(np.uint32)[:, 9]), np.arange(room.info.n_tris, dtype=np.uint32))


def test_device_builder_degenerate_inputs(aq, ao, renderer):
    for n in (1, 2, 3, 4, 9):
        pos, idx = triangle_soup(n, seed=n, r=0.2)
        sc = aq.Scene.from_arrays(pos, idx)
        ds = _with_builder("device", lambda: renderer.upload(sc))
        assert ds.accel.builder == 1
        rr = random_rays(aq, 2048, -0.2, 1.2, seed=n)
        assert hits_equal(ds.intersect(rr), ao.OracleScene(sc).intersect(rr, mode=0)), n
    # many identical triangles: duplicate Morton codes, deep index-split subtree
    pos = np.tile(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32), (500, 1))
    idx = np.arange(1500, dtype=np.uint32).reshape(-1, 3)
    sc = aq.Scene.from_arrays(pos, idx)
    ds = _with_builder("device", lambda: renderer.upload(sc))
    rr = random_rays(aq, 4096, -0.5, 1.5, seed=3)
    g = ds.intersect(rr)
    assert hits_equal(g, ao.OracleScene(sc).intersect(rr, mode=0))
    assert set(np.unique(g["prim"]).tolist()) <= {0, aq.AQ_MISS}                 # smallest id wins every tie


def test_zero_samples_and_stream_override(aq, renderer, cbox):
    ds = renderer.upload(cbox)
    film, st = ds.render(aq.Integrator(spp=0).cfg(width=16, height=16))
    assert (film == 0).all() and st["samples"] == 0 and st["n_waves"] == 0
    # run on torch's current (legacy default) stream and time with torch events
    import torch
    r2 = aq.Renderer(0)
    r2.set_stream(torch.cuda.current_stream().cuda_stream)
    d2 = r2.upload(cbox)
    cfg = aq.Integrator(spp=4, max_depth=5, seed=0).cfg(width=64, height=64)
    dfilm = torch.zeros(64, 64, 4, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    d2.render_device_async(cfg, dfilm.data_ptr())
    e1.record()
    torch.cuda.synchronize()
    assert e0.elapsed_time(e1) > 0.05  # the kernels really ran between the two torch events
    ref, _ = ds.render(cfg)
    assert np.array_equal(dfilm.cpu().numpy(), ref)
    st2 = d2.finish()
    assert 0 < st2["ms_total"] <= e0.elapsed_time(e1) + 0.05  # the library's own events sit inside the torch bracket
