"""N>1 host logic on CPU: spp partition + the film reduce over gloo with world_size 2.
Each rank renders its share with the CPU oracle (this is a test; the product renders on the
GPU) and the reduced film must equal a single full render up to fp32 summation order."""
import os
import socket
import sys

import numpy as np
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_spp_covers_range_disjointly():
    sys.path.insert(0, ROOT)
    import aqua_engine_b200  # noqa: F401
    from aqua_engine_b200.dist import partition_spp
    for (b, e) in [(0, 1024), (3, 10), (0, 1), (5, 5), (0, 4096)]:
        for world in (1, 2, 3, 4, 8):
            parts = [partition_spp(b, e, r, world) for r in range(world)]
            assert parts[0][0] == b and parts[-1][1] == e
            for (a0, a1), (b0, b1) in zip(parts, parts[1:]):
                assert a1 == b0 and a0 <= a1
            sizes = [p[1] - p[0] for p in parts]
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == e - b


def _worker(rank, world, port, out, nrc=False):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import aq_oracle as ao
    import aqua_engine_b200 as aq
    from aqua_engine_b200 import dist as aqd
    r, w, _ = aqd.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    scene = aq.Scene.load(os.path.join(aq.scenes_dir(), "cbox.json"))
    o = ao.OracleScene(scene)
    integ = aq.Integrator(spp=6, max_depth=5, seed=2, type="nrc" if nrc else "pt", batch_size=64, training_iters=12)
    b, e = aqd.partition_spp(0, integ.spp, r, w)
    cfg = integ.cfg(width=48, height=48, spp_begin=b, spp_end=e)
    if nrc:
        # every rank trains the cache itself: records and descent are deterministic, so the weights
        # are the same everywhere and the only exchange stays the film reduce
        wts, _, _, _ = o.nrc_train(cfg, integ.nrc_cfg(), n_threads=2)
        all_w = [torch.zeros(len(wts)) for _ in range(w)]
        torch.distributed.all_gather(all_w, torch.from_numpy(wts))
        assert all(torch.equal(all_w[0], x) for x in all_w)
        film, _, st = o.nrc_render(cfg, integ.nrc_cfg(), wts, n_threads=2)
    else:
        film, _, st = o.render(cfg, n_threads=2)
    t = torch.from_numpy(film)
    aqd.reduce_film(t, 0)
    if r == 0:
        np.save(out, t.numpy())
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_two_rank_film_reduce_equals_single_render(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "film.npy")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import aq_oracle as ao
    import aqua_engine_b200 as aq
    scene = aq.Scene.load(os.path.join(aq.scenes_dir(), "cbox.json"))
    full, _, _ = ao.OracleScene(scene).render(aq.Integrator(spp=6, max_depth=5, seed=2).cfg(width=48, height=48))
    got = np.load(out)
    assert np.array_equal(got[..., 3], full[..., 3])          # sample counts add exactly
    # same sample set, different summation order: fp32 tolerance ~ 1e-6 * spp
    assert np.allclose(got, full, rtol=1e-5, atol=1e-6)


def test_two_rank_nrc_render_equals_single_render(tmp_path):
    """the `nrc` integrator under the spp partition: identical caches on both ranks, reduced film ==
    one rank rendering all samples with that cache"""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "film_nrc.npy")
    mp.spawn(_worker, args=(2, port, out, True), nprocs=2, join=True)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import aq_oracle as ao
    import aqua_engine_b200 as aq
    scene = aq.Scene.load(os.path.join(aq.scenes_dir(), "cbox.json"))
    integ = aq.Integrator(spp=6, max_depth=5, seed=2, type="nrc", batch_size=64, training_iters=12)
    o = ao.OracleScene(scene)
    cfg = integ.cfg(width=48, height=48)
    wts, _, _, _ = o.nrc_train(cfg, integ.nrc_cfg())
    full, _, _ = o.nrc_render(cfg, integ.nrc_cfg(), wts)
    got = np.load(out)
    assert np.array_equal(got[..., 3], full[..., 3])
    assert np.allclose(got, full, rtol=1e-5, atol=1e-6)
