"""Multi-GPU film combine: N-GPU film == 1-GPU film up to fp32 summation order,
sample counts and path counters exact.  Needs >= 2 GPUs for the NCCL cases."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def n_gpus():
    import torch
    return torch.cuda.device_count()


def test_render_multi_single_device_equals_render(aq, renderer, cbox):
    cfg = aq.Integrator(spp=6, max_depth=5, seed=4).cfg(width=96, height=64)
    film1, st1 = renderer.upload(cbox).render(cfg)
    filmm, stm = aq.render_multi(cbox, cfg, 1)
    assert np.array_equal(film1, filmm)
    assert stm["sample_bounces"] == st1["sample_bounces"] and stm["samples"] == st1["samples"]


@pytest.mark.skipif("n_gpus() < 2")
@pytest.mark.parametrize("n", [2, 4, 8])
def test_render_multi_nccl_reduce(aq, renderer, cbox, n):
    if n_gpus() < n:
        pytest.skip(f"needs {n} GPUs")
    cfg = aq.Integrator(spp=16, max_depth=5, seed=4).cfg(width=128, height=128)
    film1, st1 = renderer.upload(cbox).render(cfg)
    filmn, stn = aq.render_multi(cbox, cfg, n)
    assert np.array_equal(filmn[..., 3], film1[..., 3])              # counts add exactly
    for k in ("samples", "sample_bounces", "rays_closest", "rays_shadow"):
        assert stn[k] == st1[k]                                       # same sample set
    assert np.allclose(filmn, film1, rtol=2e-5, atol=1e-6)            # fp32 summation order only
    assert not np.array_equal(filmn, np.zeros_like(filmn))
