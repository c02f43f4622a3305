"""ctypes binding of oracle/libaqua_oracle.so — TEST INFRASTRUCTURE ONLY (see aq_oracle.cpp).

Import from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs only."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import aqua_engine_b200 as aq  # noqa: E402  (struct definitions only)
from aqua_engine_b200 import _abi  # noqa: E402

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "libaqua_oracle.so")
        if not os.path.exists(path):
            raise ImportError(f"{path} missing: run `make -C oracle`")
        L = C.CDLL(path)
        vp, u32, i = C.c_void_p, C.c_uint32, C.c_int
        L.aqo_scene_create.argtypes = [C.POINTER(_abi.SceneDesc), i, C.POINTER(vp)]
        L.aqo_scene_destroy.argtypes = [vp]
        L.aqo_scene_destroy.restype = None
        L.aqo_intersect.argtypes = [vp, vp, u32, vp, i, i, i]
        L.aqo_bvh8_intersect.argtypes = [vp, vp, vp, u32, vp, i, i, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.aqo_bvh8_intersect_step.argtypes = [vp, vp, vp, u32, vp, i, i, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), i]
        L.aqo_camera_rays.argtypes = [vp, C.POINTER(_abi.IntegratorCfg), u32, vp]
        L.aqo_render.argtypes = [vp, C.POINTER(_abi.IntegratorCfg), vp, vp, C.POINTER(_abi.Stats), i, i]
        L.aqo_sincos_2pi.argtypes = [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.aqo_sincos_2pi.restype = None
        L.aqo_rng_key.argtypes = [u32, u32, u32]
        L.aqo_rng_key.restype = u32
        L.aqo_rng.argtypes = [u32, u32]
        L.aqo_rng.restype = C.c_float
        fp = C.POINTER(C.c_float)
        L.aqo_tri_test.argtypes = [fp, fp, C.c_float, fp, fp, fp, fp]
        L.aqo_bsdf_eval.argtypes = [fp, fp, fp, fp, fp]
        L.aqo_bsdf_sample.argtypes = [fp, fp, fp, fp, fp, fp]
        L.aqo_bsdf_eval_full.argtypes = [fp, C.c_float, fp, fp, fp, fp]
        L.aqo_bsdf_sample_full.argtypes = [fp, C.c_float, fp, fp, fp, fp, fp]
        L.aqo_bsdf_eval_full_n.argtypes = [fp, C.c_float, fp, vp, u32, vp, vp, vp]
        L.aqo_bsdf_eval_aniso_n.argtypes = [fp, fp, C.c_float, fp, vp, u32, vp, vp, vp]
        L.aqo_bsdf_eval_aniso_n.restype = None
        L.aqo_bsdf_sample_aniso_n.argtypes = [fp, fp, C.c_float, fp, vp, u32, vp, vp, vp, vp]
        L.aqo_bsdf_sample_aniso_n.restype = None
        L.aqo_bsdf_eval_full_n.restype = None
        L.aqo_bsdf_sample_full_n.argtypes = [fp, C.c_float, fp, vp, u32, vp, vp, vp, vp]
        L.aqo_bsdf_sample_full_n.restype = None
        L.aqo_fresnel_dielectric.argtypes = [C.c_float, C.c_float]
        L.aqo_fresnel_dielectric.restype = C.c_float
        L.aqo_nrc_records.argtypes = [vp, C.POINTER(_abi.IntegratorCfg), C.POINTER(_abi.NrcCfg), vp, vp, i, i]
        L.aqo_nrc_init_weights.argtypes = [u32, vp]
        L.aqo_nrc_init_weights.restype = None
        L.aqo_nrc_forward.argtypes = [vp, vp, u32, vp]
        L.aqo_nrc_forward.restype = None
        L.aqo_nrc_fit.argtypes = [C.POINTER(_abi.NrcCfg), vp, vp, vp, vp]
        L.aqo_nrc_render.argtypes = [vp, C.POINTER(_abi.IntegratorCfg), C.POINTER(_abi.NrcCfg), vp, vp, vp,
                                     C.POINTER(_abi.Stats), i, i]
        L.aqo_threads.restype = i
        _lib = L
    return _lib


def threads():
    return lib().aqo_threads()


class OracleScene:
    BRUTE, BVH = 0, 1

    def __init__(self, scene, build_bvh=False):
        self.L = lib()
        h = C.c_void_p()
        rc = self.L.aqo_scene_create(C.byref(scene.desc), int(build_bvh), C.byref(h))
        if rc:
            raise RuntimeError(f"aqo_scene_create failed: {rc}")
        self.h = h
        self.res = (scene.desc.camera.res[0], scene.desc.camera.res[1])

    def intersect(self, rays, any_hit=False, mode=0, n_threads=0):
        rays = np.ascontiguousarray(rays, dtype=aq.RAY_DTYPE)
        hits = np.zeros(rays.shape[0], aq.HIT_DTYPE)
        rc = self.L.aqo_intersect(self.h, rays.ctypes.data, rays.shape[0], hits.ctypes.data, int(any_hit), mode, n_threads)
        assert rc == 0
        return hits

    def camera_rays(self, cfg, sample=0):
        w, h = cfg.width or self.res[0], cfg.height or self.res[1]
        rays = np.zeros(w * h, aq.RAY_DTYPE)
        assert self.L.aqo_camera_rays(self.h, C.byref(cfg), sample, rays.ctypes.data) == 0
        return rays

    def render(self, cfg, mode=0, n_threads=0, want_samples=False):
        w, h = cfg.width or self.res[0], cfg.height or self.res[1]
        film = np.zeros((h, w, 4), np.float32)
        samples = np.zeros((cfg.spp_end - cfg.spp_begin, h, w, 4), np.float32) if want_samples else None
        st = _abi.Stats()
        rc = self.L.aqo_render(self.h, C.byref(cfg), film.ctypes.data,
                               samples.ctypes.data if want_samples else None, C.byref(st), mode, n_threads)
        assert rc == 0
        return film, samples, st.as_dict()

    # ---- nrc integrator (aq_nrc.h)
    def nrc_records(self, cfg, nrc, mode=0, n_threads=0):
        """training records -> (x[R,64], y[R,4] = target.rgb / fac, valid)"""
        R = nrc.batch_size * nrc.training_iters
        x, y = np.zeros((R, 64), np.float32), np.zeros((R, 4), np.float32)
        assert self.L.aqo_nrc_records(self.h, C.byref(cfg), C.byref(nrc), x.ctypes.data, y.ctypes.data, mode, n_threads) == 0
        return x, y

    def nrc_train(self, cfg, nrc, mode=0, n_threads=0):
        """records + fit -> (weights[16640], loss[training_iters], x, y)"""
        x, y = self.nrc_records(cfg, nrc, mode, n_threads)
        w = nrc_init_weights(cfg.seed)
        loss = np.zeros(nrc.training_iters, np.float32)
        assert self.L.aqo_nrc_fit(C.byref(nrc), x.ctypes.data, y.ctypes.data, w.ctypes.data, loss.ctypes.data) == 0
        return w, loss, x, y

    def nrc_render(self, cfg, nrc, weights, mode=0, n_threads=0, want_samples=False):
        w, h = cfg.width or self.res[0], cfg.height or self.res[1]
        film = np.zeros((h, w, 4), np.float32)
        samples = np.zeros((cfg.spp_end - cfg.spp_begin, h, w, 4), np.float32) if want_samples else None
        weights = np.ascontiguousarray(weights, np.float32)
        st = _abi.Stats()
        rc = self.L.aqo_nrc_render(self.h, C.byref(cfg), C.byref(nrc), weights.ctypes.data, film.ctypes.data,
                                   samples.ctypes.data if want_samples else None, C.byref(st), mode, n_threads)
        assert rc == 0
        return film, samples, st.as_dict()

    def close(self):
        if self.h:
            self.L.aqo_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def nrc_init_weights(seed):
    w = np.zeros(_abi.NRC_N_WEIGHTS, np.float32)
    lib().aqo_nrc_init_weights(seed, w.ctypes.data)
    return w


def nrc_forward(weights, x):
    """the cache's MLP on n inputs x[n,64] -> y[n,3] (before the ReLU / fac of a query)"""
    weights = np.ascontiguousarray(weights, np.float32)
    x = np.ascontiguousarray(x, np.float32).reshape(-1, 64)
    y = np.zeros((len(x), 3), np.float32)
    lib().aqo_nrc_forward(weights.ctypes.data, x.ctypes.data, len(x), y.ctypes.data)
    return y


def bsdf_eval_full(params17, eta, wo, wis):
    """FULL Principled eval for one wo and n directions -> (f*cos (n,3), pdf (n,), ok (n,))."""
    p = np.ascontiguousarray(params17, np.float32)
    wo = np.ascontiguousarray(wo, np.float32)
    wis = np.ascontiguousarray(wis, np.float32).reshape(-1, 3)
    n = len(wis)
    f, pdf, ok = np.zeros((n, 3), np.float32), np.zeros(n, np.float32), np.zeros(n, np.uint8)
    fp = C.POINTER(C.c_float)
    lib().aqo_bsdf_eval_full_n(p.ctypes.data_as(fp), float(eta), wo.ctypes.data_as(fp), wis.ctypes.data, n,
                               f.ctypes.data, pdf.ctypes.data, ok.ctypes.data)
    return f, pdf, ok.astype(bool)


def bsdf_sample_full(params17, eta, wo, u3):
    """FULL Principled sampling for one wo and n random triples -> (wi, weight, pdf, ok)."""
    p = np.ascontiguousarray(params17, np.float32)
    wo = np.ascontiguousarray(wo, np.float32)
    u3 = np.ascontiguousarray(u3, np.float32).reshape(-1, 3)
    n = len(u3)
    wi, w = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
    pdf, ok = np.zeros(n, np.float32), np.zeros(n, np.uint8)
    fp = C.POINTER(C.c_float)
    lib().aqo_bsdf_sample_full_n(p.ctypes.data_as(fp), float(eta), wo.ctypes.data_as(fp), u3.ctypes.data, n,
                                 wi.ctypes.data, w.ctypes.data, pdf.ctypes.data, ok.ctypes.data)
    return wi, w, pdf, ok.astype(bool)


def bsdf_eval_aniso(params17, aniso3, eta, wo, wis):
    """FULL Principled eval with anisotropy: aniso3 = (anisotropic, anisotropic_rotation in turns, angle of the
    surface tangent in the shading frame in radians) -> (f*cos (n,3), pdf (n,), ok (n,))."""
    p = np.ascontiguousarray(params17, np.float32)
    a = np.ascontiguousarray(aniso3, np.float32)
    wo = np.ascontiguousarray(wo, np.float32)
    wis = np.ascontiguousarray(wis, np.float32).reshape(-1, 3)
    n = len(wis)
    f, pdf, ok = np.zeros((n, 3), np.float32), np.zeros(n, np.float32), np.zeros(n, np.uint8)
    fp = C.POINTER(C.c_float)
    lib().aqo_bsdf_eval_aniso_n(p.ctypes.data_as(fp), a.ctypes.data_as(fp), float(eta), wo.ctypes.data_as(fp),
                                wis.ctypes.data, n, f.ctypes.data, pdf.ctypes.data, ok.ctypes.data)
    return f, pdf, ok.astype(bool)


def bsdf_sample_aniso(params17, aniso3, eta, wo, u3):
    p = np.ascontiguousarray(params17, np.float32)
    a = np.ascontiguousarray(aniso3, np.float32)
    wo = np.ascontiguousarray(wo, np.float32)
    u3 = np.ascontiguousarray(u3, np.float32).reshape(-1, 3)
    n = len(u3)
    wi, w = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
    pdf, ok = np.zeros(n, np.float32), np.zeros(n, np.uint8)
    fp = C.POINTER(C.c_float)
    lib().aqo_bsdf_sample_aniso_n(p.ctypes.data_as(fp), a.ctypes.data_as(fp), float(eta), wo.ctypes.data_as(fp),
                                  u3.ctypes.data, n, wi.ctypes.data, w.ctypes.data, pdf.ctypes.data, ok.ctypes.data)
    return wi, w, pdf, ok.astype(bool)


def bvh8_intersect(nodes, tris, rays, any_hit=False, n_threads=0, step=0):
    """Walk a product-built BVH8 on the CPU (same traversal template as the kernel).
    step = 0: aq_trav_step, 1: the interleaved aq_trav_step2."""
    rays = np.ascontiguousarray(rays, dtype=aq.RAY_DTYPE)
    nodes = np.ascontiguousarray(nodes)
    tris = np.ascontiguousarray(tris) if len(tris) else np.zeros((1, 12), np.float32)
    hits = np.zeros(rays.shape[0], aq.HIT_DTYPE)
    nn, nt = C.c_uint64(), C.c_uint64()
    rc = lib().aqo_bvh8_intersect_step(nodes.ctypes.data, tris.ctypes.data, rays.ctypes.data, rays.shape[0],
                                       hits.ctypes.data, int(any_hit), n_threads, C.byref(nn), C.byref(nt), step)
    assert rc == 0
    return hits, nn.value, nt.value
