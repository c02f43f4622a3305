"""Device side: aq_ctx / aq_scene handles of libaqua_cuda.so (include/aqua_cuda.h)."""
import ctypes as C

import weakref

import numpy as np

from . import _abi
from ._abi import AQ_RENDER_ACCUMULATE, AQ_RENDER_DUMP_SAMPLES, Stats

RAY_DTYPE = np.dtype([("o", np.float32, 3), ("tmin", np.float32), ("d", np.float32, 3), ("tmax", np.float32)])
HIT_DTYPE = np.dtype([("prim", np.uint32), ("t", np.float32), ("u", np.float32), ("v", np.float32)])


class Renderer:
    """One aq_ctx (one per device)."""

    def __init__(self, device=0):
        self.lib = _abi.cuda_lib()
        h = C.c_void_p()
        _abi.check(self.lib.aq_init(device, C.byref(h)))
        self.handle = h
        self.device = device
        self._scenes = weakref.WeakSet()

    def set_stream(self, cuda_stream_ptr):
        """Run on an external cudaStream_t.  torch reports the legacy default stream as 0;
        the C ABI reserves NULL for "private stream", so 0 is mapped to cudaStreamLegacy (0x1)."""
        _abi.check(self.lib.aq_set_stream(self.handle, C.c_void_p(cuda_stream_ptr or 1)), self.handle)

    def device_info(self):
        sm, ma, mi, hb = C.c_int(), C.c_int(), C.c_int(), C.c_size_t()
        _abi.check(self.lib.aq_device_info(self.handle, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(hb)), self.handle)
        return {"sm_count": sm.value, "cc": (ma.value, mi.value), "hbm_bytes": hb.value}

    def resolve(self, film, exposure=1.0, d_film_ptr=None):
        """Output stage on the device: float4 film -> uint8 sRGB image [H,W,4] (aq_resolve).
        `film` is a host array [H,W,4]; or pass (h, w) as `film` with a device pointer."""
        if d_film_ptr is not None:
            h, w = film
            src_h = None
        else:
            film = np.ascontiguousarray(film, dtype=np.float32)
            h, w = film.shape[:2]
            src_h = film.ctypes.data
        out = np.zeros((h, w, 4), np.uint8)
        _abi.check(self.lib.aq_resolve(self.handle, C.c_void_p(d_film_ptr) if d_film_ptr else None, src_h, w, h,
                                       float(exposure), out.ctypes.data), self.handle)
        return out

    def upload(self, scene, build=True):
        ds = DeviceScene(self, scene, build)
        self._scenes.add(ds)
        return ds

    def close(self):
        """Destroys the ctx; its DeviceScenes are closed first (aq_destroy would take them along
        and leave their handles dangling)."""
        if self.handle:
            for ds in list(self._scenes):
                ds.close()
            self.lib.aq_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceScene:
    """One aq_scene: geometry, materials, textures and the BVH8, resident in HBM."""

    def __init__(self, renderer, scene, build=True):
        self.r = renderer
        self.lib = renderer.lib
        h = C.c_void_p()
        _abi.check(self.lib.aq_scene_create(renderer.handle, C.byref(scene.desc), C.byref(h)), renderer.handle)
        self.handle = h
        self.res = (scene.desc.camera.res[0], scene.desc.camera.res[1])
        self.accel = None
        if build:
            self.build()

    def _ck(self, rc):
        _abi.check(rc, self.r.handle)

    def build(self):
        info = _abi.AccelInfo()
        self._ck(self.lib.aq_accel_build(self.handle, C.byref(info)))
        self.accel = info
        return info

    def accel_wait(self):
        """Hybrid build (accel.builder == 2): block until the host SAH tree has replaced the device
        LBVH tree; refreshes `self.accel` with the final tree's figures."""
        info = _abi.AccelInfo()
        self._ck(self.lib.aq_accel_wait(self.handle, C.byref(info)))
        self.accel = info
        return info

    def download_accel(self):
        self.accel_wait()
        n, t = self.accel.n_nodes, self.accel.n_tri_records
        nodes = np.zeros((n, 20), np.uint32)
        tris = np.zeros((max(t, 1), 12), np.float32)
        self._ck(self.lib.aq_accel_download(self.handle, nodes.ctypes.data, nodes.nbytes, tris.ctypes.data, tris.nbytes))
        return nodes, tris[:t]

    # ---- intersection hook
    def intersect(self, rays, any_hit=False):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.zeros(rays.shape[0], HIT_DTYPE)
        self._ck(self.lib.aq_intersect(self.handle, rays.ctypes.data, rays.shape[0], hits.ctypes.data, int(any_hit)))
        return hits

    def intersect_device(self, d_rays_ptr, n, d_hits_ptr, any_hit=False):
        self._ck(self.lib.aq_intersect_device_async(self.handle, C.c_void_p(d_rays_ptr), n, C.c_void_p(d_hits_ptr), int(any_hit)))

    def trace_counters(self, reset=True):
        """(BVH8 nodes fetched, triangle records fetched) by intersect calls since the last reset."""
        a, b = C.c_uint64(), C.c_uint64()
        self._ck(self.lib.aq_trace_counters(self.handle, C.byref(a), C.byref(b), int(reset)))
        return a.value, b.value

    def camera_rays(self, cfg, sample=0):
        w = cfg.width or self.res[0]
        h = cfg.height or self.res[1]
        rays = np.zeros(w * h, RAY_DTYPE)
        self._ck(self.lib.aq_generate_camera_rays(self.handle, C.byref(cfg), sample, rays.ctypes.data))
        return rays

    # ---- render
    def render(self, cfg, film=None):
        """Synchronous render through host buffers: returns (film[H,W,4], stats dict)."""
        w = cfg.width or self.res[0]
        h = cfg.height or self.res[1]
        if film is None:
            film = np.zeros((h, w, 4), np.float32)
        st = Stats()
        self._ck(self.lib.aq_render(self.handle, C.byref(cfg), film.ctypes.data, C.byref(st)))
        return film, st.as_dict()

    def render_device_async(self, cfg, d_film_ptr=None):
        self._ck(self.lib.aq_render_device_async(self.handle, C.byref(cfg), C.c_void_p(d_film_ptr) if d_film_ptr else None))

    def finish(self):
        st = Stats()
        self._ck(self.lib.aq_render_finish(self.handle, C.byref(st)))
        return st.as_dict()

    def samples(self, cfg):
        w = cfg.width or self.res[0]
        h = cfg.height or self.res[1]
        n = (cfg.spp_end - cfg.spp_begin) * w * h
        out = np.zeros((cfg.spp_end - cfg.spp_begin, h, w, 4), np.float32)
        self._ck(self.lib.aq_render_samples(self.handle, out.ctypes.data, n))
        return out

    # ---- nrc integrator (scenes/integrator.json:2; aq_nrc_* in include/aqua_cuda.h)
    def nrc_train(self, cfg, nrc):
        """Generate the training records and fit the scene's radiance cache; returns aq_nrc_info."""
        info = _abi.NrcInfo()
        self._ck(self.lib.aq_nrc_train(self.handle, C.byref(cfg), C.byref(nrc), C.byref(info)))
        return info.as_dict()

    def nrc_render(self, cfg, nrc, film=None):
        """Render with the trained cache: returns (film[H,W,4], stats dict)."""
        w = cfg.width or self.res[0]
        h = cfg.height or self.res[1]
        if film is None:
            film = np.zeros((h, w, 4), np.float32)
        st = Stats()
        self._ck(self.lib.aq_nrc_render(self.handle, C.byref(cfg), C.byref(nrc), film.ctypes.data, C.byref(st)))
        return film, st.as_dict()

    def nrc_render_device_async(self, cfg, nrc, d_film_ptr=None):
        self._ck(self.lib.aq_nrc_render_device_async(self.handle, C.byref(cfg), C.byref(nrc),
                                                     C.c_void_p(d_film_ptr) if d_film_ptr else None))

    def nrc_weights(self):
        w = np.zeros(_abi.NRC_N_WEIGHTS, np.float32)
        self._ck(self.lib.aq_nrc_get_weights(self.handle, w.ctypes.data, w.size))
        return w

    def nrc_set_weights(self, w):
        w = np.ascontiguousarray(w, np.float32)
        self._ck(self.lib.aq_nrc_set_weights(self.handle, w.ctypes.data, w.size))

    def nrc_loss(self, n_iters):
        l = np.zeros(n_iters, np.float32)
        self._ck(self.lib.aq_nrc_get_loss(self.handle, l.ctypes.data, n_iters))
        return l

    def nrc_records(self, n_records):
        x, y = np.zeros((n_records, _abi.NRC_IN), np.float32), np.zeros((n_records, 4), np.float32)
        self._ck(self.lib.aq_nrc_get_records(self.handle, x.ctypes.data, y.ctypes.data, n_records))
        return x, y

    def close(self):
        if self.handle:
            self.lib.aq_scene_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def render_multi(scene, cfg, n_gpus, devices=None):
    """aq_render_multi: one process, `n_gpus` devices, spp range split per device, films
    summed with one ncclReduce.  Returns (film[H,W,4], stats dict)."""
    L = _abi.cuda_lib()
    w = cfg.width or scene.desc.camera.res[0]
    h = cfg.height or scene.desc.camera.res[1]
    film = np.zeros((h, w, 4), np.float32)
    st = Stats()
    devs = (C.c_int * n_gpus)(*(devices or range(n_gpus)))
    _abi.check(L.aq_render_multi(C.byref(scene.desc), C.byref(cfg), n_gpus, devs, film.ctypes.data, C.byref(st)))
    return film, st.as_dict()


def build_accel_host(positions, indices):
    """BVH8 built on the host only (no GPU): returns (nodes[n,20] u32, tris[t,12] f32, AccelInfo)."""
    L = _abi.cuda_lib()
    pos = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
    idx = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1, 3)
    pn, pt = C.c_void_p(), C.c_void_p()
    nb, tb = C.c_size_t(), C.c_size_t()
    info = _abi.AccelInfo()
    _abi.check(L.aq_accel_build_host(pos.ctypes.data, pos.shape[0], idx.ctypes.data, idx.shape[0],
                                     C.byref(pn), C.byref(nb), C.byref(pt), C.byref(tb), C.byref(info)))
    nodes = np.frombuffer(C.string_at(pn, nb.value), np.uint32).reshape(-1, 20).copy()
    tris = np.frombuffer(C.string_at(pt, tb.value), np.float32).reshape(-1, 12).copy()
    L.aq_free(pn)
    L.aq_free(pt)
    return nodes, tris, info


def write_png(path, rgba8):
    """RGBA8 image -> PNG through libaqua_host.so (no external zlib)."""
    import os
    a = np.ascontiguousarray(rgba8, dtype=np.uint8)
    _abi.check_host(_abi.host_lib().aq_host_write_png(os.fsencode(path), a.ctypes.data, a.shape[1], a.shape[0]))


def srgb8_reference(film, exposure=1.0):
    """numpy statement of aq_resolve (same threshold table): used by the tests."""
    t = np.zeros(255, np.float32)
    _abi.host_lib().aq_host_srgb_thresholds(t.ctypes.data)
    w = np.where(film[..., 3:4] > 0, np.float32(exposure) / np.where(film[..., 3:4] > 0, film[..., 3:4], 1), 0).astype(np.float32)
    v = np.maximum(film[..., :3] * w, 0).astype(np.float32)
    rgb = np.searchsorted(t, v.ravel(), side="right").reshape(v.shape).astype(np.uint8)
    return np.concatenate([rgb, np.full(rgb.shape[:2] + (1,), 255, np.uint8)], axis=2)


def tonemap(film):
    """float4 film (sum rgb, count) -> uint8 sRGB image."""
    w = np.maximum(film[..., 3:4], 1e-20)
    c = np.clip(film[..., :3] / w, 0.0, 1.0)
    s = np.where(c <= 0.0031308, 12.92 * c, 1.055 * np.power(c, 1 / 2.4) - 0.055)
    return (s * 255.0 + 0.5).astype(np.uint8)
