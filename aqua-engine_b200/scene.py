"""Host-side scene objects: the Python face of libaqua_host.so, mirroring the reference's
serde types (Scene / Camera / Light / Shape / Bsdf / Texture, scenes/cbox.json:1-627) and
its integrator config (scenes/integrator.json:1-8)."""
import ctypes as C
import os

import numpy as np

from . import _abi

REPO = os.path.dirname(_abi.HERE)


def scenes_dir():
    """Directory holding the reference's scene assets.

    $AQUA_SCENES, else the git-ignored copy that travels to the GPU box
    (baseline/_ref/scenes, made by tools/prepare_assets.py), else /root/reference/scenes."""
    cands = [os.environ.get("AQUA_SCENES"), os.path.join(REPO, "baseline", "_ref", "scenes"),
             "/root/reference/scenes"]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "cbox.json")):
            return c
    raise FileNotFoundError("scene assets not found; run tools/prepare_assets.py or set AQUA_SCENES")


class Scene:
    """A flattened scene (aq_scene_desc) that lives in host memory."""

    def __init__(self):
        self._handle = None   # aq_host_scene* when loaded from JSON
        self._keep = []       # numpy arrays backing a desc made from arrays
        self.desc = None
        self.info = None
        self.path = None

    # ---- from the reference's JSON (Scene, scenes/cbox.json:1)
    @classmethod
    def load(cls, json_path):
        L = _abi.host_lib()
        h = C.c_void_p()
        _abi.check_host(L.aq_host_scene_load(os.fsencode(json_path), C.byref(h)))
        s = cls()
        s._handle = h
        s.desc = L.aq_host_scene_desc(h).contents
        info = _abi.HostSceneInfo()
        _abi.check_host(L.aq_host_scene_get_info(h, C.byref(info)))
        s.info = info
        s.path = json_path
        return s

    # ---- from flat arrays (synthetic scenes, config C4)
    @classmethod
    def from_arrays(cls, positions, indices, normals=None, uvs=None, tri_material=None,
                    materials=None, lights=None, camera=None, textures=None):
        """`textures`: list of uint8 arrays [H, W, 4] (8-bit sRGB RGBA, row 0 = top row), referenced by
        `Material.color_tex`."""
        s = cls()
        pos = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
        idx = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1, 3)
        d = _abi.SceneDesc()
        d.n_verts, d.n_tris = pos.shape[0], idx.shape[0]
        d.positions = pos.ctypes.data_as(C.POINTER(C.c_float))
        d.indices = idx.ctypes.data_as(C.POINTER(C.c_uint32))
        s._keep += [pos, idx]
        if normals is not None:
            nrm = np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 3)
            d.normals = nrm.ctypes.data_as(C.POINTER(C.c_float))
            s._keep.append(nrm)
        if uvs is not None:
            uv = np.ascontiguousarray(uvs, dtype=np.float32).reshape(-1, 2)
            d.uvs = uv.ctypes.data_as(C.POINTER(C.c_float))
            s._keep.append(uv)
        mats = materials or [default_material()]
        marr = (_abi.Material * len(mats))(*mats)
        d.n_materials, d.materials = len(mats), marr
        s._keep.append(marr)
        tm = np.zeros(idx.shape[0], np.uint32) if tri_material is None else \
            np.ascontiguousarray(tri_material, dtype=np.uint32)
        d.tri_material = tm.ctypes.data_as(C.POINTER(C.c_uint32))
        s._keep.append(tm)
        lights = lights or []
        if lights:
            larr = (_abi.PointLight * len(lights))(*lights)
            d.n_lights, d.lights = len(lights), larr
            s._keep.append(larr)
        if textures:
            tarr = (_abi.Texture * len(textures))()
            for k, t in enumerate(textures):
                t = np.ascontiguousarray(t, dtype=np.uint8)
                assert t.ndim == 3 and t.shape[2] == 4
                tarr[k].height, tarr[k].width = t.shape[0], t.shape[1]
                tarr[k].rgba8 = t.ctypes.data_as(C.POINTER(C.c_uint8))
                s._keep.append(t)
            d.n_textures, d.textures = len(textures), tarr
            s._keep.append(tarr)
        d.camera = camera or default_camera()
        s.desc = d
        return s

    def arrays(self):
        """numpy views of the flat geometry."""
        d = self.desc
        pos = np.ctypeslib.as_array(d.positions, shape=(d.n_verts, 3))
        idx = np.ctypeslib.as_array(d.indices, shape=(d.n_tris, 3))
        nrm = np.ctypeslib.as_array(d.normals, shape=(d.n_verts, 3)) if d.normals else None
        uv = np.ctypeslib.as_array(d.uvs, shape=(d.n_verts, 2)) if d.uvs else None
        tm = np.ctypeslib.as_array(d.tri_material, shape=(d.n_tris,))
        return pos, idx, nrm, uv, tm

    def material_names(self):
        if self._handle is None:
            return []
        L = _abi.host_lib()
        return [L.aq_host_material_name(self._handle, i).decode() for i in range(self.desc.n_materials)]

    def shape_range(self, i):
        a, b, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
        _abi.check_host(_abi.host_lib().aq_host_shape_range(self._handle, i, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def close(self):
        if self._handle is not None:
            _abi.host_lib().aq_host_scene_free(self._handle)
            self._handle = None
            self.desc = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def default_material(color=(0.5, 0.5, 0.5), metallic=0.0, roughness=0.5):
    m = _abi.Material()
    m.color[:] = color
    m.color_tex = -1
    m.metallic, m.roughness = metallic, roughness
    m.sheen_tint, m.clearcoat_roughness, m.ior = 0.5, 0.03, 1.45
    m.subsurface_radius[:] = (1.0, 0.2, 0.1)
    return m


def default_camera(res=(256, 256), fov=45.0, translate=(0, 0, 3)):
    c = _abi.Camera()
    c.res[:] = res
    c.fov, c.lens_radius, c.focal = fov, 0.0, 1.0
    c.translate[:] = translate
    c.rotate[:] = (0, 0, 0)
    c.scale[:] = (1, 1, 1)
    return c


def point_light(pos, intensity):
    l = _abi.PointLight()
    l.pos[:] = pos
    l.intensity[:] = intensity
    return l


class Integrator:
    """Integrator config (scenes/integrator.json:1-8): `type`, `spp`, `max_depth`.

    `type: "nrc"` (the only type the reference names) selects the neural radiance cache:
    `DeviceScene.nrc_train` / `nrc_render` with `nrc_cfg()`; `cfg()` alone gives the path tracer
    with the same `spp` / `max_depth`."""

    def __init__(self, spp=16, max_depth=5, seed=0, type="pt", batch_size=512, training_iters=2048,
                 learning_rate=1e-3, visualize_cache=False):
        self.spp, self.max_depth, self.seed, self.type = spp, max_depth, seed, type
        # the NRC-only keys (scenes/integrator.json:4,6-8); used when type == "nrc"
        self.batch_size, self.training_iters = batch_size, training_iters
        self.learning_rate, self.visualize_cache = learning_rate, visualize_cache

    def nrc_cfg(self):
        n = _abi.NrcCfg()
        n.batch_size, n.training_iters = self.batch_size, self.training_iters
        n.learning_rate, n.visualize_cache = self.learning_rate, int(self.visualize_cache)
        return n

    @classmethod
    def load(cls, json_path):
        cfg = _abi.IntegratorCfg()
        buf = C.create_string_buffer(32)
        _abi.check_host(_abi.host_lib().aq_host_integrator_load(os.fsencode(json_path), C.byref(cfg), buf, 32))
        n = _abi.NrcCfg()
        _abi.check_host(_abi.host_lib().aq_host_integrator_load_nrc(os.fsencode(json_path), C.byref(n)))
        return cls(cfg.spp_end, cfg.max_depth, cfg.seed, buf.value.decode(), n.batch_size, n.training_iters,
                   n.learning_rate, bool(n.visualize_cache))

    def cfg(self, width=0, height=0, spp_begin=0, spp_end=None, pool_paths=0, flags=0):
        c = _abi.IntegratorCfg()
        c.width, c.height = width, height
        c.spp_begin = spp_begin
        c.spp_end = self.spp if spp_end is None else spp_end
        c.max_depth, c.seed, c.pool_paths, c.flags = self.max_depth, self.seed, pool_paths, flags
        return c


def load_mesh(path):
    """One BSON TriangleMesh (scenes/*.mesh) -> dict of numpy arrays."""
    L = _abi.host_lib()
    name = C.create_string_buffer(256)
    nv, nt, nuv = C.c_uint32(), C.c_uint32(), C.c_uint32()
    pp, pn, pu = (C.POINTER(C.c_float)() for _ in range(3))
    pi = C.POINTER(C.c_uint32)()
    _abi.check_host(L.aq_host_mesh_load(os.fsencode(path), name, 256, C.byref(nv), C.byref(nt),
                                        C.byref(pp), C.byref(pn), C.byref(pu), C.byref(nuv), C.byref(pi)))

    def take(p, shape, dt):
        n = int(np.prod(shape))
        a = np.ctypeslib.as_array(p, shape=(n,)).astype(dt).reshape(shape) if n else np.zeros(shape, dt)
        L.aq_host_free(p)
        return a
    n_nrm = nv.value  # aq_host_mesh_load returns n_verts normals (zero vectors when the file has none)
    out = {"name": name.value.decode(),
           "vertices": take(pp, (nv.value, 3), np.float32),
           "normals": take(pn, (n_nrm, 3), np.float32),
           "texcoords": take(pu, (nuv.value, 2), np.float32),
           "indices": take(pi, (nt.value, 3), np.uint32)}
    return out


def decode_jpeg(path):
    L = _abi.host_lib()
    w, h = C.c_uint32(), C.c_uint32()
    p = C.POINTER(C.c_uint8)()
    _abi.check_host(L.aq_host_jpeg_decode(os.fsencode(path), C.byref(w), C.byref(h), C.byref(p)))
    a = np.ctypeslib.as_array(p, shape=(h.value, w.value, 4)).copy()
    L.aq_host_free(p)
    return a


def import_obj(obj_path, out_dir, scene_name):
    """Wavefront OBJ/MTL -> <out_dir>/<scene_name>.json + BSON .mesh files (the reference's asset
    format).  Returns the path of the scene JSON."""
    L = _abi.host_lib()
    os.makedirs(out_dir, exist_ok=True)
    rc = L.aq_host_import_obj(os.fsencode(obj_path), os.fsencode(out_dir), scene_name.encode())
    if rc < 0:
        raise _abi.AquaError(rc, L.aq_host_import_last_error().decode())
    return os.path.join(out_dir, scene_name + ".json")
