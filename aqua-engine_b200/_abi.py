"""ctypes mirror of include/aqua_cuda.h and include/aqua_host.h.

The native libraries are built in-tree by build.py.  Loading fails loudly: there is no
Python or CPU fallback for the render path.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))

AQ_OK = 0
AQ_MISS = 0xFFFFFFFF
AQ_RENDER_ACCUMULATE = 1
AQ_RENDER_DUMP_SAMPLES = 2
AQ_RENDER_PROFILE = 4
AQ_RENDER_MIS_NEE_ONLY = 8
AQ_RENDER_MIS_BSDF_ONLY = 16
AQ_RENDER_FORCE_FULL_BSDF = 32
AQ_RENDER_NRC_TENSOR = 64

STATUS = {0: "AQ_OK", -1: "AQ_ERR_BAD_ARG", -2: "AQ_ERR_CUDA", -3: "AQ_ERR_OOM",
          -4: "AQ_ERR_UNSUPPORTED", -5: "AQ_ERR_STATE", -6: "AQ_ERR_NCCL", -7: "AQ_ERR_IO"}


class AquaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{STATUS.get(code, code)}: {msg}")
        self.code = code


class Material(C.Structure):
    _fields_ = [("color", C.c_float * 3), ("color_tex", C.c_int32), ("metallic", C.c_float),
                ("roughness", C.c_float), ("specular", C.c_float), ("specular_tint", C.c_float),
                ("sheen", C.c_float), ("sheen_tint", C.c_float), ("clearcoat", C.c_float),
                ("clearcoat_roughness", C.c_float), ("ior", C.c_float), ("transmission", C.c_float),
                ("subsurface", C.c_float), ("anisotropic", C.c_float),
                ("anisotropic_rotation", C.c_float), ("emission", C.c_float * 3),
                ("subsurface_color", C.c_float * 3), ("subsurface_radius", C.c_float * 3),
                ("param_tex", C.c_uint8 * 16)]  # 1-based texture index per AQ_PTEX_* slot, 0 = constant


PTEX = {"metallic": 0, "roughness": 1, "specular": 2, "specular_tint": 3, "sheen": 4, "sheen_tint": 5, "transmission": 6,
        "clearcoat": 7, "clearcoat_roughness": 8, "ior": 9, "subsurface": 10, "subsurface_color": 11}


class Texture(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("rgba8", C.POINTER(C.c_uint8))]


class PointLight(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("intensity", C.c_float * 3)]


class Camera(C.Structure):
    _fields_ = [("res", C.c_uint32 * 2), ("fov", C.c_float), ("lens_radius", C.c_float),
                ("focal", C.c_float), ("translate", C.c_float * 3), ("rotate", C.c_float * 3),
                ("scale", C.c_float * 3)]


class SceneDesc(C.Structure):
    _fields_ = [("n_verts", C.c_uint32), ("n_tris", C.c_uint32),
                ("positions", C.POINTER(C.c_float)), ("normals", C.POINTER(C.c_float)),
                ("uvs", C.POINTER(C.c_float)), ("indices", C.POINTER(C.c_uint32)),
                ("tri_material", C.POINTER(C.c_uint32)),
                ("n_materials", C.c_uint32), ("materials", C.POINTER(Material)),
                ("n_textures", C.c_uint32), ("textures", C.POINTER(Texture)),
                ("n_lights", C.c_uint32), ("lights", C.POINTER(PointLight)),
                ("camera", Camera)]


class IntegratorCfg(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("spp_begin", C.c_uint32),
                ("spp_end", C.c_uint32), ("max_depth", C.c_uint32), ("seed", C.c_uint32),
                ("pool_paths", C.c_uint32), ("flags", C.c_uint32)]


class NrcCfg(C.Structure):
    """aq_nrc_cfg: the NRC-only keys of scenes/integrator.json (:4, :6-8)."""
    _fields_ = [("batch_size", C.c_uint32), ("training_iters", C.c_uint32),
                ("learning_rate", C.c_float), ("visualize_cache", C.c_uint32)]


class NrcInfo(C.Structure):
    _fields_ = [("n_weights", C.c_uint32), ("n_records", C.c_uint32), ("n_valid", C.c_uint32),
                ("loss_first", C.c_float), ("loss_last", C.c_float), ("ms_records", C.c_float),
                ("ms_train", C.c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


NRC_N_WEIGHTS = 16640
NRC_IN = 64


class Stats(C.Structure):
    _fields_ = [("samples", C.c_uint64), ("sample_bounces", C.c_uint64),
                ("rays_closest", C.c_uint64), ("rays_shadow", C.c_uint64),
                ("nodes_fetched", C.c_uint64), ("tris_fetched", C.c_uint64),
                ("ms_total", C.c_float), ("ms_raygen", C.c_float), ("ms_trace", C.c_float),
                ("ms_shade", C.c_float), ("ms_shadow", C.c_float), ("ms_film", C.c_float),
                ("n_launches", C.c_uint32), ("n_waves", C.c_uint32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class AccelInfo(C.Structure):
    _fields_ = [("n_nodes", C.c_uint32), ("n_tri_records", C.c_uint32), ("max_depth", C.c_uint32),
                ("sah_cost", C.c_float), ("build_ms", C.c_float), ("builder", C.c_uint32)]


class HostSceneInfo(C.Structure):
    _fields_ = [("n_shapes", C.c_uint32), ("n_meshes_loaded", C.c_uint32),
                ("n_meshes_missing", C.c_uint32), ("n_verts", C.c_uint32), ("n_tris", C.c_uint32),
                ("n_materials", C.c_uint32), ("n_textures", C.c_uint32), ("n_lights", C.c_uint32),
                ("bounds_min", C.c_float * 3), ("bounds_max", C.c_float * 3)]


# every symbol the headers declare (tests check the .so exports all of them)
CUDA_SYMBOLS = ["aq_abi_version", "aq_init", "aq_destroy", "aq_last_error", "aq_set_stream",
                "aq_device_info", "aq_scene_create", "aq_scene_destroy", "aq_accel_build", "aq_accel_wait",
                "aq_accel_download", "aq_accel_build_host", "aq_free", "aq_intersect", "aq_intersect_device_async", "aq_trace_counters", "aq_render",
                "aq_render_device_async", "aq_render_finish", "aq_render_samples",
                "aq_generate_camera_rays", "aq_render_multi", "aq_resolve", "aq_nrc_train", "aq_nrc_render", "aq_nrc_render_device_async",
                "aq_nrc_get_weights", "aq_nrc_set_weights", "aq_nrc_get_loss", "aq_nrc_get_records"]
HOST_SYMBOLS = ["aq_host_scene_load", "aq_host_scene_free", "aq_host_scene_desc",
                "aq_host_scene_get_info", "aq_host_material_name", "aq_host_shape_range",
                "aq_host_integrator_load", "aq_host_integrator_load_nrc", "aq_host_mesh_load", "aq_host_jpeg_decode",
                "aq_host_free", "aq_host_write_ppm", "aq_host_write_png", "aq_host_write_pfm", "aq_host_srgb_thresholds", "aq_host_import_obj", "aq_host_import_last_error", "aq_host_srgb_to_linear",
                "aq_host_last_error"]

_cuda = None
_host = None


def _load(name):
    path = os.path.join(HERE, name)
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build it with `python __graft_entry__.py build` "
            f"(aqua-engine_b200/build.py). There is no fallback implementation.")
    return C.CDLL(path)


def cuda_lib():
    """libaqua_cuda.so (the product)."""
    global _cuda
    if _cuda is None:
        L = _load(os.environ.get("AQUA_CUDA_LIB", "libaqua_cuda.so"))  # env: A/B kernel variants
        vp, u32, i = C.c_void_p, C.c_uint32, C.c_int
        L.aq_abi_version.restype = i
        L.aq_last_error.restype = C.c_char_p
        L.aq_last_error.argtypes = [vp]
        L.aq_init.argtypes = [i, C.POINTER(vp)]
        L.aq_destroy.argtypes = [vp]
        L.aq_destroy.restype = None
        L.aq_set_stream.argtypes = [vp, vp]
        L.aq_device_info.argtypes = [vp, C.POINTER(i), C.POINTER(i), C.POINTER(i), C.POINTER(C.c_size_t)]
        L.aq_scene_create.argtypes = [vp, C.POINTER(SceneDesc), C.POINTER(vp)]
        L.aq_scene_destroy.argtypes = [vp]
        L.aq_scene_destroy.restype = None
        L.aq_accel_build.argtypes = [vp, C.POINTER(AccelInfo)]
        L.aq_accel_wait.argtypes = [vp, C.POINTER(AccelInfo)]
        L.aq_accel_download.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t]
        L.aq_accel_build_host.argtypes = [vp, u32, vp, u32, C.POINTER(vp), C.POINTER(C.c_size_t),
                                          C.POINTER(vp), C.POINTER(C.c_size_t), C.POINTER(AccelInfo)]
        L.aq_free.argtypes = [vp]
        L.aq_free.restype = None
        L.aq_intersect.argtypes = [vp, vp, u32, vp, i]
        L.aq_intersect_device_async.argtypes = [vp, vp, u32, vp, i]
        L.aq_trace_counters.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), i]
        L.aq_render.argtypes = [vp, C.POINTER(IntegratorCfg), vp, C.POINTER(Stats)]
        L.aq_render_device_async.argtypes = [vp, C.POINTER(IntegratorCfg), vp]
        L.aq_render_finish.argtypes = [vp, C.POINTER(Stats)]
        L.aq_render_samples.argtypes = [vp, vp, C.c_size_t]
        L.aq_generate_camera_rays.argtypes = [vp, C.POINTER(IntegratorCfg), u32, vp]
        L.aq_resolve.argtypes = [vp, vp, vp, u32, u32, C.c_float, vp]
        L.aq_nrc_train.argtypes = [vp, C.POINTER(IntegratorCfg), C.POINTER(NrcCfg), C.POINTER(NrcInfo)]
        L.aq_nrc_render.argtypes = [vp, C.POINTER(IntegratorCfg), C.POINTER(NrcCfg), vp, C.POINTER(Stats)]
        L.aq_nrc_render_device_async.argtypes = [vp, C.POINTER(IntegratorCfg), C.POINTER(NrcCfg), vp]
        L.aq_nrc_get_weights.argtypes = [vp, vp, C.c_size_t]
        L.aq_nrc_set_weights.argtypes = [vp, vp, C.c_size_t]
        L.aq_nrc_get_loss.argtypes = [vp, vp, C.c_size_t]
        L.aq_nrc_get_records.argtypes = [vp, vp, vp, C.c_size_t]
        L.aq_render_multi.argtypes = [C.POINTER(SceneDesc), C.POINTER(IntegratorCfg), i,
                                      C.POINTER(i), vp, C.POINTER(Stats)]
        _cuda = L
    return _cuda


def host_lib():
    """libaqua_host.so (scene ingest, no GPU)."""
    global _host
    if _host is None:
        L = _load("libaqua_host.so")
        vp, u32 = C.c_void_p, C.c_uint32
        L.aq_host_last_error.restype = C.c_char_p
        L.aq_host_scene_load.argtypes = [C.c_char_p, C.POINTER(vp)]
        L.aq_host_scene_free.argtypes = [vp]
        L.aq_host_scene_free.restype = None
        L.aq_host_scene_desc.argtypes = [vp]
        L.aq_host_scene_desc.restype = C.POINTER(SceneDesc)
        L.aq_host_scene_get_info.argtypes = [vp, C.POINTER(HostSceneInfo)]
        L.aq_host_material_name.argtypes = [vp, u32]
        L.aq_host_material_name.restype = C.c_char_p
        L.aq_host_shape_range.argtypes = [vp, u32, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]
        L.aq_host_integrator_load.argtypes = [C.c_char_p, C.POINTER(IntegratorCfg), C.c_char_p, C.c_size_t]
        L.aq_host_integrator_load_nrc.argtypes = [C.c_char_p, C.POINTER(NrcCfg)]
        L.aq_host_mesh_load.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t, C.POINTER(u32),
                                        C.POINTER(u32), C.POINTER(C.POINTER(C.c_float)),
                                        C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.POINTER(C.c_float)),
                                        C.POINTER(u32), C.POINTER(C.POINTER(u32))]
        L.aq_host_jpeg_decode.argtypes = [C.c_char_p, C.POINTER(u32), C.POINTER(u32),
                                          C.POINTER(C.POINTER(C.c_uint8))]
        L.aq_host_free.argtypes = [vp]
        L.aq_host_free.restype = None
        L.aq_host_write_ppm.argtypes = [C.c_char_p, vp, u32, u32]
        L.aq_host_write_png.argtypes = [C.c_char_p, vp, u32, u32]
        L.aq_host_write_pfm.argtypes = [C.c_char_p, vp, u32, u32]
        L.aq_host_srgb_thresholds.argtypes = [vp]
        L.aq_host_srgb_thresholds.restype = None
        L.aq_host_import_obj.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p]
        L.aq_host_import_last_error.restype = C.c_char_p
        L.aq_host_srgb_to_linear.argtypes = [C.c_float]
        L.aq_host_srgb_to_linear.restype = C.c_float
        _host = L
    return _host


def check(rc, ctx=None):
    if rc != AQ_OK:
        msg = cuda_lib().aq_last_error(ctx)
        raise AquaError(rc, msg.decode() if msg else "")


def check_host(rc):
    if rc != AQ_OK:
        msg = host_lib().aq_host_last_error()
        raise AquaError(rc, msg.decode() if msg else "")
