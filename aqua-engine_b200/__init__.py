"""aqua-engine_b200 — B200 (sm_100a) render hot path for aqua-engine scenes.

Python face of two in-tree native libraries:
  libaqua_cuda.so  hand-written CUDA wavefront path tracer behind the C ABI include/aqua_cuda.h
  libaqua_host.so  scene ingest (JSON / BSON .mesh / JPEG), include/aqua_host.h
The directory name carries a hyphen (it mirrors the reference repo's name), so import it
through the `aqua_engine_b200` shim at the repo root.
"""
from . import _abi
from ._abi import (AQ_MISS, AQ_RENDER_ACCUMULATE, AQ_RENDER_DUMP_SAMPLES, AQ_RENDER_PROFILE, AQ_RENDER_MIS_NEE_ONLY, AQ_RENDER_MIS_BSDF_ONLY, AQ_RENDER_FORCE_FULL_BSDF, AQ_RENDER_NRC_TENSOR, AquaError, IntegratorCfg,
                   Stats)
from .render import HIT_DTYPE, RAY_DTYPE, DeviceScene, Renderer, build_accel_host, render_multi, srgb8_reference, tonemap, write_png
from .scene import (Integrator, Scene, decode_jpeg, import_obj, default_camera, default_material, load_mesh,
                    point_light, scenes_dir)

__all__ = ["Scene", "Integrator", "Renderer", "DeviceScene", "AquaError", "IntegratorCfg", "Stats",
           "RAY_DTYPE", "HIT_DTYPE", "AQ_MISS", "AQ_RENDER_ACCUMULATE", "AQ_RENDER_DUMP_SAMPLES", "AQ_RENDER_PROFILE", "AQ_RENDER_MIS_NEE_ONLY", "AQ_RENDER_MIS_BSDF_ONLY", "AQ_RENDER_FORCE_FULL_BSDF", "AQ_RENDER_NRC_TENSOR",
           "scenes_dir", "load_mesh", "decode_jpeg", "default_material", "default_camera",
           "point_light", "tonemap", "build_accel_host", "render_multi", "import_obj", "write_png", "srgb8_reference"]
