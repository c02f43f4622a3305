"""The drop-in boundary: both shared libraries load, export every symbol the headers
declare, and the ctypes mirror has the same struct layout as the C headers."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(aq_[a-z0-9_]+)\s*\(", src)))


def test_cuda_lib_exports_every_declared_symbol(aq):
    L = aq._abi.cuda_lib()
    names = declared_functions("aqua_cuda.h")
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), f"libaqua_cuda.so does not export {n}"
    assert sorted(aq._abi.CUDA_SYMBOLS) == names
    assert L.aq_abi_version() == 1


def test_host_lib_exports_every_declared_symbol(aq):
    L = aq._abi.host_lib()
    names = declared_functions("aqua_host.h")
    for n in names:
        assert hasattr(L, n), f"libaqua_host.so does not export {n}"
    assert sorted(aq._abi.HOST_SYMBOLS) == names


def test_struct_layout_matches_header(aq):
    prog = r"""
    #include <stdio.h>
    #include "aqua_host.h"
    int main(void){
      printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(aq_material), sizeof(aq_texture),
        sizeof(aq_point_light), sizeof(aq_camera), sizeof(aq_scene_desc), sizeof(aq_integrator_cfg),
        sizeof(aq_ray), sizeof(aq_hit), sizeof(aq_stats), sizeof(aq_accel_info), sizeof(aq_host_scene_info),
        sizeof(aq_nrc_cfg), sizeof(aq_nrc_info));
      return 0; }"""
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "s.c")
        open(src, "w").write(prog)
        exe = os.path.join(td, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    A = aq._abi
    mirror = [A.Material, A.Texture, A.PointLight, A.Camera, A.SceneDesc, A.IntegratorCfg, None, None,
              A.Stats, A.AccelInfo, A.HostSceneInfo, A.NrcCfg, A.NrcInfo]
    for s, m in zip(sizes, mirror):
        if m is not None:
            assert C.sizeof(m) == s, (m, C.sizeof(m), s)
    assert sizes[6] == 32 and aq.RAY_DTYPE.itemsize == 32
    assert sizes[7] == 16 and aq.HIT_DTYPE.itemsize == 16


def test_no_cpu_fallback_without_gpu(aq):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(aq.AquaError) as e:
        aq.Renderer(0)
    assert e.value.code == -2 and "no CPU path" in str(e.value)


def test_product_never_references_the_oracle():
    """oracle/ is test infrastructure: nothing in the product tree may import or link it."""
    pkg = os.path.join(ROOT, "aqua-engine_b200")
    for root, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".inl")):
                txt = open(os.path.join(root, f), errors="ignore").read()
                assert "aq_oracle" not in txt and "libaqua_oracle" not in txt and "aqo_" not in txt, f
    out = subprocess.check_output(["ldd", os.path.join(pkg, "libaqua_cuda.so")]).decode()
    assert "oracle" not in out
