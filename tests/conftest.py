import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the native libraries once (no-op when up to date; the GPU box uses the prebuilt .so)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("aq_build", os.path.join(ROOT, "aqua-engine_b200", "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    import shutil
    if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
        b.build_all()
        import subprocess
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    yield


@pytest.fixture(scope="session")
def aq():
    import aqua_engine_b200
    return aqua_engine_b200


@pytest.fixture(scope="session")
def ao():
    import aq_oracle
    return aq_oracle


@pytest.fixture(scope="session")
def scenes(aq):
    return aq.scenes_dir()


@pytest.fixture(scope="session")
def cbox(aq, scenes):
    return aq.Scene.load(os.path.join(scenes, "cbox.json"))


@pytest.fixture(scope="session")
def room(aq, scenes):
    return aq.Scene.load(os.path.join(scenes, "room.json"))


@pytest.fixture(scope="session")
def renderer(aq):
    return aq.Renderer(0)


def random_rays(aq, n, lo, hi, seed=1, tmax=3.0e38):
    r = np.random.default_rng(seed)
    rays = np.zeros(n, aq.RAY_DTYPE)
    rays["o"] = r.uniform(lo, hi, (n, 3)).astype(np.float32)
    d = r.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays["d"] = d.astype(np.float32)
    rays["tmin"] = 0.0
    rays["tmax"] = tmax
    return rays


def triangle_soup(n, seed=12345, r=0.005):
    """BASELINE config C4: centres U[0,1]^3, vertices = centre + U[-r,r]^3."""
    g = np.random.default_rng(seed)
    c = g.uniform(0, 1, (n, 1, 3))
    v = (c + g.uniform(-r, r, (n, 3, 3))).astype(np.float32)
    return v.reshape(-1, 3), np.arange(3 * n, dtype=np.uint32).reshape(-1, 3)


def hits_equal(a, b):
    return (np.array_equal(a["prim"], b["prim"]) and np.array_equal(a["t"].view(np.uint32), b["t"].view(np.uint32))
            and np.array_equal(a["u"].view(np.uint32), b["u"].view(np.uint32))
            and np.array_equal(a["v"].view(np.uint32), b["v"].view(np.uint32)))
