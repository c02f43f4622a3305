"""CPU replay of the device builder's tree stage (row a5): the per-element functions of
csrc/aq_bvh_ploc.h — PLOC nearest / merge / first-assignment and the cost-optimal collapse tables —
are the code the kernels of aq_bvh_build_gpu.cu wrap; tools/experimental/ploc_emulate.cpp runs them
in loops the way the kernels are launched and checks the tree: n-1 internal nodes, every leaf
reachable once, boxes enclose children, counts add up, the leaves of every subtree contiguous in
the new primitive order, and the DP emit walk covers every triangle exactly once with leaf groups
of <= 3 triangles.  (The GPU side of the same code is covered by the device-builder tests of
tests/test_gpu_parity.py.)"""
import os
import subprocess

import numpy as np
import pytest

from conftest import triangle_soup

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emulate(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("ploc") / "ploc_emulate")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-w", "-I", os.path.join(ROOT, "aqua-engine_b200", "csrc"),
                           "-I", os.path.join(ROOT, "include"), "-o", exe, os.path.join(ROOT, "tools", "experimental", "ploc_emulate.cpp")])
    return exe


def _run(exe, tmp_path, pos, idx, radius):
    p, i = str(tmp_path / "pos.bin"), str(tmp_path / "idx.bin")
    np.ascontiguousarray(pos, np.float32).tofile(p)
    np.ascontiguousarray(idx, np.uint32).tofile(i)
    r = subprocess.run([exe, p, i, str(radius)], capture_output=True, text=True, timeout=600)
    return r.returncode, r.stdout


@pytest.mark.parametrize("radius", [1, 4, 16])
def test_ploc_replay_on_cbox(aq, cbox, emulate, tmp_path, radius):
    pos, idx, *_ = cbox.arrays()
    rc, out = _run(emulate, tmp_path, pos, idx, radius)
    assert rc == 0 and "tree ok" in out and "emit ok" in out, out
    assert f"internal={len(idx) - 1}" in out and f"{len(idx)} triangles" in out


def test_ploc_replay_on_room_and_soup(aq, room, emulate, tmp_path):
    pos, idx, *_ = room.arrays()
    rc, out = _run(emulate, tmp_path, pos, idx, 16)
    assert rc == 0 and "tree ok" in out and "emit ok" in out, out
    sah = float(out.split("sah=")[1].split()[0])
    assert sah < 50.0  # the radix tree over the same order costs 55.9, PLOC 43.2 (profiles/r02c_ploc_prototype_cpu.log)
    pos, idx = triangle_soup(50_000)
    rc, out = _run(emulate, tmp_path, pos, idx, 8)
    assert rc == 0 and "tree ok" in out and "emit ok" in out and "50000 triangles" in out, out


def test_ploc_replay_on_degenerate_input(emulate, tmp_path):
    # coincident triangles (identical boxes: every merge is a tie) and a single pair
    pos = np.tile(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32), (64, 1))
    idx = np.arange(192, dtype=np.uint32).reshape(-1, 3)
    rc, out = _run(emulate, tmp_path, pos, idx, 16)
    assert rc == 0 and "emit ok" in out and "64 triangles" in out, out
    rc, out = _run(emulate, tmp_path, pos[:6], idx[:2], 16)
    assert rc == 0 and "emit ok" in out and "2 triangles" in out, out


def test_warp_cost_model_policies_walk_the_same_tree(aq, ao, cbox, tmp_path):
    """tools/experimental/warp_sim.cpp replays the traversal kernel's per-lane-refill scheduling on the CPU with
    the product's traversal template; the balanced triangle phase the kernels ship (policy 5: a lane with
    triangles left does not open a node) must give every ray the same walk as the plain step (policy 0): same
    hit checksum, same node visits and triangle tests per ray — and cost fewer warp instructions."""
    exe = str(tmp_path / "warp_sim")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-w", "-I", os.path.join(ROOT, "aqua-engine_b200", "csrc"),
                           "-I", os.path.join(ROOT, "include"), "-o", exe, os.path.join(ROOT, "tools", "experimental", "warp_sim.cpp")])
    pos, idx, *_ = cbox.arrays()
    nodes, tris, info = aq.build_accel_host(pos, idx)
    o = ao.OracleScene(cbox, build_bvh=True)
    cam = o.camera_rays(aq.Integrator(spp=1).cfg(width=96, height=96), 0)
    h = o.intersect(cam, mode=1)
    ok = h["prim"] != aq.AQ_MISS
    g = np.random.default_rng(3)
    P = cam["o"][ok] + h["t"][ok, None] * cam["d"][ok]
    d = g.normal(size=P.shape)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.zeros(len(P), aq.RAY_DTYPE)
    rays["o"], rays["d"], rays["tmin"], rays["tmax"] = P + 1e-3 * d, d, 0.0, 3e38
    f = {k: str(tmp_path / (k + ".bin")) for k in ("nodes", "tris", "rays")}
    np.ascontiguousarray(nodes).tofile(f["nodes"])
    np.ascontiguousarray(tris).tofile(f["tris"])
    rays.tofile(f["rays"])

    def run(policy, k, any_hit):
        out = subprocess.run([exe, f["nodes"], f["tris"], f["rays"], str(policy), str(k), str(any_hit)], capture_output=True, text=True, timeout=300).stdout
        num = lambda key: float(out.split(key + "=")[1].split()[0])
        return out.split("checksum=")[1].split()[0], num("nodes/ray"), num("tris/ray"), num("warp-instr/ray")

    for any_hit in (0, 1):
        base, bal = run(0, 0, any_hit), run(5, 6, any_hit)
        assert base[:3] == bal[:3], (base, bal)
        assert bal[3] < base[3]
