#!/usr/bin/env python
"""bench.py — the driver's measurement contract for the render hot path.

A *step* is one full render of BASELINE.json's headline configuration C2:
scenes/cbox.json at 1024x1024, 1024 spp, max depth 5, NEE on the point light
(~1.07e9 camera paths, ~2.6e9 shaded path vertices, ~5.4e9 rays per step and GPU).

  python bench.py --gpus N --steps K --warmup W           (N>1: launched under torchrun)
  python bench.py --impl reference ...                    (CPU oracle on the host cores)

`value` times the device-resident path (scene + BVH already in HBM, film stays in HBM);
`e2e` times the C-ABI call a user makes with HOST buffers: aq_scene_create (H2D of the scene
arrays) + aq_accel_build + aq_render (film D2H into pinned memory) every step.
With N GPUs every rank renders its own 1024-spp sample range of the same image (weak
scaling: per-GPU work fixed) and the float4 films are summed with one NCCL reduce.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "C2: scenes/cbox.json 1024x1024, 1024 spp, max_depth 5, point-light NEE (MIS weight 1)"


def workload_name(args):
    """The default arguments are BASELINE.json's config C2; --scene/--width/--height/--spp select
    the other configs (C1: cbox 256x256x16, C3: room 1920x1080x256) for the tables in profiles/."""
    if (args.scene, args.width, args.height, args.spp) == ("cbox", 1024, 1024, 1024):
        return WORKLOAD
    return f"scenes/{args.scene}.json {args.width}x{args.height}, {args.spp} spp, max_depth 5, point-light NEE"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "200"], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons, smax = [], set(), None
        try:
            for line in open(self.path):
                t = [x.strip() for x in line.split(",")]
                if len(t) < 9:
                    continue
                try:
                    sm.append(float(t[1]))
                    smax = float(t[2])
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), t[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=smax, reasons=sorted(reasons), samples=len(sm))
        return out


def stage_bytes(st, spw=8):
    """Algorithmic bytes each stage must move per step, from the step's
    own device counters.  Scene/BVH bytes are not counted (cbox: 3 KB, cache resident)."""
    # spw = samples of one pixel held by one wave (pool / tile pixels)
    rc, rs, sb, smp = st["rays_closest"], st["rays_shadow"], st["sample_bounces"], st["samples"]
    n_next = rc - smp  # continuation rays written by shade
    return {
        "raygen": smp * (48 + 16),                       # ray(32)+beta/slot(16) out, L init(16)
        "closest": rc * (32 + 16),                       # ray in, hit out
        "shade": rc * 16 + sb * 32 + n_next * 48 + rs * 48,  # hit in; dir+beta in (hits only); queues out
        "shadow": rs * 32 + rs * (16 + 32),              # ray in; contribution in + L RMW (upper bound: all visible)
        "film": smp * 16 + smp * 32 // max(1, spw),      # L in; film RMW once per pixel per wave
    }


def run_reference(args):
    """The reference has no CPU implementation (src/lib.rs is empty): the timed CPU arm is
    this repo's oracle port on all host cores, on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import aq_oracle as ao
    import aqua_engine_b200 as aq
    scene = aq.Scene.load(os.path.join(aq.scenes_dir(), args.scene + ".json"))
    o = ao.OracleScene(scene, build_bvh=True)
    cores = ao.threads()
    spp = args.cpu_spp
    cfg = aq.Integrator(spp=spp, max_depth=5, seed=0).cfg(width=args.width, height=args.height)
    for _ in range(args.warmup):
        o.render(aq.Integrator(spp=1, max_depth=5, seed=0).cfg(width=args.width, height=args.height), mode=1)
    t0 = time.perf_counter()
    tot = {"rays": 0, "samples": 0, "sb": 0}
    for _ in range(args.steps):
        _, _, st = o.render(cfg, mode=1)
        tot["rays"] += st["rays_closest"] + st["rays_shadow"]
        tot["samples"] += st["samples"]
        tot["sb"] += st["sample_bounces"]
    dt = time.perf_counter() - t0
    v = tot["rays"] / dt / 1e6
    sample = f"{spp} of {args.spp} spp of the same {args.width}x{args.height} image per step (oracle BVH2, {cores} threads)"
    line = {"impl": "reference", "metric": "Mrays/s", "value": v, "unit": "Mrays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": f"reference scene assets (scenes/{args.scene}.json), no synthetic substitutes",
            "config": {"workload": workload_name(args), "sample": sample},
            "samples_per_s": tot["samples"] / dt, "sample_bounces_per_s": tot["sb"] / dt,
            "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scene", default="cbox")
    ap.add_argument("--width", type=int, default=1024)
    ap.add_argument("--height", type=int, default=1024)
    ap.add_argument("--spp", type=int, default=1024)
    ap.add_argument("--pool", type=int, default=0)
    ap.add_argument("--cpu-spp", type=int, default=16, help="spp per step of the CPU arms (16 spp of 1024^2 = ~20 CPU-seconds)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every rank renders --spp samples per pixel; strong: --spp is split across the ranks (config C5)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import aqua_engine_b200 as aq
    from aqua_engine_b200 import dist as aqd

    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: the render path has no CPU fallback"}))
        return 1
    rank, world, local = aqd.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    scene = aq.Scene.load(os.path.join(aq.scenes_dir(), args.scene + ".json"))
    integ = aq.Integrator(spp=args.spp, max_depth=5, seed=0)
    W, H = args.width, args.height
    r = aq.Renderer(local)
    r.set_stream(torch.cuda.current_stream().cuda_stream)
    ds = r.upload(scene)
    film = torch.zeros(H, W, 4, device=dev)
    # weak scaling: rank k renders samples [k*spp, (k+1)*spp) of the same image
    if args.scaling == "strong":
        sb, se = aqd.partition_spp(0, args.spp, rank, world)
    else:  # weak scaling: rank k renders samples [k*spp, (k+1)*spp) of the same image
        sb, se = rank * args.spp, (rank + 1) * args.spp
    cfg = integ.cfg(width=W, height=H, spp_begin=sb, spp_end=se, pool_paths=args.pool, flags=aq.AQ_RENDER_PROFILE)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        ds.render_device_async(cfg, film.data_ptr())
        aqd.reduce_film(film, 0)
        return ds.finish()

    # ---------------- value: device-resident
    for _ in range(args.warmup):
        step()
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stats = []
    e0.record()
    for _ in range(args.steps):
        stats.append(step())
    e1.record()
    barrier()
    clk = clocks.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    tot = torch.tensor([sum(s["rays_closest"] + s["rays_shadow"] for s in stats), sum(s["samples"] for s in stats),
                        sum(s["sample_bounces"] for s in stats), sum(s["n_launches"] for s in stats)],
                       device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_total = float(ms.item())
    rays, samples, bounces, launches = (float(x) for x in tot.tolist())
    secs = ms_total * 1e-3

    # ---------------- e2e: host buffers through the C ABI, every step
    pinned = torch.zeros(H, W, 4, dtype=torch.float32).pin_memory()
    d = scene.desc
    h2d = (d.n_verts * 3 * 4 * (2 if d.normals else 1) + (d.n_verts * 2 * 4 if d.uvs else 0) + d.n_tris * 16
           + d.n_materials * 64 + 256 * 4 + d.n_lights * 24
           + sum(d.textures[i].width * d.textures[i].height * 4 for i in range(d.n_textures)))
    cfg_e = integ.cfg(width=W, height=H, spp_begin=sb, spp_end=se, pool_paths=args.pool)

    dbg = os.environ.get("AQ_BENCH_DEBUG")

    def e2e_step():
        t_a = time.perf_counter()
        ds2 = r.upload(scene)  # aq_scene_create (H2D) + aq_accel_build
        t_b = time.perf_counter()
        if world == 1:
            _, st = ds2.render(cfg_e, film=pinned.numpy())  # aq_render: render + film D2H
        else:
            ds2.render_device_async(cfg_e, film.data_ptr())
            aqd.reduce_film(film, 0)
            st = ds2.finish()
            if rank == 0:
                pinned.copy_(film, non_blocking=False)
        t_c = time.perf_counter()
        ds2.close()
        if dbg:
            print(f"[e2e] upload+build {1e3 * (t_b - t_a):.1f} ms, render+readback {1e3 * (t_c - t_b):.1f} ms "
                  f"(device {st['ms_total']:.1f}, {st['n_waves']} waves, {st['n_launches']} launches), destroy {1e3 * (time.perf_counter() - t_c):.1f} ms", file=sys.stderr)
        return st

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    est = [e2e_step() for _ in range(args.steps)]
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    e2e_rays = torch.tensor([sum(s["rays_closest"] + s["rays_shadow"] for s in est)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_rays, op=dist.ReduceOp.SUM)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---------------- roofline of the dominant kernel (rank 0's own launches)
    peak, peak_src = load_peaks()
    stage_ms = {"raygen": sum(s["ms_raygen"] for s in stats), "closest": sum(s["ms_trace"] for s in stats),
                "shade": sum(s["ms_shade"] for s in stats), "shadow": sum(s["ms_shadow"] for s in stats),
                "film": sum(s["ms_film"] for s in stats)}
    dom = max(stage_ms, key=stage_ms.get)
    pool = args.pool or (1 << 24)
    spw = max(1, pool // min(W * H, pool))  # samples of one pixel held by one wave
    nb = {k: sum(stage_bytes(s, spw)[k] for s in stats) for k in stage_ms}
    n_launch = {"raygen": sum(s["n_waves"] for s in stats), "film": sum(s["n_waves"] for s in stats)}
    for k in ("closest", "shade", "shadow"):
        n_launch[k] = sum(s["n_waves"] for s in stats) * integ.max_depth
    achieved = nb[dom] / (stage_ms[dom] * 1e-3) / 1e9 if stage_ms[dom] > 0 else 0.0
    traffic, ncu = None, {}
    prof = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(prof):  # one `ncu --set full` capture of the same kernels (tools/run_captures.sh)
        try:
            ncu = json.load(open(prof)).get(dom, {})
            traffic = ncu.get("dram_bytes_per_launch")
        except Exception:
            traffic, ncu = None, {}
    roofline = {"bound": "hbm", "kernel": {"closest": "aq_k_trace<3> (closest hit)", "shadow": "aq_k_trace<1> (any hit)", "shade": "aq_k_shade",
                                           "raygen": "aq_k_raygen", "film": "aq_k_film"}[dom],
                "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "ncu_issue_active_pct": ncu.get("issue_active_pct"), "ncu_threads_per_inst": ncu.get("threads_per_inst"),
                "algorithmic_bytes_per_launch": nb[dom] / max(1, n_launch[dom]),
                "avg_launch_ms": stage_ms[dom] / max(1, n_launch[dom]),
                "stage_ms_per_step": {k: v / args.steps for k, v in stage_ms.items()},
                "stage_gbs": {k: (nb[k] / (stage_ms[k] * 1e-3) / 1e9 if stage_ms[k] > 0 else 0.0) for k in stage_ms},
                "note": "cbox is issue/latency bound (36 triangles, working set on chip): the HBM fraction is low by "
                        "construction; issue utilisation per kernel is in profiles/ (ncu)"}

    # ---------------- CPU baseline (oracle port) on a bounded sample
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import aq_oracle as ao
        o = ao.OracleScene(scene, build_bvh=True)
        ccfg = aq.Integrator(spp=args.cpu_spp, max_depth=5, seed=0).cfg(width=W, height=H)
        o.render(aq.Integrator(spp=1, max_depth=5, seed=0).cfg(width=W, height=H), mode=1)
        t0 = time.perf_counter()
        _, _, cst = o.render(ccfg, mode=1)
        cdt = time.perf_counter() - t0
        cpu = {"value": (cst["rays_closest"] + cst["rays_shadow"]) / cdt / 1e6, "unit": "Mrays/s", "cores": ao.threads(),
               "kind": "port", "sample": f"{args.cpu_spp} of {args.spp} spp of the same {W}x{H} image (oracle BVH2)",
               "samples_per_s": cst["samples"] / cdt}

    line = {
        "metric": "Mrays/s", "value": rays / secs / 1e6, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32",
        "data": f"reference scene assets (scenes/{args.scene}.json, {scene.desc.n_tris} triangles); no dataset substitution needed",
        "config": {"workload": workload_name(args), "scene": args.scene, "width": W, "height": H, "spp_per_gpu": se - sb,
                   "max_depth": 5, "pool_paths": args.pool or (1 << 24), "parallelism": f"spp-partition x{world}, film reduce",
                   "l2": "no flush: the wavefront pool rewritten every wave (11 x 16 B x pool = 2.95 GB) exceeds the 126 MB L2"},
        "samples_per_s": samples / secs, "sample_bounces_per_s": bounces / secs,
        "rays_per_sample": rays / samples, "bounces_per_sample": bounces / samples,
        "clocks": clk,
        "e2e": {"value": float(e2e_rays.item()) / float(e2e_s.item()) / 1e6, "unit": "Mrays/s",
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(W * H * 16),
                "ms_per_step": float(e2e_s.item()) / args.steps * 1e3},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
