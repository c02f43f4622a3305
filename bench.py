#!/usr/bin/env python
"""bench.py — the driver's measurement contract for the render hot path.

A *step* is one full render of BASELINE.json's headline configuration C2:
scenes/cbox.json at 1024x1024, 1024 spp, max depth 5, NEE on the point light
(~1.07e9 camera paths, ~2.6e9 shaded path vertices, ~5.4e9 rays per step and GPU).

  python bench.py --gpus N --steps K --warmup W           (N>1: launched under torchrun)
  python bench.py --impl reference ...                    (CPU oracle on the host cores)

`value` times the device-resident path (scene + BVH already in HBM, film stays in HBM), with no
profiling events in the stream; the per-stage split behind `roofline` comes from ONE extra step
with AQ_RENDER_PROFILE after the timed region.  `e2e` times the C-ABI call a user makes with HOST
buffers: aq_scene_create (H2D of the scene arrays) + aq_accel_build + aq_render (film D2H into
pinned memory) every step.
With N GPUs every rank renders its own 1024-spp sample range of the same image (weak
scaling: per-GPU work fixed) and the float4 films are summed with one NCCL reduce; the line also
carries `strong`: a short slice of config C5 (scenes/room.json at 3840x2160, --strong-spp samples
per pixel SPLIT over the N ranks + one film reduce), so the per-N lines of a scaling run hold the
north-star spp split as well.

`roofline` is reported for the traversal kernel aq_k_trace (its closest-hit and any-hit
instantiations together: one kernel source, 60 % of the step) — fixed, not "whichever stage
happened to be longest in this run" — with every stage's HBM fraction and, where profiles/
ncu_summary.json holds an ncu capture of the same wave shape, its issue fraction next to it.
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


ISSUE_ROOF = 148 * 4 * 1.965e9  # warp instructions / s: 148 SMs x 4 schedulers x 1.965 GHz


def load_ncu_summary(scene, W, H, pool):
    """profiles/ncu_summary.json: per-stage totals of ONE wave from an `ncu --set full` capture
    (tools/ncu_wave_summary.py).  Used only when this run renders waves of the captured shape."""
    import glob
    for p in sorted(glob.glob(os.path.join(ROOT, "profiles", "ncu_summary*.json"))):
        try:
            d = json.load(open(p))
        except Exception:
            continue
        wv = d.get("wave", {})
        if (wv.get("scene"), wv.get("width"), wv.get("height"), wv.get("pool")) == (scene, W, H, pool):
            return d
    return None


def run_strong_slice(args, aq, aqd, torch, dist, r, rank, world, dev):
    """Short slice of config C5: room.json at 3840x2160, `--strong-spp` samples per pixel split
    [k*spp/N, (k+1)*spp/N) over the N ranks, films summed with one reduce.  Device-timed, max over
    ranks; the driver's per-N lines then carry the strong-scaling number of the north star."""
    W, H, spp = 3840, 2160, args.strong_spp
    try:
        scene = aq.Scene.load(os.path.join(aq.scenes_dir(), "room.json"))
    except Exception as e:  # assets missing on this box: report, do not fail the bench line
        return {"unavailable": f"{type(e).__name__}: {e}"}
    ds = r.upload(scene)
    ds.accel_wait()  # hybrid build: time the final (SAH) tree
    film = torch.zeros(H, W, 4, device=dev)
    sb, se = aqd.partition_spp(0, spp, rank, world)
    cfg = aq.Integrator(spp=spp, max_depth=5, seed=0).cfg(width=W, height=H, spp_begin=sb, spp_end=se)

    def step():
        ds.render_device_async(cfg, film.data_ptr())
        aqd.reduce_film(film, 0)
        return ds.finish()

    step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sts = [step() for _ in range(args.strong_steps)]
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    rays = torch.tensor([sum(s["rays_closest"] + s["rays_shadow"] for s in sts), sum(s["samples"] for s in sts)],
                        device=dev, dtype=torch.float64)
    per_rank = [torch.zeros_like(ms) for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, ms)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(rays, op=dist.ReduceOp.SUM)
    else:
        per_rank = [ms]
    ds.close()
    del film
    secs = float(ms.item()) * 1e-3
    return {"workload": f"C5 slice: scenes/room.json 3840x2160, {spp} spp split over {world} GPU(s), one film reduce (132.7 MB)",
            "scaling": "strong", "value": float(rays[0].item()) / secs / 1e6, "unit": "Mrays/s",
            "samples_per_s": float(rays[1].item()) / secs, "steps": args.strong_steps,
            "ms_per_step": float(ms.item()) / args.strong_steps,
            "ms_per_step_per_rank": [float(x.item()) / args.strong_steps for x in per_rank]}


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
    ap.add_argument("--strong-spp", type=int, default=64, help="spp of the C5 slice in the `strong` sub-record (0 = skip)")
    ap.add_argument("--strong-steps", type=int, default=2)
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
    ds.accel_wait()  # `value` is the scene-resident number: final tree (e2e below takes the hybrid build as a user gets it)
    film = torch.zeros(H, W, 4, device=dev)
    # weak scaling: rank k renders samples [k*spp, (k+1)*spp) of the same image
    if args.scaling == "strong":
        sb, se = aqd.partition_spp(0, args.spp, rank, world)
    else:  # weak scaling: rank k renders samples [k*spp, (k+1)*spp) of the same image
        sb, se = rank * args.spp, (rank + 1) * args.spp
    cfg = integ.cfg(width=W, height=H, spp_begin=sb, spp_end=se, pool_paths=args.pool)
    cfg_prof = integ.cfg(width=W, height=H, spp_begin=sb, spp_end=se, pool_paths=args.pool, flags=aq.AQ_RENDER_PROFILE)

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
    ms_ranks = [torch.zeros_like(ms) for _ in range(world)]
    if world > 1:
        dist.all_gather(ms_ranks, ms)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    else:
        ms_ranks = [ms.clone()]
    ms_per_rank = [float(x.item()) / args.steps for x in ms_ranks]
    ms_total = float(ms.item())
    # stage split: one extra step with events after every launch of every 8th wave (outside the timed region)
    ds.render_device_async(cfg_prof, film.data_ptr())
    prof_stats = [ds.finish()]
    barrier()
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

    # ---------------- strong-scaling sub-record (all ranks take part)
    strong = None
    if args.strong_spp > 0 and args.scaling == "weak":
        strong = run_strong_slice(args, aq, aqd, torch, dist, r, rank, world, dev)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---------------- roofline (rank 0's own launches): the traversal kernel, plus every stage
    peak, peak_src = load_peaks()
    ps = prof_stats[0]
    stage_ms = {"raygen": ps["ms_raygen"], "closest": ps["ms_trace"], "shade": ps["ms_shade"], "shadow": ps["ms_shadow"],
                "film": ps["ms_film"]}
    # the profiled step is a few % slower than a timed one (events in the stream): scale its split to the timed step
    k_scale = (ms_total / args.steps) / max(1e-9, sum(stage_ms.values()))
    stage_ms = {k: v * k_scale for k, v in stage_ms.items()}
    pool = args.pool or (1 << 24)
    spw = max(1, pool // min(W * H, pool))  # samples of one pixel held by one wave
    nb = stage_bytes(ps, spw)
    waves = ps["n_waves"]
    n_launch = {"raygen": waves, "film": waves, "closest": waves * integ.max_depth, "shade": waves * integ.max_depth,
                "shadow": waves * integ.max_depth}
    gbs = {k: (nb[k] / (stage_ms[k] * 1e-3) / 1e9 if stage_ms[k] > 0 else 0.0) for k in stage_ms}
    ncu = load_ncu_summary(args.scene, W, H, pool)
    full_waves = (se - sb) % spw == 0
    stages = {}
    for k in stage_ms:
        d = {"ms_per_step": stage_ms[k], "share": stage_ms[k] / max(1e-9, sum(stage_ms.values())), "algorithmic_gbs": gbs[k],
             "hbm_frac": gbs[k] / peak}
        if ncu and full_waves and k in ncu.get("stages", {}):
            n = ncu["stages"][k]
            d["issue_frac"] = n["warp_inst_per_wave"] * waves / (stage_ms[k] * 1e-3) / ISSUE_ROOF
            d["ncu_threads_per_inst"] = n.get("threads_per_inst")
            d["ncu_issue_active_pct"] = n.get("issue_active_pct")
            d["ncu_dram_bytes_per_wave"] = n.get("dram_bytes_per_wave")
        stages[k] = d
    t_ms = stage_ms["closest"] + stage_ms["shadow"]
    t_bytes = nb["closest"] + nb["shadow"]
    t_launch = n_launch["closest"] + n_launch["shadow"]
    achieved = t_bytes / (t_ms * 1e-3) / 1e9 if t_ms > 0 else 0.0
    traffic = issue_frac = None
    if ncu and full_waves:
        st_ = ncu.get("stages", {})
        if "closest" in st_ and "shadow" in st_:
            traffic = (st_["closest"]["dram_bytes_per_wave"] + st_["shadow"]["dram_bytes_per_wave"]) * waves / max(1, t_launch)
            issue_frac = (st_["closest"]["warp_inst_per_wave"] + st_["shadow"]["warp_inst_per_wave"]) * waves / (t_ms * 1e-3) / ISSUE_ROOF
    roofline = {"bound": "hbm", "kernel": "aq_k_trace (closest-hit <3> + any-hit <1> instantiations; fixed choice: the traversal kernel is 60 % of the step)",
                "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "issue_frac": issue_frac, "issue_roof_warp_inst_per_s": ISSUE_ROOF,
                "issue_frac_source": (f"warp instructions per wave from the ncu --set full capture {ncu.get('source')} of this wave shape x waves of this run / "
                                      "CUDA-event stage time of this run" if issue_frac is not None else None),
                "algorithmic_bytes_per_launch": t_bytes / max(1, t_launch),
                "avg_launch_ms": t_ms / max(1, t_launch),
                "share_of_step": t_ms / max(1e-9, sum(stage_ms.values())),
                "stages": stages,
                "stage_split_source": "one extra step with AQ_RENDER_PROFILE (events around every launch of every 8th wave), scaled to the timed ms_per_step",
                "note": "cbox is issue bound (36 triangles, working set on chip): the HBM fraction is low by construction, "
                        "issue_frac is the binding roof (SURVEY 8d)"}

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
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "ms_per_step_per_rank": ms_per_rank,
        "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32",
        "data": f"reference scene assets (scenes/{args.scene}.json, {scene.desc.n_tris} triangles); no dataset substitution needed",
        "config": {"workload": workload_name(args), "scene": args.scene, "width": W, "height": H, "spp_per_gpu": se - sb,
                   "max_depth": 5, "pool_paths": args.pool or (1 << 24), "parallelism": f"spp-partition x{world}, film reduce",
                   "l2": "no flush: the wavefront queues rewritten every wave (11 x 16 B x pool = 2.95 GB) exceed the 126 MB L2"},
        "samples_per_s": samples / secs, "sample_bounces_per_s": bounces / secs,
        "rays_per_sample": rays / samples, "bounces_per_sample": bounces / samples,
        "clocks": clk,
        "e2e": {"value": float(e2e_rays.item()) / float(e2e_s.item()) / 1e6, "unit": "Mrays/s",
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(W * H * 16),
                "ms_per_step": float(e2e_s.item()) / args.steps * 1e3},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "strong": strong,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
