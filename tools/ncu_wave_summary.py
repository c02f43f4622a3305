"""Per-stage totals of ONE wavefront wave from an `ncu --set full` capture (run here, no GPU):
launches, time, warp instructions, DRAM bytes, SIMD efficiency, issue utilisation per stage.
bench.py multiplies `warp_inst_per_wave` by the waves of its own run to state the issue-roof
fraction next to the HBM fraction.

  python tools/ncu_wave_summary.py gpurun_out/prof_cbox.ncu-rep --scene cbox --width 1024 --height 1024 \
         --pool 16777216 --source r02_cbox_ncu_full.json --out profiles/ncu_summary.json
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_summary import load_rep

ISSUE_ROOF = 148 * 4 * 1.965e9


def stage_of(kernel):
    if "aq_k_raygen" in kernel:
        return "raygen"
    if "aq_k_film" in kernel:
        return "film"
    if "aq_k_shade" in kernel:
        return "shade"
    if "aq_k_trace<3" in kernel or "aq_k_trace<(int)3" in kernel:
        return "closest"
    if "aq_k_trace<1" in kernel or "aq_k_trace<(int)1" in kernel:
        return "shadow"
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--scene", default="cbox")
    ap.add_argument("--width", type=int, default=1024)
    ap.add_argument("--height", type=int, default=1024)
    ap.add_argument("--pool", type=int, default=1 << 24)
    ap.add_argument("--source", default="")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    rows = load_rep(a.rep)
    stages = {}
    for d in rows:
        st = stage_of(d["kernel"])
        if st is None:
            continue
        s = stages.setdefault(st, {"launches": 0, "us": 0.0, "warp_inst": 0.0, "dram": 0.0, "thr_w": 0.0, "iss_w": 0.0, "l2": 0.0,
                                   "regs": d.get("regs"), "occ_w": 0.0})
        s["launches"] += 1
        s["us"] += d["duration_us"]
        s["warp_inst"] += d["warp_inst"]
        s["dram"] += d["dram_read_bytes"] + d.get("dram_write_bytes", 0.0)
        s["l2"] += d.get("l2_sectors", 0.0) * 32.0
        s["thr_w"] += d["threads_per_inst"] * d["warp_inst"]
        s["iss_w"] += d["issue_active_pct"] * d["duration_us"]
        s["occ_w"] += d.get("achieved_occupancy_pct", 0.0) * d["duration_us"]
    out = {"wave": {"scene": a.scene, "width": a.width, "height": a.height, "pool": a.pool},
           "source": a.source, "issue_roof_warp_inst_per_s": ISSUE_ROOF, "stages": {}}
    for k, s in stages.items():
        out["stages"][k] = {
            "launches_per_wave": s["launches"], "us_per_wave_under_ncu": round(s["us"], 2),
            "warp_inst_per_wave": s["warp_inst"], "dram_bytes_per_wave": s["dram"], "l2_bytes_per_wave": s["l2"],
            "threads_per_inst": round(s["thr_w"] / max(1.0, s["warp_inst"]), 3),
            "issue_active_pct": round(s["iss_w"] / max(1e-9, s["us"]), 2),
            "occupancy_pct": round(s["occ_w"] / max(1e-9, s["us"]), 2),
            "issue_frac_under_ncu": round(s["warp_inst"] / (s["us"] * 1e-6) / ISSUE_ROOF, 4),
            "dram_gbs_under_ncu": round(s["dram"] / (s["us"] * 1e-6) / 1e9, 1), "regs": s["regs"]}
    txt = json.dumps(out, indent=1)
    print(txt)
    if a.out:
        open(a.out, "w").write(txt + "\n")


if __name__ == "__main__":
    main()
