"""profiles/ncu_summary.json (read by bench.py for roofline.traffic) and a markdown table from
the per-launch JSON files written by tools/ncu_summary.py.
Usage: python tools/make_profile_summary.py profiles/r01_cbox_ncu_full.json [profiles/r01_room_ncu_full.json ...]"""
import json
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PEAK_HBM = 6555.8  # GB/s, MEASURED_PEAKS.json
ISSUE_PEAK = 148 * 4 * 1.965e9  # warp-inst/s


def family(name):
    if "aq_k_trace<3" in name or "aq_k_trace<0" in name:
        return "closest"
    if "aq_k_trace<1" in name or "aq_k_trace<2" in name:
        return "shadow"
    for k in ("shade", "raygen", "film"):
        if "aq_k_" + k in name:
            return k
    return "other"


def main():
    out_md = ["| capture | kernel | launches | avg us | DRAM GB/s (% of 6555.8 measured) | L2 GB/s | L2 hit % | issue active % | warp-inst/s vs 1.163e12 roof | threads/inst | occupancy % | regs | warp-inst/launch | DRAM bytes/launch |",
              "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
    summary = {}
    for path in sys.argv[1:]:
        rows = json.load(open(path))
        fam = defaultdict(list)
        for r in rows:
            fam[family(r["kernel"])].append(r)
        tag = os.path.basename(path).replace("_ncu_full.json", "")
        for k, rs in fam.items():
            n = len(rs)
            avg = lambda key: sum(float(r.get(key, 0) or 0) for r in rs) / n
            dram = avg("dram_read_bytes") + avg("dram_write_bytes")
            gbs = sum(r.get("dram_gbs", 0) for r in rs) / n
            out_md.append(f"| {tag} | {k} | {n} | {avg('duration_us'):.1f} | {gbs:.0f} ({gbs / PEAK_HBM * 100:.1f} %) | {avg('l2_gbs'):.0f} | {avg('l2_hit_pct'):.1f} | "
                          f"{avg('issue_active_pct'):.1f} | {sum(r.get('warp_inst', 0) / r['duration_us'] for r in rs) / n * 1e6 / ISSUE_PEAK * 100:.1f} % | {avg('threads_per_inst'):.1f} | {avg('achieved_occupancy_pct'):.1f} | {avg('regs'):.0f} | "
                          f"{avg('warp_inst'):.3g} | {dram:.3g} |")
            if tag.endswith("cbox"):
                summary[k] = {"dram_bytes_per_launch": dram, "avg_us": avg("duration_us"), "launches": n,
                              "issue_active_pct": avg("issue_active_pct"), "threads_per_inst": avg("threads_per_inst"),
                              "source": os.path.basename(path)}
    json.dump(summary, open(os.path.join(ROOT, "profiles", "ncu_summary.json"), "w"), indent=1)
    print("\n".join(out_md))


if __name__ == "__main__":
    main()
