"""Top SASS/source hot spots of one kernel from an .ncu-rep (source page, needs -lineinfo).
Usage: python tools/ncu_source_top.py rep.ncu-rep <kernel regex> [launch index] [topN]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    topn = int(sys.argv[4]) if len(sys.argv) > 4 else 25
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda" if False else "sass",
                                   "--kernel-name", f"regex:{kern}"], stderr=subprocess.DEVNULL).decode()
    # the output holds one table per launch, each starting with a "Kernel Name" row
    blocks = raw.split('"Kernel Name"')[1:]
    blk = '"Kernel Name"' + blocks[which]
    rows = list(csv.reader(io.StringIO(blk)))
    hdr = rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    tot_samples = sum(float(r[col["# Samples"]] or 0) for r in data)
    tot_inst = sum(float(r[col["Instructions Executed"]] or 0) for r in data)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = defaultdict(float)
    for r in data:
        for s in stalls:
            agg[s] += float(r[col[s]] or 0)
    print(f"launch {which}: samples={tot_samples:.0f} warp-inst={tot_inst:.0f} SASS lines={len(data)}")
    print("stall totals:", {k: round(v / tot_samples * 100, 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
    data.sort(key=lambda r: -float(r[col["# Samples"]] or 0))
    for r in data[:topn]:
        top = sorted(((float(r[col[s]] or 0), s) for s in stalls), reverse=True)[:2]
        print(f"{float(r[col['# Samples']])/tot_samples*100:5.1f}%  inst={float(r[col['Instructions Executed']] or 0):9.0f} "
              f"thr={r[col['Avg. Threads Executed']]:>5}  {r[col['Source']][:70]:70s} {[(s, int(v)) for v, s in top]}")


if __name__ == "__main__":
    main()
