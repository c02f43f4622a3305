"""Attribute the executed warp instructions of one kernel to SOURCE lines (including the inlined
call chain), from an .ncu-rep taken with --import-source on and a cubin of the same source built
with -lineinfo.  ncu's CSV source page has SASS rows only; `nvdisasm -gi` has the line table; the
two are joined by instruction index.

usage: python tools/ncu_line_attrib.py rep.ncu-rep <ncu kernel regex> <mangled-name prefix> [launch] [file-for-lines]
  e.g. python tools/ncu_line_attrib.py gpurun_out/prof_room.ncu-rep aq_k_trace _Z10aq_k_traceILi3ELb0 0 aq_bvh.h"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "aqua-engine_b200", "csrc")


def main():
    rep, kre, mangled = sys.argv[1:4]
    launch = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    focus = sys.argv[5] if len(sys.argv) > 5 else "aq_core.h"
    tmp = tempfile.mkdtemp()
    cubin = os.path.join(tmp, "k.cubin")
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
                           "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false", "-I", os.path.join(ROOT, "include"),
                           "-I", CSRC, "-cubin", "-o", cubin, os.path.join(CSRC, "aq_cuda.cu")]
                          + os.environ.get("AQ_LINE_DEFS", "").split())  # e.g. AQ_LINE_DEFS="-DAQ_TRI_DYN=8" for an A/B variant
    txt = subprocess.check_output(["nvdisasm", "-gi", cubin]).decode().split("\n")
    start = [i for i, l in enumerate(txt) if l.startswith(".text." + mangled)][0]
    insts, stack, fresh = [], [], True
    for l in txt[start + 1:]:
        if l.startswith("//---------------------"):
            break
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
        if m:
            if not fresh:
                stack, fresh = [], True
            stack.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+.*?;", l):
            insts.append(list(stack))
            fresh = False
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass",
                                   "--kernel-name", f"regex:{kre}"], stderr=subprocess.DEVNULL).decode()
    blocks = ['"Kernel Name"' + b for b in raw.split('"Kernel Name"')[1:]]
    want = os.environ.get("AQ_LINE_KERNEL")  # e.g. "aq_k_trace<(int)3, (bool)0, (int)6>": the launch-th launch of THAT instantiation
    if want:
        blocks = [b for b in blocks if want in b.split("\n", 1)[0]]
    blk = blocks[launch]
    rows = list(csv.reader(io.StringIO(blk)))
    hdr = rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    print(f"{rows[0][1][:70]}: {len(data)} SASS rows in the report, {len(insts)} in the rebuilt cubin")
    src = {f: open(os.path.join(CSRC, f)).read().split("\n") for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))}
    outer, inner, thr = collections.Counter(), collections.Counter(), collections.Counter()
    # warp-state sampling columns, when the report has them: where the warps WAIT (not where they issue)
    samp_col = next((h for h in hdr if "Sampling" in h and "All" in h), None)
    stall_cols = [h for h in hdr if h.startswith("stall_")]
    samp_outer, stall_tot, samp_inner = collections.Counter(), collections.Counter(), collections.Counter()
    tot = 0.0
    for k in range(min(len(data), len(insts))):
        ie = float(data[k][col["Instructions Executed"]] or 0)
        te = float(data[k][col["Thread Instructions Executed"]] or 0)
        tot += ie
        st = insts[k]
        if not st:
            continue
        inner[st[0]] += ie
        key = next((e for e in reversed(st) if e[0] == focus), st[-1])  # outermost frame inside `focus`
        outer[key] += ie
        thr[key] += te
        if samp_col:
            try:
                sv = float(data[k][col[samp_col]] or 0)
            except ValueError:
                sv = 0.0
            samp_outer[key] += sv
            samp_inner[st[0]] += sv
        for h in stall_cols:
            try:
                stall_tot[h] += float(data[k][col[h]] or 0)
            except ValueError:
                pass
    for title, agg in (("by outermost line in " + focus, outer), ("by innermost line", inner)):
        print("---", title)
        for k, v in agg.most_common(30):
            line = src.get(k[0], [""] * (k[1] + 1))[k[1] - 1].strip()[:90]
            extra = f" thr={thr[k] / max(v, 1):4.1f}" if agg is outer else ""
            print(f"{v / tot * 100:5.1f}%{extra}  {k[0]}:{k[1]}  {line}")


    if samp_col and sum(samp_outer.values()) > 0:
        ts = sum(samp_outer.values())
        print(f"--- warp samples ({samp_col}) by outermost line in {focus}")
        for k, v in samp_outer.most_common(15):
            line = src.get(k[0], [""] * (k[1] + 1))[k[1] - 1].strip()[:90]
            print(f"{v / ts * 100:5.1f}%  {k[0]}:{k[1]}  {line}")
        print("--- warp samples by innermost line")
        for k, v in samp_inner.most_common(15):
            line = src.get(k[0], [""] * (k[1] + 1))[k[1] - 1].strip()[:90]
            print(f"{v / ts * 100:5.1f}%  {k[0]}:{k[1]}  {line}")
    if stall_tot:
        tt = sum(stall_tot.values()) or 1.0
        print("--- stall reasons over the kernel (sampled): " + ", ".join(f"{h[6:]} {v / tt * 100:.1f}%" for h, v in stall_tot.most_common(8)))


if __name__ == "__main__":
    main()
