#!/bin/bash
# quick check after a change: full GPU suite, stage bench, C3 e2e (hybrid build) with the builder's timings
mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/chk_pytest_gpu.log 2>&1
( python tools/stage_bench.py; python tools/stage_bench.py --scene room --res 1920 1080 --spp 8 ) 2>/dev/null | grep prof=4 > gpurun_out/chk_stage.log
AQ_BUILD_VERBOSE=1 AQ_BENCH_DEBUG=1 python bench.py --scene room --width 1920 --height 1080 --spp 256 --steps 3 --warmup 3 --no-cpu-baseline --strong-spp 0 > gpurun_out/chk_bench_c3.json 2> gpurun_out/chk_bench_c3.err
cat gpurun_out/chk_pytest_gpu.log gpurun_out/chk_stage.log; grep -v warning gpurun_out/chk_bench_c3.err | tail -n 9
python - <<'PY'
import json
d = json.load(open("gpurun_out/chk_bench_c3.json"))
print("C3 value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ratio", round(d["e2e"]["value"] / d["value"], 3), "ms", round(d["ms_per_step"], 1), round(d["e2e"]["ms_per_step"], 1))
PY
