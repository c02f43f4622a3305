#!/bin/bash
# one 8-GPU box: NCCL parity tests, then bench.py at N = 1, 2, 4, 8 (weak C2 + strong C5-slice sub-record)
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_headline.py -m gpu -q -k "multi or nccl" 2>&1 | tail -6 ) > gpurun_out/r02_pytest_multi_gpu.log 2>&1
python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_scale_n1.json 2> gpurun_out/r02_scale.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) \
      bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/r02_scale_n$n.json 2>> gpurun_out/r02_scale.err
done
cat gpurun_out/r02_pytest_multi_gpu.log
python - <<'PY'
import json
b = None
for n in (1, 2, 4, 8):
    try:
        d = json.loads([l for l in open(f"gpurun_out/r02_scale_n{n}.json") if l.startswith("{")][-1])
    except Exception as e:
        print(n, "failed", e); continue
    s = d.get("strong") or {}
    if n == 1: b = (d["value"], s.get("ms_per_step"))
    print(n, "weak", round(d["value"]), "Mrays/s", round(d["ms_per_step"], 1), "ms eff", round(d["value"] / (n * b[0]), 4), "e2e", round(d["e2e"]["value"]),
          "| strong", round(s.get("value", 0)), round(s.get("ms_per_step", 0), 1), "ms eff", round(b[1] / (n * s["ms_per_step"]), 4) if s.get("ms_per_step") else None,
          "| per-rank ms", [round(x, 1) for x in d["ms_per_step_per_rank"]])
PY
tail -n 5 gpurun_out/r02_scale.err
