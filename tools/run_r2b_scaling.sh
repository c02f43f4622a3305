#!/bin/bash
# one multi-GPU box: usage run_r2b_scaling.sh N [tests]  — optional NCCL film-equality tests, then bench.py at N GPUs
# (weak C2 + strong C5-slice sub-record); outputs gpurun_out/r02c_scale_n$N.json
n=$1
mkdir -p gpurun_out
if [ "$2" = tests ]; then
  ( timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_headline.py -m gpu -q -k "multi or nccl" 2>&1 | tail -6 ) > gpurun_out/r02c_pytest_multi_gpu_n$n.log 2>&1
  cat gpurun_out/r02c_pytest_multi_gpu_n$n.log
fi
if [ "$n" = 1 ]; then
  python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_scale_n1.json 2> gpurun_out/r02c_scale_n1.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) \
      bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_scale_n$n.json 2> gpurun_out/r02c_scale_n$n.err
fi
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r02c_scale_n$n.json") if l.startswith("{")][-1])
s = d.get("strong") or {}
print($n, "weak", round(d["value"]), "Mrays/s", round(d["ms_per_step"], 1), "ms e2e", round(d["e2e"]["value"]), "| strong", round(s.get("value", 0)), round(s.get("ms_per_step", 0), 1), "ms | per-rank ms", [round(x, 1) for x in d["ms_per_step_per_rank"]])
PY
tail -n 3 gpurun_out/r02c_scale_n$n.err
