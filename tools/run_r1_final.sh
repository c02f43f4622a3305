#!/bin/bash
# Final verification of the round's tree: GPU tests, smoke, the C2 bench line (both arms).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/final_pytest.log
python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/final_smoke.log
python bench.py > gpurun_out/bench_c2_final.json 2> gpurun_out/bench_final.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_c2_ref_final.json 2>> gpurun_out/bench_final.err
python -c "
import json
d=json.load(open('gpurun_out/bench_c2_final.json')); r=json.load(open('gpurun_out/bench_c2_ref_final.json'))
print('value',d['value'],'sb/s',d['sample_bounces_per_s'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'frac',d['roofline']['frac'],'cpu',d['cpu_baseline']['value'],'ref arm',r['value'],'clocks',d['clocks'])"
