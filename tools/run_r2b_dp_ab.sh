#!/bin/bash
# device builder: cost-optimal (DP) collapse vs the greedy one — parity on the variant library, room on the
# device-built tree, C4 soup.  usage: run_r2b_dp_ab.sh LIB
lib=${1:-libaqua_cuda.so}
mkdir -p gpurun_out
( echo "== device-builder parity tests on $lib"; AQUA_CUDA_LIB=$lib timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "device or lbvh or hybrid or soup" 2>&1 | tail -3 ) > gpurun_out/r02b_dp_verify.log 2>&1
cat gpurun_out/r02b_dp_verify.log
SKIP_PYTEST=1 bash tools/run_r2b_ab2.sh r02b_dp "$lib|AQUA_ACCEL_BUILDER=device AQUA_COLLAPSE=greedy" "$lib|AQUA_ACCEL_BUILDER=device AQUA_COLLAPSE=dp" | grep room
for c in greedy dp; do AQUA_CUDA_LIB=$lib AQUA_COLLAPSE=$c python tools/bench_soup.py --brief --rays 33554432 --check-bvh 262144 2>/dev/null | sed "s/^/[collapse=$c] /"; done | tee gpurun_out/r02b_dp_soup.log
