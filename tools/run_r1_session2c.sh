#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_nrc.py -m gpu -x -q 2>&1 | tail -3
python tools/bench_nrc.py --scene room > gpurun_out/nrc_room.json 2> gpurun_out/nrc.err
python tools/bench_nrc.py --scene cbox --res 1024 1024 > gpurun_out/nrc_cbox.json 2>> gpurun_out/nrc.err
cat gpurun_out/nrc_room.json gpurun_out/nrc_cbox.json
ncu --set full --clock-control none --import-source on -k regex:"aq_k_nrc_(query|train)" -c 3 -f -o gpurun_out/prof_nrc2 \
    python tools/bench_nrc.py --scene room --quick > gpurun_out/ncu_nrc2.log 2>&1
