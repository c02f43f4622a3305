#!/bin/bash
# BASELINE config C5 exactly: room.json 3840x2160, 4096 spp split over the ranks, one NCCL film reduce.
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 1 --warmup 1 --no-cpu-baseline --scaling strong --scene room --width 3840 --height 2160 --spp 4096 2>/dev/null | tail -1 > gpurun_out/c5_n1.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29588 bench.py --gpus 8 --steps 1 --warmup 1 --no-cpu-baseline --scaling strong --scene room --width 3840 --height 2160 --spp 4096 2>/dev/null | tail -1 > gpurun_out/c5_n8.json
python tools/show_bench.py gpurun_out/c5_n1.json gpurun_out/c5_n8.json
