#!/bin/bash
# 8-GPU box: weak scaling of the headline config (C2) and strong scaling of C5 (room 3840x2160,
# spp split across ranks; 256 spp instead of 4096 to keep the run short), plus the CLI driver.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -1
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n"; fi
  $L bench.py --gpus $n --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/scale_weak_n$n.json
  $L bench.py --gpus $n --steps 2 --warmup 1 --no-cpu-baseline --scaling strong --scene room --width 3840 --height 2160 --spp 256 2>/dev/null | tail -1 > gpurun_out/scale_strong_c5_n$n.json
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/render.py \
  --scene baseline/_ref/scenes/room.json --res 1280 720 --spp 512 --output gpurun_out/room_8gpu.png 2>/dev/null | tail -1
python tools/render.py --scene baseline/_ref/scenes/cbox.json --integrator baseline/_ref/scenes/integrator.json --res 512 512 --spp 256 --output gpurun_out/cbox_cli.png | tail -1
python tools/show_bench.py gpurun_out/scale_weak_n*.json gpurun_out/scale_strong_c5_n*.json
