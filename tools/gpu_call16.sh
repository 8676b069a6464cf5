#!/bin/bash
# round 2, GPU call 16 (2 GPUs): where does the mirrored-halo solve lose its time? (device stamps)
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
FLUID_HALO_MIRROR=1 FLUID_DEBUG_HALO_TIMING=1 timeout 300 $R --nproc-per-node=2 --master-port 29771 bench.py --gpus 2 --steps 200 --warmup 10 --quick > gpurun_out/c16_mirror.log 2>&1
FLUID_HALO_MIRROR=1 FLUID_PDL=0 FLUID_DEBUG_HALO_TIMING=1 timeout 300 $R --nproc-per-node=2 --master-port 29772 bench.py --gpus 2 --steps 200 --warmup 10 --quick > gpurun_out/c16_mirror_nopdl.log 2>&1
FLUID_DEBUG_HALO_TIMING=1 timeout 300 $R --nproc-per-node=2 --master-port 29773 bench.py --gpus 2 --steps 200 --warmup 10 --quick > gpurun_out/c16_base.log 2>&1
FLUID_PDL=0 timeout 300 $R --nproc-per-node=2 --master-port 29774 bench.py --gpus 2 --steps 200 --warmup 10 --quick > gpurun_out/c16_base_nopdl.log 2>&1
for n in mirror mirror_nopdl base base_nopdl; do echo "== $n"; grep -E "mirror rank|halo rank" gpurun_out/c16_$n.log | grep -v '^{' | cut -c1-420; grep -oE '"ms_per_step": [0-9.]+' gpurun_out/c16_$n.log | head -1; done
