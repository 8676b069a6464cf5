#!/bin/bash
# Runs inside gpurun on an 8-GPU box: bitwise slab check at 8 GPUs (both halo paths), then the
# bench metric at N = 1, 2, 4, 8 (peer-memory path) and N = 8 with NCCL.  Outputs -> gpurun_out/.
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for m in p2p nccl; do
  FLUID_HALO=$m SLAB_H=2048 SLAB_HD=4096 SLAB_ITERS=50 timeout 180 $R --nproc-per-node=8 --master-port 29751 tools/slab_check.py 2>&1 | grep -E "SLAB_CHECK|rror" | head -3
done | tee gpurun_out/slab_check8.log
python bench.py --steps 200 --warmup 10 --quick > gpurun_out/scale_1.json 2> gpurun_out/scale_1.err
for n in 2 4 8; do
  FLUID_DEBUG_HALO_TIMING=1 timeout 180 $R --nproc-per-node=$n --master-port 2976$n bench.py --gpus $n --steps 200 --warmup 10 --quick > gpurun_out/scale_$n.log 2>&1
  grep -E "^\{" gpurun_out/scale_$n.log > gpurun_out/scale_$n.json
done
FLUID_HALO=nccl timeout 180 $R --nproc-per-node=8 --master-port 29769 bench.py --gpus 8 --steps 200 --warmup 10 --quick 2>&1 | grep -E "^\{" > gpurun_out/scale_8_nccl.json
python - <<'P'
import json
for n in ("1", "2", "4", "8", "8_nccl"):
    try:
        d = json.load(open(f"gpurun_out/scale_{n}.json"))
        print(n, "ms/step", round(d["ms_per_step"], 4), "T updates/s", round(d["value"] / 1e12, 3), "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    except Exception as e:
        print(n, "failed", e)
P
grep -h "halo rank" gpurun_out/scale_8.log | head -8
