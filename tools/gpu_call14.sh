#!/bin/bash
# round 2, GPU call 14 (8 GPUs): weak scaling of the headline workload (N = 1, 2, 4, 8; N = 8 with the full
# parity / strong / e2e blocks), strong scaling of BASELINE configs[3] (8192^2, 40 it) and configs[4]
# (16384^2, 80 it), NCCL transport and mirrored halo at 8, slab parity at 8, blocking depth for short slabs.
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() { # name nproc args...
  local name=$1 n=$2; shift 2
  if [ $n = 1 ]; then timeout 240 python bench.py --gpus 1 "$@" 2> gpurun_out/c14_$name.err | grep -E "^\{" > gpurun_out/c14_$name.json
  else timeout 300 $R --nproc-per-node=$n --master-port $((29800 + RANDOM % 100)) bench.py --gpus $n "$@" 2> gpurun_out/c14_$name.err | grep -E "^\{" > gpurun_out/c14_$name.json; fi
}
SLAB_W=2048 SLAB_H=4096 SLAB_WD=2048 SLAB_HD=4096 SLAB_ITERS=50 timeout 200 $R --nproc-per-node=8 --master-port 29791 tools/slab_check.py 2>&1 | grep -E "SLAB_CHECK|rror" | head -3 > gpurun_out/c14_slab_check.log
run weak1 1 --steps 200 --warmup 10 --quick
run weak2 2 --steps 200 --warmup 10 --quick
run weak4 4 --steps 200 --warmup 10 --quick
run weak8 8 --steps 200 --warmup 10 --no-cpu
FLUID_HALO=nccl run weak8_nccl 8 --steps 200 --warmup 10 --quick
FLUID_HALO_MIRROR=1 run weak8_mirror 8 --steps 200 --warmup 10 --quick
for n in 1 2 4 8; do run c3_strong$n $n --grid 8192 --iters 40 --strong --steps 50 --warmup 5 --quick; done
for kb in 5 7 10; do run s4096_8_k$kb 8 --strong --jacobi-block $kb --steps 200 --warmup 10 --quick; done
for n in 1 8; do run c4_strong$n $n --grid 16384 --iters 80 --strong --steps 5 --warmup 3 --quick; done
cat gpurun_out/c14_slab_check.log
python - <<'P'
import json, glob, os
for f in sorted(glob.glob("gpurun_out/c14_*.json")):
    try:
        d = json.load(open(f))
        print(os.path.basename(f), "n", d["n_gpus"], "ms/step", round(d["ms_per_step"], 4), "T/s", round(d["value"] / 1e12, 3), d.get("scaling"), d.get("parity"), (d.get("strong") or {}).get("ms_per_step"))
    except Exception as e:
        print(f, "failed", e)
P
