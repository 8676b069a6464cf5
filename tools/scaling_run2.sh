#!/bin/bash
# 8-GPU box: strong scaling of BASELINE configs[2] (4096^2/50), configs[3] (8192^2/40) and
# configs[4] (16384^2/80).  Outputs -> gpurun_out/strong_*.json
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() {  # grid iters n
  if [ "$3" = "1" ]; then
    timeout 300 python bench.py --grid $1 --iters $2 --steps 30 --warmup 3 --quick 2>&1 | grep -E "^\{" > gpurun_out/strong_$1_$3.json
  else
    timeout 300 $R --nproc-per-node=$3 --master-port 298$3$((RANDOM % 10)) bench.py --gpus $3 --grid $1 --iters $2 --strong --steps 30 --warmup 3 --quick 2>&1 | grep -E "^\{" > gpurun_out/strong_$1_$3.json
  fi
}
for n in 2 4 8; do run 4096 50 $n; done      # N=1: 0.287 ms (profiles/r01_bench_n1.json)
for n in 1 8; do run 8192 40 $n; done
for n in 8; do run 16384 80 $n; done
python - <<'P'
import json, glob
for g in (4096, 8192, 16384):
    base = None
    for n in (1, 2, 4, 8):
        try:
            d = json.load(open(f"gpurun_out/strong_{g}_{n}.json"))
        except Exception as e:
            continue
        if n == 1: base = d["value"]
        if base is None and g == 4096: base = 4096 * 4096 * 50 / 0.287e-3
        print(g, n, "ms", round(d["ms_per_step"], 4), "T/s", round(d["value"] / 1e12, 3), "speedup", round(d["value"] / base, 2) if base else None)
P
