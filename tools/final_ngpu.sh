#!/bin/bash
# Round-end bench line on N GPUs of one box (torchrun, one rank per GPU): bash tools/final_ngpu.sh N [extra env as VAR=VALUE ...]
N=$1; shift
mkdir -p gpurun_out
for kv in "$@"; do export "$kv"; done
TAG=${TAG:-}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_r02_${N}gpu${TAG}.json 2> gpurun_out/r2f_b${N}${TAG}.err
python - gpurun_out/bench_r02_${N}gpu${TAG}.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d["n_gpus"], round(d["value"]), round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), d["e2e"].get("form"),
      "slide", d["slide"]["seconds"], "train", d["train"]["ms_per_step"], round(d["train"]["value"]), d["train"].get("ddp_check", {}).get("ok"))
PY
