#!/bin/bash
# one 8-GPU box: the bench line at N = 8 and at N = 1 back to back (same box: efficiency without box-to-box spread)
set -x
mkdir -p gpurun_out
F="--steps 20 --warmup 5 --no-modes --no-cpu-baseline --no-reuse-line"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 $F > gpurun_out/r2c_bench_n8.json 2> gpurun_out/r2c_bench_n8.err; echo "bench n8 rc=$?"
timeout 600 python bench.py --gpus 1 $F > gpurun_out/r2c_bench_n1_samebox.json 2> gpurun_out/r2c_bench_n1_samebox.err; echo "bench n1 rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2c_bench_n8.json","gpurun_out/r2c_bench_n1_samebox.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, {k:d.get(k) for k in ("value","ms_per_step","n_gpus","ms_per_step_per_rank","sustained")})
    print({k:(v.get("ms_per_step") or v.get("ms_per_render")) for k,v in d["configs"].items()})
    print(json.dumps(d["configs"]["config5"].get("volume_grad_exchange"))[:500])
PY
