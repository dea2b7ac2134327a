#!/bin/bash
# 2 GPUs: the multi-GPU tests (NCCL) and the bench line at N = 2
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2c_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/r2c_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2c_bench_n2.json 2> gpurun_out/r2c_bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c_bench_n2.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","n_gpus","ms_per_step_per_rank","multi_gpu_check")})
print({k:(v.get("ms_per_step") or v.get("ms_per_render")) for k,v in d["configs"].items()})
print(json.dumps(d["configs"]["config5"])[:900])
PY
