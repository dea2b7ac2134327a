#!/bin/bash
set -x
mkdir -p gpurun_out
B="--steps 20 --warmup 5 --no-modes --no-cpu-baseline --sustain-steps 0 --no-reuse-line --no-extra"
for i in 1 2 3; do
  NRF_LIB_PATH=$PWD/real-robot-nerf-actor_b200/ab/libnrf_b200_directsave.so timeout 300 python bench.py $B > gpurun_out/r2z9_bench_direct_$i.json 2>> gpurun_out/r2z9_bench.err
  timeout 300 python bench.py $B > gpurun_out/r2z9_bench_staged_$i.json 2>> gpurun_out/r2z9_bench.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2z9_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d["kernel_ms_per_step"]
        print(f, d["ms_per_step"], "fwd", k["fused_fwd"], "bwd", k["fused_bwd"], "wgrad", k["wgrad_tc"], "frac", d["roofline"]["frac"], d["roofline"].get("frac_executed"))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -5 gpurun_out/r2z9_bench.err
NRF_LIB_PATH=$PWD/real-robot-nerf-actor_b200/ab/libnrf_b200_directsave.so timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "fused" > gpurun_out/r2z9_pytest_direct.log 2>&1; echo "pytest direct rc=$?"; tail -2 gpurun_out/r2z9_pytest_direct.log
timeout 200 python scripts/raygen_probe.py gpurun_out/r2c_raygen_probe.json > gpurun_out/r2c_raygen_probe.log 2>&1; echo "probe rc=$?"; tail -4 gpurun_out/r2c_raygen_probe.log
