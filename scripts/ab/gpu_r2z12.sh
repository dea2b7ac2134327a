#!/bin/bash
# fc_0 epilogue of the fused forward: ReLU as x & (x > 0) mask shared with the gate bits (default) vs max.bf16x2 + a second HSET2 (-DNRF_RELU_MAX build)
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "fused" > gpurun_out/r2z12_pytest_a.log 2>&1; echo "pytest a rc=$?"; tail -2 gpurun_out/r2z12_pytest_a.log
B="--steps 20 --warmup 5 --no-modes --no-cpu-baseline --sustain-steps 0 --no-reuse-line --no-extra"
for i in 1 2 3; do
  NRF_LIB_PATH=$PWD/real-robot-nerf-actor_b200/ab/libnrf_b200_relumax.so timeout 300 python bench.py $B > gpurun_out/r2z12_bench_hmnmx_$i.json 2>> gpurun_out/r2z12_bench.err
  timeout 300 python bench.py $B > gpurun_out/r2z12_bench_mask_$i.json 2>> gpurun_out/r2z12_bench.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2z12_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d["kernel_ms_per_step"]
        print(f, d["ms_per_step"], "fwd", k["fused_fwd"], "bwd", k["fused_bwd"], "wgrad", k["wgrad_tc"], "frac", d["roofline"]["frac"])
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/r2z12_bench.err
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2z12_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2z12_pytest.log
