#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -s -k "skips_latent or fused_mlp_forward" > gpurun_out/r2z6_pytest_a.log 2>&1; echo "pytest a rc=$?"
tail -12 gpurun_out/r2z6_pytest_a.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2z6_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2z6_pytest.log
B="--steps 20 --warmup 5 --no-modes --no-cpu-baseline --sustain-steps 0 --no-reuse-line --no-configs"
for i in 1 2 3; do
  NRF_SKIP_EMPTY_TILES=0 timeout 300 python bench.py $B > gpurun_out/r2z6_bench_dense_$i.json 2>> gpurun_out/r2z6_bench.err
  timeout 300 python bench.py $B > gpurun_out/r2z6_bench_skip_$i.json 2>> gpurun_out/r2z6_bench.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2z6_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d["kernel_ms_per_step"]
        print(f, d["ms_per_step"], "fwd", k["fused_fwd"], "bwd", k["fused_bwd"], "wgrad", k["wgrad_tc"], "gemm", k["gemm_tc"], "frac", d["roofline"]["frac"], "inbox", d["in_box"]["ms_per_step"])
    except Exception as e:
        print(f, "ERR", e)
PY
tail -5 gpurun_out/r2z6_bench.err
