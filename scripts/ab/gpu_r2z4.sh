#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -s -k "tma_box or encode" > gpurun_out/r2z4_pytest_enc.log 2>&1; echo "pytest enc rc=$?"
tail -4 gpurun_out/r2z4_pytest_enc.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2z4_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2z4_pytest.log
B="--steps 20 --warmup 5 --no-modes --no-cpu-baseline --sustain-steps 0 --no-reuse-line --no-configs"
for i in 1 2; do
  NRF_ENCODE_TMA=1 timeout 300 python bench.py $B > gpurun_out/r2z4_bench_tma_$i.json 2> gpurun_out/r2z4_bench_tma_$i.err
  timeout 300 python bench.py $B > gpurun_out/r2z4_bench_ldg_$i.json 2> gpurun_out/r2z4_bench_ldg_$i.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2z4_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], "encode", d["kernel_ms_per_step"]["encode"], "inbox", d["in_box"]["ms_per_step"], "encode", d["in_box"]["kernel_ms_per_step"]["encode"], d["in_box"]["encode_frac_of_hbm"], "scatter", d["in_box"]["kernel_ms_per_step"]["scatter"])
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/r2z4_bench_tma_1.err
