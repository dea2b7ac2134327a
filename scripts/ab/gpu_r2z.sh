#!/bin/bash
# round 2, call z: full GPU suite, raygen rounding probe, scatter-reduce A/B (in-box)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2z_pytest.log
timeout 200 python scripts/raygen_probe.py gpurun_out/r2z_raygen_probe.json > gpurun_out/r2z_raygen_probe.log 2>&1; echo "probe rc=$?"
cat gpurun_out/r2z_raygen_probe.log | tail -8
B="--steps 20 --warmup 5 --no-modes --no-cpu-baseline --sustain-steps 0 --no-reuse-line --no-configs"
for i in 1 2; do
  NRF_SCATTER_RCF=1 timeout 300 python bench.py $B > gpurun_out/r2z_bench_old_$i.json 2> gpurun_out/r2z_bench_old_$i.err
  timeout 300 python bench.py $B > gpurun_out/r2z_bench_new_$i.json 2> gpurun_out/r2z_bench_new_$i.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2z_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["kernel_ms_per_step"]["scatter"], d["in_box"]["ms_per_step"], d["in_box"]["kernel_ms_per_step"]["scatter"], d["in_box"]["scatter_frac_of_hbm"], d["in_box"]["kernel_ms_per_step"]["encode"])
    except Exception as e:
        print(f, "ERR", e)
PY
