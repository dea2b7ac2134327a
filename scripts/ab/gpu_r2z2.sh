#!/bin/bash
# round 2, call z2: scatter-reduce A/B (in-box) after interleaving the voxel ownership, ncu of the new kernel
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "scatter" > gpurun_out/r2z2_pytest.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/r2z2_pytest.log
B="--steps 20 --warmup 5 --no-modes --no-cpu-baseline --sustain-steps 0 --no-reuse-line --no-configs"
for i in 1 2; do
  NRF_SCATTER_RCF=1 timeout 300 python bench.py $B > gpurun_out/r2z2_bench_old_$i.json 2> gpurun_out/r2z2_bench_old_$i.err
  timeout 300 python bench.py $B > gpurun_out/r2z2_bench_new_$i.json 2> gpurun_out/r2z2_bench_new_$i.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2z2_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["kernel_ms_per_step"]["scatter"], d["in_box"]["ms_per_step"], d["in_box"]["kernel_ms_per_step"]["scatter"], d["in_box"]["scatter_frac_of_hbm"], d["in_box"]["kernel_ms_per_step"]["encode"])
    except Exception as e:
        print(f, "ERR", e)
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:scatter_reduce_cf2 -c 1 -s 1 -o gpurun_out/prof_scatter_cf2_r2z -f python scripts/one_step.py bf16 inbox 2 > gpurun_out/ncu_cf2.log 2>&1; echo "ncu rc=$?"
