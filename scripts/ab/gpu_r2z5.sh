#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "tma_box or encode" > gpurun_out/r2z5_pytest_enc.log 2>&1; echo "pytest enc rc=$?"
tail -2 gpurun_out/r2z5_pytest_enc.log
B="--steps 20 --warmup 5 --no-modes --no-cpu-baseline --sustain-steps 0 --no-reuse-line --no-configs"
for st in 2 3 4 3 2 4; do
  NRF_ENCODE_TMA_STAGES=$st timeout 300 python bench.py $B > gpurun_out/r2z5_bench_st${st}_$RANDOM.json 2>> gpurun_out/r2z5_bench.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2z5_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], "encode", d["kernel_ms_per_step"]["encode"], "inbox", d["in_box"]["ms_per_step"], "encode", d["in_box"]["kernel_ms_per_step"]["encode"], d["in_box"]["encode_frac_of_hbm"])
    except Exception as e:
        print(f, "ERR", e)
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:encode_points_tma -c 2 -s 2 -o gpurun_out/prof_encode_tma_r2z -f python scripts/one_step.py bf16 inbox 2 > gpurun_out/ncu_enc_tma.log 2>&1; echo "ncu rc=$?"
