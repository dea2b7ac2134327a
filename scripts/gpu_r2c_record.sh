#!/bin/bash
# round 2, record call: captures at HEAD (ncu --set full of the dominant kernels, launch list), smoke, both bench arms
set -x
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2c_smoke.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c_bench_ref.json 2> gpurun_out/r2c_bench_ref.err; echo "ref rc=$?"
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:mlp_fused -s 4 -c 4 -o gpurun_out/prof_fused_r2c python scripts/one_step.py bf16 x 2 > gpurun_out/ncu_fused_r2c.log 2>&1; echo "ncu fused rc=$?"
timeout 300 $NCU -k regex:wgrad_multi -s 2 -c 2 -o gpurun_out/prof_wgradm_r2c python scripts/one_step.py bf16 x 2 > gpurun_out/ncu_wgradm_r2c.log 2>&1; echo "ncu wgrad rc=$?"
timeout 300 $NCU -k "regex:encode_points_tma|composite_|scatter_reduce_cf2|gemm_tc" -s 9 -c 9 -o gpurun_out/prof_hbm_inbox_r2c python scripts/one_step.py bf16 inbox 2 > gpurun_out/ncu_hbm_r2c.log 2>&1; echo "ncu hbm rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2c.csv python bench.py --steps 2 --warmup 1 --no-extra --no-modes --no-cpu-baseline --sustain-steps 0 --no-reuse-line > gpurun_out/ncu_launches_r2c.log 2>&1; echo "ncu list rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c_bench_n1.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches","kernel_ms_per_step","sustained","clocks")})
print(d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["roofline"].get("frac_executed"), d["in_box"]["encode_frac_of_hbm"], d["in_box"]["scatter_frac_of_hbm"])
print(open("gpurun_out/r2c_bench_ref.json").read()[:400])
PY
