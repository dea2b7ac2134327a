#!/bin/bash
set -x
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:mlp_fused -s 4 -c 4 -o gpurun_out/prof_fused_r2c python scripts/one_step.py bf16 x 2 > gpurun_out/ncu_fused_r2c.log 2>&1; echo "ncu fused rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2c.csv python bench.py --steps 2 --warmup 1 --no-extra --no-modes --no-cpu-baseline --sustain-steps 0 --no-reuse-line > gpurun_out/ncu_launches_r2c.log 2>&1; echo "ncu list rc=$?"
