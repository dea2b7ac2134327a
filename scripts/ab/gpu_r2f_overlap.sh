#!/bin/bash
# A/B of renderer.overlap_scatter (the merged volume scatter on a side stream under the last pass's weight gradients):
# the bit-identity test, then three alternating pairs of the config-2 bench line, all on one box.
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q -k "scatter_under or cuda_graph or reproducible or full_s32" > gpurun_out/r2f_pytest_overlap.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2f_pytest_overlap.log
F="--steps 20 --warmup 5 --no-modes --no-cpu-baseline --no-reuse-line --no-configs --no-extra"
for i in 1 2 3; do
  for v in 0 1; do
    NRF_SCATTER_OVERLAP=$v timeout 120 python bench.py $F > gpurun_out/r2f_bench_ov${v}_$i.json 2>> gpurun_out/r2f_bench_ov.err
    python - <<P
import json
d=json.loads(open("gpurun_out/r2f_bench_ov${v}_$i.json").read().strip().splitlines()[-1])
print("overlap=$v run $i", d["ms_per_step"], d["e2e"]["ms_per_step"], d.get("sustained",{}).get("ms_per_step"), {k:d["kernel_ms_per_step"][k] for k in ("wgrad_tc","scatter","fused_bwd")})
P
  done
done
