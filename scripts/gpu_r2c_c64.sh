#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "tma_box or encode" > gpurun_out/r2c_pytest_enc64.log 2>&1; echo "pytest enc rc=$?"; tail -3 gpurun_out/r2c_pytest_enc64.log
NRF_ENCODE_TMA=0 python scripts/small_step_probe.py > gpurun_out/r2c_small_step_ldg.json 2>/dev/null
python scripts/small_step_probe.py > gpurun_out/r2c_small_step_tma.json 2>/dev/null
NRF_ENCODE_TMA=0 python scripts/small_step_probe.py > gpurun_out/r2c_small_step_ldg2.json 2>/dev/null
python scripts/small_step_probe.py > gpurun_out/r2c_small_step_tma2.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c_small_step_*.json")):
    d=json.loads(open(f).read().strip().splitlines()[-1])["nerfact.conf shape, SB=1"]
    print(f, d["reference schedule"]["ms_per_step"], d["reference schedule"]["kernel_ms"]["encode"], d["reference schedule"]["kernel_ms_sum"], d["cuda graph (reference schedule)"]["ms_per_step"])
PY
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2c_pytest.log
