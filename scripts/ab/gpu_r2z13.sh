#!/bin/bash
# re-layout of the touched voxels (first version: whole 32-voxel tiles, at every samples-per-voxel ratio below 1) vs the dense (C,V)->(V,C) pass (NRF_SPARSE_RELAYOUT=0)
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_render.py -m gpu -x -q -k "touched_voxel" > gpurun_out/r2z13_pytest_a.log 2>&1; echo "pytest a rc=$?"; tail -5 gpurun_out/r2z13_pytest_a.log
B="--steps 20 --warmup 5 --no-modes --no-cpu-baseline --sustain-steps 0 --no-reuse-line"
for i in 1 2; do
  NRF_SPARSE_RELAYOUT=0 timeout 300 python bench.py $B > gpurun_out/r2z13_bench_dense_$i.json 2>> gpurun_out/r2z13_bench.err
  timeout 300 python bench.py $B > gpurun_out/r2z13_bench_sparse_$i.json 2>> gpurun_out/r2z13_bench.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2z13_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d["kernel_ms_per_step"]
        print(f, d["ms_per_step"], "transpose", k["transpose"], "encode", k["encode"], "| inbox", d["in_box"]["ms_per_step"], d["in_box"]["kernel_ms_per_step"]["transpose"], "| cfg", {c:(v.get("ms_per_step") or v.get("ms_per_render")) for c,v in d["configs"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/r2z13_bench.err
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2z13_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2z13_pytest.log
NRF_SPARSE_RELAYOUT=0 python scripts/small_step_probe.py 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1])['nerfact.conf shape, SB=1']['reference schedule']; print('dense ', d['ms_per_step'], d['kernel_ms']['transpose'], d['kernel_ms_sum'])"
python scripts/small_step_probe.py 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1])['nerfact.conf shape, SB=1']['reference schedule']; print('sparse', d['ms_per_step'], d['kernel_ms']['transpose'], d['kernel_ms_sum'])"
