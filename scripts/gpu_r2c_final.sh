#!/bin/bash
# round 2 final record at HEAD: full GPU suite, smoke, the default bench line, the reference arm
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2c_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2c_smoke.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c_bench_ref.json 2> gpurun_out/r2c_bench_ref.err; echo "ref rc=$?"
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err; echo "bench rc=$?"
python scripts/small_step_probe.py > gpurun_out/r2c_small_step.json 2> gpurun_out/r2c_small_step.err; echo "small rc=$?"; tail -3 gpurun_out/r2c_small_step.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c_bench_n1.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches","kernel_ms_per_step","sustained","clocks")})
r=d["roofline"]; print(d["e2e"]["value"], d["e2e"]["ms_per_step"], r["frac"], r["frac_executed"], r["traffic"], r["step_frac_of_tensor_peak"])
print(d["in_box"]["ms_per_step"], d["in_box"]["encode_frac_of_hbm"], d["in_box"]["scatter_frac_of_hbm"], d["in_box"]["kernel_ms_per_step"])
print({k:(v.get("ms_per_step") or v.get("ms_per_render")) for k,v in d["configs"].items()})
print(d.get("reuse_coarse_evals"))
PY
