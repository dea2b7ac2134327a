#!/bin/bash
# the record at HEAD: full GPU suite, smoke, both bench arms, small-step probe, ncu of the fused kernels + launch list
set -x
bash scripts/gpu_r2c_final.sh
bash scripts/gpu_r2c_ncu.sh
