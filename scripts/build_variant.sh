#!/bin/bash
# A/B builds: scripts/build_variant.sh <tag> <extra nvcc flags...>  ->  real-robot-nerf-actor_b200/ab/libnrf_b200_<tag>.so
# (use with NRF_LIB_PATH=... ; *.so is git-ignored but travels to the GPU box)
set -e
tag=$1; shift
cd "$(dirname "$0")/../real-robot-nerf-actor_b200"
mkdir -p ab/obj_$tag
for f in csrc/*.cu; do
  o=ab/obj_$tag/$(basename ${f%.cu}).o
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c $f -o $o &
done
wait
/usr/local/cuda/bin/nvcc -shared -o ab/libnrf_b200_$tag.so ab/obj_$tag/*.o -gencode arch=compute_100a,code=sm_100a -lcudart
rm -rf ab/obj_$tag
echo built ab/libnrf_b200_$tag.so
