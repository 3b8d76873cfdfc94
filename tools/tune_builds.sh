#!/bin/bash
# Builds the tuning matrix timed by tools/tune_libs.py into noahmp_b200/exp/ (git-ignored *.so): CTA geometries and
# single-barrier-removed variants of the fast build, benchmark kernel only (-DNMP_EXPERIMENT_ONE_KERNEL, ~6 s per build).
# Round-1 result (profiles/r01_tuning.md): 256 x 2 is the best geometry, no single barrier matters by more than 0.5 %.
set -e
cd "$(dirname "$0")/../noahmp_b200/csrc"
mkdir -p ../exp
export B="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -DNMP_PHASE_BARRIERS -Xcompiler -fPIC -shared --use_fast_math -DNMP_EXPERIMENT_ONE_KERNEL"
build() { nvcc $B ${*:2} -o ../exp/$1.so noahmp_abi.cu; }
export -f build
{
  echo "base -DNMP_BLOCK=256 -DNMP_MIN_BLOCKS=2"
  for g in 128x4 192x2 192x3 320x2 384x1 512x1; do echo "g$g -DNMP_BLOCK=${g%x*} -DNMP_MIN_BLOCKS=${g#*x}"; done
  for i in $(seq 0 18); do echo "m$i -DNMP_BLOCK=256 -DNMP_MIN_BLOCKS=2 -DNMP_SYNC_MASK=$(printf '0x%xu' $((0x7ffff & ~(1 << i))))"; done
} | xargs -P "$(nproc)" -L 1 bash -c 'build "$@"' _
ls ../exp
