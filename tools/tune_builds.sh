#!/bin/bash
# Builds tuning variants of the fast build into noahmp_b200/exp/ (git-ignored *.so), benchmark kernel only
# (-DNMP_EXPERIMENT_ONE_KERNEL, ~6 s per build, all cores); tools/tune_run.sh then times them on a B200.
#
#   tools/tune_builds.sh                       # the geometry matrix of profiles/r01_tuning.md
#   tools/tune_builds.sh - <<< "name -Dflag ... "   # one variant per line on stdin: NAME followed by extra nvcc flags
#
# Base = the shipped fast build (noahmp_b200/build.py) without its geometry; every variant adds its own
# -DNMP_BLOCK / -DNMP_MIN_BLOCKS (or -DNMP_MIN_BLOCKS=0 -DNMP_MAXNREG=n for an explicit register cap).
set -e
cd "$(dirname "$0")/../noahmp_b200/csrc"
mkdir -p ../exp
export B="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -DNMP_PHASE_BARRIERS -DNMP_PREFETCH_FORCING -DNMP_SYNC_MASK=0x7f7afu -Xcompiler -fPIC -shared --use_fast_math -DNMP_FAST_ALGEBRA -DNMP_EXPERIMENT_ONE_KERNEL"
build() { nvcc $B ${*:2} -o ../exp/$1.so noahmp_abi.cu; }
export -f build
if [ "$1" = "-" ]; then
  cat
else
  echo "base -DNMP_BLOCK=256 -DNMP_MIN_BLOCKS=3"
  for g in 128x6 256x2 256x4 384x2 512x1 768x1; do echo "g$g -DNMP_BLOCK=${g%x*} -DNMP_MIN_BLOCKS=${g#*x}"; done
fi | xargs -P "$(nproc)" -L 1 bash -c 'build "$@"' _
ls ../exp
