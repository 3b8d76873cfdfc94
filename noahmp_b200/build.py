"""Build the CUDA shared library in-tree (sm_100a only).

    python -m noahmp_b200.build [--fast] [--force]

* ``libnoahmp_b200.so``       parity build: ``-fmad=false`` -- every +,-,*,/ and sqrt is the same IEEE
                              operation, in the same order, as the reference's fp32 Fortran; only libm
                              calls (expf/logf/powf/...) may differ in the last ulp.
* ``libnoahmp_b200_fast.so``  ``--use_fast_math`` -- FMA contraction, approximate division/sqrt and the MUFU-based
                              __expf/__logf/__powf; 2.6x faster than the parity build and statistically as
                              close to the oracle (the error floor is the physics' sensitivity, not the ulp).
* ``libnoahmp_b200_strict.so`` ``-fmad=false -DNMP_CR_LIBM`` -- libm calls evaluated in fp64 and rounded once;
                              must agree BIT FOR BIT with the oracle's ORACLE_CR_LIBM build (verification only).

The built ``.so`` files are git-ignored but travel to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["noahmp_abi.cu"]
HEADERS = ["common.cuh", "energy.cuh", "water.cuh", "column_step.cuh", "forcing_prep.cuh", os.path.join("..", "..", "include", "noahmp_b200.h")]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


# All variants are the SAME source.  NMP_PHASE_BARRIERS puts a CTA barrier at every phase boundary so the warps
# of a CTA stay inside one instruction-cache window (the step is ~16k SASS instructions; without the
# barriers ncu shows 70-80% of warp samples in stall_no_inst).  Block sizes were swept on B200
# (profiles/r01_block_sweep.md, r01_tuning.md): the IEEE builds (3x the code) want many warps in lock-step (768 threads,
# 80 registers; 256 x 3 is 37 % slower there), the smaller fast-math code is best at 3 CTAs x 256 threads, 80 registers.
# NMP_SYNC_MASK=0x7f7af: every numbered barrier site except the three inside fixed-count loops (4, 6, 11), +2 %.
COMMON = ["-DNMP_PHASE_BARRIERS", "-DNMP_PREFETCH_FORCING", "-DNMP_SYNC_MASK=0x7f7afu"]
VARIANTS = {
    "parity": ("libnoahmp_b200.so", ["-fmad=false", "-DNMP_BLOCK=768", "-DNMP_MIN_BLOCKS=1"]),
    "fast": ("libnoahmp_b200_fast.so", ["--use_fast_math", "-DNMP_FAST_ALGEBRA", "-DNMP_BLOCK=256", "-DNMP_MIN_BLOCKS=3"]),
    "strict": ("libnoahmp_b200_strict.so", ["-fmad=false", "-DNMP_CR_LIBM", "-DNMP_BLOCK=768", "-DNMP_MIN_BLOCKS=1"]),
    # the strict build with every data-dependent index range-checked on the device (common.cuh NMP_CHK): debug only
    "bounds": ("libnoahmp_b200_bounds.so", ["-fmad=false", "-DNMP_CR_LIBM", "-DNMP_BOUNDS", "-DNMP_BLOCK=768", "-DNMP_MIN_BLOCKS=1"]),
}


def _variant(fast) -> str:
    if isinstance(fast, str):
        return fast
    return "fast" if fast else "parity"


def lib_path(fast=False) -> str:
    override = os.environ.get("NOAHMP_B200_LIB")   # experiments only: absolute path of an alternative build
    if override:
        return override
    return os.path.join(HERE, VARIANTS[_variant(fast)][0])


def _stale(out: str) -> bool:
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_library(fast=False, force: bool = False, verbose: bool = False) -> str:
    out = lib_path(fast)
    if not force and not _stale(out):
        return out
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc, *ARCH, "-lineinfo", "-O3", "-std=c++17", *COMMON, *VARIANTS[_variant(fast)][1],
           "-Xcompiler", "-fPIC", "-shared", "-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return out


if __name__ == "__main__":
    which = [v for v in VARIANTS if f"--{v}" in sys.argv] or ["parity"]
    if "--all" in sys.argv:
        which = list(VARIANTS)
    for v in which:
        print(build_library(fast=v, force="--force" in sys.argv, verbose="-v" in sys.argv))
