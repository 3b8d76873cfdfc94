"""C ABI surface on a machine without a GPU: every symbol declared in include/noahmp_b200.h is exported, the
layout constants of header / CUDA library / Python / oracle agree, and the product fails loudly without CUDA."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT
from noahmp_b200 import layout as L
from noahmp_b200 import _lib
from noahmp_b200.build import VARIANTS, lib_path

HEADER = os.path.join(ROOT, "include", "noahmp_b200.h")


def header_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(noahmp_[a-z0-9_]+)\s*\(", text)))


@pytest.mark.parametrize("variant", list(VARIANTS))
def test_library_exports_every_declared_symbol(variant):
    path = lib_path(variant)
    assert os.path.exists(path), f"{path} missing: run `python -m noahmp_b200.build --all`"
    lib = C.CDLL(path)
    declared = header_symbols()
    assert len(declared) >= 16
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert sorted(_lib.SYMBOLS) == declared


def test_layout_constants_agree_everywhere(oracle_libs):
    lib = _lib.load()
    hdr = dict(re.findall(r"#define\s+NOAHMP_(N[A-Z_]+)\s+(\d+)", open(HEADER).read()))
    expect = {"NSTATE": L.NSTATE, "NSTATE_CORE": L.NSTATE_CORE, "NFORCING": L.NFORCING, "NSTATIC_F": L.NSTATIC_F,
              "NDIAG": L.NDIAG, "NBUDGET": L.NBUDGET}
    for k, v in expect.items():
        assert int(hdr[k]) == v, k
    got = [lib.noahmp_layout(i) for i in range(12)]
    assert got[:5] == [L.NSTATE, L.NSTATE_CORE, L.NFORCING, L.NSTATIC_F, L.NDIAG] and got[11] == L.NBUDGET
    assert got[10] == C.sizeof(_lib.RunParams)
    # names plane by plane
    assert [lib.noahmp_state_plane_name(i).decode() for i in range(L.NSTATE)] == L.STATE
    assert [lib.noahmp_forcing_plane_name(i).decode() for i in range(L.NFORCING)] == L.FORCING
    assert [lib.noahmp_diag_plane_name(i).decode() for i in range(L.NDIAG)] == L.DIAG
    assert [lib.noahmp_budget_name(i).decode() for i in range(L.NBUDGET)] == L.BUDGET
    assert lib.noahmp_state_plane_name(L.NSTATE) is None
    # the oracle states its layout independently (oracle/src/layout.h)
    orc = C.CDLL(os.path.join(oracle_libs, "liboracle.so"))
    assert [orc.oracle_layout(i) for i in range(10)] == got[:10]
    from oracle.oracle import RunParams as ORP
    assert orc.oracle_layout(10) == C.sizeof(ORP) == got[10]


def test_type_packing_round_trip():
    import numpy as np
    lut, slt, slp, isc = np.array([1, 27, 14]), np.array([1, 19, 6]), np.array([9, 1, 3]), np.array([9, 1, 4])
    ist, ice = np.array([1, 2, 1]), np.array([0, 0, 1])
    t = L.pack_types(lut, slt, slp, isc, ist, ice)
    assert ((t & 0xff) == lut).all() and (((t >> 8) & 0xff) == slt).all() and (((t >> 16) & 0xff) == slp).all()
    assert (((t >> 24) & 0xf) == isc).all() and ((((t >> 28) & 1) + 1) == ist).all() and (((t >> 29) & 1) == ice).all()


def test_no_cpu_fallback():
    """Without a CUDA device the engine must raise, not compute on the CPU; a missing library must raise too."""
    import torch
    from noahmp_b200.engine import NoahMP, NoahMPError
    if not torch.cuda.is_available():
        with pytest.raises(NoahMPError, match="noahmp_init failed"):
            NoahMP()
    os.environ["NOAHMP_B200_LIB"] = "/nonexistent/libnoahmp.so"
    try:
        _lib._cache.pop("ghost", None)
        with pytest.raises(_lib.NoahMPLibraryError, match="no CPU fallback"):
            _lib.load("ghost")
    finally:
        del os.environ["NOAHMP_B200_LIB"]


def test_product_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/."""
    pkg = os.path.join(ROOT, "noahmp_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "oracle/src" not in text.replace(
                    "oracle/src/common.h, ORACLE_CR_LIBM", ""), os.path.join(base, f)
    for f in ("run/main.py", "offline/noahmp_config.py"):
        assert "oracle" not in open(os.path.join(ROOT, f)).read()
