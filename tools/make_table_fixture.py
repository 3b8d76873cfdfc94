#!/usr/bin/env python3
"""tests/golden/tables_reference.npz: every module array (`LK_*`, `KK_*`, flags, counts) as the reference's own
`noahmp_{gen,soil,veg}_param_readptable` routines leave it, for the four tag combinations of the shipped tables - produced by
executing those routines from their Fortran source text (oracle/f90_tables.py) on /root/reference/tbl.  Needs /root/reference."""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from oracle import f90_tables  # noqa: E402

COMBOS = [("STAS", "USGS"), ("STAS", "MODIFIED_IGBP_MODIS_NOAH"), ("STAS-RUC", "USGS"), ("STAS-RUC", "MODIFIED_IGBP_MODIS_NOAH")]


def main():
    out = {}
    for st, vt in COMBOS:
        for k, v in f90_tables.read_all("/root/reference/tbl", st, vt).items():
            out[f"{st}|{vt}|{k}"] = np.asarray(v)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "tables_reference.npz"), **out)
    print(len(out), "arrays")


if __name__ == "__main__":
    main()
