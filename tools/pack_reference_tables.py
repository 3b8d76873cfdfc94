#!/usr/bin/env python3
"""Parse the reference's tbl/*.TBL with noahmp_b200.tables and write the packed JSON copies shipped in
noahmp_b200/data/ (the GPU box has no /root/reference).  Re-run whenever the reference tables change:

    python tools/pack_reference_tables.py [/root/reference/tbl]
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from noahmp_b200.tables import ParamTables  # noqa: E402


def main():
    tbl = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/tbl"
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "noahmp_b200", "data")
    os.makedirs(out, exist_ok=True)
    for soil_tag in ("STAS", "STAS-RUC"):
        for veg_tag in ("USGS", "MODIFIED_IGBP_MODIS_NOAH"):
            t = ParamTables.from_tbl_dir(tbl, soil_tag, veg_tag)
            path = os.path.join(out, f"tables_{soil_tag}_{veg_tag}.json")
            with open(path, "w") as fh:
                fh.write(t.to_json())
            print(path, t.nsoiltyp, t.ncolor, t.nvegtyp)


if __name__ == "__main__":
    main()
