#!/usr/bin/env python3
# -*- coding: utf8 -*-
"""Noah-MP offline driver: same entry point as the reference's ``run/main.py`` (positional ``nmlfile``,
default ``case.nml``, ``main(nmlfile)``), but instead of printing ``dir(cfg)`` (main:12-14) it loads the
parameter tables, builds the columns and advances them on the GPU through the C ABI.

The reference defines no forcing reader, no ``init.nc`` reader and no cosine-zenith computation.  When the
namelist's ``initialization_file`` exists (NetCDF-3 with WRF/HRLDAS names, or ``.npz``) the file-driven loop of
``noahmp_b200.driver.run_files`` runs: forcing records from ``input_directory``, outputs and restarts at the namelist's
frequencies.  Otherwise the synthetic single-point workload of ``noahmp_b200.synth`` is run over the namelist's period.
"""
import argparse
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(_HERE, '..'), os.path.join(_HERE, '..', 'offline')):
    if os.path.abspath(_p) not in sys.path:
        sys.path.insert(0, os.path.abspath(_p))

import noahmp_config  # noqa: E402

DEFAULT_NAMELIST_FILE = 'case.nml'


def main(nmlfile, ncol=1, verbose=True):
    cfg = noahmp_config.Config(nmlfile)
    from noahmp_b200.driver import run_case
    summary = run_case(cfg, ncol=ncol)
    if verbose:
        for key, val in summary.items():
            print('{:>22s} : {}'.format(key, val))
    return summary


if __name__ == '__main__':
    parser = argparse.ArgumentParser(description='Noah-MP Land Surface Model (B200 column-step engine)')
    parser.add_argument('nmlfile', nargs='?', type=str, default=DEFAULT_NAMELIST_FILE, help='configuration file')
    parser.add_argument('--ncol', type=int, default=1, help='replicate the single point over this many columns')
    args = parser.parse_args()
    main(args.nmlfile, ncol=args.ncol)
