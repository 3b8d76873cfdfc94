"""Helper of test_driver_gpu.py::test_two_rank_file_driven_run: run under torchrun, one rank per GPU."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "offline"))
import noahmp_config  # noqa: E402
from noahmp_b200.driver import run_files  # noqa: E402

cfg = noahmp_config.Config(sys.argv[1])
out = run_files(cfg, variant="parity")
if int(os.environ.get("RANK", "0")) == 0:
    print("RANKS", out["ranks"], "STEPS", out["steps"], "FLAGGED", out["flagged_column_launches"], "COLUMNS", out["columns"])
