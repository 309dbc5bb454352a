"""scripts/exp_run_with_lib.py <lib.so> <script.py> [args...] -- run a profiles/ script against an experimental build of the library
(timing experiments only; never used by tests or bench)."""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gvamp_b200 import capi  # noqa: E402

capi.LIB_PATH = os.path.abspath(sys.argv[1])
script = sys.argv[2]
sys.argv = [script] + sys.argv[3:]
runpy.run_path(script, run_name="__main__")
