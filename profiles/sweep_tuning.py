"""profiles/sweep_tuning.py -- SUSTAINED X.v / X^T.u rates of the tile-kernel variants on one HBM-resident shard.

    python profiles/sweep_tuning.py [--M 275000] [--pairs 40] [--configs name=ENV1:val,ENV2:val ...]

Every configuration runs `pairs` back-to-back (X.v, X^T.u) sweep pairs (about half a second for the 27.5 GB shard: long enough for the
1 kW power cap to settle the SM clock, which is what the sweeps see inside a VAMP iteration) and reports the mean per-sweep time of
each kind from CUDA events around all launches of a sweep.  Not a bench value: a tuning aid whose output is kept under profiles/."""
import argparse
import os
import subprocess
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gvamp_b200 import capi  # noqa: E402

DEFAULT = ["cpasync15x2=GVB_TAB:cpasync", "tma11x2=GVB_TAB:tma,GVB_PAIR_SHAPE:0", "tma12x2=GVB_TAB:tma,GVB_PAIR_SHAPE:1", "tma7x3=GVB_TAB:tma,GVB_PAIR_SHAPE:2"]
ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=400_000)
ap.add_argument("--M", type=int, default=275_000)
ap.add_argument("--pairs", type=int, default=40)
ap.add_argument("--miss", type=float, default=0.0)
ap.add_argument("--twin", default=None, help="GVB_TWIN for every configuration (0: X.v gathers)")
ap.add_argument("--configs", nargs="*", default=DEFAULT)
a = ap.parse_args()
if a.twin is not None:
    os.environ["GVB_TWIN"] = a.twin


def sm_clock():
    try:
        return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True).stdout.strip()
    except Exception:
        return "?"


bed = a.M * ((a.N + 3) // 4)
ctx = capi.Context(0)
ctx.synth(1, a.N, a.M, 0, a.M, a.miss)
ctx.compute_stats(1.0)
rng = np.random.default_rng(0)
v, u = ctx.vecM(rng.normal(size=a.M)), ctx.vecN(rng.normal(size=a.N))
ov, ou = ctx.vecN(), ctx.vecM()
ctx.dAx(v, ov)
ctx.dATx(u, ou)
ref_ax, ref_atx = ov.download(), ou.download()
print(f"shard {a.N} x {a.M} = {bed / 1e9:.1f} GB, twin state {ctx.twin_state()} ({ctx.twin_stripes()} stripes), {a.pairs} sweep pairs per configuration")
for spec in a.configs:
    name, _, envs = spec.partition("=")
    keys = []
    for kv in filter(None, envs.split(",")):
        k, _, val = kv.partition(":")
        os.environ[k] = val
        keys.append(k)
    clk = []
    stop = False

    def sample():
        while not stop:
            clk.append(sm_clock())
            time.sleep(0.1)

    th = threading.Thread(target=sample, daemon=True)
    th.start()
    ctx.sync()
    ctx.profile(True)
    for _ in range(a.pairs):
        ctx.dAx(v, ov)
        ctx.dATx(u, ou)
    p = ctx.profile_read()
    stop = True
    th.join()
    same = bool(np.array_equal(ov.download(), ref_ax) and np.array_equal(ou.download(), ref_atx))
    mhz = sorted(float(c.split(",")[0]) for c in clk if "," in c)
    print(f"{name:>14s}: X.v {p['ax_ms'] / p['ax_n']:7.3f} ms = {bed / (p['ax_ms'] / p['ax_n']) / 1e6:6.0f} GB/s | X^T.u {p['atx_ms'] / p['atx_n']:7.3f} ms = "
          f"{bed / (p['atx_ms'] / p['atx_n']) / 1e6:6.0f} GB/s | SM clock median {mhz[len(mhz) // 2] if mhz else 0:.0f} MHz | bit-identical to the first configuration: {same}", flush=True)
    for k in keys:
        os.environ.pop(k, None)
ctx.close()
