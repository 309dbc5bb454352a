"""profiles/run_sweeps.py -- a few X.v / X^T.u sweeps on a synthetic HBM-resident matrix; the command that the
ncu captures under profiles/ were taken with (see profiles/README.md)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gvamp_b200 import capi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=400_000)
ap.add_argument("--M", type=int, default=275_000)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--miss", type=float, default=0.0)
a = ap.parse_args()
ctx = capi.Context(0)
ctx.synth(1, a.N, a.M, 0, a.M, a.miss)
import time  # noqa: E402
for _ in range(2):
    ctx.sync()
    t0 = time.perf_counter()
    ctx.compute_stats(1.0)
    dt = time.perf_counter() - t0
print(f"compute_stats: {dt * 1e3:.3f} ms = {a.M * ((a.N + 3) // 4) / dt / 1e9:.0f} GB/s (host wall clock, includes the final sync)")
rng = np.random.default_rng(0)
v, u = ctx.vecM(rng.normal(size=a.M)), ctx.vecN(rng.normal(size=a.N))
ov, ou = ctx.vecN(), ctx.vecM()
bed = a.M * ((a.N + 3) // 4)
for r in range(a.reps):
    ctx.profile(True)
    ctx.dAx(v, ov)
    ctx.dATx(u, ou)
    p = ctx.profile_read()
    print(f"rep {r}: X.v {p['ax_ms']:.3f} ms = {bed / p['ax_ms'] / 1e6:.0f} GB/s | X^T.u {p['atx_ms']:.3f} ms = {bed / p['atx_ms'] / 1e6:.0f} GB/s")
print(f"twin layout state: {ctx.twin_state()} (1 = X.v walks the individual-major twin)")
ctx.close()
