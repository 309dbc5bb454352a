"""profiles/dual_timing.py -- sustained cost of the dual sweeps (gvb_dAx2 / gvb_dATx2: two products, one bed read) against two single sweeps.

    python profiles/dual_timing.py [--M 275000] [--pairs 40] [--twin 0]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gvamp_b200 import capi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=400_000)
ap.add_argument("--M", type=int, default=275_000)
ap.add_argument("--pairs", type=int, default=40)
ap.add_argument("--twin", default=None)
a = ap.parse_args()
if a.twin is not None:
    os.environ["GVB_TWIN"] = a.twin
bed = a.M * ((a.N + 3) // 4)
ctx = capi.Context(0)
ctx.synth(1, a.N, a.M, 0, a.M, 0.0)
ctx.compute_stats(1.0)
rng = np.random.default_rng(0)
v0, v1 = ctx.vecM(rng.normal(size=a.M)), ctx.vecM(rng.normal(size=a.M))
o0, o1, p0, p1 = ctx.vecN(), ctx.vecN(), ctx.vecN(), ctx.vecN()
u0, u1 = ctx.vecN(rng.normal(size=a.N)), ctx.vecN(rng.normal(size=a.N))
m0, m1, q0, q1 = ctx.vecM(), ctx.vecM(), ctx.vecM(), ctx.vecM()
ctx.dAx(v0, o0), ctx.dAx(v1, o1), ctx.dAx2(v0, v1, p0, p1)
same = np.array_equal(o0.download(), p0.download()) and np.array_equal(o1.download(), p1.download())
ctx.dATx(u0, m0), ctx.dATx(u1, m1), ctx.dATx2(u0, u1, q0, q1)
same_t = np.array_equal(m0.download(), q0.download()) and np.array_equal(m1.download(), q1.download())
print(f"shard {a.N} x {a.M} = {bed / 1e9:.1f} GB, twin state {ctx.twin_state()} ({ctx.twin_stripes()} stripes); dual == two singles bit for bit: "
      f"X.v {same}, X^T.u {same_t}")
for rep in range(2):
    ctx.profile(True)
    for _ in range(a.pairs):
        ctx.dAx(v0, o0)
        ctx.dAx(v1, o1)
    for _ in range(a.pairs):
        ctx.dATx(u0, m0)
        ctx.dATx(u1, m1)
    single = ctx.profile_read()
    for _ in range(a.pairs):
        ctx.dAx2(v0, v1, p0, p1)
    dual = ctx.profile_read_dual()
    for _ in range(a.pairs):
        ctx.dATx2(u0, u1, q0, q1)
    dual_t = ctx.profile_read_dual()
    ctx.profile(False)
    ms1, ms2 = single["ax_ms"] / single["ax_n"], dual["dual_ms"] / dual["dual_n"]
    print(f"rep {rep}: single X.v {ms1:.3f} ms = {bed / ms1 / 1e6:.0f} GB/s | dual {ms2:.3f} ms = {ms2 / ms1:.3f} x a single sweep "
          f"({2 * bed / ms2 / 1e6:.0f} GB/s of product bytes, {bed / ms2 / 1e6:.0f} GB/s of bed bytes)")
    ms1, ms2 = single["atx_ms"] / single["atx_n"], dual_t["dual_ms"] / dual_t["dual_n"]
    print(f"rep {rep}: single X^T.u {ms1:.3f} ms = {bed / ms1 / 1e6:.0f} GB/s | dual {ms2:.3f} ms = {ms2 / ms1:.3f} x a single sweep "
          f"({2 * bed / ms2 / 1e6:.0f} GB/s of product bytes, {bed / ms2 / 1e6:.0f} GB/s of bed bytes)")
ctx.close()
