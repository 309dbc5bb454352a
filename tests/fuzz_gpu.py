"""tests/fuzz_gpu.py -- randomized differential run of the sweep kernels against the oracle (not collected by pytest; run by hand on a GPU box):

    python tests/fuzz_gpu.py [--cases 300] [--seed 1]

Every case draws a shape (N from 5 to 6000 with every N % 4, M from 1 to 3000), a missing rate, phenotype NAs, a work-item length, an
amount of twin (none / all / a random prefix of the stripes), the table staging mode and CTA shape, the gather form of the
missing-genotype list, and input vectors with a random dynamic range (outliers up to 1e6 x), and checks: layout round trip and counts
bit-exact, statistics 1e-12, X.v and X^T.u within 1e-6 of the oracle in the 2-norm and the infinity norm, list == second walk bit for
bit when every stripe is class 0, results identical across the staging modes."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gvamp_b200 import capi  # noqa: E402
from oracle import oracle as O  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cases", type=int, default=300)
ap.add_argument("--seed", type=int, default=1)
a = ap.parse_args()
rng = np.random.default_rng(a.seed)
rel = lambda x, y: float(np.linalg.norm(x - y) / max(np.linalg.norm(y), 1e-300))
linf = lambda x, y: float(np.max(np.abs(x - y)) / max(np.max(np.abs(y)), 1e-300))
KEYS = ("GVB_TWIN", "GVB_TWIN_STRIPES", "GVB_TAB", "GVB_GATHER_TAB", "GVB_PAIR_SHAPE", "GVB_MISS", "GVB_MISS_SUM", "GVB_AX_TPC", "GVB_ATX_SPC")
worst = {"ax": 0.0, "atx": 0.0, "ax_inf": 0.0, "atx_inf": 0.0}
for case in range(a.cases):
    N = int(rng.integers(5, 6000))
    M = int(rng.integers(1, 3000))
    miss = float(rng.choice([0.0, 0.0, 0.01, 0.05, 0.2]))
    bed = O.synth_bed(int(rng.integers(1, 1 << 30)), 0, M, N, miss_rate=miss)
    present = rng.random(N) > float(rng.choice([0.0, 0.0, 0.02, 0.5]))
    if present.sum() < 3:
        present[:3] = True
    mask4 = O.make_mask4(N, present)
    ds = O.Dataset(bed, N, mask4=mask4, nonas=int(present.sum()))
    n_stripes = ((N + 3) // 4 + 31) // 32
    env = {}
    tw = rng.integers(0, 3)
    if tw == 0:
        env["GVB_TWIN"] = "0"
    elif tw == 2:
        env["GVB_TWIN_STRIPES"] = str(int(rng.integers(0, n_stripes + 1)))
    env["GVB_TAB"] = str(rng.choice(["tma", "cpasync"]))
    env["GVB_GATHER_TAB"] = str(rng.choice(["tma", "cpasync"]))
    env["GVB_PAIR_SHAPE"] = str(int(rng.integers(0, 3)))
    env["GVB_MISS_SUM"] = str(rng.choice(["warp", "lane"]))
    if rng.random() < 0.5:
        env["GVB_AX_TPC"], env["GVB_ATX_SPC"] = str(int(rng.integers(1, 9))), str(int(rng.integers(1, 9)))
    v, u = rng.normal(size=M), rng.normal(size=N)
    kind = rng.integers(0, 4)
    if kind == 1:
        v[rng.integers(0, M)] *= 10.0 ** rng.integers(1, 7)
        u[rng.integers(0, N)] *= 10.0 ** rng.integers(1, 7)
    elif kind == 2:
        v *= np.where(rng.random(M) < 0.05, 1.0, 1e-5)
        u[: N // 2] *= 1e-4
    elif kind == 3:
        v *= 10.0 ** rng.integers(-100, 100)
        u *= 10.0 ** rng.integers(-100, 100)
    out = {}
    for mode in ("list", "twopass") if miss > 0 else ("list",):
        for k in KEYS:
            os.environ.pop(k, None)
        os.environ.update(env)
        os.environ["GVB_MISS"] = mode
        with capi.Context(0) as ctx:
            ctx.load_host(bed, N).set_mask(mask4, int(present.sum())).compute_stats(1.0)
            assert np.array_equal(ctx.decode(0, M), bed), (case, "decode")
            assert np.array_equal(ctx.counts(), ds.counts()), (case, "counts")
            mave, msig = ctx.stats()
            assert rel(mave, ds.mave) < 1e-12 and rel(msig, ds.msig) < 1e-10, (case, "stats")
            out[mode] = (ctx.Ax(v), ctx.ATx(u))
    ax, atx = out["list"]
    ax_ref, atx_ref = ds.Ax(v), ds.ATx(u)
    e = (rel(ax, ax_ref), rel(atx, atx_ref), linf(ax, ax_ref), linf(atx, atx_ref))
    worst = {"ax": max(worst["ax"], e[0]), "atx": max(worst["atx"], e[1]), "ax_inf": max(worst["ax_inf"], e[2]), "atx_inf": max(worst["atx_inf"], e[3])}
    assert max(e) < 1e-6, (case, N, M, miss, env, kind, e)
    if "twopass" in out:
        assert np.array_equal(out["twopass"][0], ax), (case, "X.v differs between the list and the second walk")
        assert rel(out["twopass"][1], atx_ref) < 1e-6
        if kind == 0:   # no outliers: every stripe is class 0 and the two forms are the same integers
            assert np.array_equal(out["twopass"][1], atx), (case, N, M, miss, env)
    if case % 25 == 24:
        print(f"{case + 1} cases ok; worst so far {worst}", flush=True)
for k in KEYS:
    os.environ.pop(k, None)
print(f"fuzz_gpu: {a.cases} cases passed (seed {a.seed}); worst errors {worst}")
