"""profiles/sanitize_small.py -- a small pass over every hand-written kernel family for compute-sanitizer:

    compute-sanitizer --tool memcheck  python profiles/sanitize_small.py
    compute-sanitizer --tool racecheck python profiles/sanitize_small.py

Shapes are small (racecheck slows a kernel by two orders of magnitude) but cover what matters for the mbarrier / TMA pipeline of the
tile kernels: several work items per CTA, ragged stripes and marker tiles, both table buffers, the producer warp, X.v on the twin, on
the one matrix and on a partial twin (two launches), X^T.u with the missing-genotype list (both gather forms) and with the second
walk, the statistics walk (packed counters), the CG driver with a speculative iteration, the batched reductions.  Results are
checked against the oracle so that a "clean" run is also a correct one."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gvamp_b200 import capi  # noqa: E402
from oracle import oracle as O  # noqa: E402

rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
N, M = 9001, 1301          # 71 stripes (ragged), 11 marker tiles (ragged)
bed = O.synth_bed(3, 0, M, N, miss_rate=0.01)
ds = O.Dataset(bed, N)
rng = np.random.default_rng(0)
v, u = rng.normal(size=M), rng.normal(size=N)
ax_ref, atx_ref = ds.Ax(v), ds.ATx(u)
os.environ["GVB_AX_TPC"], os.environ["GVB_ATX_SPC"] = "3", "5"          # many work items per CTA
for twin, stripes in (("0", None), ("1", None), (None, "30")):
    for miss, form in (("list", "lane"), ("list", "warp"), ("twopass", "lane")):
        for k, val in (("GVB_TWIN", twin), ("GVB_TWIN_STRIPES", stripes), ("GVB_MISS", miss), ("GVB_MISS_SUM", form)):
            if val is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = val
        with capi.Context(0) as ctx:
            ctx.load_host(bed, N).compute_stats(1.0)
            assert np.array_equal(ctx.counts(), ds.counts())
            assert rel(ctx.Ax(v), ax_ref) < 1e-6 and rel(ctx.ATx(u), atx_ref) < 1e-6
            if miss == "list" and form == "lane":
                rhs, mu, ax = ctx.vecM(v), ctx.vecM(), ctx.vecN()
                its, _, d3 = ctx.cg_solve_ex(rhs, mu, 2.0, 0.7, 6, 1, ax)
                mu_ref, its_ref = O.precond_cg(ds, v, np.zeros(M), 2.0, 0.7, 6, 1)
                assert its == its_ref and rel(mu.download(), mu_ref) < 1e-5
                its, _ = ctx.cg_solve(rhs, mu, 2.0, 5.0, 30, 1)          # converges: one speculative iteration runs as empty launches
                assert its < 30
                r = ctx.reduce_batch([(capi.RED_DOT, rhs, mu, 0, 0, 1), (capi.RED_SQ, ax, None, 1.0, 0, 0)])
                assert np.all(np.isfinite(r))
                x1 = ctx.vecM()
                ctx.denoise(rhs, 3.0, [0.9, 0.06, 0.04], [0.0, 0.1, 1.0], x1)
                v2, o1, o2, p1, p2 = ctx.vecM(0.5 * v + 1.0), ctx.vecN(), ctx.vecN(), ctx.vecN(), ctx.vecN()     # dual sweep, prepared solve
                ctx.dAx(rhs, o1), ctx.dAx(v2, o2), ctx.dAx2(rhs, v2, p1, p2)
                assert np.array_equal(o1.download(), p1.download()) and np.array_equal(o2.download(), p2.download())
                mu2, ata = ctx.vecM(), ctx.vecM()
                ctx.cg_prepare(rhs, mu2, 2.0, 0.7, 6, ax, ata, 2, v2, p2)
                its2, _, _ = ctx.cg_solve_prepared(rhs, mu2, 2.0, 0.7, 6, 1, ax, ata, 2)
                assert its2 == its_ref and np.array_equal(p2.download(), o2.download())
                stage = ctx.vecM()
                assert rhs.upload_changed(v, stage) is False and rhs.upload_changed(-v, stage) is True
                ctx.snapshot_begin(rhs, M, 9)
                assert np.array_equal(ctx.snapshot_wait(9), -v)
        print(f"ok twin={twin} stripes={stripes} miss={miss}/{form}", flush=True)
print("sanitize_small: all checks passed")
