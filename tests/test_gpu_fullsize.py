"""Parity at BASELINE.json's FULL per-GPU sizes (configs 2, 3, 4-shard, the 160 GB single-GPU cut of config 4, the 105 GB shard of
config 5 with 1 % missing), where the oracle cannot
sweep the whole matrix: the matrix is generated directly in HBM by the counter-based generator, which the oracle can
regenerate marker by marker (gvamp_oracle.c:orc_synth_bed is byte-identical), so

  * the bytes, genotype counts (bit-exact) and statistics of a random sample of markers are checked against the oracle,
  * X^T.u is checked on the sampled markers against the oracle's dot products,
  * X.v is checked EXACTLY against the oracle for vectors v supported on the sampled markers (all N outputs),
  * and the size-independent properties of the operator pair tie the rest of the matrix to those samples:
    adjointness <Xv, u> = <v, X^T u> for dense random v, u, linearity, a checksum of the counts
    (n00 + n01 + n10 + n11 = N for every marker) and bit-reproducibility of the fixed-point sweeps.

Tolerances as in test_gpu_kernels.py (north_star): integers bit-exact, mat-vecs 1e-6 norm-wise."""
import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu
TOL_MATVEC = 1e-6
SEED = 4242


@pytest.fixture(scope="module")
def C():
    from gvamp_b200 import capi
    capi.load()
    assert capi.device_count() > 0
    return capi


CASES = [
    pytest.param(100_000, 500_000, 0.0, id="config2-100kx500k"),
    pytest.param(400_000, 275_000, 0.0, id="config4-shard-400kx275k"),
    pytest.param(400_000, 120_000, 0.01, id="config5-like-400kx120k-1pct-missing"),
    pytest.param(200_000, 1_000_000, 0.0, id="config3-200kx1M"),
    # the two shards that do not leave room for a whole individual-major twin: X.v runs on a partial twin + the gathering kernel
    pytest.param(400_000, 1_050_000, 0.01, id="config5-shard-400kx1.05M-1pct-missing-105GB"),
    pytest.param(400_000, 1_600_000, 0.0, id="config4-single-gpu-cut-400kx1.6M-160GB"),
]


@pytest.mark.parametrize("N,M,miss", CASES)
def test_full_size_properties(C, oracle, N, M, miss, monkeypatch):
    monkeypatch.delenv("GVB_KERNELS", raising=False)
    rng = np.random.default_rng(N + M)
    sample = np.sort(rng.choice(M, size=48, replace=False))
    sample[0], sample[-1] = 0, M - 1                      # first and last marker of the shard
    cols = np.concatenate([oracle.synth_bed(SEED, int(j), 1, N, miss_rate=miss) for j in sample])
    ds = oracle.Dataset(cols, N)                           # the sampled columns as their own little dataset
    u = rng.normal(size=N)
    v = rng.normal(size=M) * np.where(rng.random(M) < 0.02, 20.0, 1.0)
    v_sparse = np.zeros(M)
    v_sparse[sample] = rng.normal(size=len(sample))
    with C.Context(0) as ctx:
        ctx.synth(SEED, N, M, 0, M, miss)
        ctx.compute_stats(1.0)
        # --- bytes, counts, statistics of the sampled markers: bit-exact / 1e-12
        for j, col in zip(sample, cols):
            assert np.array_equal(ctx.decode(int(j), 1)[0], col), j
        counts = ctx.counts()
        assert np.array_equal(counts[sample], ds.counts())
        assert np.all(counts[:, :4].sum(axis=1) == N) and np.all(counts[:, 4:].sum(axis=1) == N)      # checksum over ALL markers
        if miss == 0.0:
            assert not counts[:, 1].any()
        else:
            assert abs(counts[:, 1].mean() / N - miss) < 0.05 * miss
        mave, msig = ctx.stats()
        # the oracle sums N squared deviations like the reference; the closed form from the exact counts differs by its rounding
        assert relerr(mave[sample], ds.mave) < 1e-13 and relerr(msig[sample], ds.msig) < 1e-11
        # --- products
        atx = ctx.ATx(u)
        ax = ctx.Ax(v)
        ax_sparse = ctx.Ax(v_sparse)
        ax_again, atx_again = ctx.Ax(v), ctx.ATx(u)
        ax_scaled = ctx.Ax(2.5 * v)
    # X^T.u on the sample against the oracle; the error scale is the typical size of an output (norm-wise tolerance)
    ref = ds.ATx(u)
    assert np.linalg.norm(atx[sample] - ref) / np.linalg.norm(ref) < TOL_MATVEC
    # X.v for v supported on the sample: every one of the N outputs against the oracle
    assert relerr(ax_sparse, ds.Ax(v_sparse[sample])) < TOL_MATVEC
    assert np.all(ax[N:] == 0)
    # adjointness ties the unsampled part of the matrix to the checked one
    lhs, rhs = float(np.dot(ax[:N], u)), float(np.dot(v, atx))
    assert abs(lhs - rhs) <= TOL_MATVEC * np.linalg.norm(ax) * np.linalg.norm(u)
    assert relerr(ax_scaled, 2.5 * ax) < TOL_MATVEC
    # a standardised column has unit variance: ||X^T u||^2 / M ~ ||u||^2 / N * (N-1)/N ... only a sanity band
    assert 0.5 < (atx @ atx) / M / (u @ u / N) < 2.0
    # fixed point: bit-reproducible
    assert np.array_equal(ax, ax_again) and np.array_equal(atx, atx_again)
