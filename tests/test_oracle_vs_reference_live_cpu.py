"""The oracle against the LIVE reference (oracle/_ref/libgvamp_ref.so = the unmodified data / vamp classes of the reference behind
oracle/ref_harness.cpp) on seeded random shapes: ragged N and M, shards of a larger matrix, missing genotypes, phenotype NAs,
alpha_scale, byte sub-ranges.  The committed golden vectors (tests/golden) pin a handful of shapes; this sweep pins the corner
cases wherever the compiled reference exists (the build container; the GPU box when the snapshot carried oracle/_ref).
CPU only.  Skipped where oracle/_ref is absent."""
import numpy as np
import pytest

from conftest import relerr


@pytest.fixture(scope="module")
def R():
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref/libgvamp_ref.so is not present on this machine")
    return ref


# (N, M, Mt, S, miss_rate, phenotype-NA rate, alpha_scale)
CASES = [
    (5, 1, 1, 0, 0.0, 0.0, 1.0),          # smaller than one byte column + one marker
    (8, 3, 3, 0, 0.2, 0.0, 1.0),
    (63, 17, 40, 11, 0.05, 0.1, 1.0),     # N % 4 == 3, a shard in the middle of the matrix
    (130, 64, 64, 0, 0.0, 0.0, 0.3),      # alpha_scale != 1 (data.cpp:479-483)
    (257, 9, 100, 91, 0.3, 0.25, 1.0),    # heavy missingness and many phenotype NAs, last shard
    (1001, 33, 33, 0, 0.01, 0.02, 1.0),
    (1024, 5, 5, 0, 0.0, 0.5, 0.7),       # N % 4 == 0, half of the phenotypes missing
]


@pytest.mark.parametrize("N,M,Mt,S,miss,na_rate,alpha", CASES)
def test_data_class_against_the_live_reference(oracle, R, tmp_path, N, M, Mt, S, miss, na_rate, alpha):
    rng = np.random.default_rng(N * 1000 + M)
    bed_all = oracle.synth_bed(N + M, 0, Mt, N, miss_rate=miss)
    bedp, phenp = str(tmp_path / "a.bed"), str(tmp_path / "a.phen")
    oracle.write_bed(bedp, bed_all)
    y = rng.normal(size=N) * 3.0 + 1.5
    na_idx = np.flatnonzero(rng.random(N) < na_rate)
    if len(na_idx) > N - 3:
        na_idx = na_idx[: N - 3]
    oracle.write_phen(phenp, y, na_idx=na_idx)

    rd = R.RefData(bedp, N, M, Mt=Mt, S=S, phen_path=phenp, alpha_scale=alpha)
    try:
        # phenotype reader: mask, NA count, intercept / scale, filtered phenotype (data.cpp:136-186, 1065-1079)
        phen, mask4, nonas, avg, sqn = oracle.read_phen(phenp, N)
        assert np.array_equal(mask4, rd.mask4()) and nonas == rd.nonas()
        ref_avg, ref_sqn = rd.intercept_scale()
        # the sum of squares is `sqn += d * d` in the reference (data.cpp:176): its -march=native build contracts that into an FMA,
        # numpy rounds the product first, so the scale may differ in the last bit
        assert avg == ref_avg and abs(sqn - ref_sqn) <= 4e-16 * abs(ref_sqn)
        ds = oracle.Dataset(bed_all[S:S + M], N, phen=phen, mask4=mask4, nonas=nonas, alpha_scale=alpha)
        assert relerr(ds.filter_pheno(), rd.filter_pheno()) < 1e-15
        # marker statistics: same scalar loop order => bit-identical (data.cpp:447-485)
        mave, msig = rd.stats()
        assert np.array_equal(ds.mave, mave) and np.array_equal(ds.msig, msig)
        # X.v: bit-identical incl. zeros at NA / padded individuals; X^T.u: OpenMP reduction order differs
        v, u = rng.normal(size=M), rng.normal(size=4 * ((N + 3) // 4))
        u[N:] = 0.0
        assert np.array_equal(ds.Ax(v), rd.Ax(v))
        assert relerr(ds.ATx(u), rd.ATx(u)) < 1e-13
        mbytes = (N + 3) // 4
        if mbytes >= 3:   # byte sub-range [SB, SB+LB) with its own 1/sqrt(4 LB) scaling (data.cpp:998-1005)
            SB, LB = 1, mbytes - 2
            assert np.array_equal(ds.Ax(v, SB, LB), rd.Ax(v, SB, LB))
            assert relerr(ds.ATx(u[4 * SB:4 * (SB + LB)], SB, LB), rd.ATx(u[4 * SB:4 * (SB + LB)], SB, LB)) < 1e-13
    finally:
        rd.close()


# (number of mixture components, gam1, scale of r1)
DENOISER_CASES = [(2, 1e-6, 1.0), (3, 0.8, 0.05), (5, 40.0, 0.3), (23, 2.0e3, 0.01), (4, 5.0e10, 1.0)]   # last: the |1/gam1| < 1e-10 shortcut


@pytest.mark.parametrize("L,gam1,rscale", DENOISER_CASES)
def test_denoiser_and_em_against_the_live_reference(oracle, R, L, gam1, rscale):
    """vamp::g1 / g1d (vamp.cpp:805-869) and one vamp::updatePrior call (vamp.cpp:929-1072) on random priors, from the
    near-cancelling regime of iteration 1 (gam1 = 1e-6) to the 1/gam1 < 1e-10 shortcut."""
    rng = np.random.default_rng(L * 7 + int(abs(np.log10(gam1))))
    M = 4001
    probs = rng.random(L) + 0.05
    probs[0] += L          # a spike-and-slab prior: most of the mass on the null component
    probs /= probs.sum()
    vars_ = np.concatenate(([0.0], np.sort(10.0 ** rng.uniform(-5, 0, size=L - 1))))
    r1 = rng.normal(size=M) * rscale * np.where(rng.random(M) < 0.03, 20.0, 1.0)
    rv = R.RefVamp(2000, M, M, probs, vars_, gam1=gam1, EM_max_iter=3, EM_err_thr=1e-4)
    try:
        g, gd = rv.g1(r1, gam1)
        # g1 = r + s * pkd / pk cancels for tiny gam1: measure against the input scale like the golden test does
        assert np.linalg.norm(oracle.g1(r1, gam1, probs, vars_) - g) / np.linalg.norm(r1) < 1e-13
        assert np.max(np.abs(oracle.g1d(r1, gam1, probs, vars_) - gd)) < 1e-11 * max(1.0, np.max(np.abs(gd)))
        if gam1 < 1e9:
            rp, rvv = rv.update_prior(r1, gam1)
            p, v = oracle.update_prior(r1, gam1, probs, vars_, M, 3, 1e-4)
            assert len(p) == len(rp)          # the same components merged (vamp.cpp:1054-1071)
            assert relerr(p, rp) < 1e-10 and relerr(v, rvv) < 1e-10
    finally:
        rv.close()


@pytest.mark.parametrize("N,M,miss,tau,gam2", [(400, 150, 0.0, 2.0, 0.7), (203, 301, 0.05, 0.4, 30.0), (64, 9, 0.0, 10.0, 1e-3)])
def test_cg_and_onsager_against_the_live_reference(oracle, R, tmp_path, N, M, miss, tau, gam2):
    """vamp::lmmse_mult / precondCG_solver / g2d_onsager (vamp.cpp:871-889, 1074-1229) incl. a warm start, M > N and M < N."""
    rng = np.random.default_rng(N + M)
    bed = oracle.synth_bed(N * M, 0, M, N, miss_rate=miss)
    bedp = str(tmp_path / "c.bed")
    oracle.write_bed(bedp, bed)
    ds = oracle.Dataset(bed, N)
    rd = R.RefData(bedp, N, M, y=np.zeros(N))
    K = 25
    rv = R.RefVamp(N, M, M, [0.9, 0.1], [0.0, 0.1], CG_max_iter=K, seed=5)
    try:
        rv.set_state(1.0, gam2, tau)
        rhs = rng.normal(size=M)
        assert relerr(oracle.lmmse_mult(ds, rhs, tau, gam2), rv.lmmse_mult(rd, rhs, tau)) < 1e-12
        mu, _ = oracle.precond_cg(ds, rhs, np.zeros(M), tau, gam2, K, 1)
        mu_ref = rv.cg(rd, rhs, np.zeros(M), tau, 1)
        assert relerr(mu, mu_ref) < 1e-9
        mu_w, _ = oracle.precond_cg(ds, rhs, 0.9 * mu_ref, tau, gam2, K, 1)
        assert relerr(mu_w, rv.cg(rd, rhs, 0.9 * mu_ref, tau, 1)) < 1e-9
        alpha2, bern, invq = rv.onsager(rd, gam2, tau)
        probe = oracle.bernoulli_probe(5, 0, M, M)
        assert np.array_equal(probe, bern)      # mt19937{seed + S} + bernoulli_distribution: the identical stream
        q, _ = oracle.precond_cg(ds, probe, np.zeros(M), tau, gam2, K, 0)
        assert relerr(q, invq) < 1e-9 and abs(gam2 * probe.dot(q) / alpha2 - 1) < 1e-10
    finally:
        rv.close()
        rd.close()


def test_probit_pieces_against_the_live_reference(oracle, R):
    """erfcx (utilities.cpp:345-409) over its three branches and g1_bin_class / g1d_bin_class (vamp_probit.cpp:661-726)."""
    x = np.concatenate((np.linspace(-30, 30, 601), [-1e3, 1e3, 0.0, 50.0, -26.7]))
    ref = np.asarray(R.erfcx(x), dtype=float).reshape(-1)
    mine = oracle.erfcx(x)
    fin = np.isfinite(ref)
    rel = np.abs(mine[fin] / ref[fin] - 1)
    # for x << 0 erfcx ~ 2 exp(x^2): one ulp of x^2 (up to 900) is amplified by the exponential in both implementations
    assert np.max(rel[np.abs(x[fin]) <= 5]) < 1e-14 and np.max(rel) < 1e-12 and np.array_equal(np.isinf(mine), np.isinf(ref))
    rng = np.random.default_rng(3)
    n = 3000
    p, y, mcov = rng.normal(size=n) * 2.0, (rng.random(n) < 0.4).astype(float), rng.normal(size=n) * 0.5
    rv = R.RefVamp(n, 10, 10, [0.9, 0.1], [0.0, 0.1], model="bin_class")
    try:
        for tau1 in (1e-3, 0.7, 25.0):
            g, gd = rv.g1_bin_class(p, tau1, y, mcov)
            assert relerr(oracle.g1_bin_class(p, tau1, y, mcov), g) < 1e-13
            assert relerr(oracle.g1d_bin_class(p, tau1, y, mcov), gd) < 1e-12
    finally:
        rv.close()


def test_sharded_probe_against_the_live_reference(oracle, R, tmp_path):
    """The emulated R-rank probe (oracle.sharded_probe) is, shard by shard, what the reference's g2d_onsager draws when it
    is built on shard (S_r, M_r) of the matrix: mt19937{seed + S_r}, M_r draws, scaled by 1/sqrt(Mt) (vamp.cpp:875-882)."""
    N, Mt, seed, nranks = 120, 301, 9, 3
    bed = oracle.synth_bed(4, 0, Mt, N)
    bedp = str(tmp_path / "s.bed")
    oracle.write_bed(bedp, bed)
    whole = oracle.sharded_probe(seed, 0, Mt, Mt, nranks)
    assert whole.shape == (Mt,) and np.array_equal(oracle.sharded_probe(seed, 0, Mt, Mt, 1), oracle.bernoulli_probe(seed, 0, Mt, Mt))
    for r in range(nranks):
        Mr, Sr = oracle.divide_work(Mt, nranks, r)
        rd = R.RefData(bedp, N, Mr, Mt=Mt, S=Sr, y=np.zeros(N))
        rv = R.RefVamp(N, Mr, Mt, [0.9, 0.1], [0.0, 0.1], CG_max_iter=2, seed=seed)
        try:
            rv.set_state(1.0, 0.5, 2.0)
            _, bern, _ = rv.onsager(rd, 0.5, 2.0)
            assert np.array_equal(bern, whole[Sr:Sr + Mr]), r
        finally:
            rv.close()
            rd.close()
