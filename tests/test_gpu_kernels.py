"""Parity of the CUDA path (through the C ABI) against the oracle and the golden vectors.
Integer / byte work must be bit-exact; floating point tolerances are written at each assert.
The matvec tolerance is the north_star's 1e-6, measured norm-wise (||out-ref|| / ||ref||)."""
import hashlib
import math
import os

import numpy as np
import pytest

from conftest import golden, relerr

pytestmark = pytest.mark.gpu

TOL_MATVEC = 1e-6     # BASELINE.json north_star: "Matvec outputs must agree within 1e-6 relative"
KERNEL_GENS = ("simple", "lut1", "lut")     # FP64 cross-check kernels, gen-1 table kernels, gen-2 tile kernels (default)


@pytest.fixture(scope="module")
def C():
    from gvamp_b200 import capi
    capi.load()
    assert capi.device_count() > 0, "no CUDA device: -m gpu tests need the B200 box"
    return capi


def make_ctx(C, gen):
    """"lut": the product library.  "simple" / "lut1": the on-device cross-check kernels, which exist only in the test build of the
    library (gvamp_b200/lib/xcheck/libgvamp_b200.so)."""
    os.environ["GVB_KERNELS"] = gen
    return C.Context(0, xcheck=gen in ("simple", "lut1"))


@pytest.fixture(scope="module")
def case_na(oracle):
    """N % 4 != 0, 2 % missing genotypes, phenotype NAs (the golden matvec case)."""
    g = golden("matvec_n1003.npz")
    N, M = int(g["N"]), int(g["M"])
    bed = oracle.synth_bed(int(g["seed"]), 0, M, N, miss_rate=float(g["miss_rate"]))
    present = np.ones(N, bool)
    present[g["na_idx"]] = False
    mask4 = oracle.make_mask4(N, present)
    ds = oracle.Dataset(bed, N, mask4=mask4, nonas=int(g["nonas"]))
    return g, bed, mask4, ds


def test_layout_roundtrip_bit_exact(C, oracle):
    """PLINK bytes -> striped interleaved HBM layout -> PLINK bytes, for ragged shapes."""
    for N, M in ((1003, 401), (128, 4), (5, 1), (4099, 67), (640, 130)):
        bed = oracle.synth_bed(7, 0, M, N, miss_rate=0.05)
        with C.Context(0) as ctx:
            ctx.load_host(bed, N)
            assert np.array_equal(ctx.decode(0, M), bed), (N, M)
            if M > 3:
                assert np.array_equal(ctx.decode(2, M - 3), bed[2:M - 1])


def test_load_file_shard(C, oracle, tmp_path):
    N, Mt = 1003, 400
    bed = oracle.synth_bed(11, 0, Mt, N, miss_rate=0.02)
    path = str(tmp_path / "x.bed")
    oracle.write_bed(path, bed)
    M, S = C.divide_work(Mt, 3, 1)
    with C.Context(0) as ctx:
        ctx.load_file(path, N, Mt, S, M)
        assert np.array_equal(ctx.decode(0, M), bed[S:S + M])
    with C.Context(0) as ctx:
        with pytest.raises(C.GvbError):
            ctx.load_file(str(tmp_path / "nope.bed"), N, Mt, S, M)


def test_synth_matches_oracle_bytes(C, oracle):
    for N, Mt, S, M, miss in ((1003, 900, 300, 257, 0.0), (2048, 64, 0, 64, 0.01), (777, 5000, 4100, 33, 0.3)):
        with C.Context(0) as ctx:
            ctx.synth(5, N, Mt, S, M, miss)
            got = ctx.decode(0, M)
        ref = oracle.synth_bed(5, S, M, N, miss_rate=miss)
        assert np.array_equal(got, ref), (N, Mt, S, M)


def test_counts_and_stats(C, case_na):
    g, bed, mask4, ds = case_na
    with C.Context(0) as ctx:
        ctx.load_host(bed, ds.N).set_mask(mask4, ds.nonas).compute_stats(1.0)
        assert np.array_equal(ctx.counts(), ds.counts())          # integers: bit-exact
        mave, msig = ctx.stats()
    assert relerr(mave, g["mave"]) < 1e-14 and relerr(msig, g["msig"]) < 1e-12
    assert np.max(np.abs(mave - g["mave"])) < 1e-14


def test_stats_alpha_scale_and_degenerate_columns(C, oracle):
    N, M = 403, 12
    bed = oracle.synth_bed(3, 0, M, N)
    bed[0, :] = 0x55          # all missing   -> mean 0, inverse sd 1 (data.cpp:462-483)
    bed[1, :] = 0xFF          # constant 0    -> inverse sd 1
    bed[2, :] = 0x00          # constant 2
    ds = oracle.Dataset(bed, N, alpha_scale=0.3)
    with C.Context(0) as ctx:
        ctx.load_host(bed, N).compute_stats(0.3)
        mave, msig = ctx.stats()
    assert relerr(mave, ds.mave) < 1e-14 and relerr(msig, ds.msig) < 1e-12
    assert mave[0] == 0 and msig[0] == 1 and msig[1] == 1 and msig[2] == 1


@pytest.mark.parametrize("gen", KERNEL_GENS)
def test_ax_atx_golden(C, case_na, gen):
    g, bed, mask4, ds = case_na
    N = ds.N
    with make_ctx(C, gen) as ctx:
        ctx.load_host(bed, N).set_mask(mask4, ds.nonas).compute_stats(1.0)
        ax = ctx.Ax(g["v"])
        atx = ctx.ATx(g["u"])
        SB, LB = int(g["SB"]), int(g["LB"])
        ax_sub = ctx.Ax(g["v"], SB, LB)
        atx_sub = ctx.ATx(g["u"][:4 * LB], SB, LB)
    assert relerr(ax, g["Ax"]) < TOL_MATVEC and relerr(atx, g["ATx"]) < TOL_MATVEC
    assert relerr(ax_sub, g["Ax_sub"]) < TOL_MATVEC and relerr(atx_sub, g["ATx_sub"]) < TOL_MATVEC
    assert np.all(ax[N:] == 0) and np.all(ax[g["na_idx"]] == 0)      # pads and phenotype NAs are exactly zero
    if gen == "simple":                                               # FP64 kernels: only the summation order differs
        assert relerr(ax, g["Ax"]) < 1e-13 and relerr(atx, g["ATx"]) < 1e-13


@pytest.mark.parametrize("gen", KERNEL_GENS)
@pytest.mark.parametrize("N,M,miss", [(4096, 1024, 0.0), (1000, 2000, 0.0), (2500, 333, 0.01), (129, 5, 0.2), (20000, 4000, 0.0)])
def test_ax_atx_vs_oracle(C, oracle, gen, N, M, miss):
    bed = oracle.synth_bed(17, 0, M, N, miss_rate=miss)
    ds = oracle.Dataset(bed, N)
    rng = np.random.default_rng(N + M)
    v = rng.normal(size=M) * np.where(rng.random(M) < 0.05, 30.0, 1.0)     # heavy-tailed like a sparse effect vector
    u = rng.normal(size=N)
    with make_ctx(C, gen) as ctx:
        ctx.load_host(bed, N).compute_stats(1.0)
        ax, atx = ctx.Ax(v), ctx.ATx(u)
        # linearity and adjointness are size-independent properties of the operator pair
        ax2 = ctx.Ax(2.5 * v)
        lhs = float(np.dot(ax[:N], u))
        rhs = float(np.dot(v, atx))
    assert relerr(ax, ds.Ax(v)) < TOL_MATVEC
    assert relerr(atx, ds.ATx(u)) < TOL_MATVEC
    assert relerr(ax2, 2.5 * ax) < TOL_MATVEC
    assert abs(lhs - rhs) <= 1e-6 * (np.linalg.norm(ax) * np.linalg.norm(u))


@pytest.mark.parametrize("staging,shape", [("tma", 0), ("tma", 1), ("tma", 2), ("cpasync", 0), ("cpasync", 1)])
@pytest.mark.parametrize("tpc,spc", [(3, 5), (2, 7), (64, 64)])
def test_tile_kernels_many_work_items(C, oracle, staging, shape, tpc, spc, monkeypatch):
    """The tile kernels with small work items: several marker chunks / stripe chunks per CTA, partial last chunks (odd numbers of steps:
    a half-filled table pair), inactive warps (stripes and marker tiles that are not multiples of the CTA shape), in both table staging
    modes and every CTA shape: pair mode (tables by TMA bulk copy) with 11x2 / 12x2 / 7x3 consumer warps x bed stages, producer-warp
    cp.async with 15x2 / 12x3; X.v on the twin and (second context) gathering from the one matrix."""
    monkeypatch.setenv("GVB_TAB", staging)
    monkeypatch.setenv("GVB_PAIR_SHAPE" if staging == "tma" else "GVB_TILE_VARIANT", str(shape))
    monkeypatch.setenv("GVB_AX_TPC", str(tpc))
    monkeypatch.setenv("GVB_ATX_SPC", str(spc))
    N, M = 6100, 9000          # 48 stripes (3 CTA rows of 16), 71 marker tiles
    bed = oracle.synth_bed(23, 0, M, N)
    ds = oracle.Dataset(bed, N)
    rng = np.random.default_rng(1)
    v, u = rng.normal(size=M), rng.normal(size=N)
    with make_ctx(C, "lut") as ctx:
        ctx.load_host(bed, N).compute_stats(1.0)
        ax, atx = ctx.Ax(v), ctx.ATx(u)
        ax_again, atx_again = ctx.Ax(v), ctx.ATx(u)
    monkeypatch.setenv("GVB_TWIN", "0")
    with make_ctx(C, "lut") as ctx:
        ctx.load_host(bed, N).compute_stats(1.0)
        ax_gather = ctx.Ax(v)
    assert relerr(ax, ds.Ax(v)) < TOL_MATVEC and relerr(atx, ds.ATx(u)) < TOL_MATVEC
    # fixed-point accumulation: bit-reproducible, whatever order the work items ran in and whichever kernel walked the matrix
    assert np.array_equal(ax, ax_again) and np.array_equal(atx, atx_again) and np.array_equal(ax, ax_gather)


def test_tile_kernels_reproducible_across_chunking(C, oracle, monkeypatch):
    """integer accumulation => the result does not depend on the tiling of the work at all (bit-exact)."""
    N, M = 3000, 5000
    bed = oracle.synth_bed(29, 0, M, N)
    rng = np.random.default_rng(2)
    v, u = rng.normal(size=M), rng.normal(size=N)
    res = []
    for staging, shape, tpc, spc in (("tma", 1, 64, 64), ("tma", 0, 1, 1), ("cpasync", 0, 5, 3), ("cpasync", 1, 2, 2), ("tma", 2, 7, 9)):
        monkeypatch.setenv("GVB_TAB", staging)
        monkeypatch.setenv("GVB_PAIR_SHAPE" if staging == "tma" else "GVB_TILE_VARIANT", str(shape))
        monkeypatch.setenv("GVB_AX_TPC", str(tpc))
        monkeypatch.setenv("GVB_ATX_SPC", str(spc))
        with make_ctx(C, "lut") as ctx:
            ctx.load_host(bed, N).compute_stats(1.0)
            res.append((ctx.Ax(v), ctx.ATx(u)))
    for ax, atx in res[1:]:
        assert np.array_equal(ax, res[0][0]) and np.array_equal(atx, res[0][1])


@pytest.mark.parametrize("N,M,miss", [(70_001, 301, 0.02), (33_000, 130, 0.001), (2500, 333, 0.05)])
def test_missing_list_equals_second_walk(C, oracle, N, M, miss, monkeypatch):
    """X^T.u on a shard with missing genotypes: the sparse missing-genotype list (misslist.cu; several 32768-individual
    blocks, M % 4 != 0, padded segments) and the second table walk are two integer evaluations of the same sums, so
    they must agree BIT FOR BIT; both are within the mat-vec tolerance of the oracle."""
    bed = oracle.synth_bed(31, 0, M, N, miss_rate=miss)
    ds = oracle.Dataset(bed, N)
    u = np.random.default_rng(3).normal(size=N)
    res = {}
    for mode in ("list", "list_warp", "twopass"):
        monkeypatch.setenv("GVB_MISS", mode.split("_")[0])
        monkeypatch.setenv("GVB_MISS_SUM", "warp" if mode == "list_warp" else "lane")   # both forms of the gather kernel
        with make_ctx(C, "lut") as ctx:
            ctx.load_host(bed, N).compute_stats(1.0)
            res[mode] = (ctx.ATx(u), ctx.ATx(u), ctx.missing_list_entries())
    assert res["list"][2] > 0 and res["twopass"][2] == 0          # the list was really built / really bypassed
    assert np.array_equal(res["list"][0], res["list_warp"][0])
    assert res["list"][2] % 4 == 0 and res["list"][2] >= ds.counts()[:, 5].sum()
    assert np.array_equal(res["list"][0], res["twopass"][0]) and np.array_equal(res["list"][0], res["list"][1])
    assert relerr(res["list"][0], ds.ATx(u)) < TOL_MATVEC


@pytest.mark.parametrize("N,M,miss", [(70_001, 301, 0.02), (4099, 1030, 0.0), (2500, 333, 0.05)])
def test_twin_layout_is_bit_identical(C, oracle, N, M, miss, monkeypatch):
    """X.v on the individual-major twin of the matrix (twin.cu: byte k of a word = the table index of individual k) and
    X.v on the one matrix (index gathered with mask + multiply) walk the same tables and add the same integers: the
    results must agree BIT FOR BIT, with phenotype NAs, ragged N / M and missing genotypes; people statistics (the same
    walk with other tables) likewise.  Both are within the mat-vec tolerance of the oracle."""
    bed = oracle.synth_bed(37, 0, M, N, miss_rate=miss)
    present = np.ones(N, bool)
    present[np.random.default_rng(5).choice(N, N // 50, replace=False)] = False
    mask4 = oracle.make_mask4(N, present)
    ds = oracle.Dataset(bed, N, mask4=mask4, nonas=int(present.sum()))
    v = np.random.default_rng(4).normal(size=M)
    res = {}
    n_stripes = ((N + 3) // 4 + 31) // 32
    for mode in ("0", "1", "partial"):
        if mode == "partial":                                        # a twin for the first third of the stripes only: two launches per X.v
            monkeypatch.delenv("GVB_TWIN")
            monkeypatch.setenv("GVB_TWIN_STRIPES", str(n_stripes // 3))
        else:
            monkeypatch.setenv("GVB_TWIN", mode)
        with make_ctx(C, "lut") as ctx:
            ctx.load_host(bed, N).set_mask(mask4, int(present.sum())).compute_stats(1.0)
            assert ctx.twin_state() == 0                             # decided by the first X.v
            res[mode] = (ctx.Ax(v), ctx.Ax(2.5 * v), ctx.twin_state(), ctx.twin_stripes())
            if mode == "1":                                          # giving the twin back (what a failed allocation triggers) changes nothing
                ctx.twin_release()
                assert ctx.twin_state() == -1 and ctx.twin_stripes() == 0 and np.array_equal(ctx.Ax(v), res[mode][0])
    monkeypatch.delenv("GVB_TWIN_STRIPES")
    assert res["0"][2] == 0 and res["1"][2] == 1                     # really bypassed / really built
    assert res["partial"][2] == 2 and res["partial"][3] == n_stripes // 3 and res["1"][3] == n_stripes
    assert np.array_equal(res["0"][0], res["1"][0]) and np.array_equal(res["0"][1], res["1"][1])
    assert np.array_equal(res["0"][0], res["partial"][0]) and np.array_equal(res["0"][1], res["partial"][1])
    assert relerr(res["1"][0], ds.Ax(v)) < TOL_MATVEC


@pytest.mark.parametrize("N,M,miss", [(70_001, 301, 0.02), (4099, 1030, 0.0), (2500, 333, 0.05), (1200, 9000, 0.0)])
def test_dual_ax_equals_two_sweeps(C, oracle, N, M, miss, monkeypatch):
    """gvb_dAx2: two products from one pass over the bed (tables interleaved [entry][slot][rhs], one LDS.64 per lookup).  Every int32
    window, shift and int64 sum is the one the single-product kernel forms, so both outputs must equal two gvb_dAx calls BIT FOR BIT -
    on the twin, on the one matrix (gathered indices), on a partial twin, with vectors of very different magnitude (separate scale
    classes per right-hand side), a zero vector and a non-finite entry (NaN stays with its own product)."""
    bed = oracle.synth_bed(41, 0, M, N, miss_rate=miss)
    present = np.ones(N, bool)
    present[np.random.default_rng(6).choice(N, N // 40, replace=False)] = False
    mask4 = oracle.make_mask4(N, present)
    rng = np.random.default_rng(8)
    v0 = rng.normal(size=M)
    v1 = rng.normal(size=M) * np.where(rng.random(M) < 0.01, 1e6, 1e-3)     # outliers: other scale classes than v0
    n_stripes = ((N + 3) // 4 + 31) // 32
    for mode in ("0", "1", "partial"):
        if mode == "partial":
            monkeypatch.delenv("GVB_TWIN")
            monkeypatch.setenv("GVB_TWIN_STRIPES", str(max(1, n_stripes // 3)))
        else:
            monkeypatch.setenv("GVB_TWIN", mode)
        with C.Context(0) as ctx:
            ctx.load_host(bed, N).set_mask(mask4, int(present.sum())).compute_stats(1.0)
            a, b, oa, ob, pa, pb = ctx.vecM(v0), ctx.vecM(v1), ctx.vecN(), ctx.vecN(), ctx.vecN(), ctx.vecN()
            for x, y in ((v0, v1), (v1, v0), (v0, np.zeros(M)), (v0, v0)):
                a.upload(x), b.upload(y)
                ctx.dAx(a, oa), ctx.dAx(b, ob)
                s0, d0 = ctx.sweeps(), ctx.dual_sweeps()
                ctx.dAx2(a, b, pa, pb)
                assert ctx.sweeps() == s0 + 1 and ctx.dual_sweeps() == d0 + 1
                assert np.array_equal(oa.download(), pa.download()) and np.array_equal(ob.download(), pb.download())
            bad = v1.copy()
            bad[M // 2] = np.nan
            a.upload(v0), b.upload(bad)
            ctx.dAx(a, oa)
            ctx.dAx2(a, b, pa, pb)
            assert np.array_equal(oa.download(), pa.download())
            assert np.all(np.isnan(pb.download()[:N][present]))
            ctx.dAx2(a, b, pa, pb)                                         # and the flags re-arm
            assert np.array_equal(oa.download(), pa.download())
            assert ctx.twin_state() == {"0": 0, "1": 1, "partial": 2}[mode]
    monkeypatch.delenv("GVB_TWIN_STRIPES")


@pytest.mark.parametrize("N,M,miss", [(70_001, 301, 0.0), (4099, 1030, 0.0), (1200, 9000, 0.0), (2500, 333, 0.05)])
def test_dual_atx_equals_two_sweeps(C, oracle, N, M, miss):
    """gvb_dATx2: X^T.u0 and X^T.u1 from one pass over the bed, bit for bit two gvb_dATx calls - with phenotype NAs, ragged shapes,
    vectors in different scale classes, a zero vector, a non-finite entry; on a shard with missing genotypes the call is two single
    sweeps (and says so in its counters)."""
    bed = oracle.synth_bed(43, 0, M, N, miss_rate=miss)
    present = np.ones(N, bool)
    present[np.random.default_rng(6).choice(N, N // 40, replace=False)] = False
    mask4 = oracle.make_mask4(N, present)
    rng = np.random.default_rng(9)
    u0 = rng.normal(size=N) * present
    u1 = rng.normal(size=N) * np.where(rng.random(N) < 0.01, 1e6, 1e-3) * present
    with C.Context(0) as ctx:
        ctx.load_host(bed, N).set_mask(mask4, int(present.sum())).compute_stats(1.0)
        a, b, oa, ob, pa, pb = ctx.vecN(u0), ctx.vecN(u1), ctx.vecM(), ctx.vecM(), ctx.vecM(), ctx.vecM()
        for x, y in ((u0, u1), (u1, u0), (u0, np.zeros(N)), (u0, u0)):
            a.upload(x), b.upload(y)
            ctx.dATx(a, oa), ctx.dATx(b, ob)
            s0, d0 = ctx.sweeps(), ctx.dual_sweeps()
            ctx.dATx2(a, b, pa, pb)
            assert (ctx.sweeps() - s0, ctx.dual_sweeps() - d0) == ((1, 1) if miss == 0.0 else (2, 0))
            assert np.array_equal(oa.download(), pa.download()) and np.array_equal(ob.download(), pb.download())
        bad = u1.copy()
        bad[N // 2] = np.inf
        a.upload(u0), b.upload(bad)
        ctx.dATx(a, oa)
        ctx.dATx2(a, b, pa, pb)
        assert np.array_equal(oa.download(), pa.download()) and np.all(np.isnan(pb.download()[:M]))
        ctx.dATx2(a, b, pa, pb)
        assert np.array_equal(oa.download(), pa.download())


def test_dual_buffers_follow_a_reload(C, oracle):
    """The second product's tables are sized by the marker tiles, its accumulators by the individuals: a context that reloads a
    matrix with fewer markers but more individuals (and the other way round) must resize both."""
    rng = np.random.default_rng(3)
    with C.Context(0) as ctx:
        for N, M in ((1200, 9000), (9000, 1200), (20_000, 700)):
            bed = oracle.synth_bed(5, 0, M, N)
            ctx.load_host(bed, N).compute_stats(1.0)
            a, b, oa, ob, pa, pb = ctx.vecM(rng.normal(size=M)), ctx.vecM(rng.normal(size=M)), ctx.vecN(), ctx.vecN(), ctx.vecN(), ctx.vecN()
            ctx.dAx(a, oa), ctx.dAx(b, ob)
            ctx.dAx2(a, b, pa, pb)
            assert np.array_equal(oa.download(), pa.download()) and np.array_equal(ob.download(), pb.download())
            mu, ax, ata, z = ctx.vecM(), ctx.vecN(), ctx.vecM(), ctx.vecN()
            ctx.cg_prepare(a, mu, 1.5, 0.9, 5, ax, ata, 2, b, z)
            its, _, _ = ctx.cg_solve_prepared(a, mu, 1.5, 0.9, 5, 1, ax, ata, 2)
            assert its >= 1 and np.array_equal(z.download(), ob.download())


def test_nonfinite_inputs_turn_the_outputs_into_nan(C, oracle):
    """A NaN / infinity in the input vector: the reference's FP64 LUT products (0 * NaN, data.cpp:766 / :975) turn EVERY
    output into NaN.  The fixed-point sweeps cannot carry a NaN through their integer sums, so they flag it (bound kernel ->
    finish kernel): all markers NaN for X^T.u; all present individuals NaN, masked and padded ones 0 for X.v; and the next
    sweep on finite data is clean again (nothing sticks in padded entries or scales)."""
    N, M = 1003, 517
    bed = oracle.synth_bed(41, 0, M, N, miss_rate=0.01)
    present = np.ones(N, bool)
    present[[3, 500, 1002]] = False
    mask4 = oracle.make_mask4(N, present)
    ds = oracle.Dataset(bed, N, mask4=mask4, nonas=int(present.sum()))
    rng = np.random.default_rng(9)
    v, u = rng.normal(size=M), rng.normal(size=N)
    with make_ctx(C, "lut") as ctx:
        ctx.load_host(bed, N).set_mask(mask4, int(present.sum())).compute_stats(1.0)
        for bad in (np.nan, np.inf, -np.inf):
            ub, vb = u.copy(), v.copy()
            ub[77], vb[M - 1] = bad, bad
            atx, ax = ctx.ATx(ub), ctx.Ax(vb)
            assert np.isnan(atx).all() and atx.shape == (M,)
            assert np.isnan(ax[:N][present]).all() and not ax[:N][~present].any() and not ax[N:].any()
            assert relerr(ctx.ATx(u), ds.ATx(u)) < TOL_MATVEC and relerr(ctx.Ax(v), ds.Ax(v)) < TOL_MATVEC
        # and through the device-vector entry points used by the solver: padded entries of the M-vector stay finite
        du, dv, oM, oN = ctx.vecN(np.where(np.arange(N) == 5, np.nan, u)), ctx.vecM(v), ctx.vecM(), ctx.vecN()
        ctx.dATx(du, oM)
        ctx.dAx(oM, oN)                       # NaN in -> NaN out
        assert np.isnan(oN.download(N)[present]).all()
        ctx.dAx(dv, oN)
        assert relerr(oN.download(), ds.Ax(v)) < TOL_MATVEC


def test_snapshot_is_ordered_with_the_stream(C, oracle):
    """gvb_snapshot_begin / _wait: the snapshot holds the vector as it was when begin was called even if it is overwritten
    right afterwards (staging copy on the library stream), slots are independent, a slot can be reused."""
    N, M = 512, 30011
    bed = oracle.synth_bed(1, 0, M, N)
    a = np.random.default_rng(0).normal(size=M)
    with C.Context(0) as ctx:
        ctx.load_host(bed, N)
        va, vb = ctx.vecM(a), ctx.vecM(-a)
        ctx.snapshot_begin(va, M, 0)
        ctx.axpby(va, 3.0, va)                     # overwrite after begin
        ctx.snapshot_begin(va, M - 7, 1)
        ctx.snapshot_begin(vb, M, 2)
        assert np.array_equal(ctx.snapshot_wait(1), (3.0 * a)[:M - 7])
        assert np.array_equal(ctx.snapshot_wait(0), a)
        assert np.array_equal(ctx.snapshot_wait(2), -a)
        ctx.snapshot_begin(vb, 5, 0)               # reuse with another length
        assert np.array_equal(ctx.snapshot_wait(0), -a[:5])
        with pytest.raises(RuntimeError):
            ctx.snapshot_begin(va, M, 16)          # slots are 0..15


def test_upload_changed_reports_bitwise_changes(C, oracle):
    """gvb_vec_upload_changed: the vector ends up equal to the uploaded data and the flag is set exactly when a bit differs
    (a sign of zero counts, a re-upload of the same data does not)."""
    N, M = 512, 30011
    bed = oracle.synth_bed(1, 0, M, N)
    a = np.random.default_rng(5).normal(size=M)
    with C.Context(0) as ctx:
        ctx.load_host(bed, N)
        v, stage = ctx.vecM(a), ctx.vecM()
        assert v.upload_changed(a, stage) is False
        assert np.array_equal(v.download(M), a)
        b = a.copy()
        b[M - 1] = np.nextafter(b[M - 1], np.inf)          # one ulp in the last entry
        assert v.upload_changed(b, stage) is True
        assert np.array_equal(v.download(M), b)
        assert v.upload_changed(b, stage) is False
        z = b.copy()
        z[17] = 0.0
        assert v.upload_changed(z, stage) is True
        z2 = z.copy()
        z2[17] = -0.0
        assert v.upload_changed(z2, stage) is True         # bitwise, not numeric
        assert np.signbit(v.download(M)[17])
        assert v.upload_changed(z2[:100], stage) is False  # prefix upload
        with pytest.raises(RuntimeError):
            v.upload_changed(a, v)                          # the staging vector must be another vector


def test_vector_ops(C, oracle):
    N, M = 512, 3001
    bed = oracle.synth_bed(1, 0, M, N)
    rng = np.random.default_rng(0)
    a, b = rng.normal(size=M), rng.normal(size=M)
    with C.Context(0) as ctx:
        ctx.load_host(bed, N)
        va, vb, vo = ctx.vecM(a), ctx.vecM(b), ctx.vecM()
        ctx.axpby(vo, 0.3, va, -1.7, vb)
        assert relerr(vo.download(), 0.3 * a - 1.7 * b) < 1e-15
        ctx.axpby(vo, 2.0, va)
        assert np.array_equal(vo.download(), 2.0 * a)
        d = ctx.dots([va, va, vb], [vb, None, None])
        assert np.allclose(d, [a @ b, a @ a, b @ b], rtol=1e-13)
        assert abs(ctx.dist2(va, vb) - ((a - b) ** 2).sum()) < 1e-10


def test_denoiser_and_em_golden(C, oracle):
    g = golden("denoiser.npz")
    r1, probs, vars_ = g["r1"], g["probs"], g["vars"]
    M = len(r1)
    bed = oracle.synth_bed(1, 0, M, 64)
    with C.Context(0) as ctx:
        ctx.load_host(bed, 64)
        vr, vx = ctx.vecM(r1), ctx.vecM()
        for tag in "abc":
            gam1 = float(g["gam1_" + tag])
            sums = ctx.denoise(vr, gam1, probs, vars_, vx)
            x1 = vx.download()
            assert np.linalg.norm(x1 - g["g1_" + tag]) / np.linalg.norm(r1) < 1e-13
            assert abs(sums[0] - g["g1d_" + tag].sum()) < 1e-9 * M
            assert abs(sums[1] - ((g["g1_" + tag] - r1) ** 2).sum()) <= 1e-10 * max(1.0, sums[1])
        # one E-step against the oracle restatement of vamp.cpp:944-1008
        lam = 1 - probs[0]
        om = probs.copy()
        om[1:] /= lam
        sums = ctx.em_stats(vr, 3.0, lam, om, vars_)
        spin, res, resg = oracle.em_sufficient_stats(r1, 3.0, list(probs), list(vars_), lam, list(om))
        assert relerr(sums, np.concatenate([[spin], res, resg])) < 1e-12


def test_lmmse_and_cg_golden(C, oracle):
    g = golden("cg.npz")
    N, M = int(g["N"]), int(g["M"])
    bed = oracle.synth_bed(int(g["seed"]), 0, M, N)
    tau, gam2, K = float(g["tau"]), float(g["gam2"]), int(g["CG_max_iter"])
    with C.Context(0) as ctx:
        ctx.load_host(bed, N).compute_stats(1.0)
        rhs, out, mu = ctx.vecM(g["rhs"]), ctx.vecM(), ctx.vecM()
        ctx.lmmse_mult(rhs, tau, gam2, out)
        assert relerr(out.download(), g["lmmse"]) < TOL_MATVEC
        its, log = ctx.cg_solve(rhs, mu, tau, gam2, K, 1)
        assert relerr(mu.download(), g["mu"]) < 1e-5            # CG stops at ||r||/||rhs|| < 1e-5 (vamp.cpp:1217)
        mu.upload(g["mu"] * 0.9)
        ctx.cg_solve(rhs, mu, tau, gam2, K, 1)
        assert relerr(mu.download(), g["mu_warm"]) < 1e-5
        bern, q = ctx.vecM(g["bern"]), ctx.vecM()
        ctx.cg_solve(bern, q, tau, gam2, K, 0)
        alpha2 = gam2 * ctx.dots([bern], [q])[0]
        assert abs(alpha2 / float(g["alpha2"]) - 1) < 1e-6
        # zero right-hand side start vector shortcut: lmmse_mult(0) == 0 (vamp.cpp:1079)
        z = ctx.vecM()
        ctx.lmmse_mult(z, tau, gam2, out)
        assert not out.download().any()


def test_probit_denoiser_golden(C, oracle):
    g = golden("probit_pieces.npz")
    n = len(g["p"])
    bed = oracle.synth_bed(1, 0, 8, n)
    with C.Context(0) as ctx:
        ctx.load_host(bed, n)
        p, y, mc, z = ctx.vecN(g["p"]), ctx.vecN(g["y"]), ctx.vecN(g["mcov"]), ctx.vecN()
        sums = ctx.probit_denoise(p, y, mc, float(g["tau1"]), 1.0, z)
        assert relerr(z.download(n), g["g"]) < 1e-14
        assert abs(sums[0] - g["gd"].sum()) < 1e-10 * n
        assert abs(sums[1] - ((g["g"] - g["p"]) ** 2).sum()) < 1e-10 * n


def test_assoc_pvals_golden(C, oracle):
    """gvb_assoc_pvals (LOO with the marker's own effect added back; LOCO per chromosome with a marker selection) against
    the p-values of the UNMODIFIED reference (data::pvals_calc / pvals_calc_LOCO): N % 4 != 0, 2 % missing genotypes,
    phenotype NAs, p from 1e-68 to 1.  Tolerance 1e-4 relative on p (north_star's bound for final outputs; the sums come
    out of the fixed-point X^T.u sweep and d ln p / d ln t reaches ~400 at the small end)."""
    g = golden("pvals.npz")
    N, M = int(g["N"]), int(g["M"])
    bed = oracle.synth_bed(int(g["seed"]), 0, M, N, miss_rate=float(g["miss_rate"]))
    with make_ctx(C, "lut") as ctx:
        ctx.load_host(bed, N).set_mask(g["mask4"], int(g["nonas"])).compute_stats(1.0)
        ymod = np.zeros(4 * ((N + 3) // 4))
        ymod[:N] = g["y_filtered"] - g["z1"][:N]
        yres, coef, pv, sel = ctx.vecN(ymod), ctx.vecM(g["x1"]), ctx.vecM(), ctx.vecM()
        ctx.assoc_pvals(yres, coef, None, pv)
        loo = pv.download()
        pv.fill(0.0)
        for ch in range(1, 24):
            pick = g["chrom"] == ch
            if not pick.any():
                continue
            pred = ctx.Ax(np.where(pick, g["x1"], 0.0))
            pred[N:] = 0.0
            yres.upload(ymod + pred)
            sel.upload(pick.astype(np.float64))
            ctx.assoc_pvals(yres, None, sel, pv)
        loco = pv.download()
    assert np.allclose(loo, g["pvals_loo"], rtol=1e-4, atol=0), np.max(np.abs(loo / g["pvals_loo"] - 1))
    assert np.allclose(loco, g["pvals_loco"], rtol=1e-4, atol=0), np.max(np.abs(loco / g["pvals_loco"] - 1))


@pytest.mark.parametrize("C_cov", [3, 20, 32])
def test_probit_covariate_pass(C, oracle, C_cov):
    """gvb_probit_cov_pass / gvb_probit_cov_apply (the N x C passes of vamp::Newton_method_cov, grad_cov, mlogL_probit)
    against the numpy restatement; N not a multiple of the 128-individual chunk.  FP64 sums in a different order: 1e-12."""
    N, M = 3001, 64
    rng = np.random.default_rng(C_cov)
    bed = oracle.synth_bed(5, 0, M, N)
    Z = rng.normal(size=(N, C_cov))
    eta = rng.normal(scale=0.3, size=C_cov)
    gg = rng.normal(scale=0.5, size=N)
    y = (rng.random(N) < 0.4).astype(np.float64)
    with make_ctx(C, "lut") as ctx:
        ctx.load_host(bed, N).compute_stats(1.0)
        dy, dgg, dZ, dm = ctx.vecN(y), ctx.vecN(gg), ctx.vec(N * C_cov), ctx.vecN()
        dZ.upload(Z.reshape(-1))
        for pv in (1.0, 0.7):
            got = ctx.probit_cov_pass(dy, dgg, dZ, C_cov, eta, pv)
            want = oracle.probit_cov_pass(y, gg, Z, eta, pv)
            for a, b in zip(got, want):
                assert np.allclose(a, b, rtol=1e-12, atol=1e-12 * np.max(np.abs(b)))
        got0 = ctx.probit_cov_pass(dy, None, dZ, C_cov, eta, 1.0, what=0)
        assert np.isclose(got0[0], oracle.probit_cov_pass(y, np.zeros(N), Z, eta)[0], rtol=1e-12)
        ctx.probit_cov_apply(dZ, C_cov, eta, dm)
        assert np.allclose(dm.download()[:N], Z @ eta, rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("denoiser,warm", [(1, False), (1, True), (0, False)])
def test_cg_by_products(C, oracle, denoiser, warm):
    """gvb_cg_solve_ex: A.mu accumulated from the A p_k of the iterations equals a fresh X.v of the solution, and
    (<rhs,rhs> - gam2 <rhs,mu> - <rhs,r>)/tau equals <rhs, A^T A mu> by explicit sweeps -- in the LMMSE mode (cold and warm
    start) and in the Onsager mode, whose exit leaves the residual one update behind (vamp.cpp:1174-1193).  The solution and
    the iteration count are those of gvb_cg_solve."""
    N, M = 3000, 2600
    bed = oracle.synth_bed(41, 0, M, N, miss_rate=0.01)
    rng = np.random.default_rng(8)
    tau, gam2 = 2.0, 0.7
    rhs_h = rng.normal(size=M) if denoiser else (2.0 * rng.integers(0, 2, size=M) - 1.0) / math.sqrt(M)
    mu0 = rng.normal(size=M) * 0.1 if warm else np.zeros(M)
    with make_ctx(C, "lut") as ctx:
        ctx.load_host(bed, N).compute_stats(1.0)
        rhs, mu, mu_b, ax = ctx.vecM(rhs_h), ctx.vecM(mu0), ctx.vecM(mu0), ctx.vecN()
        its, _ = ctx.cg_solve(rhs, mu, tau, gam2, 30, denoiser)
        s0 = ctx.sweeps()
        its_b, _, d3 = ctx.cg_solve_ex(rhs, mu_b, tau, gam2, 30, denoiser, ax)
        assert ctx.sweeps() - s0 == 2 * its_b + (2 if warm else 0)            # the by-products cost no sweep
        assert its_b == its and np.array_equal(mu.download(), mu_b.download())
        ax_fresh, tmpN, tmpM = ctx.vecN(), ctx.vecN(), ctx.vecM()
        ctx.dAx(mu_b, ax_fresh)
        ctx.dATx(ax_fresh, tmpM)
        explicit = float(np.dot(rhs_h, tmpM.download()[:M]))
        a, b = ax.download(), ax_fresh.download()
    assert relerr(a, b) < 1e-6                                                  # two fixed-point evaluations of A.mu
    assert np.isclose(d3[0], rhs_h @ rhs_h, rtol=1e-12)
    assert abs((d3[0] - gam2 * d3[1] - d3[2]) / tau - explicit) < 1e-6 * abs(explicit)


@pytest.mark.parametrize("start", [2, 1])
def test_prepared_solve_is_bit_identical(C, oracle, start):
    """gvb_cg_prepare + gvb_cg_solve_prepared: the solve's first product A p0 shares one bed read with a companion product (dual sweep).
    Solution, iteration count, log, by-products and the companion product are the bits of gvb_dAx + gvb_cg_solve_warm; the pair of
    calls costs one sweep less; other work may be enqueued between the two calls; a second solve_prepared without prepare is refused."""
    N, M = 3000, 2600
    bed = oracle.synth_bed(43, 0, M, N, miss_rate=0.01)
    rng = np.random.default_rng(10)
    rhs1, rhs2, x = rng.normal(size=M), rng.normal(size=M), rng.normal(size=M)
    with C.Context(0) as ctx:
        ctx.load_host(bed, N).compute_stats(1.0)
        res = {}
        for mode in ("plain", "prepared"):
            r1, r2, mu, ax, ata, xv, z = ctx.vecM(rhs1), ctx.vecM(rhs2), ctx.vecM(), ctx.vecN(), ctx.vecM(), ctx.vecM(x), ctx.vecN()
            if start == 1:                                     # a first solve leaves a start vector with its by-products
                ctx.cg_solve_warm(r1, mu, 2.0, 0.7, 30, 1, ax, ata, 2)
            s0 = ctx.sweeps()
            if mode == "plain":
                ctx.dAx(xv, z)
                its, log, d3 = ctx.cg_solve_warm(r2, mu, 1.5, 0.9, 30, 1, ax, ata, start)
            else:
                ctx.cg_prepare(r2, mu, 1.5, 0.9, 30, ax, ata, start, xv, z)
                tmp = ctx.vecM(rhs1)                            # unrelated work between the two calls
                ctx.axpby(tmp, 2.0, tmp)
                assert np.isfinite(ctx.reduce_batch([(C.RED_DOT, tmp, None, 0, 0, 1)])[0])
                its, log, d3 = ctx.cg_solve_prepared(r2, mu, 1.5, 0.9, 30, 1, ax, ata, start)
                with pytest.raises(RuntimeError):
                    ctx.cg_solve_prepared(r2, mu, 1.5, 0.9, 30, 1, ax, ata, start)
            res[mode] = (its, log.copy(), d3.copy(), mu.download(), ax.download(), ata.download(), z.download(), ctx.sweeps() - s0)
        a, b = res["plain"], res["prepared"]
        assert a[0] == b[0] and a[0] > 3
        for k in range(1, 7):
            assert np.array_equal(a[k], b[k]), k
        assert b[7] == a[7] - 1


def test_solve_with_a_companion_product(C, oracle):
    """gvb_cg_set_companion: a companion's product A v rides on the A p of the solver's iterations (one dual sweep each).  The solver's
    solution, log and iteration count do not change by a bit, the companion receives exactly gvb_dAx(v) every time it asks, the hook is
    consumed by one solve, and an iteration enqueued speculatively after the exit still delivers the companion's product."""
    import ctypes
    N, M = 3000, 2600
    bed = oracle.synth_bed(47, 0, M, N)          # no missing genotypes: the dual X^T.u is a real dual sweep
    rng = np.random.default_rng(12)
    rhs_h = rng.normal(size=M)
    vs = [rng.normal(size=M) for _ in range(3)]
    with C.Context(0) as ctx:
        ctx.load_host(bed, N).compute_stats(1.0)
        rhs, mu, mu_c = ctx.vecM(rhs_h), ctx.vecM(), ctx.vecM()
        its, log = ctx.cg_solve(rhs, mu, 2.0, 0.7, 30, 1)
        v, av, ref, w, wref = ctx.vecM(), ctx.vecN(), ctx.vecN(), ctx.vecM(), ctx.vecM()
        seen = []
        PV = ctypes.POINTER(ctypes.c_void_p)
        FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, PV, PV, PV)

        def companion(user, stage, i, pv, pav, pw):
            if stage == 0:
                if i % 2 == 1:                          # every other iteration goes without
                    return 0
                v.upload(vs[i % 3])
                pv[0], pav[0] = v.h, av.h
                if i % 4 == 0:                          # ... and every other companion iteration also wants w = A^T av
                    pw[0] = w.h
                return 1
            got = av.download()                         # stage 1: the products are there (the download synchronises)
            ctx.dAx(v, ref)                             # ... and sweeps of the companion's own are allowed here
            ok = np.array_equal(got, ref.download())
            if pw[0]:
                ctx.dATx(ref, wref)
                ok = ok and np.array_equal(w.download(), wref.download())
            seen.append((i, ok))
            return 0

        cb = FN(companion)
        assert ctx.L.gvb_cg_set_companion(ctx.h, ctypes.cast(cb, ctypes.c_void_p), None) == 0
        s0, d0 = ctx.sweeps(), ctx.dual_sweeps()
        its_c, log_c = ctx.cg_solve(rhs, mu_c, 2.0, 0.7, 30, 1)
        n_dual = ctx.dual_sweeps() - d0
        assert its_c == its and np.array_equal(log_c, log) and np.array_equal(mu.download(), mu_c.download())
        assert [i for i, _ in seen] == [i for i in range(its + 1) if i % 2 == 0][:len(seen)] and len(seen) >= (its + 1) // 2
        assert all(ok for _, ok in seen)
        n_w = sum(1 for i, _ in seen if i % 4 == 0)
        assert n_dual == len(seen) + n_w                                       # one dual X.v per companion iteration, a dual X^T.u where w was asked for
        assert ctx.sweeps() - s0 >= 2 * its + len(seen) + n_w                  # the solver's own sweeps + the sweeps of the companion's checks
        seen.clear()
        its_d, _ = ctx.cg_solve(rhs, mu_c, 2.0, 0.7, 30, 1)   # the hook served one solve only
        assert seen == [] and its_d <= 1                       # (mu_c is the solution: the solve stops at once)


def test_cg_warm_start_products(C, oracle):
    """gvb_cg_solve_warm: a second solve (other rhs, tau, gam2) started from the first solve's solution with its A.mu / A^T A.mu
    by-products forms its initial residual without a sweep, runs the same iterations as gvb_cg_solve from the same start and
    ends within the fixed-point error of the sweeps; its own by-products match fresh sweeps of its solution."""
    N, M = 3000, 2600
    bed = oracle.synth_bed(43, 0, M, N, miss_rate=0.01)
    rng = np.random.default_rng(9)
    rhs1, rhs2 = rng.normal(size=M), rng.normal(size=M)
    with make_ctx(C, "lut") as ctx:
        ctx.load_host(bed, N).compute_stats(1.0)
        r1, r2, mu, ax, ata = ctx.vecM(rhs1), ctx.vecM(rhs2), ctx.vecM(), ctx.vecN(), ctx.vecM()
        ctx.cg_solve_warm(r1, mu, 2.0, 0.7, 30, 1, ax, ata, False)
        start = mu.download()
        mu_plain = ctx.vecM(start)
        its_plain, _ = ctx.cg_solve(r2, mu_plain, 1.5, 0.9, 30, 1)
        s0 = ctx.sweeps()
        its, _, _ = ctx.cg_solve_warm(r2, mu, 1.5, 0.9, 30, 1, ax, ata, True)
        assert ctx.sweeps() - s0 == 2 * its and its == its_plain
        assert relerr(mu.download(), mu_plain.download()) < 1e-6
        ax_f, ata_f = ctx.vecN(), ctx.vecM()
        ctx.dAx(mu, ax_f)
        ctx.dATx(ax_f, ata_f)
        assert relerr(ax.download(), ax_f.download()) < 1e-6 and relerr(ata.download(), ata_f.download()) < 1e-6


def test_more_than_2M_individuals(C, oracle):
    """N >= 2^21: the packed 21-bit counters of the statistics walk do not apply and the popcount kernel takes over; 17188
    stripes, 68 individual blocks of the missing-genotype list, M = 9 (one full marker group + one ragged)."""
    N, M = 2_200_003, 9
    bed = oracle.synth_bed(51, 0, M, N, miss_rate=0.005)
    ds = oracle.Dataset(bed, N)
    rng = np.random.default_rng(51)
    v, u = rng.normal(size=M), rng.normal(size=N)
    with make_ctx(C, "lut") as ctx:
        ctx.load_host(bed, N).compute_stats(1.0)
        assert np.array_equal(ctx.counts(), ds.counts())
        mave, msig = ctx.stats()
        ax, atx = ctx.Ax(v), ctx.ATx(u)
        assert ctx.missing_list_entries() >= ds.counts()[:, 5].sum() > 0
        assert np.array_equal(ctx.decode(0, M), bed)
    # the oracle adds 2.2M squared deviations one by one like the reference; the closed form from the exact counts is the more
    # accurate of the two, hence 1e-10 here instead of the 1e-11 of the small cases
    assert relerr(mave, ds.mave) < 1e-13 and relerr(msig, ds.msig) < 1e-10
    assert relerr(ax, ds.Ax(v)) < TOL_MATVEC and relerr(atx, ds.ATx(u)) < TOL_MATVEC


def test_people_statistics_and_xxt_solver(C, oracle):
    """gvb_people_stats (three X.v-type walks with per-code tables: value, 1, value^2) and gvb_cg_solve_aat against the reference's
    compute_people_statistics / CG_solverAAT (tests/golden/xxt.npz): N % 4 != 0, 2 % missing genotypes, phenotype NAs.
    Counts bit-exact; the FP sums carry the fixed-point error of a sweep (1e-6 bound), the solve its own 1e-4 exit tolerance."""
    g, gm = golden("xxt.npz"), golden("matvec_n1003.npz")
    N, M = int(g["N"]), int(g["M"])
    bed = oracle.synth_bed(int(gm["seed"]), 0, M, N, miss_rate=float(gm["miss_rate"]))
    with make_ctx(C, "lut") as ctx:
        ctx.load_host(bed, N).set_mask(gm["mask4"], int(gm["nonas"])).compute_stats(1.0)
        pe = ctx.people_stats()
        mave_p, msig_p, numb_p = (v.download() for v in pe)
        rhs, mu = ctx.vecN(np.concatenate([g["cg_rhs"], np.zeros(len(mave_p) - N)])), ctx.vecN()
        its, log = ctx.cg_solve_aat(rhs, mu, float(g["cg_tau"]), float(g["cg_gam2"]), pe, int(g["cg_max_iter"]))
        u = mu.download()
    assert np.array_equal(numb_p, g["numb_people"])
    assert relerr(mave_p, g["mave_people"]) < TOL_MATVEC and relerr(msig_p, g["msig_people"]) < TOL_MATVEC
    assert relerr(u, g["cg_u"]) < 1e-4 and log[-1, 0] < 1e-4


@pytest.mark.parametrize("denoiser", [1, 0])
def test_cg_device_resident_scalars(C, oracle, denoiser, monkeypatch):
    """The CG driver keeps alpha / beta / the exit tests on the device and enqueues iteration i+1 before it has seen the flag of
    iteration i (cg.cu).  The speculative form (default) and the wait-every-iteration form (GVB_CG_LAG=0) must run the same number of
    iterations as the oracle's host loop (vamp.cpp:1161-1224), give bit-identical solutions and the same per-iteration log; the
    speculative iteration is not counted as sweeps; max_iter == 0 returns the start vector and <rhs, mu_start>."""
    N, M = 2000, 3100
    bed = oracle.synth_bed(17, 0, M, N, miss_rate=0.005)
    ds = oracle.Dataset(bed, N)
    rng = np.random.default_rng(3)
    tau, gam2, K = 1.7, 0.4, 40
    rhs_h = rng.normal(size=M) if denoiser else oracle.bernoulli_probe(3, 0, M, M)
    log_ref = []
    mu_ref, its_ref = oracle.precond_cg(ds, rhs_h, np.zeros(M), tau, gam2, K, denoiser, log=log_ref)
    out = {}
    for lag in ("1", "0"):
        monkeypatch.setenv("GVB_CG_LAG", lag)
        with make_ctx(C, "lut") as ctx:
            ctx.load_host(bed, N).compute_stats(1.0)
            rhs, mu = ctx.vecM(rhs_h), ctx.vecM()
            s0, h0 = ctx.sweeps(), ctx.host_syncs()
            its, log = ctx.cg_solve(rhs, mu, tau, gam2, K, denoiser)
            out[lag] = (its, log.copy(), mu.download(), ctx.sweeps() - s0, ctx.host_syncs() - h0)
            if lag == "1":
                mu0 = rng.normal(size=M)
                mu.upload(mu0)
                ax = ctx.vecN()
                its0, _, d3 = ctx.cg_solve_ex(rhs, mu, tau, gam2, 0, denoiser, ax)
                assert its0 == 0 and np.array_equal(mu.download(), mu0)
                assert np.isclose(d3[1], rhs_h @ mu0, rtol=1e-12) and np.isclose(d3[0], rhs_h @ rhs_h, rtol=1e-12)
    its, log, mu_d, sweeps, syncs = out["1"]
    assert its == its_ref == out["0"][0], (its, its_ref, out["0"][0])
    assert np.array_equal(mu_d, out["0"][2]) and np.array_equal(log, out["0"][1])
    assert sweeps == 2 * its == out["0"][3]                      # the iteration enqueued after the exit is not a sweep
    assert relerr(mu_d, mu_ref) < 1e-5
    n_res = len(log_ref)                                         # an Onsager exit leaves no residual line for its last iteration
    assert np.allclose(log[:n_res, 0], log_ref, rtol=1e-3)       # ||r||/||rhs|| per iteration: fixed-point sweeps vs FP64 oracle
    assert syncs <= its + 3


def test_reduce_batch(C, oracle):
    """gvb_vec_reduce_batch: M- and N-vectors mixed, dot / squared-norm-of-combination kinds, one host synchronisation."""
    N, M = 777, 3001
    bed = oracle.synth_bed(1, 0, M, N)
    rng = np.random.default_rng(1)
    a, b, u, w = rng.normal(size=M), rng.normal(size=M), rng.normal(size=N), rng.normal(size=N)
    with C.Context(0) as ctx:
        ctx.load_host(bed, N)
        va, vb, vu, vw = ctx.vecM(a), ctx.vecM(b), ctx.vecN(np.pad(u, (0, 4 * ((N + 3) // 4) - N))), ctx.vecN(np.pad(w, (0, 4 * ((N + 3) // 4) - N)))
        h0 = ctx.host_syncs()
        res = ctx.reduce_batch([(C.RED_DOT, va, vb, 0, 0, 1), (C.RED_DOT, va, None, 0, 0, 1), (C.RED_SQ, va, vb, 0.5, -2.0, 1), (C.RED_SQ, vb, None, 3.0, 0, 1),
                                (C.RED_SQ, vu, vw, 1.0, -1.0, 0), (C.RED_DOT, vu, vw, 0, 0, 0)])
        assert ctx.host_syncs() - h0 == 1
        ref = [a @ b, a @ a, ((0.5 * a - 2.0 * b) ** 2).sum(), 9.0 * (b @ b), ((u - w) ** 2).sum(), u @ w]
        assert np.allclose(res, ref, rtol=1e-13)
        with pytest.raises(C.GvbError):
            ctx.reduce_batch([(C.RED_DOT, vu, vw, 0, 0, 0), (C.RED_DOT, va, vb, 0, 0, 1)])   # rank-summed operations come first


ADVERSARIAL = ["u_outlier_1e6", "u_outlier_1e3", "u_two_scales", "v_outlier_1e6", "v_clustered_tile", "v_sparse", "both_tiny_and_huge"]


@pytest.mark.parametrize("case", ADVERSARIAL)
def test_fixed_point_dynamic_range(C, oracle, case):
    """The fixed-point sweeps under adversarial dynamic range: a single outlier 10^3 / 10^6 times the rest in u,
    stripes of u on two scales 10^4 apart, a single outlier in v, 128 clustered large effects in ONE marker tile over a tiny dense
    background, a 99.9 % sparse v.  With one global scale the outlier's magnitude would set the resolution of every table entry
    (norm-wise error ~ 3.5e-8 * max|u| / rms(u), i.e. 2e-5 for the 10^6 outlier at this N); the per-stripe / per-tile scale classes
    (matvec_tile.cu "Scale classes") keep BOTH the norm-wise error ||d||_2 / ||ref||_2 and the worst element against the largest
    output, ||d||_inf / ||ref||_inf, below the north_star's 1e-6.  N = 131072 (1024 stripes): the error grows with sqrt(N)."""
    N, M = 131072, 2048
    bed = oracle.synth_bed(23, 0, M, N, miss_rate=0.0)
    ds = oracle.Dataset(bed, N)
    rng = np.random.default_rng(len(case))
    u, v = rng.normal(size=N), rng.normal(size=M)
    if case == "u_outlier_1e6":
        u[54321] = 1e6
    elif case == "u_outlier_1e3":
        u[777] = -1e3
    elif case == "u_two_scales":
        u[: N // 2] *= 1e-4
    elif case == "v_outlier_1e6":
        v[77] = 1e6
    elif case == "v_clustered_tile":
        v *= 1e-3
        v[256:384] = rng.normal(size=128) * 1e3
    elif case == "v_sparse":
        v[rng.random(M) > 0.001] = 0.0
        v[5] = 3.0
    elif case == "both_tiny_and_huge":
        u *= 1e-150
        v *= 1e150
    with make_ctx(C, "lut") as ctx:
        ctx.load_host(bed, N).compute_stats(1.0)
        ax, atx = ctx.Ax(v)[:N], ctx.ATx(u)
    ax_ref, atx_ref = ds.Ax(v)[:N], ds.ATx(u)
    linf = lambda a, b: float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
    assert relerr(ax, ax_ref) < TOL_MATVEC and linf(ax, ax_ref) < TOL_MATVEC, (relerr(ax, ax_ref), linf(ax, ax_ref))
    assert relerr(atx, atx_ref) < TOL_MATVEC and linf(atx, atx_ref) < TOL_MATVEC, (relerr(atx, atx_ref), linf(atx, atx_ref))


def test_fixed_point_dynamic_range_with_missing_genotypes(C, oracle):
    """The same outlier on a shard WITH missing genotypes: the missing-genotype correction sum_{i missing} u_i is gathered at the common
    class-0 scale (misslist.cu sums indices of many stripes in one window), so its share of the result keeps the single-scale
    resolution; with 2 % missing genotypes and a 10^3 outlier both norms stay below 1e-6, and the list and the second walk (which
    carries the classes) agree to that accuracy instead of bit for bit."""
    N, M = 131072, 1024
    bed = oracle.synth_bed(29, 0, M, N, miss_rate=0.02)
    ds = oracle.Dataset(bed, N)
    u = np.random.default_rng(2).normal(size=N)
    u[4242] = 1e3
    with make_ctx(C, "lut") as ctx:
        ctx.load_host(bed, N).compute_stats(1.0)
        atx = ctx.ATx(u)
        assert ctx.missing_list_entries() > 0
    ref = ds.ATx(u)
    assert relerr(atx, ref) < TOL_MATVEC and float(np.max(np.abs(atx - ref)) / np.max(np.abs(ref))) < TOL_MATVEC


def test_product_library_refuses_cross_check_kernels(C, monkeypatch):
    """GVB_KERNELS=simple|lut1 selects a kernel generation of the TEST build only; the product library fails loudly instead of
    silently running something else."""
    monkeypatch.setenv("GVB_KERNELS", "simple")
    with pytest.raises(C.GvbError) as e:
        C.Context(0)
    assert "not part of the product library" in str(e.value)
    monkeypatch.setenv("GVB_KERNELS", "lut")
    C.Context(0).close()


def test_layout_generation_counts_reloads(C, oracle):
    """gvb_layout_generation is bumped by every (re)load of the matrix and every gvb_set_mask: the host layer drops its cached
    by-products (A mu, A^T A mu, A^T y) and re-creates its device vectors when it changes (vamp::dev_open)."""
    bed = oracle.synth_bed(2, 0, 300, 500)
    with C.Context(0) as ctx:
        g0 = ctx.layout_generation()
        ctx.load_host(bed, 500)
        g1 = ctx.layout_generation()
        ctx.set_mask(oracle.make_mask4(500, np.arange(500) % 7 != 0), int((np.arange(500) % 7 != 0).sum()))
        g2 = ctx.layout_generation()
        ctx.load_host(bed[:200], 500)
        g3 = ctx.layout_generation()
    assert g0 < g1 < g2 < g3


def test_cg_cached_operator_product(C, oracle):
    """gvb_cg_solve_cached: a zero-start solve against a recurring right-hand side (the Onsager probe) takes the operator product of
    its first iteration from a cache of A^T A rhs.  First call: fills the cache, same sweeps as gvb_cg_solve; later calls (other tau /
    gam2): two sweeps less, the same iteration count and the same solution up to the fixed-point error of the sweeps."""
    N, M = 3000, 2600
    bed = oracle.synth_bed(47, 0, M, N, miss_rate=0.01)
    probe = oracle.bernoulli_probe(5, 0, M, M)
    with make_ctx(C, "lut") as ctx:
        ctx.load_host(bed, N).compute_stats(1.0)
        rhs, mu, mu_plain, cache = ctx.vecM(probe), ctx.vecM(), ctx.vecM(), ctx.vecM()
        state = [0]
        for k, (tau, gam2) in enumerate(((2.0, 0.7), (1.3, 0.2), (3.1, 5.0))):
            mu_plain.fill(0.0)
            s0 = ctx.sweeps()
            its_plain, log_plain = ctx.cg_solve(rhs, mu_plain, tau, gam2, 30, 0)
            s1 = ctx.sweeps()
            its, log, d3 = ctx.cg_solve_cached(rhs, mu, tau, gam2, 30, 0, cache, state)
            s2 = ctx.sweeps()
            assert state[0] == 1 and its == its_plain
            assert (s2 - s1) == (s1 - s0) - (0 if k == 0 else 2), (k, s1 - s0, s2 - s1)
            assert relerr(mu.download(), mu_plain.download()) < 1e-6
            assert np.isclose(d3[1], probe @ mu.download()[:M], rtol=1e-10)
        ata = cache.download()[:M]
        tmpN, tmpM = ctx.vecN(), ctx.vecM()
        ctx.dAx(rhs, tmpN)
        ctx.dATx(tmpN, tmpM)
        assert relerr(ata, tmpM.download()[:M]) < 1e-6
