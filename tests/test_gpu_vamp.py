"""End-to-end parity of the B200 gVAMP driver (gvamp_b200/bin/main_real, same command line as the
reference's main_real.exe) against the golden output files of the unmodified reference.
Tolerances: BASELINE.json north_star -- final signal estimates, learned prior variances within 1e-4."""
import math
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden, relerr

pytestmark = pytest.mark.gpu
EXE = os.path.join(ROOT, "gvamp_b200", "bin", "main_real")
TOL_FINAL = 1e-4


def _kernel_env(gen, extra=None):
    """Environment of a driver run with kernel generation `gen`: the cross-check generations live in the test build of the library,
    which the executables pick up through LD_LIBRARY_PATH (their RUNPATH to gvamp_b200/lib is searched after it)."""
    env = dict(os.environ, GVB_KERNELS=gen, **(extra or {}))
    if gen in ("simple", "lut1"):
        env["LD_LIBRARY_PATH"] = os.path.join(ROOT, "gvamp_b200", "lib", "xcheck") + ":" + env.get("LD_LIBRARY_PATH", "")
    return env


def _run_case(oracle, tmp_path, g, gen, env_extra=None):
    N, M, iters = int(g["N"]), int(g["M"]), int(g["iterations"])
    bed = oracle.synth_bed(int(g["seed"]), 0, M, N)
    bedp, phenp = str(tmp_path / "v.bed"), str(tmp_path / "v.phen")
    oracle.write_bed(bedp, bed)
    oracle.write_phen(phenp, g["y"])
    outd = str(tmp_path / f"out_{gen}") + "/"
    args = ["--run-mode", "infere", "--model", "linear", "--bed-file", bedp, "--phen-files", phenp, "--N", str(N), "--Mt", str(M),
            "--out-dir", outd, "--out-name", "g"]
    # the remaining options exactly as the reference run used them (make_golden.py), minus its own paths / sizes
    extra = [str(a) for a in g["args"]]
    skip = {"--N", "--Mt", "--out-dir", "--out-name"}
    for k in range(0, len(extra), 2):
        if extra[k] not in skip:
            args += [extra[k], extra[k + 1]]
    env = _kernel_env(gen, env_extra)
    r = subprocess.run([EXE] + args, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return outd, r.stdout, iters


@pytest.mark.parametrize("gen", ["simple", "lut"])
def test_linear_vamp_matches_reference_files(oracle, tmp_path, gen):
    g = golden("vamp_linear.npz")
    outd, log, iters = _run_case(oracle, tmp_path, g, gen)
    for it in range(1, iters + 1):
        for key, fn in (("x1", f"g_it_{it}.bin"), ("r1", f"g_r1_it_{it}.bin"), ("r2", f"g_r2_it_{it}.bin"), ("x2", f"g_it_{it}_x2_hat.bin")):
            ref = g[f"{key}_{it}"]
            got = np.fromfile(outd + fn)
            assert got.shape == ref.shape
            if np.linalg.norm(ref) > 0:
                assert relerr(got, ref) < TOL_FINAL, (key, it, relerr(got, ref))
    for nm in ("gam1s", "gam2s", "R2trains"):
        assert np.allclose(np.loadtxt(f"{outd}g_{nm}.csv"), g[nm], rtol=TOL_FINAL, atol=1e-6), nm
    assert np.allclose(np.loadtxt(f"{outd}g_z1_it_{iters}.csv"), g["z1_text_last"], rtol=1e-3, atol=1e-5)
    gamw = [float(l.split("=")[1]) for l in log.splitlines() if l.startswith("gamw = ")]
    assert np.allclose(gamw, g["gamw_log"], rtol=TOL_FINAL)
    pv = [np.array(l.split("=")[1].split(), dtype=float) for l in log.splitlines() if l.startswith("prior variances")]
    pp = [np.array(l.split("=")[1].split(), dtype=float) for l in log.splitlines() if l.startswith("prior probabilities")]
    assert np.allclose(pv[-1], g["prior_vars_last"], rtol=TOL_FINAL) and np.allclose(pp[-1], g["prior_probs_last"], rtol=TOL_FINAL)
    alpha2 = [float(l.split("=")[1]) for l in log.splitlines() if l.startswith("alpha2 = ")]
    assert np.allclose(alpha2, g["alpha2_log"], rtol=TOL_FINAL)


def test_async_outputs_write_the_same_files(oracle, tmp_path):
    """The iteration outputs leave the device as asynchronous snapshots and are written at the end of the iteration
    (vamp::emit_output / flush_outputs); GVB_ASYNC_OUT=0 downloads and writes each one in place like the reference.  Same
    arithmetic, same bytes: every output file of the two runs must be identical."""
    g = golden("vamp_linear.npz")
    outs = {}
    for mode in ("1", "0"):
        (tmp_path / mode).mkdir()
        outs[mode], _, iters = _run_case(oracle, tmp_path / mode, g, "lut", {"GVB_ASYNC_OUT": mode})
    names = sorted(os.listdir(outs["1"]))
    assert names == sorted(os.listdir(outs["0"])) and len(names) >= 5 * iters
    for fn in names:
        assert open(outs["1"] + fn, "rb").read() == open(outs["0"] + fn, "rb").read(), fn


def test_dual_sweep_writes_the_same_files(oracle, tmp_path):
    """z1 = A x1_hat shares a bed read with the first product of the LMMSE solve (gvb_cg_prepare, a dual sweep; the solve's inputs
    are formed before z1 instead of after it).  Same integer sums, same kernels on the same data: every output file and every
    value line of the log must be identical to a run with GVB_DUAL_SWEEP=0, which costs one sweep more per iteration."""
    g = golden("vamp_linear.npz")
    outs, logs = {}, {}
    for mode in ("1", "0"):
        (tmp_path / mode).mkdir()
        outs[mode], logs[mode], iters = _run_case(oracle, tmp_path / mode, g, "lut", {"GVB_DUAL_SWEEP": mode, "GVB_LOG_SWEEPS": "1"})
    names = sorted(os.listdir(outs["1"]))
    assert names == sorted(os.listdir(outs["0"])) and len(names) >= 5 * iters
    for fn in names:
        assert open(outs["1"] + fn, "rb").read() == open(outs["0"] + fn, "rb").read(), fn
    def keep(mode):    # the log without its wall-clock lines and without the run's own directory in the paths it prints
        lines = [l.replace(str(tmp_path / mode), "") for l in logs[mode].splitlines()]
        return [l for l in lines if not any(w in l for w in ("seconds", "took", "time", "bed sweeps"))]
    assert keep("1") == keep("0")
    sw = {m: [int(l.split("=")[1]) for l in logs[m].splitlines() if l.startswith("bed sweeps this iteration")] for m in ("0", "1")}
    assert len(sw["1"]) == iters and all(a <= b for a, b in zip(sw["1"], sw["0"])) and sum(sw["0"]) - sum(sw["1"]) >= iters - 2


def test_config1_matches_reference(oracle, tmp_path):
    """BASELINE.json configs[0] end to end: N=10,000 x M=20,000, h2=0.5, CV=2,000, linear, 10 iterations, against the output of
    the reference's own main_real.exe (MANVECT build, tests/golden/config1.npz): signal estimate of the first, a middle and the
    final iteration, per-iteration gam1 / gam2 / R2, gamw, Onsager alpha2 and the learned prior within 1e-4 (north_star)."""
    g = golden("config1.npz")
    N, M, iters = int(g["N"]), int(g["M"]), int(g["iterations_done"])
    bed = oracle.synth_bed(int(g["seed"]), 0, M, N)
    bedp, phenp = str(tmp_path / "c1.bed"), str(tmp_path / "c1.phen")
    oracle.write_bed(bedp, bed)
    oracle.write_phen(phenp, g["y"])
    outd = str(tmp_path / "c1out") + "/"
    args = ["--run-mode", "infere", "--model", "linear", "--bed-file", bedp, "--phen-files", phenp, "--N", str(N), "--Mt", str(M),
            "--out-dir", outd, "--out-name", "c1"]
    extra = [str(a) for a in g["args"]]
    skip = {"--N", "--Mt", "--out-dir", "--out-name"}
    for k in range(0, len(extra), 2):
        if extra[k] not in skip:
            args += [extra[k], extra[k + 1]]
    r = subprocess.run([EXE] + args, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    log = r.stdout
    for key, it in (("x1_first", 1), ("x1_mid", max(1, iters // 2)), ("x1_last", iters)):
        got = np.fromfile(outd + f"c1_it_{it}.bin")
        assert relerr(got, g[key]) < TOL_FINAL, (key, relerr(got, g[key]))
    for nm in ("gam1s", "gam2s", "R2trains"):
        assert np.allclose(np.loadtxt(f"{outd}c1_{nm}.csv"), g[nm], rtol=TOL_FINAL, atol=1e-6), nm
    gamw = [float(l.split("=")[1]) for l in log.splitlines() if l.startswith("gamw = ")]
    alpha2 = [float(l.split("=")[1]) for l in log.splitlines() if l.startswith("alpha2 = ")]
    assert np.allclose(gamw, g["gamw_log"], rtol=TOL_FINAL) and np.allclose(alpha2, g["alpha2_log"], rtol=TOL_FINAL)
    pv = [np.array(l.split("=")[1].split(), dtype=float) for l in log.splitlines() if l.startswith("prior variances")]
    pp = [np.array(l.split("=")[1].split(), dtype=float) for l in log.splitlines() if l.startswith("prior probabilities")]
    assert np.allclose(pv[-1], g["prior_vars_last"], rtol=TOL_FINAL) and np.allclose(pp[-1], g["prior_probs_last"], rtol=TOL_FINAL)


def test_xxt_denoiser_matches_reference_files(oracle, tmp_path):
    """--use-XXT-denoiser 1 (N-space LMMSE step, denoiserXXT.cpp) end to end against the reference's files.  The N-space solve
    stops at ||r||/||rhs|| < 1e-4 (denoiserXXT.cpp:125), so two correct implementations agree to about that; the bound here is
    5e-3 on the estimates (the 1e-4 of the north_star applies to the default M-space path, which the other tests hold)."""
    g, gl = golden("xxt.npz"), golden("vamp_linear.npz")
    N, M, iters = int(gl["N"]), int(gl["M"]), int(g["e2e_iterations"])
    bed = oracle.synth_bed(int(gl["seed"]), 0, M, N)
    bedp, phenp = str(tmp_path / "x.bed"), str(tmp_path / "x.phen")
    oracle.write_bed(bedp, bed)
    oracle.write_phen(phenp, gl["y"])
    outd = str(tmp_path / "xout") + "/"
    args = ["--run-mode", "infere", "--model", "linear", "--bed-file", bedp, "--phen-files", phenp, "--N", str(N), "--Mt", str(M),
            "--out-dir", outd, "--out-name", "x"]
    extra = [str(a) for a in g["e2e_args"]]
    for k in range(0, len(extra), 2):
        if extra[k] not in {"--N", "--Mt", "--out-dir", "--out-name"}:
            args += [extra[k], extra[k + 1]]
    r = subprocess.run([EXE] + args, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    for it in range(1, iters + 1):
        for key, fn in (("x1", f"x_it_{it}.bin"), ("x2", f"x_it_{it}_x2_hat.bin")):
            ref, got = g[f"{key}_{it}"], np.fromfile(outd + fn)
            if np.linalg.norm(ref) > 0:
                assert relerr(got, ref) < 5e-3, (key, it, relerr(got, ref))
    gamw = [float(l.split("=")[1]) for l in r.stdout.splitlines() if l.startswith("gamw = ")]
    assert np.allclose(gamw, g["gamw_log"], rtol=5e-3)


def test_cg_by_products_leave_the_outputs_unchanged(oracle, tmp_path, monkeypatch):
    """Default run (A x2_hat and the trace term of updateNoisePrec as by-products of the two CG solves) against
    GVB_REFERENCE_SWEEPS=1 (their own bed sweeps, as the reference does, vamp.cpp:897-915): same files and gamw far inside the
    tolerance, 3 sweeps less in the cold-started first iteration and 5 less in the warm-started ones (their initial residual comes
    from the previous solve's A^T A x2_hat)."""
    g = golden("vamp_linear.npz")
    outs = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("GVB_REFERENCE_SWEEPS", mode)
        monkeypatch.setenv("GVB_LOG_SWEEPS", "1")      # the per-iteration sweep count is not a line of the reference's log
        (tmp_path / mode).mkdir()
        outd, log, iters = _run_case(oracle, tmp_path / mode, g, "lut")
        x1 = np.fromfile(outd + f"g_it_{iters}.bin")
        gamw = [float(l.split("=")[1]) for l in log.splitlines() if l.startswith("gamw = ")]
        sweeps = [int(l.split("=")[1]) for l in log.splitlines() if l.startswith("bed sweeps this iteration")]
        outs[mode] = (x1, np.array(gamw), np.loadtxt(f"{outd}g_R2trains.csv"), np.array(sweeps))
    assert relerr(outs["0"][0], outs["1"][0]) < 1e-5 and np.allclose(outs["0"][1], outs["1"][1], rtol=1e-5)
    assert np.allclose(outs["0"][2], outs["1"][2], rtol=1e-5, atol=1e-9)
    saved = outs["1"][3] - outs["0"][3]
    # first iteration: the 3 by-product sweeps, minus the two Lanczos steps that the projected Onsager solve builds beyond the solve's own
    # iteration count; later iterations: 5 (by-products + the sweep-free initial residual of the warm-started LMMSE solve) plus ALL
    # sweeps of the Onsager solve, which runs on the cached Lanczos projection of A^T A (vamp::onsager_projected)
    assert saved[0] >= 0 and np.all(saved[1:] >= 5 + 4), saved


def test_test_mode_r2(oracle, tmp_path):
    """--run-mode test: R2 of an estimate file on a second (test) bed, against the oracle's arithmetic."""
    g = golden("vamp_linear.npz")
    N, M = 800, int(g["M"])
    bed = oracle.synth_bed(99, 0, M, N)
    ds = oracle.Dataset(bed, N)
    beta = g["beta"]
    y = ds.Ax(beta * math.sqrt(N))[:N] + oracle.synth_noise(99, N, 0.5)
    bedp, phenp, estp = str(tmp_path / "t.bed"), str(tmp_path / "t.phen"), str(tmp_path / "est_it_3.bin")
    oracle.write_bed(bedp, bed)
    oracle.write_phen(phenp, y)
    g["x1_3"].astype(np.float64).tofile(estp)
    r = subprocess.run([EXE, "--run-mode", "test", "--bed-file-test", bedp, "--phen-files-test", phenp, "--N-test", str(N), "--Mt-test", str(M),
                        "--estimate-file", estp], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:]
    got = float([l for l in r.stdout.splitlines() if l.startswith("test R2 = ")][0].split("=")[1])
    phen, mask4, nonas, avg, sqn = oracle.read_phen(phenp, N)
    dsp = oracle.Dataset(bed, N, phen=phen, mask4=mask4, nonas=nonas)
    z = dsp.Ax(g["x1_3"] * math.sqrt(N))[:N]
    sd2 = (np.sum(phen ** 2) - N * phen.mean() ** 2) / (N - 1)
    want = 1 - np.sum((phen - z) ** 2) / (sd2 * N)
    assert abs(got - want) < TOL_FINAL


def test_pvals_calc_mode(oracle, tmp_path):
    """--run-mode pvals-calc (main_real.cpp:331-452): LOO + LOCO p-value files of an estimate file against the reference's."""
    g = golden("pvals.npz")
    N, M = int(g["N"]), int(g["M"])
    bed = oracle.synth_bed(int(g["seed"]), 0, M, N, miss_rate=float(g["miss_rate"]))
    bedp, phenp, bimp, estp = str(tmp_path / "pv.bed"), str(tmp_path / "pv.phen"), str(tmp_path / "pv.bim"), str(tmp_path / "est.bin")
    oracle.write_bed(bedp, bed)
    oracle.write_phen(phenp, g["y"], na_idx=[int(i) for i in g["na_idx"]])
    with open(bimp, "w") as fh:
        for j, ch in enumerate(g["chrom"]):
            fh.write(f"{'X' if ch == 23 else ch}\trs{j}\t0\t{1000 + j}\tA\tG\n")
    (g["x1"] / math.sqrt(N)).astype(np.float64).tofile(estp)        # the driver multiplies the file by sqrt(N)
    outd = str(tmp_path / "pvout") + "/"
    r = subprocess.run([EXE, "--run-mode", "pvals-calc", "--bed-file", bedp, "--phen-files", phenp, "--bim-file", bimp, "--N", str(N), "--Mt", str(M),
                        "--estimate-file", estp, "--out-dir", outd, "--out-name", "pv"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    loo = np.fromfile(outd + "pv_pvals.bin")
    loco = np.fromfile(outd + "pv_pvals_LOCO.bin_pvals_LOCO.bin")     # the reference appends the suffix twice (data.cpp:1348)
    # y in the golden was scaled by read_phen before the reference used it; the driver re-reads the same phen file
    assert np.allclose(loo, g["pvals_loo"], rtol=1e-4, atol=0) and np.allclose(loco, g["pvals_loco"], rtol=1e-4, atol=0)
    assert os.path.exists(outd + "pv_pvals_LOCO.bin_LOCO_chr_7.csv")


EXE_PROBIT = os.path.join(ROOT, "gvamp_b200", "bin", "main_real_probit")


@pytest.mark.parametrize("gen", ["simple", "lut"])
def test_probit_vamp_matches_reference_files(oracle, tmp_path, gen):
    """main_real_probit --model bin_class with C = 3 covariates against the files and log values of the UNMODIFIED
    reference (tests/golden/vamp_probit.npz, made by make_golden.py:case_vamp_probit; vamp_probit.cpp:20-658)."""
    g = golden("vamp_probit.npz")
    N, M, iters, C = int(g["N"]), int(g["M"]), int(g["iterations"]), int(g["C"])
    bed = oracle.synth_bed(int(g["seed"]), 0, M, N)
    bedp, phenp, covp = str(tmp_path / "p.bed"), str(tmp_path / "p.phen"), str(tmp_path / "p.cov")
    oracle.write_bed(bedp, bed)
    oracle.write_phen(phenp, g["y"])
    with open(covp, "w") as fh:
        for row in g["Z"]:
            fh.write(" ".join(repr(float(z)) for z in row) + "\n")
    outd = str(tmp_path / f"outp_{gen}") + "/"
    args = ["--run-mode", "infere", "--model", "bin_class", "--bed-file", bedp, "--phen-files", phenp, "--N", str(N), "--Mt", str(M),
            "--out-dir", outd, "--out-name", "p", "--cov-file", covp, "--C", str(C)]
    extra = [str(a) for a in g["args"]]
    skip = {"--N", "--Mt", "--out-dir", "--out-name", "--cov-file", "--C"}
    for k in range(0, len(extra), 2):
        if extra[k] not in skip:
            args += [extra[k], extra[k + 1]]
    r = subprocess.run([EXE_PROBIT] + args, capture_output=True, text=True, env=_kernel_env(gen), timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert int(g["iterations_done"]) == iters
    for it in range(1, iters + 1):
        for key, fn in (("x1", f"p_probit_it_{it}.bin"), ("r1", f"p_probit_r1_it_{it}.bin")):
            ref, got = g[f"{key}_{it}"], np.fromfile(outd + fn)
            assert got.shape == ref.shape
            if np.linalg.norm(ref) > 0:
                assert relerr(got, ref) < TOL_FINAL, (key, it, relerr(got, ref))

    def grab(prefix):
        return np.array([float(l.split("=")[-1]) for l in r.stdout.splitlines() if l.startswith(prefix)])
    for key, prefix in (("gam1_log", "gam1 = "), ("gam2_log", "gam2 = "), ("alpha2_log", "alpha2 = "), ("tau1_log", "tau1 = "), ("tau2_log", "tau2 = "),
                        ("beta1_log", "beta1 = "), ("beta2_log", "beta2 = "), ("eta1_log", "eta1 = ")):
        got = grab(prefix)
        assert got.shape == g[key].shape, (key, got, g[key])
        assert np.allclose(got, g[key], rtol=5 * TOL_FINAL), (key, got, g[key])     # log lines carry 6 significant digits
    cov = [l for l in r.stdout.splitlines() if l.startswith("cov_eff[")]
    got_cov = np.array([float(tok.split("=")[1]) for l in cov for tok in l.split(",") if "=" in tok])
    assert np.allclose(got_cov, g["cov_eff_log"], rtol=5 * TOL_FINAL)


def _stress_prior():
    p = min(50000.0 / 500000, 1.0) / (2 - 1.0 / 2 ** 21)
    probs = [1 - min(50000.0 / 500000, 1.0)]
    for _ in range(22):
        probs.append(p)
        p /= 2
    step = 10 ** (math.log10(1e2 / 1e-5) / 21)
    vars_, v = [0.0], 1e-5
    for _ in range(22):
        vars_.append(v)
        v *= step
    return probs, vars_


@pytest.mark.parametrize("mode", ["default", "reference_sweeps", "onsager_sweeps", "onsager_warm", "cg_lag0"])
def test_stress_case_matches_reference(oracle, tmp_path, mode):
    """An ill-conditioned run against the reference's files (tests/golden/make_golden.py:case_vamp_stress): M/N = 5, rho = 0.05 (the
    dnanexus damping), 30 iterations, the reference's 23-component prior shape.  The default path (CG by-products instead of the
    three extra sweeps, the Onsager solve on the Lanczos projection of A^T A - no bed sweep after the first iteration -, device-resident
    CG scalars) and the variants GVB_REFERENCE_SWEEPS=1 (the reference's own sweeps), GVB_ONSAGER_LANCZOS=0 (the Onsager solve by bed
    sweeps, A^T A probe cached), GVB_CG_LAG=0 (no speculative iteration) must stay within the north_star's 1e-4 over all 30 iterations
    and run the reference's number of CG iterations; GVB_ONSAGER_WARM=1 changes the solver's path on purpose (opt-in)
    and is held to a looser 5e-4."""
    g = golden("vamp_stress.npz")
    env = {"reference_sweeps": {"GVB_REFERENCE_SWEEPS": "1"}, "onsager_warm": {"GVB_ONSAGER_WARM": "1"}, "cg_lag0": {"GVB_CG_LAG": "0"},
           "onsager_sweeps": {"GVB_ONSAGER_LANCZOS": "0"}}.get(mode, {})
    outd, log, _ = _run_case(oracle, tmp_path, g, "lut", env)
    iters = int(g["iterations_done"])
    tol = 5e-4 if mode == "onsager_warm" else TOL_FINAL
    for it in (2, 3, 5, 10, 20, iters):
        for key, fn in (("x1", f"g_it_{it}.bin"), ("x2", f"g_it_{it}_x2_hat.bin"), ("r1", f"g_r1_it_{it}.bin")):
            assert relerr(np.fromfile(outd + fn), g[f"{key}_{it}"]) < tol, (key, it, relerr(np.fromfile(outd + fn), g[f"{key}_{it}"]))
    for nm in ("gam1s", "gam2s", "R2trains"):
        assert np.allclose(np.loadtxt(f"{outd}g_{nm}.csv"), g[nm], rtol=tol, atol=1e-6), nm
    gamw = [float(l.split("=")[1]) for l in log.splitlines() if l.startswith("gamw = ")]
    alpha2 = [float(l.split("=")[1]) for l in log.splitlines() if l.startswith("alpha2 = ")]
    assert np.allclose(gamw, g["gamw_log"], rtol=tol) and np.allclose(alpha2, g["alpha2_log"], rtol=tol)
    pv = [np.array(l.split("=")[1].split(), dtype=float) for l in log.splitlines() if l.startswith("prior variances")]
    pp = [np.array(l.split("=")[1].split(), dtype=float) for l in log.splitlines() if l.startswith("prior probabilities")]
    assert np.allclose(pv[-1], g["prior_vars_last"], rtol=tol) and np.allclose(pp[-1], g["prior_probs_last"], rtol=tol, atol=1e-12)
    if mode != "onsager_warm":
        assert len([l for l in log.splitlines() if l.startswith("[CG] it = ")]) == int(g["cg_lines"])   # the reference's CG iteration count


def _modes_inputs(oracle, tmp_path):
    g, gl = golden("modes.npz"), golden("vamp_linear.npz")
    N, M, Nt = int(g["N"]), int(g["M"]), int(g["N_test"])
    bedp, phenp, bedt, phent = (str(tmp_path / n) for n in ("m.bed", "m.phen", "t.bed", "t.phen"))
    oracle.write_bed(bedp, oracle.synth_bed(int(g["seed"]), 0, M, N))
    oracle.write_phen(phenp, gl["y"])
    oracle.write_bed(bedt, oracle.synth_bed(int(g["seed_test"]), 0, M, Nt))
    oracle.write_phen(phent, g["y_test"])
    est = str(tmp_path / "est") + "/"
    os.makedirs(est, exist_ok=True)
    for it in range(1, 5):
        g[f"est_{it}"].astype(np.float64).tofile(f"{est}g_it_{it}.bin")
    g["r1_3"].astype(np.float64).tofile(f"{est}g_r1_it_3.bin")
    train = ["--bed-file", bedp, "--phen-files", phenp, "--N", str(N), "--Mt", str(M)]
    test = ["--bed-file-test", bedt, "--phen-files-test", phent, "--N-test", str(Nt), "--Mt-test", str(M)]
    return g, train, test, est, [str(a) for a in g["common"]]


def _val(log, key):
    return float([l for l in log.splitlines() if l.startswith(key)][-1].split("=")[1])


def _run(args):
    r = subprocess.run([EXE] + args, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


def test_run_mode_test_matches_reference(oracle, tmp_path):
    """--run-mode test (main_real.cpp:129-254) against the printed numbers of the reference's executable: one estimate file, and an
    iteration range with its 'max R2' / 'max ind' lines."""
    g, _, test, est, _ = _modes_inputs(oracle, tmp_path)
    log = _run(["--run-mode", "test", "--estimate-file", f"{est}g_it_4.bin"] + test)
    assert abs(_val(log, "test R2 = ") - float(g["test_R2"])) < TOL_FINAL
    assert abs(_val(log, "test l2 pred err^2 = ") / float(g["test_err2"]) - 1) < TOL_FINAL
    assert abs(_val(log, "y stdev^2 = ") / float(g["test_sd2"]) - 1) < 1e-5
    log = _run(["--run-mode", "test", "--estimate-file", f"{est}g_it_4.bin", "--test-iter-range", "1,4"] + test)
    line = [l for l in log.splitlines() if l.rstrip().endswith(",")][-1]
    assert np.allclose([float(x) for x in line.split(",") if x.strip()], g["range_R2"], atol=TOL_FINAL)
    assert abs(_val(log, "max R2 = ") - float(g["range_max_R2"])) < TOL_FINAL and int(_val(log, "max ind = ")) == int(g["range_max_ind"])


def test_run_mode_both_matches_reference(oracle, tmp_path):
    """--run-mode both (main_real.cpp:255-330): inference on the training bed, then the test-set R2 of the final estimate with the
    phenotype's intercept / scale restored; the test-set matrix is loaded while the training matrix (and its twin) is resident."""
    g, train, test, _, common = _modes_inputs(oracle, tmp_path)
    outd = str(tmp_path / "outb") + "/"
    log = _run(["--run-mode", "both", "--out-dir", outd, "--out-name", "g", "--iterations", "4"] + train + common + test)
    assert abs(_val(log, "test R2 = ") - float(g["both_R2"])) < TOL_FINAL
    assert abs(_val(log, "test l2 pred err^2 = ") / float(g["both_err2"]) - 1) < TOL_FINAL
    assert abs(_val(log, "intercept = ") - float(g["both_intercept"])) < 1e-5 and abs(_val(log, "scale = ") / float(g["both_scale"]) - 1) < 1e-5
    assert relerr(np.fromfile(outd + "g_it_4.bin"), g["both_x1_last"]) < TOL_FINAL


def test_run_mode_restart_matches_reference(oracle, tmp_path):
    """--run-mode restart (main_real.cpp:453-486): gam1 / gamw from the command line and r1 from a stored r1 file, which the
    reference divides by sqrt(N) once more on the way in (vamp.cpp:226-233) -- reproduced, not corrected."""
    g, train, _, est, common = _modes_inputs(oracle, tmp_path)
    outd = str(tmp_path / "outr") + "/"
    log = _run(["--run-mode", "restart", "--out-dir", outd, "--out-name", "g", "--iterations", "3", "--estimate-file", f"{est}g_r1_it_3.bin",
                "--gam1-init", repr(float(g["restart_gam1_init"])), "--gamw-init", repr(float(g["restart_gamw_init"]))] + train + common)
    for it in range(1, 4):
        ref = g[f"restart_x1_{it}"]
        got = np.fromfile(f"{outd}g_it_{it}.bin")
        assert relerr(got, ref) < TOL_FINAL, (it, relerr(got, ref))
    assert np.allclose(np.loadtxt(outd + "g_gam1s.csv"), g["restart_gam1s"], rtol=TOL_FINAL)
    assert np.allclose(np.loadtxt(outd + "g_R2trains.csv"), g["restart_R2trains"], rtol=TOL_FINAL, atol=1e-6)
    gamw = [float(l.split("=")[1]) for l in log.splitlines() if l.startswith("gamw = ")]
    assert np.allclose(gamw, g["restart_gamw_log"], rtol=TOL_FINAL)


def test_run_mode_predict_single_matches_reference(oracle, tmp_path):
    """--run-mode predict_single (main_real.cpp:555-594): X_test . estimate written as <out>_predict.csv."""
    g, _, test, est, _ = _modes_inputs(oracle, tmp_path)
    outd = str(tmp_path / "outp") + "/"
    os.makedirs(outd, exist_ok=True)
    args = [a for a in test if a not in ("--phen-files-test",) and not a.endswith("t.phen")]
    _run(["--run-mode", "predict_single", "--estimate-file", f"{est}g_it_4.bin", "--out-dir", outd, "--out-name", "g"] + args)
    got = np.loadtxt(outd + "g_predict.csv")
    ref = g["predict_single"]
    assert got.shape == ref.shape and np.allclose(got, ref, rtol=1e-4, atol=1e-5 * np.abs(ref).max())


def test_reuploaded_phenotype_keeps_aty_only_while_unchanged(oracle, tmp_path, monkeypatch):
    """The stepping interface (bench.py's end-to-end leg) re-sends y before every iteration.  vamp::upload_iteration_inputs keeps the
    cached A^T y (vamp.cpp:588 computes it every iteration) while the uploaded y is bit for bit the resident one and drops it as soon
    as one entry differs: same results to the last bit as a run that recomputes A^T y after every upload (GVB_ATY_CACHE=0), one sweep
    per iteration less while y is constant."""
    import ctypes
    import sys
    sys.path.insert(0, ROOT)
    import bench
    from gvamp_b200 import capi as C
    H = bench.load_host_lib()
    N, M, iters = 2048, 6000, 6
    bed = oracle.synth_bed(3, 0, M, N)
    rng = np.random.default_rng(9)
    y = rng.normal(size=N)
    y2 = y.copy()
    y2[5] += 1e-3
    f64p = ctypes.POINTER(ctypes.c_double)
    monkeypatch.setenv("GVB_NO_FILES", "1")
    argv = ["t", "--bed-file", "in-hbm", "--N", str(N), "--Mt", str(M), "--iterations", str(iters), "--CG-max-iter", "10", "--rho", "0.5",
            "--probs", "0.9,0.1", "--vars", "0,0.001", "--h2", "0.5", "--stop-criteria-thr", "1e-12", "--out-dir", str(tmp_path) + "/",
            "--out-name", "t", "--model", "linear", "--run-mode", "infere"]
    carr = (ctypes.c_char_p * len(argv))(*[a.encode() for a in argv])

    def run(ctx, dat, opt, cache, uploads):
        monkeypatch.setenv("GVB_ATY_CACHE", cache)
        vmp = H.gvbh_vamp_create(opt, M, 1e-6, 2.0)
        H.gvbh_vamp_linear_begin(vmp, dat)
        s0 = ctx.sweeps()
        for it in range(1, iters + 1):
            u = uploads.get(it)
            H.gvbh_vamp_linear_iteration(vmp, dat, it, u.ctypes.data_as(f64p) if u is not None else None, None)
        x = np.zeros(M)
        H.gvbh_vamp_linear_end(vmp, x.ctypes.data_as(f64p), M)
        sweeps = ctx.sweeps() - s0
        H.gvbh_vamp_destroy(vmp)
        return x, sweeps

    with C.Context(0) as ctx:
        ctx.load_host(bed, N)
        opt = H.gvbh_options_create(len(argv), carr)
        dat = H.gvbh_data_create_resident(ctx.h, y.ctypes.data_as(f64p), N, M, M, 0, 1.0)
        same = {it: y for it in range(1, iters + 1)}
        x_none, s_none = run(ctx, dat, opt, "1", {})
        x_keep, s_keep = run(ctx, dat, opt, "1", same)
        x_drop, s_drop = run(ctx, dat, opt, "0", same)
        assert np.array_equal(x_none, x_keep) and np.array_equal(x_none, x_drop)
        assert s_keep == s_none and s_drop == s_none + (iters - 1)
        change = dict(same)
        for it in range(4, iters + 1):
            change[it] = y2
        x_c1, s_c1 = run(ctx, dat, opt, "1", change)
        x_c0, s_c0 = run(ctx, dat, opt, "0", change)
        assert np.array_equal(x_c1, x_c0) and not np.array_equal(x_c1, x_none)
        assert s_c0 - s_c1 == iters - 2   # the only recomputation left is the one at the iteration that brought the new y
        H.gvbh_data_destroy(dat)
