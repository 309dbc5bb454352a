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


def _run_case(oracle, tmp_path, g, gen):
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
    env = dict(os.environ, GVB_KERNELS=gen)
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
