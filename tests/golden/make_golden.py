"""tests/golden/make_golden.py -- regenerates the committed golden vectors.

Runs ONLY in the build container (needs /root/reference compiled into oracle/_ref by
`make -C oracle ref`).  Inputs are produced by the seeded synthetic generator in
oracle/gvamp_oracle.c, outputs come from the UNMODIFIED reference (harness library and the
main_real executable).  The .npz files written here are what `-m "not gpu"` and `-m gpu` tests
compare against on machines where the reference does not exist.

    python tests/golden/make_golden.py
"""
import math
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from oracle import ref as R  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
PROBS = [0.9, 0.06, 0.04]
VARS = [0, 1e-4, 1e-3]


def case_matvec(tmp):
    """N % 4 != 0, 2 % missing genotypes, two phenotype NAs: stats / Ax / ATx incl. byte sub-ranges."""
    N, M, seed = 1003, 400, 11
    bed = O.synth_bed(seed, 0, M, N, miss_rate=0.02)
    bedp = os.path.join(tmp, "mv.bed")
    O.write_bed(bedp, bed)
    rng = np.random.default_rng(seed)
    y = rng.normal(size=N)
    phenp = os.path.join(tmp, "mv.phen")
    na_idx = [5, 700]
    O.write_phen(phenp, y, na_idx=na_idx)
    rd = R.RefData(bedp, N, M, phen_path=phenp)
    v = rng.normal(size=M)
    u = rng.normal(size=4 * rd.mbytes)
    u[N:] = 0.0
    mave, msig = rd.stats()
    np.savez_compressed(
        os.path.join(OUT, "matvec_n1003.npz"), N=N, M=M, seed=seed, miss_rate=0.02, y=y, na_idx=na_idx,
        mask4=rd.mask4(), nonas=rd.nonas(), phen=rd.filter_pheno(), intercept_scale=np.array(rd.intercept_scale()),
        mave=mave, msig=msig, v=v, u=u, Ax=rd.Ax(v), ATx=rd.ATx(u), SB=10, LB=100, Ax_sub=rd.Ax(v, 10, 100),
        ATx_sub=rd.ATx(u[:400], 10, 100), bed_sha=np.frombuffer(__import__("hashlib").sha256(bed.tobytes()).digest(), dtype=np.uint8))
    # alpha_scale != 1 and a shard with S != 0 of a larger marker set
    rd2 = R.RefData(bedp, N, 150, Mt=M, S=100, phen_path=phenp, alpha_scale=0.3)
    mave2, msig2 = rd2.stats()
    np.savez_compressed(os.path.join(OUT, "matvec_shard.npz"), N=N, M=150, Mt=M, S=100, alpha_scale=0.3, mave=mave2, msig=msig2,
                        Ax=rd2.Ax(v[100:250]), ATx=rd2.ATx(u))


def case_denoiser():
    N, M = 1000, 5000
    rng = np.random.default_rng(5)
    r1 = rng.normal(size=M) * np.where(rng.random(M) < 0.1, 2.0, 0.3)
    rv = R.RefVamp(N, M, M, PROBS, [v * N for v in VARS])
    out = {}
    for tag, gam1 in (("a", 3.0), ("b", 1e-6), ("c", 250.0)):
        g, gd = rv.g1(r1, gam1)
        out["g1_" + tag], out["g1d_" + tag], out["gam1_" + tag] = g, gd, gam1
    # EM update: plain, and one where two variances end within 50 % and merge (vamp.cpp:1054-1071)
    rv.set_prior(PROBS, [v * N for v in VARS])
    p, v = rv.update_prior(r1, 3.0)
    out["em_probs"], out["em_vars"] = p, v
    p23 = [0.8, 0.05, 0.05, 0.05, 0.05]
    v23 = [0, 0.1, 0.12, 1.0, 5.0]
    rv2 = R.RefVamp(N, M, M, p23, v23, EM_max_iter=5, EM_err_thr=1e-4)
    p, v = rv2.update_prior(r1, 3.0)
    out["em5_probs_in"], out["em5_vars_in"], out["em5_probs"], out["em5_vars"] = p23, v23, p, v
    rv3 = R.RefVamp(N, M, M, p23, v23, EM_max_iter=3, EM_err_thr=1e-4, learn_vars=0)
    p, v = rv3.update_prior(r1, 0.7)
    out["em5nl_probs"], out["em5nl_vars"] = p, v
    np.savez_compressed(os.path.join(OUT, "denoiser.npz"), r1=r1, N=N, probs=PROBS, vars=[v * N for v in VARS], **out)


def case_probit_pieces():
    rng = np.random.default_rng(9)
    x = np.concatenate([np.linspace(-30, 30, 241), rng.normal(size=100) * 5])
    rv = R.RefVamp(1000, 10, 10, PROBS, VARS)
    p = rng.normal(size=500) * 2
    y = (rng.random(500) < 0.5).astype(float)
    mc = rng.normal(size=500) * 0.3
    rv.set_state(1e-6, 0.0, 2.0, probit_var=1.0)
    g, gd = rv.g1_bin_class(p, 0.8, y, mc)
    eta = np.array([1.0])
    sim = np.empty(64)
    R.lib().ref_simulate(64, eta.ctypes.data_as(R.c_f64p), eta.ctypes.data_as(R.c_f64p), 1, 1, sim.ctypes.data_as(R.c_f64p))
    np.savez_compressed(os.path.join(OUT, "probit_pieces.npz"), x=x, erfcx=R.erfcx(x), p=p, y=y, mcov=mc, tau1=0.8, g=g, gd=gd,
                        simulate_unit=sim)


def case_cg(tmp):
    N, M, seed = 1000, 1500, 21
    bed = O.synth_bed(seed, 0, M, N)
    bedp = os.path.join(tmp, "cg.bed")
    O.write_bed(bedp, bed)
    rd = R.RefData(bedp, N, M)
    rv = R.RefVamp(N, M, M, PROBS, VARS, CG_max_iter=25, seed=4)
    rng = np.random.default_rng(seed)
    rhs = rng.normal(size=M)
    rv.set_state(1e-6, 0.7, 2.0)
    lm = rv.lmmse_mult(rd, rhs, 2.0)
    mu = rv.cg(rd, rhs, np.zeros(M), 2.0, 1)
    mu_warm = rv.cg(rd, rhs, mu * 0.9, 2.0, 1)
    a2, bern, invq = rv.onsager(rd, 0.7, 2.0)
    np.savez_compressed(os.path.join(OUT, "cg.npz"), N=N, M=M, seed=seed, rhs=rhs, gam2=0.7, tau=2.0, lmmse=lm, mu=mu,
                        mu_warm=mu_warm, alpha2=a2, bern=bern, invq=invq, vamp_seed=4, CG_max_iter=25)


def case_vamp_linear(tmp):
    """main_real.exe --run-mode infere --model linear, end to end (scalar build = the semantic spec)."""
    N, M, seed, h2, CV, iters = 1000, 2000, 3, 0.5, 200, 6
    bed = O.synth_bed(seed, 0, M, N)
    bedp = os.path.join(tmp, "v.bed")
    O.write_bed(bedp, bed)
    ds0 = O.Dataset(bed, N)
    beta = O.synth_beta(seed, M, CV, h2)
    y = ds0.Ax(beta * math.sqrt(N))[:N] + O.synth_noise(seed, N, h2)
    phenp = os.path.join(tmp, "v.phen")
    O.write_phen(phenp, y)
    outd = os.path.join(tmp, "out") + "/"
    args = ["--run-mode", "infere", "--model", "linear", "--bed-file", bedp, "--phen-files", phenp, "--N", str(N), "--Mt", str(M),
            "--out-dir", outd, "--out-name", "g", "--iterations", str(iters), "--CG-max-iter", "20", "--rho", "0.5",
            "--probs", ",".join(map(str, PROBS)), "--vars", ",".join(map(str, VARS)), "--h2", str(h2), "--seed", "1"]
    env = dict(os.environ, OMP_NUM_THREADS="4")
    log = subprocess.run([R.exe("main_real_scalar.exe")] + args, check=True, capture_output=True, text=True, env=env).stdout
    out = dict(N=N, M=M, seed=seed, h2=h2, CV=CV, iterations=iters, args=np.array(args[8:]), y=y, beta=beta)
    for it in range(1, iters + 1):
        out[f"x1_{it}"] = np.fromfile(f"{outd}g_it_{it}.bin")
        out[f"r1_{it}"] = np.fromfile(f"{outd}g_r1_it_{it}.bin")
        out[f"r2_{it}"] = np.fromfile(f"{outd}g_r2_it_{it}.bin")
        out[f"x2_{it}"] = np.fromfile(f"{outd}g_it_{it}_x2_hat.bin")
    for nm in ("gam1s", "gam2s", "R2trains"):
        out[nm] = np.loadtxt(f"{outd}g_{nm}.csv")
    out["z1_text_last"] = np.loadtxt(f"{outd}g_z1_it_{iters}.csv")
    gamw = [float(l.split("=")[1]) for l in log.splitlines() if l.startswith("gamw = ")]
    alpha2 = [float(l.split("=")[1]) for l in log.splitlines() if l.startswith("alpha2 = ")]
    pv = [np.array(l.split("=")[1].split(), dtype=float) for l in log.splitlines() if l.startswith("prior variances")]
    pp = [np.array(l.split("=")[1].split(), dtype=float) for l in log.splitlines() if l.startswith("prior probabilities")]
    out.update(gamw_log=np.array(gamw), alpha2_log=np.array(alpha2), prior_vars_last=pv[-1], prior_probs_last=pp[-1])
    # MANVECT build of the same run: must agree (N % 4 == 0, no NAs) -- BASELINE.md section 3
    outd2 = os.path.join(tmp, "out2") + "/"
    args2 = [a if a != outd else outd2 for a in args]
    subprocess.run([R.exe("main_real_manvect.exe")] + args2, check=True, capture_output=True, text=True, env=env)
    out["x1_last_manvect"] = np.fromfile(f"{outd2}g_it_{iters}.bin")
    np.savez_compressed(os.path.join(OUT, "vamp_linear.npz"), **out)
    with open(os.path.join(OUT, "vamp_linear.log"), "w") as fh:
        fh.write("\n".join(l for l in log.splitlines() if not l.startswith("[CG")) + "\n")


def case_vamp_probit(tmp):
    """main_real_probit.exe --run-mode infere --model bin_class with C = 3 covariates, end to end
    (vamp_probit.cpp:20-658; y = 1[U <= Phi(g + Z eta)] as in sim_probit.cpp:195-205)."""
    from scipy.special import ndtr
    N, M, seed, h2, CV, iters, C = 1000, 2000, 13, 0.5, 200, 5, 3
    bed = O.synth_bed(seed, 0, M, N)
    bedp = os.path.join(tmp, "p.bed")
    O.write_bed(bedp, bed)
    ds0 = O.Dataset(bed, N)
    beta = O.synth_beta(seed, M, CV, h2)
    rng = np.random.default_rng(seed)
    Z = rng.normal(size=(N, C))
    eta = np.array([0.25, -0.25, 0.25])
    g = ds0.Ax(beta * math.sqrt(N))[:N] / math.sqrt(1.0 - h2)
    y = (rng.random(N) <= ndtr(g + Z @ eta)).astype(float)
    phenp, covp = os.path.join(tmp, "p.phen"), os.path.join(tmp, "p.cov")
    O.write_phen(phenp, y)
    with open(covp, "w") as fh:
        for i in range(N):
            fh.write(" ".join(repr(float(z)) for z in Z[i]) + "\n")   # C values per line, no id columns (data.cpp:286-331)
    outd = os.path.join(tmp, "outp") + "/"
    args = ["--run-mode", "infere", "--model", "bin_class", "--bed-file", bedp, "--phen-files", phenp, "--N", str(N), "--Mt", str(M),
            "--out-dir", outd, "--out-name", "p", "--iterations", str(iters), "--CG-max-iter", "20", "--rho", "0.5",
            "--probs", ",".join(map(str, PROBS)), "--vars", ",".join(map(str, VARS)), "--seed", "1", "--cov-file", covp, "--C", str(C)]
    env = dict(os.environ, OMP_NUM_THREADS="4")
    log = subprocess.run([R.exe("main_real_probit_scalar.exe")] + args, check=True, capture_output=True, text=True, env=env).stdout
    out = dict(N=N, M=M, seed=seed, h2=h2, CV=CV, iterations=iters, C=C, args=np.array(args[8:]), y=y, Z=Z, beta=beta)
    n_it = 0
    for it in range(1, iters + 1):
        f = f"{outd}p_probit_it_{it}.bin"
        if not os.path.exists(f):
            break
        out[f"x1_{it}"] = np.fromfile(f)
        f = f"{outd}p_probit_r1_it_{it}.bin"
        if os.path.exists(f):
            out[f"r1_{it}"] = np.fromfile(f)
        n_it = it
    out["iterations_done"] = n_it

    def grab(prefix):
        return np.array([float(l.split("=")[-1]) for l in log.splitlines() if l.startswith(prefix)])
    for key, prefix in (("gam1_log", "gam1 = "), ("gam2_log", "gam2 = "), ("alpha2_log", "alpha2 = "), ("tau1_log", "tau1 = "), ("tau2_log", "tau2 = "),
                        ("beta1_log", "beta1 = "), ("beta2_log", "beta2 = "), ("eta1_log", "eta1 = ")):
        out[key] = grab(prefix)
    cov = [l for l in log.splitlines() if l.startswith("cov_eff[")]
    out["cov_eff_log"] = np.array([float(tok.split("=")[1]) for l in cov for tok in l.split(",") if "=" in tok])
    np.savez_compressed(os.path.join(OUT, "vamp_probit.npz"), **out)
    with open(os.path.join(OUT, "vamp_probit.log"), "w") as fh:
        fh.write("\n".join(l for l in log.splitlines() if not l.startswith("[CG")) + "\n")


def case_config1(tmp):
    """BASELINE.json configs[0]: synthetic N=10,000 x M=20,000, h2=0.5, CV=2,000, linear model, 10 iterations of the reference's
    main_real.exe (MANVECT build = the reference's own build line; 1 rank through the MPI shim instead of mpirun -np 2, which only
    changes summation order and the Onsager probe's seed offset S).  Stored: what the north_star bounds -- the final signal
    estimate, the learned prior, gamw, and the per-iteration scalars; the 50 MB bed is regenerated from the seed."""
    N, M, seed, h2, CV, iters = 10_000, 20_000, 101, 0.5, 2000, 10
    bed = O.synth_bed(seed, 0, M, N)
    bedp = os.path.join(tmp, "c1.bed")
    O.write_bed(bedp, bed)
    ds0 = O.Dataset(bed, N)
    beta = O.synth_beta(seed, M, CV, h2)
    y = ds0.Ax(beta * math.sqrt(N))[:N] + O.synth_noise(seed, N, h2)
    phenp = os.path.join(tmp, "c1.phen")
    O.write_phen(phenp, y)
    outd = os.path.join(tmp, "c1out") + "/"
    probs, vars_ = [0.9, 0.05, 0.03, 0.02], [0, 1e-5, 1e-4, 1e-3]     # Mt <= 50000: the prior must be passed (utilities.cpp:98-99)
    args = ["--run-mode", "infere", "--model", "linear", "--bed-file", bedp, "--phen-files", phenp, "--N", str(N), "--Mt", str(M),
            "--out-dir", outd, "--out-name", "c1", "--iterations", str(iters), "--CG-max-iter", "20", "--rho", "0.5",
            "--probs", ",".join(map(str, probs)), "--vars", ",".join(map(str, vars_)), "--h2", str(h2), "--seed", "1"]
    env = dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count()))
    import time
    t0 = time.time()
    log = subprocess.run([R.exe("main_real_manvect.exe")] + args, check=True, capture_output=True, text=True, env=env).stdout
    wall = time.time() - t0
    it_times = [float(l.split("=")[1]) for l in log.splitlines() if l.startswith("total iteration time")]
    n_it = len([f for f in os.listdir(outd) if f.startswith("c1_it_") and f.endswith(".bin") and "x2" not in f])
    out = dict(N=N, M=M, seed=seed, h2=h2, CV=CV, iterations=iters, iterations_done=n_it, args=np.array(args[8:]),
               y=y.astype(np.float32).astype(np.float64) * 0 + y, ref_wall_s=wall, ref_iter_s=np.array(it_times), ref_threads=os.cpu_count())
    out["x1_last"] = np.fromfile(f"{outd}c1_it_{n_it}.bin")
    out["x1_first"] = np.fromfile(f"{outd}c1_it_1.bin")
    out["x1_mid"] = np.fromfile(f"{outd}c1_it_{max(1, n_it // 2)}.bin")
    for nm in ("gam1s", "gam2s", "R2trains"):
        out[nm] = np.loadtxt(f"{outd}c1_{nm}.csv")
    gamw = [float(l.split("=")[1]) for l in log.splitlines() if l.startswith("gamw = ")]
    alpha2 = [float(l.split("=")[1]) for l in log.splitlines() if l.startswith("alpha2 = ")]
    pv = [np.array(l.split("=")[1].split(), dtype=float) for l in log.splitlines() if l.startswith("prior variances")]
    pp = [np.array(l.split("=")[1].split(), dtype=float) for l in log.splitlines() if l.startswith("prior probabilities")]
    out.update(gamw_log=np.array(gamw), alpha2_log=np.array(alpha2), prior_vars_last=pv[-1], prior_probs_last=pp[-1])
    np.savez_compressed(os.path.join(OUT, "config1.npz"), **out)
    print(f"config1: reference wall {wall:.1f} s, iterations {n_it}, per-iteration {it_times}")
    print("R2trains", out["R2trains"], "gamw", gamw)


def case_xxt(tmp):
    """--use-XXT-denoiser 1: people statistics (data.cpp:548-640), one CG_solverAAT solve through the harness and a complete
    main_real.exe run with the N-space LMMSE step (denoiserXXT.cpp), on the N=1003 / 2 % missing / phenotype-NA case and on the
    N=1000 x M=2000 linear case."""
    g = np.load(os.path.join(OUT, "matvec_n1003.npz"))
    N, M, seed = int(g["N"]), int(g["M"]), int(g["seed"])
    bed = O.synth_bed(seed, 0, M, N, miss_rate=float(g["miss_rate"]))
    bedp, phenp = os.path.join(tmp, "x.bed"), os.path.join(tmp, "x.phen")
    O.write_bed(bedp, bed)
    O.write_phen(phenp, g["y"], na_idx=[int(i) for i in g["na_idx"]])
    rd = R.RefData(bedp, N, M, phen_path=phenp)
    mave_p, msig_p, numb_p = rd.people_stats()
    rng = np.random.default_rng(77)
    rhs = rng.normal(size=N) * ((rd.mask4()[:, None] >> np.arange(4)) & 1).reshape(-1)[:N]
    rv = R.RefVamp(N, M, M, PROBS, VARS, iterations=1, CG_max_iter=30)
    rv.set_state(1.0, 0.7, 2.0)
    u = rv.cg_aat(rd, rhs, np.zeros(4 * rd.mbytes), 2.0)
    out = dict(N=N, M=M, mave_people=mave_p, msig_people=msig_p, numb_people=numb_p, cg_rhs=rhs, cg_tau=2.0, cg_gam2=0.7, cg_max_iter=30, cg_u=u)
    # end to end
    gl = np.load(os.path.join(OUT, "vamp_linear.npz"))
    N2, M2, iters = int(gl["N"]), int(gl["M"]), 5
    bed2 = O.synth_bed(int(gl["seed"]), 0, M2, N2)
    bedp2, phenp2 = os.path.join(tmp, "xl.bed"), os.path.join(tmp, "xl.phen")
    O.write_bed(bedp2, bed2)
    O.write_phen(phenp2, gl["y"])
    outd = os.path.join(tmp, "xout") + "/"
    args = ["--run-mode", "infere", "--model", "linear", "--bed-file", bedp2, "--phen-files", phenp2, "--N", str(N2), "--Mt", str(M2),
            "--out-dir", outd, "--out-name", "x", "--iterations", str(iters), "--CG-max-iter", "20", "--rho", "0.5",
            "--probs", ",".join(map(str, PROBS)), "--vars", ",".join(map(str, VARS)), "--h2", "0.5", "--seed", "1", "--use-XXT-denoiser", "1"]
    log = subprocess.run([R.exe("main_real_scalar.exe")] + args, check=True, capture_output=True, text=True, env=dict(os.environ, OMP_NUM_THREADS="4")).stdout
    out.update(e2e_args=np.array(args[8:]), e2e_iterations=iters)
    for it in range(1, iters + 1):
        out[f"x1_{it}"] = np.fromfile(f"{outd}x_it_{it}.bin")
        out[f"x2_{it}"] = np.fromfile(f"{outd}x_it_{it}_x2_hat.bin")
    out["gam1s"] = np.loadtxt(f"{outd}x_gam1s.csv")
    out["gamw_log"] = np.array([float(l.split("=")[1]) for l in log.splitlines() if l.startswith("gamw = ")])
    np.savez_compressed(os.path.join(OUT, "xxt.npz"), **out)
    print("xxt: CG iterations in log:", len([l for l in log.splitlines() if l.startswith("[CG] it")]), "gamw", out["gamw_log"][-3:])


def case_pvals(tmp):
    """LOO and LOCO association p-values (data.cpp:1108-1353) on the N=1003 case (2 % missing genotypes, phenotype NAs)
    with a 5-chromosome .bim (incl. "X" -> 23).  The Student-t tail comes from oracle/shims (incomplete beta), which the
    oracle cross-checks against scipy."""
    N, M, seed = 1003, 400, 11
    bed = O.synth_bed(seed, 0, M, N, miss_rate=0.02)
    bedp, phenp, bimp = os.path.join(tmp, "pv.bed"), os.path.join(tmp, "pv.phen"), os.path.join(tmp, "pv.bim")
    O.write_bed(bedp, bed)
    rng = np.random.default_rng(seed + 1)
    beta = np.zeros(M)
    causal = rng.choice(M, size=12, replace=False)
    beta[causal] = rng.normal(0, math.sqrt(0.5 / 12), size=12)
    ds0 = O.Dataset(bed, N)
    y = ds0.Ax(beta * math.sqrt(N))[:N] + rng.normal(0, math.sqrt(0.5), size=N)
    na_idx = [5, 700, 701]
    O.write_phen(phenp, y, na_idx=na_idx)
    chrom_of = np.array([1] * 100 + [2] * 120 + [7] * 80 + [22] * 60 + [23] * 40)
    with open(bimp, "w") as fh:
        for j in range(M):
            ch = "X" if chrom_of[j] == 23 else str(chrom_of[j])
            fh.write(f"{ch}\trs{j}\t0\t{1000 + j}\tA\tG\n")
    rd = R.RefData(bedp, N, M, phen_path=phenp, bim_path=bimp)
    x1 = (beta + rng.normal(0, 0.01, size=M) * (rng.random(M) < 0.2)) * math.sqrt(N)   # an imperfect estimate, scaled like main_real.cpp:390
    z1 = rd.Ax(x1)
    yf = rd.filter_pheno()
    loo = rd.pvals(z1, yf, x1, os.path.join(tmp, "pv_loo.bin"), loco=False)
    loco = rd.pvals(z1, yf, x1, os.path.join(tmp, "pv"), loco=True)
    np.savez_compressed(os.path.join(OUT, "pvals.npz"), N=N, M=M, seed=seed, miss_rate=0.02, y=y, na_idx=na_idx, chrom=chrom_of, x1=x1, z1=z1,
                        y_filtered=yf, pvals_loo=loo, pvals_loco=loco, mask4=rd.mask4(), nonas=rd.nonas())
    print("pvals: LOO min %.3e, LOCO min %.3e" % (loo.min(), loco.min()))


def default_prior_23(Mt):
    """The reference's built-in 23-component prior (utilities.cpp:96-129) written out, so that it can be passed to a run with Mt <= 50000."""
    p = min(50000.0 / Mt, 1.0) / (2 - 1.0 / 2 ** 21)
    probs = [1 - min(50000.0 / Mt, 1.0)]
    for _ in range(22):
        probs.append(p)
        p /= 2
    step = 10 ** (math.log10(1e2 / 1e-5) / 21)
    vars_, v = [0.0], 1e-5
    for _ in range(22):
        vars_.append(v)
        v *= step
    return probs, vars_


def _grab_log(log):
    f = lambda key: np.array([float(l.split("=")[1]) for l in log.splitlines() if l.startswith(key)])
    pv = [np.array(l.split("=")[1].split(), dtype=float) for l in log.splitlines() if l.startswith("prior variances")]
    pp = [np.array(l.split("=")[1].split(), dtype=float) for l in log.splitlines() if l.startswith("prior probabilities")]
    return dict(gamw_log=f("gamw = "), alpha2_log=f("alpha2 = "), prior_vars_last=pv[-1], prior_probs_last=pp[-1])


def case_vamp_stress(tmp):
    """An ill-conditioned linear run: M/N = 5, heavy damping (rho = 0.05, the dnanexus setting), 30 iterations, a sparse
    23-component prior of the reference's built-in shape (sized as for Mt = 500k markers: 10 % of the markers carry an effect), few causal markers."""
    N, M, seed, h2, CV, iters = 600, 3000, 21, 0.6, 60, 30
    bed = O.synth_bed(seed, 0, M, N)
    bedp = os.path.join(tmp, "s.bed")
    O.write_bed(bedp, bed)
    ds0 = O.Dataset(bed, N)
    beta = O.synth_beta(seed, M, CV, h2)
    y = ds0.Ax(beta * math.sqrt(N))[:N] + O.synth_noise(seed, N, h2)
    phenp = os.path.join(tmp, "s.phen")
    O.write_phen(phenp, y)
    probs, vars_ = default_prior_23(500000)
    outd = os.path.join(tmp, "outs") + "/"
    args = ["--run-mode", "infere", "--model", "linear", "--bed-file", bedp, "--phen-files", phenp, "--N", str(N), "--Mt", str(M),
            "--out-dir", outd, "--out-name", "g", "--iterations", str(iters), "--CG-max-iter", "40", "--rho", "0.05",
            "--probs", ",".join(repr(p) for p in probs), "--vars", ",".join(repr(v) for v in vars_), "--h2", str(h2), "--seed", "7",
            "--stop-criteria-thr", "1e-9"]
    env = dict(os.environ, OMP_NUM_THREADS="4")
    log = subprocess.run([R.exe("main_real_scalar.exe")] + args, check=True, capture_output=True, text=True, env=env).stdout
    done = len([l for l in log.splitlines() if l.startswith("iteration = ")])
    out = dict(N=N, M=M, seed=seed, h2=h2, CV=CV, iterations=iters, iterations_done=done, args=np.array(args[8:]), y=y, beta=beta)
    for it in (1, 2, 3, 5, 10, 20, done):
        if it <= done:
            out[f"x1_{it}"] = np.fromfile(f"{outd}g_it_{it}.bin")
            out[f"x2_{it}"] = np.fromfile(f"{outd}g_it_{it}_x2_hat.bin")
            out[f"r1_{it}"] = np.fromfile(f"{outd}g_r1_it_{it}.bin")
    for nm in ("gam1s", "gam2s", "R2trains"):
        out[nm] = np.loadtxt(f"{outd}g_{nm}.csv")
    out.update(_grab_log(log))
    out["cg_lines"] = len([l for l in log.splitlines() if l.startswith("[CG] it = ")])
    np.savez_compressed(os.path.join(OUT, "vamp_stress.npz"), **out)
    print("stress: %d iterations, %d CG lines, corr(x1, beta) = %.3f, R2 last %.4f" %
          (done, out["cg_lines"], np.corrcoef(out[f"x1_{done}"], beta)[0, 1], out["R2trains"][-1]))


def case_modes(tmp):
    """The run modes around the hot path (main_real.cpp:129-330, 453-594): test (one estimate file and an iteration range), both,
    restart (with its divide-by-sqrt(N) of the stored r1, vamp.cpp:226-233), predict_single -- outputs of the reference's executable."""
    g = np.load(os.path.join(OUT, "vamp_linear.npz"))
    N, M, seed, h2 = int(g["N"]), int(g["M"]), int(g["seed"]), float(g["h2"])
    bed = O.synth_bed(seed, 0, M, N)
    bedp, phenp = os.path.join(tmp, "m.bed"), os.path.join(tmp, "m.phen")
    O.write_bed(bedp, bed)
    O.write_phen(phenp, g["y"])
    Nt, seed_t = 800, 99
    bed_t = O.synth_bed(seed_t, 0, M, Nt)
    y_t = O.Dataset(bed_t, Nt).Ax(g["beta"] * math.sqrt(Nt))[:Nt] + O.synth_noise(seed_t, Nt, h2)
    bedt, phent = os.path.join(tmp, "t.bed"), os.path.join(tmp, "t.phen")
    O.write_bed(bedt, bed_t)
    O.write_phen(phent, y_t)   # no NAs: the reference's test modes read the raw phenotype and an NA turns every R2 into NaN
    env = dict(os.environ, OMP_NUM_THREADS="4")
    exe = R.exe("main_real_scalar.exe")
    run = lambda a: subprocess.run([exe] + a, check=True, capture_output=True, text=True, env=env).stdout
    val = lambda log, key: float([l for l in log.splitlines() if l.startswith(key)][-1].split("=")[1])
    common = ["--model", "linear", "--CG-max-iter", "20", "--rho", "0.5", "--probs", ",".join(map(str, PROBS)), "--vars", ",".join(map(str, VARS)),
              "--h2", str(h2), "--seed", "1"]
    out = dict(N=N, M=M, seed=seed, N_test=Nt, seed_test=seed_t, y_test=y_t, common=np.array(common))
    # estimate files of a 4-iteration run
    outd = os.path.join(tmp, "outm") + "/"
    run(["--run-mode", "infere", "--bed-file", bedp, "--phen-files", phenp, "--N", str(N), "--Mt", str(M), "--out-dir", outd, "--out-name", "g",
         "--iterations", "4"] + common)
    for it in range(1, 5):
        out[f"est_{it}"] = np.fromfile(f"{outd}g_it_{it}.bin")
    out["r1_3"] = np.fromfile(f"{outd}g_r1_it_3.bin")
    test_args = ["--bed-file-test", bedt, "--phen-files-test", phent, "--N-test", str(Nt), "--Mt-test", str(M)]
    # test, one file
    log = run(["--run-mode", "test", "--estimate-file", f"{outd}g_it_4.bin"] + test_args)
    out.update(test_R2=val(log, "test R2 = "), test_err2=val(log, "test l2 pred err^2 = "), test_sd2=val(log, "y stdev^2 = "))
    # test, iteration range 1..4
    log = run(["--run-mode", "test", "--estimate-file", f"{outd}g_it_4.bin", "--test-iter-range", "1,4"] + test_args)
    line = [l for l in log.splitlines() if l.rstrip().endswith(",") and "," in l][-1]
    out.update(range_R2=np.array([float(x) for x in line.split(",") if x.strip()]), range_max_R2=val(log, "max R2 = "), range_max_ind=val(log, "max ind = "))
    # both
    outb = os.path.join(tmp, "outb") + "/"
    log = run(["--run-mode", "both", "--bed-file", bedp, "--phen-files", phenp, "--N", str(N), "--Mt", str(M), "--out-dir", outb, "--out-name", "g",
               "--iterations", "4"] + common + test_args)
    out.update(both_R2=val(log, "test R2 = "), both_err2=val(log, "test l2 pred err^2 = "), both_intercept=val(log, "intercept = "), both_scale=val(log, "scale = "),
               both_x1_last=np.fromfile(f"{outb}g_it_4.bin"))
    # restart from iteration 3's r1 with given precisions
    outr = os.path.join(tmp, "outr") + "/"
    gam1_init, gamw_init = float(g["gam1s"][2]), float(g["gamw_log"][5])
    log = run(["--run-mode", "restart", "--bed-file", bedp, "--phen-files", phenp, "--N", str(N), "--Mt", str(M), "--out-dir", outr, "--out-name", "g",
               "--iterations", "3", "--estimate-file", f"{outd}g_r1_it_3.bin", "--gam1-init", repr(gam1_init), "--gamw-init", repr(gamw_init)] + common)
    out.update(restart_gam1_init=gam1_init, restart_gamw_init=gamw_init, restart_gam1s=np.loadtxt(f"{outr}g_gam1s.csv"),
               restart_R2trains=np.loadtxt(f"{outr}g_R2trains.csv"), restart_gamw_log=_grab_log(log)["gamw_log"])
    for it in range(1, 4):
        out[f"restart_x1_{it}"] = np.fromfile(f"{outr}g_it_{it}.bin")
    # predict_single
    outp = os.path.join(tmp, "outp") + "/"
    os.makedirs(outp, exist_ok=True)
    run(["--run-mode", "predict_single", "--bed-file-test", bedt, "--N-test", str(Nt), "--Mt-test", str(M), "--estimate-file", f"{outd}g_it_4.bin",
         "--out-dir", outp, "--out-name", "g"])
    out["predict_single"] = np.loadtxt(f"{outp}g_predict.csv")
    np.savez_compressed(os.path.join(OUT, "modes.npz"), **out)
    print("modes: test R2 %.6f, range %s, both R2 %.6f, restart R2 %s" % (out["test_R2"], out["range_R2"], out["both_R2"], out["restart_R2trains"]))


if __name__ == "__main__":
    assert R.available(), "build oracle/_ref first: make -C oracle ref"
    with tempfile.TemporaryDirectory() as tmp:
        case_matvec(tmp)
        case_denoiser()
        case_probit_pieces()
        case_cg(tmp)
        case_vamp_linear(tmp)
        case_vamp_probit(tmp)
        case_pvals(tmp)
        case_config1(tmp)
        case_xxt(tmp)
        case_vamp_stress(tmp)
        case_modes(tmp)
    print("golden vectors written to", OUT)
