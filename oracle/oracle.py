"""oracle/oracle.py -- TEST INFRASTRUCTURE (the parity oracle); never imported by the product path.

numpy / ctypes restatement of the gVAMP reference's hot path.  The bed sweeps (stats, Ax, ATx,
counts, the synthetic generator) live in oracle/gvamp_oracle.c (liboracle.so); the VAMP control
flow, denoiser, EM prior update and CG are restated here in numpy.  Every function cites the
reference file:line (relative to /root/reference) it follows.

Pinning: tests/test_oracle_vs_ref.py checks every function here against the UNMODIFIED reference
compiled into oracle/_ref (in the build container), and tests/test_oracle_golden.py checks it
against the committed outputs of that reference in tests/golden/ (runs anywhere).
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_f64p = ctypes.POINTER(ctypes.c_double)
c_i64p = ctypes.POINTER(ctypes.c_int64)


def build(force: bool = False) -> str:
    """Compile liboracle.so (plain C, gcc) next to this file."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "gvamp_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-o", so, src, "-lm"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        L.orc_counts.argtypes = [c_u8p, ctypes.c_long, ctypes.c_long, ctypes.c_long, c_u8p, c_i64p]
        L.orc_stats.argtypes = [c_u8p, ctypes.c_long, ctypes.c_long, c_u8p, ctypes.c_int, ctypes.c_double, c_f64p, c_f64p]
        L.orc_ATx.argtypes = [c_u8p, ctypes.c_long, ctypes.c_long, ctypes.c_long, c_f64p, c_f64p, c_f64p,
                              ctypes.c_long, ctypes.c_long, c_f64p]
        L.orc_Ax.argtypes = [c_u8p, ctypes.c_long, ctypes.c_long, ctypes.c_long, c_f64p, c_f64p, c_f64p, c_u8p,
                             ctypes.c_long, ctypes.c_long, c_f64p]
        L.orc_filter_pheno.argtypes = [c_f64p, ctypes.c_long, c_u8p, c_f64p]
        L.orc_synth_bed.argtypes = [ctypes.c_uint64, ctypes.c_long, ctypes.c_long, ctypes.c_long, ctypes.c_uint, c_u8p]
        L.orc_marker_freq.argtypes = [ctypes.c_uint64, ctypes.c_uint64]
        L.orc_marker_freq.restype = ctypes.c_double
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(t)


# --------------------------------------------------------------------------------------------
# work partition, utilities.cpp:259-291
# --------------------------------------------------------------------------------------------
def divide_work(Mt: int, nranks: int, rank: int):
    """(M, S) of `rank`: the first Mt % nranks ranks own Mt // nranks + 1 contiguous markers."""
    modu, size = Mt % nranks, Mt // nranks
    lens = [size + 1 if i < modu else size for i in range(nranks)]
    return lens[rank], sum(lens[:rank])


# --------------------------------------------------------------------------------------------
# phenotype file, data.cpp:128-192
# --------------------------------------------------------------------------------------------
def read_phen(path: str, N: int):
    """PLINK .phen: FID IID value ('NA' allowed).  Returns (phen, mask4, nonas, intercept, scale).
    y is multiplied by 1/sd but NOT centred (data.cpp:181-182); NA entries hold DBL_MAX * scale."""
    vals, na = [], []
    with open(path) as fh:
        for line in fh:
            tok = line.split()
            if not tok:
                continue
            if tok[2] == "NA":
                vals.append(np.finfo(np.float64).max)
                na.append(True)
            else:
                vals.append(float(tok[2]))
                na.append(False)
    assert len(vals) == N
    phen = np.array(vals, dtype=np.float64)
    na = np.array(na)
    mask4 = make_mask4(N, ~na)
    nonas = int((~na).sum())
    # sequential (left-to-right) sums like the reference's scalar loops, so intercept/scale are bit-identical
    avg = float(np.cumsum(phen[~na])[-1]) / nonas
    sqn = math.sqrt((nonas - 1) / float(np.cumsum((phen[~na] - avg) * (phen[~na] - avg))[-1]))
    with np.errstate(over="ignore"):
        phen = phen * sqn
    return phen, mask4, nonas, avg, sqn


def make_mask4(N: int, present=None):
    """mask4 (data.cpp:136-169 / :86-100): low nibble, bit k <=> individual 4j+k has a phenotype; pads cleared."""
    mbytes = (N + 3) // 4
    pres = np.zeros(4 * mbytes, dtype=np.uint8)
    pres[:N] = 1 if present is None else np.asarray(present, dtype=np.uint8)
    pres = pres.reshape(mbytes, 4)
    return (pres[:, 0] | (pres[:, 1] << 1) | (pres[:, 2] << 2) | (pres[:, 3] << 3)).astype(np.uint8)


# --------------------------------------------------------------------------------------------
# synthetic data; byte-identical to gvamp_b200/csrc/synth.cu
# --------------------------------------------------------------------------------------------
def synth_bed(seed: int, j0: int, M: int, N: int, miss_rate: float = 0.0) -> np.ndarray:
    mbytes = (N + 3) // 4
    out = np.empty((M, mbytes), dtype=np.uint8)
    lib().orc_synth_bed(seed, j0, M, N, int(round(miss_rate * 65536)), _p(out, c_u8p))
    return out


def write_bed(path: str, bed: np.ndarray):
    with open(path, "wb") as fh:
        fh.write(bytes([0x6C, 0x1B, 0x01]))
        fh.write(np.ascontiguousarray(bed).tobytes())


def write_phen(path: str, y, na_idx=()):
    na = set(int(i) for i in na_idx)
    with open(path, "w") as fh:
        for i, v in enumerate(y):
            fh.write(f"{i} {i} " + ("NA" if i in na else repr(float(v))) + "\n")


def synth_beta(seed: int, Mt: int, CV: int, h2: float) -> np.ndarray:
    """CV causal markers chosen uniformly, effects ~ N(0, h2/CV) (sim.cpp:78-79)."""
    rng = np.random.Generator(np.random.Philox(seed + 7919))
    idx = rng.choice(Mt, size=CV, replace=False)
    beta = np.zeros(Mt)
    beta[idx] = rng.normal(0.0, math.sqrt(h2 / CV), size=CV)
    return beta


def synth_noise(seed: int, N: int, h2: float) -> np.ndarray:
    rng = np.random.Generator(np.random.Philox(seed + 104729))
    return rng.normal(0.0, math.sqrt(1.0 - h2), size=N)


# --------------------------------------------------------------------------------------------
# class data (data.hpp / data.cpp), bed path only
# --------------------------------------------------------------------------------------------
class Dataset:
    """Shard [S, S+M) of an N x Mt genotype matrix, SNP-major packed bytes `bed` of shape (M, mbytes)."""

    def __init__(self, bed: np.ndarray, N: int, phen=None, mask4=None, nonas=None, alpha_scale: float = 1.0,
                 Mt: int | None = None, S: int = 0):
        self.bed = np.ascontiguousarray(bed, dtype=np.uint8)
        self.M, self.mbytes = self.bed.shape
        assert self.mbytes == (N + 3) // 4
        self.N, self.Mt, self.S = N, (Mt if Mt is not None else self.M), S
        self.phen = np.zeros(N) if phen is None else np.asarray(phen, dtype=np.float64)
        self.mask4 = make_mask4(N) if mask4 is None else np.ascontiguousarray(mask4, dtype=np.uint8)
        self.nonas = N if nonas is None else nonas
        self.alpha_scale = alpha_scale
        self.mave = np.empty(self.M)
        self.msig = np.empty(self.M)
        self.compute_markers_statistics()

    # data.cpp:392-485 (scalar branch)
    def compute_markers_statistics(self):
        lib().orc_stats(_p(self.bed, c_u8p), self.M, self.mbytes, _p(self.mask4, c_u8p), self.nonas, self.alpha_scale,
                        _p(self.mave, c_f64p), _p(self.msig, c_f64p))

    def counts(self) -> np.ndarray:
        out = np.empty((self.M, 8), dtype=np.int64)
        lib().orc_counts(_p(self.bed, c_u8p), self.M, self.mbytes, self.N, _p(self.mask4, c_u8p), _p(out, c_i64p))
        return out

    # data.cpp:810-835
    def ATx(self, u, SB: int = 0, LB: int | None = None) -> np.ndarray:
        LB = self.mbytes if LB is None else LB
        uu = np.zeros(4 * LB)
        u = np.asarray(u, dtype=np.float64)
        uu[: min(len(u), 4 * LB)] = u[: 4 * LB]
        out = np.empty(self.M)
        lib().orc_ATx(_p(self.bed, c_u8p), self.M, self.mbytes, self.N, _p(uu, c_f64p), _p(self.mave, c_f64p),
                      _p(self.msig, c_f64p), SB, LB, _p(out, c_f64p))
        return out

    # data.cpp:848-1007 (scalar branch: phenotype mask applied)
    def Ax(self, v, SB: int = 0, LB: int | None = None) -> np.ndarray:
        LB = self.mbytes if LB is None else LB
        v = np.ascontiguousarray(v, dtype=np.float64)
        assert len(v) == self.M
        out = np.empty(4 * LB)
        lib().orc_Ax(_p(self.bed, c_u8p), self.M, self.mbytes, self.N, _p(v, c_f64p), _p(self.mave, c_f64p),
                     _p(self.msig, c_f64p), _p(self.mask4, c_u8p), SB, LB, _p(out, c_f64p))
        return out

    # data.cpp:1065-1079
    def filter_pheno(self) -> np.ndarray:
        out = np.empty(self.N)
        lib().orc_filter_pheno(_p(np.ascontiguousarray(self.phen), c_f64p), self.N, _p(self.mask4, c_u8p), _p(out, c_f64p))
        return out


# --------------------------------------------------------------------------------------------
# association tests: LOO / LOCO p-values, data.cpp:1108-1353 + utilities.cpp:321-334
# --------------------------------------------------------------------------------------------
_DOSAGE = np.array([2.0, 0.0, 1.0, 0.0])      # dotp_lut_a: code 00 -> 2, 01 (missing) -> 0, 10 -> 1, 11 -> 0
_PRESENT = np.array([1.0, 0.0, 1.0, 1.0])     # dotp_lut_b


def _column_values(ds: "Dataset", j: int):
    """(value, bm) of marker j over the 4*mbytes slots: value = (a - mu) sigma b m, bm = b m (data.cpp:1150, 1326)."""
    codes = ((ds.bed[j][:, None] >> np.array([0, 2, 4, 6])) & 3).reshape(-1)
    m = ((ds.mask4[:, None] >> np.arange(4)) & 1).reshape(-1).astype(np.float64)
    bm = _PRESENT[codes] * m
    return (_DOSAGE[codes] - ds.mave[j]) * ds.msig[j] * bm, bm


def linear_reg1d_pvals(sumx, sumsqx, sumxy, sumy, sumsqy, n):
    """utilities.cpp:321-334; the Student-t tail through the regularised incomplete beta function (scipy)."""
    from scipy.special import betainc
    s2y = (sumsqy - sumy * sumy / n) / (n - 1)
    s2x = (sumsqx - sumx * sumx / n) / (n - 1)
    sxy = (sumxy - sumx * sumy / n) / (n - 1)
    rxy = sxy / math.sqrt(s2x * s2y)
    t = rxy * math.sqrt((n - 2) / (1 - rxy * rxy))
    nu = n - 2
    return float(betainc(0.5 * nu, 0.5, nu / (nu + t * t)))


def _reg_pvalue(value, bm, ymark):
    return linear_reg1d_pvals(value.sum(), (value * value).sum(), (value * ymark).sum(), (ymark * bm).sum(), (ymark * ymark * bm).sum(),
                              int(round(bm.sum())))


def pvals_loo(ds: "Dataset", z1, y, x1_hat) -> np.ndarray:
    """data::pvals_calc, data.cpp:1108-1180 (one estimator): y - z1 with the marker's own effect added back."""
    ymod = np.zeros(4 * ds.mbytes)
    ymod[: ds.N] = np.asarray(y)[: ds.N] - np.asarray(z1)[: ds.N]
    out = np.empty(ds.M)
    for j in range(ds.M):
        value, bm = _column_values(ds, j)
        out[j] = _reg_pvalue(value, bm, ymod + value / math.sqrt(ds.N) * x1_hat[j])
    return out


def pvals_loco(ds: "Dataset", z1, y, x1_hat, chroms, allreduce=lambda x: x) -> np.ndarray:
    """data::pvals_calc_LOCO, data.cpp:1220-1353: per chromosome the predictor of its own markers is added back."""
    ymod = np.zeros(4 * ds.mbytes)
    ymod[: ds.N] = np.asarray(y)[: ds.N] - np.asarray(z1)[: ds.N]
    chroms = np.asarray(chroms)
    out = np.zeros(ds.M)
    for ch in range(1, 24):
        sel = np.flatnonzero(chroms == ch)
        pred = np.zeros(4 * ds.mbytes)
        for j in sel:
            value, _ = _column_values(ds, j)
            pred += value / math.sqrt(ds.N) * x1_hat[j]
        pred = allreduce(pred)
        pred[ds.N:] = 0.0                  # the reference all-reduces N entries only (data.cpp:1287)
        ych = pred + ymod
        for j in sel:
            value, bm = _column_values(ds, j)
            out[j] = _reg_pvalue(value, bm, ych)
    return out


# --------------------------------------------------------------------------------------------
# denoiser, vamp.cpp:805-869
# --------------------------------------------------------------------------------------------
def g1(y, gam1, probs, vars_):
    """Posterior mean of a spike + Gaussian-mixture prior under AWGN of precision gam1 (vamp.cpp:805-834)."""
    y = np.asarray(y, dtype=np.float64)
    probs, vars_ = np.asarray(probs, float), np.asarray(vars_, float)
    sigma = 1.0 / gam1
    if -1e-10 < sigma < 1e-10:
        return y.copy()
    eta_max = vars_.max()
    pk = np.zeros_like(y)
    pkd = np.zeros_like(y)
    for p, v in zip(probs, vars_):
        expe = -0.5 * y ** 2 * (eta_max - v) / (v + sigma) / (eta_max + sigma)
        z = p / math.sqrt(v + sigma) * np.exp(expe)
        pk = pk + z
        z = z / (v + sigma) * y
        pkd = pkd - z
    return y + sigma * pkd / pk


def g1d(y, gam1, probs, vars_):
    """Derivative of g1 (vamp.cpp:836-869)."""
    y = np.asarray(y, dtype=np.float64)
    probs, vars_ = np.asarray(probs, float), np.asarray(vars_, float)
    sigma = 1.0 / gam1
    if -1e-10 < sigma < 1e-10:
        return np.ones_like(y)
    eta_max = vars_.max()
    pk = np.zeros_like(y)
    pkd = np.zeros_like(y)
    pkdd = np.zeros_like(y)
    for p, v in zip(probs, vars_):
        expe = -0.5 * y ** 2 * (eta_max - v) / (v + sigma) / (eta_max + sigma)
        ex = np.exp(expe)
        z = p / math.sqrt(v + sigma) * ex
        pk = pk + z
        z = z / (v + sigma) * y
        pkd = pkd - z
        z2 = z / (v + sigma) * y
        pkdd = pkdd - p / (v + sigma) ** 1.5 * ex + z2
    return 1.0 + sigma * (pkdd / pk - (pkd / pk) ** 2)


# --------------------------------------------------------------------------------------------
# EM prior update, vamp.cpp:929-1072
# --------------------------------------------------------------------------------------------
def em_sufficient_stats(r1, gam1, probs, vars_, lam, omegas):
    """One E-step over the local markers (vamp.cpp:944-1008): returns (sum pin, res[L-1], res_gammas[L-1])."""
    r1 = np.asarray(r1, float)
    noise_var = 1.0 / gam1
    max_sigma = max(vars_)
    L = len(probs)
    num = np.empty((L - 1, len(r1)))
    gam = np.empty((L - 1, len(r1)))
    for j in range(1, L):
        num[j - 1] = (lam * omegas[j] * np.exp(-r1 ** 2 / 2 * (max_sigma - vars_[j]) / (vars_[j] + noise_var)
                                                / (max_sigma + noise_var)) / math.sqrt(vars_[j] + noise_var)
                      / math.sqrt(2 * math.pi))
        gam[j - 1] = gam1 * r1 / (1 / vars_[j] + gam1)
    s = num.sum(axis=0)
    beta = num / s
    pin = 1 / (1 + (1 - lam) / math.sqrt(2 * math.pi * noise_var)
               * np.exp(-r1 ** 2 / 2 * max_sigma / noise_var / (noise_var + max_sigma)) / s)
    v = np.array([1.0 / (1.0 / vars_[j] + gam1) for j in range(1, L)])
    gam2 = beta * (gam * gam + v[:, None])
    return pin.sum(), (beta * pin).sum(axis=1), (gam2 * pin).sum(axis=1)


def merge_close_variances(probs, vars_):
    """vamp.cpp:1054-1071: merge components whose variances differ by less than 50 %."""
    probs, vars_ = list(probs), list(vars_)
    j = 0
    while j < len(vars_):
        k = j + 1
        while k < len(vars_):
            denom = min(vars_[j], vars_[k]) if vars_[j] != 0 else 1e-7
            if abs(vars_[j] - vars_[k]) / denom < 5e-1:
                s = probs[j] + probs[k]
                del vars_[k]
                del probs[k]
                probs[j] = s
                k -= 1
            k += 1
        j += 1
    return probs, vars_


def update_prior(r1, gam1, probs, vars_, Mt, EM_max_iter, EM_err_thr, learn_vars=1, allreduce=lambda x: x):
    """vamp::updatePrior (vamp.cpp:929-1072).  `allreduce` sums a numpy vector over marker shards."""
    probs, vars_ = list(map(float, probs)), list(map(float, vars_))
    lam = 1 - probs[0]
    omegas = list(probs)
    for j in range(1, len(omegas)):
        omegas[j] /= lam
    L = len(probs)
    for _ in range(EM_max_iter):
        probs_prev, vars_prev = list(probs), list(vars_)
        spin, res, resg = em_sufficient_stats(r1, gam1, probs, vars_, lam, omegas)
        tot = allreduce(np.concatenate([[spin], res, resg]))
        spin, res, resg = tot[0], tot[1:L], tot[L:]
        lam = spin / Mt
        for j in range(L - 1):
            if learn_vars == 1:
                vars_[j + 1] = resg[j] / res[j]
            omegas[j + 1] = res[j] / spin
            probs[j + 1] = lam * omegas[j + 1]
        probs[0] = 1 - lam
        dp = sum((a - b) ** 2 for a, b in zip(probs, probs_prev))
        npr = sum(a * a for a in probs)
        dv = sum((a - b) ** 2 for a, b in zip(vars_, vars_prev))
        nv = sum(a * a for a in vars_)
        if math.sqrt(dp / npr) < EM_err_thr and math.sqrt(dv / nv) < EM_err_thr:
            break
    return merge_close_variances(probs, vars_)


# --------------------------------------------------------------------------------------------
# Onsager probe, vamp.cpp:875-882 (libstdc++ mt19937 + bernoulli_distribution(0.5))
# --------------------------------------------------------------------------------------------
def bernoulli_probe(seed: int, S: int, M: int, Mt: int) -> np.ndarray:
    """(2*bern - 1)/sqrt(Mt) with std::mt19937{seed + S}; libstdc++ draws one canonical double
    (two 32-bit outputs, low word first) per Bernoulli sample."""
    raw = np.random.RandomState((seed + S) & 0xFFFFFFFF)._bit_generator.random_raw(2 * M).astype(np.float64)
    canon = (raw[0::2] + raw[1::2] * 4294967296.0) / 18446744073709551616.0
    return (2.0 * (canon < 0.5) - 1.0) / math.sqrt(Mt)


def sharded_probe(seed: int, S: int, M: int, Mt: int, nranks: int = 1) -> np.ndarray:
    """The probe of an R-rank run laid end to end: shard r (divide_work, utilities.cpp:266-281) holds
    bernoulli_probe(seed, S_r, M_r, Mt).  nranks == 1 is bernoulli_probe(seed, S, M, Mt) itself."""
    if nranks <= 1:
        return bernoulli_probe(seed, S, M, Mt)
    assert S == 0 and M == Mt, "the emulated multi-rank probe covers the whole marker range"
    parts = []
    for r in range(nranks):
        Mr, Sr = divide_work(Mt, nranks, r)
        parts.append(bernoulli_probe(seed, Sr, Mr, Mt))
    return np.concatenate(parts)


# --------------------------------------------------------------------------------------------
# LMMSE operator and preconditioned CG, vamp.cpp:1074-1229
# --------------------------------------------------------------------------------------------
def lmmse_mult(ds: Dataset, v, tau, gam2):
    v = np.asarray(v, float)
    if not v.any():  # vamp.cpp:1079-1080
        return np.zeros_like(v)
    return tau * ds.ATx(ds.Ax(v)) + gam2 * v


def precond_cg(ds: Dataset, v, mu_start, tau, gam2, CG_max_iter, denoiser, log=None):
    """vamp::precondCG_solver (vamp.cpp:1130-1229).  Returns (mu, iterations run)."""
    N = ds.N
    diag = tau * (N - 1) / N + gam2
    mu = np.array(mu_start, float)
    r = v - lmmse_mult(ds, mu, tau, gam2)
    z = r / diag
    p = z.copy()
    prev_onsager = 0.0
    its = 0
    for i in range(CG_max_iter):
        its = i + 1
        d = lmmse_mult(ds, p, tau, gam2)
        alpha = r.dot(z) / d.dot(p)
        mu = mu + alpha * p
        if denoiser == 0:
            onsager = gam2 * v.dot(mu)
            rel = abs((onsager - prev_onsager) / onsager) if onsager != 0 else 1.0
            if rel < 1e-8:
                break
            prev_onsager = onsager
        beta = 1.0 / r.dot(z)
        r = r - d * alpha
        z = r / diag
        beta *= r.dot(z)
        p = z + beta * p
        rel_err = math.sqrt(r.dot(r)) / math.sqrt(v.dot(v))
        if log is not None:
            log.append(rel_err)
        if rel_err < 1e-5:
            break
    return mu, its


# --------------------------------------------------------------------------------------------
# linear VAMP loop, vamp.cpp:149-803 (live branches: use_cross_val=0, use_freeze=0, reverse=0,
# use_lmmse_damp=0, init_est=0, no restart)
# --------------------------------------------------------------------------------------------
@dataclass
class VampConfig:
    iterations: int = 10
    rho: float = 0.15
    probs: tuple = ()
    vars: tuple = ()  # unscaled; multiplied by N at the start of infere (vamp.cpp:154-155)
    EM_max_iter: int = 2
    EM_err_thr: float = 1e-2
    CG_max_iter: int = 60
    stop_criteria_thr: float = 1e-4
    learn_vars: int = 1
    seed: int = 1
    gam1: float = 1e-6
    gamw: float = 2.0
    gamma_min: float = 1e-11
    gamma_max: float = 1e11
    auto_var_max_iter: int = 5
    # emulation of an R-rank run of the reference on one process: the only rank-dependent input of the algorithm is the
    # Onsager probe, which rank r draws from mt19937{seed + S_r} for its own M_r markers (vamp.cpp:875-882); everything
    # else is a sum over shards (MPI_Allreduce) and differs from the single-rank run by summation order only
    nranks: int = 1


@dataclass
class VampTrace:
    x1_hat: list = field(default_factory=list)  # per iteration, / sqrt(N) like the _it_k.bin files
    r1: list = field(default_factory=list)
    r2: list = field(default_factory=list)
    x2_hat: list = field(default_factory=list)
    z1: list = field(default_factory=list)
    gam1s: list = field(default_factory=list)
    gam2s: list = field(default_factory=list)
    R2trains: list = field(default_factory=list)
    gamw: list = field(default_factory=list)
    alpha2: list = field(default_factory=list)
    cg_iters: list = field(default_factory=list)
    probs: list = field(default_factory=list)
    vars: list = field(default_factory=list)


def infere_linear(ds: Dataset, cfg: VampConfig) -> VampTrace:
    N, M, Mt = ds.N, ds.M, ds.Mt
    tr = VampTrace()
    probs = list(cfg.probs)
    vars_ = [v * N for v in cfg.vars]
    gam1, gamw, rho = cfg.gam1, cfg.gamw, cfg.rho
    gam2 = 0.0
    alpha1 = 0.0
    alpha2 = 0.0  # the reference reads an uninitialised member at it=1 (vamp.hpp:14, vamp.cpp:501); 0 here
    y = ds.filter_pheno()
    r1 = np.zeros(M)
    x1_hat = np.zeros(M)
    r2 = np.zeros(M)
    mu_CG_last = np.zeros(M)
    clamp = lambda g: min(max(g, cfg.gamma_min), cfg.gamma_max)
    sqrtN = math.sqrt(N)
    for it in range(1, cfg.iterations + 1):
        x1_hat_prev = x1_hat.copy()
        alpha1_prev = alpha1
        # ---- denoising, vamp.cpp:289-338
        for it_revar in range(1, cfg.auto_var_max_iter + 1):
            x1_hat = g1(r1, gam1, probs, vars_)
            alpha1 = g1d(r1, gam1, probs, vars_).sum() / Mt
            eta1 = gam1 / alpha1
            if it <= 1:
                break
            gam1_prev = gam1
            gam1 = clamp(1.0 / (1.0 / eta1 + ((x1_hat - r1) ** 2).sum() / Mt))
            probs, vars_ = update_prior(r1, gam1, probs, vars_, Mt, cfg.EM_max_iter, cfg.EM_err_thr, cfg.learn_vars)
            if abs(gam1 - gam1_prev) < 1e-3:
                break
        tr.gam1s.append(gam1)
        if it > 1:  # damping, vamp.cpp:348-414
            x1_hat = rho * x1_hat + (1 - rho) * x1_hat_prev
            alpha1 = rho * alpha1 + (1 - rho) * alpha1_prev
        z1 = ds.Ax(x1_hat)  # vamp.cpp:429
        tr.z1.append(z1.copy())
        tr.x1_hat.append(x1_hat / sqrtN)
        tr.r1.append(r1 / sqrtN)
        gam2 = clamp(eta1 - gam1)  # vamp.cpp:472
        r2_prev = r2
        r2 = (eta1 * x1_hat - gam1 * r1) / gam2
        rho = max(rho, min(2 * min(alpha1, alpha2), 1.0))  # vamp.cpp:501-502
        if it <= 1:  # vamp.cpp:518-519
            probs, vars_ = update_prior(r1, gam1, probs, vars_, Mt, cfg.EM_max_iter, cfg.EM_err_thr, cfg.learn_vars)
        # err_measures(1): vamp.cpp:1290-1317 (uses the cached z1)
        tr.R2trains.append(1 - ((y - z1[:N]) ** 2).sum() / (y ** 2).sum())
        tr.r2.append(r2 / sqrtN)
        # ---- LMMSE, vamp.cpp:584-596
        v = gamw * ds.ATx(y) + gam2 * r2
        x2_hat, k1 = precond_cg(ds, v, np.zeros(M) if it == 1 else mu_CG_last, gamw, gam2, cfg.CG_max_iter, 1)
        mu_CG_last = x2_hat.copy()
        tr.x2_hat.append(x2_hat / sqrtN)
        # Onsager, vamp.cpp:871-889
        bern = sharded_probe(cfg.seed, ds.S, M, Mt, cfg.nranks)
        invQ, k2 = precond_cg(ds, bern, np.zeros(M), gamw, gam2, cfg.CG_max_iter, 0)
        alpha2 = gam2 * bern.dot(invQ)
        tr.alpha2.append(alpha2)
        tr.cg_iters.append((k1, k2))
        eta2 = gam2 / alpha2
        if it > 2:  # vamp.cpp:691-693
            gam2 = clamp(1 / (1 / eta2 + ((x2_hat - r2) ** 2).sum() / Mt))
        tr.gam2s.append(gam2)
        gam1 = clamp(eta2 - gam2)
        r1 = (eta2 * x2_hat - gam2 * r2) / gam1
        # updateNoisePrec, vamp.cpp:892-927
        temp = ds.Ax(x2_hat)[:N] - y
        trace_corr = bern.dot(ds.ATx(ds.Ax(invQ))) * Mt
        gamw = N / (temp.dot(temp) + trace_corr)
        tr.gamw.append(gamw)
        # err_measures(2): another Ax(x2_hat), vamp.cpp:1301
        tr.R2trains.append(1 - ((y - ds.Ax(x2_hat)[:N]) ** 2).sum() / (y ** 2).sum())
        tr.probs.append(list(probs))
        tr.vars.append(list(vars_))
        # stopping rule, vamp.cpp:741-749
        with np.errstate(divide="ignore", invalid="ignore"):  # x1_hat_prev == 0 after iteration 1 -> inf, like the reference
            rel_change = np.sqrt(((x1_hat_prev - x1_hat) ** 2).sum() / (x1_hat_prev ** 2).sum())
        if it > 1 and rel_change < cfg.stop_criteria_thr:
            break
    return tr


# --------------------------------------------------------------------------------------------
# probit pieces: utilities.cpp:345-409 (erfcx) and vamp_probit.cpp:661-726
# --------------------------------------------------------------------------------------------
_ERFCX_COEF = [float.fromhex(h) for h in (
    "0x1.edcad78fc8044p-31", "0x1.b1548f14735d1p-30", "-0x1.a1ad2e6c4a7a8p-27", "-0x1.1985b48f08574p-26",
    "0x1.c6a8093ac4f83p-24", "0x1.31c2b2b44b731p-24", "-0x1.b87373facb29fp-21", "0x1.3fef1358803b7p-22",
    "0x1.7eec072bb0be3p-18", "-0x1.78a680a741c4ap-17", "-0x1.9951f39295cf4p-16", "0x1.3be1255ce180bp-13",
    "-0x1.a1df71176b791p-13", "-0x1.8d4aaa0099bc8p-11", "0x1.49c673066c831p-8", "-0x1.0962386ea02b7p-6",
    "0x1.3079edf465cc3p-5", "-0x1.0fb06dfedc4ccp-4", "0x1.7fee004e266dfp-4", "-0x1.9ddb23c3e14d2p-4",
    "0x1.16ecefcfa4865p-4", "0x1.f7f5df66fc349p-7", "-0x1.1df1ad154a27fp-3", "0x1.dd2c8b74febf6p-3")]


def erfcx(x):
    """Scaled complementary error function exp(x^2) erfc(x), same polynomial as utilities.cpp:345-409
    (numpy has no fma, so agreement with the reference is ~1e-15 relative, not bitwise)."""
    x = np.asarray(x, dtype=np.float64)
    a = np.abs(x)
    q = (a - 4.0) / (a + 4.0)
    p = np.full_like(a, _ERFCX_COEF[0])
    for c in _ERFCX_COEF[1:]:
        p = p * q + c
    r = (p + 1.0) / (1.0 + 2.0 * a)
    with np.errstate(over="ignore"):
        e = np.exp(x * x)
    neg = 2.0 * e - r
    neg = np.where(np.isinf(e), e, neg)
    return np.where(x < 0, neg, r)


def g1_bin_class(p, tau1, y, m_cov, probit_var=1.0):
    """vamp_probit.cpp:661-690."""
    p, y = np.asarray(p, float), np.asarray(y, float)
    s = math.sqrt(probit_var + 1.0 / tau1)
    c = (p + m_cov) / s
    ratio = 2.0 / math.sqrt(2 * math.pi) / erfcx(-(2 * y - 1) * c / math.sqrt(2))
    return p + (2 * y - 1) * ratio / tau1 / s


def g1d_bin_class(p, tau1, y, m_cov, probit_var=1.0):
    """vamp_probit.cpp:693-726."""
    p, y = np.asarray(p, float), np.asarray(y, float)
    s = math.sqrt(probit_var + 1.0 / tau1)
    c = (p + m_cov) / s
    ratio = 2.0 / math.sqrt(2 * math.pi) / erfcx(-(2 * y - 1) * c / math.sqrt(2))
    return 1 - ratio / (1 + tau1 * probit_var) * ((2 * y - 1) * c + ratio)


def probit_cov_pass(y, gg, Z, eta, probit_var=1.0):
    """The N x C sums of the covariate Newton step: mlogL_probit (vamp_probit.cpp:841-858), grad_cov (:814-839), and the
    numerator / Hessian of Newton_method_cov (:944-970; note that these two do not carry probit_var)."""
    from scipy.special import erfc
    y, gg, Z, eta = np.asarray(y, float), np.asarray(gg, float), np.asarray(Z, float), np.asarray(eta, float)
    n = len(y)
    g = gg + Z @ eta
    sgn = 2 * y - 1
    arg = sgn / math.sqrt(probit_var) * g
    mlogl = -np.sum(np.log(0.5 * erfc(-arg * math.sqrt(0.5)))) / n
    ratio = 2.0 / math.sqrt(2 * math.pi) / erfcx(-arg / math.sqrt(2))
    grad = -(ratio * sgn / math.sqrt(probit_var)) @ Z / n
    lam = 2.0 / math.sqrt(2 * math.pi) / erfcx(-(sgn * g) / math.sqrt(2)) * sgn
    step = Z.T @ lam
    H = Z.T @ (Z * (lam * (lam + g))[:, None])
    return mlogl, grad, step, H


# --------------------------------------------------------------------------------------------
# XXT form of the LMMSE step: people statistics (data.cpp:548-640) and the N-space solver (denoiserXXT.cpp:15-135)
# --------------------------------------------------------------------------------------------
def people_statistics(ds: "Dataset", allreduce=lambda x: x):
    """(mave_people, msig_people, numb_people), each 4*mbytes long."""
    n4 = 4 * ds.mbytes
    s1, s2, numb = np.zeros(n4), np.zeros(n4), np.zeros(n4)
    for j in range(ds.M):
        value, bm = _column_values(ds, j)
        s1 += value
        numb += bm
        s2 += value * value
    s1, s2, numb = allreduce(s1), allreduce(s2), allreduce(numb)
    m = ((ds.mask4[:, None] >> np.arange(4)) & 1).reshape(-1).astype(bool)
    m[ds.N:] = False
    mave, msig = np.zeros(n4), np.zeros(n4)
    mave[m] = s1[m] / numb[m]
    msig[m] = np.sqrt((numb[m] - 1) / (s2[m] - numb[m] * mave[m] ** 2))
    return mave, msig, numb


def lmmse_mult_aat(ds: "Dataset", u, tau, gam2):
    """denoiserXXT.cpp:15-29."""
    n4 = 4 * ds.mbytes
    uu = np.zeros(n4)
    uu[: ds.N] = np.asarray(u)[: ds.N]
    if not uu.any():
        return np.zeros(n4)
    res = ds.Ax(ds.ATx(uu))
    res[: ds.N] = res[: ds.N] * tau + gam2 * uu[: ds.N]
    return res


def cg_solver_aat(ds: "Dataset", v, mu_start, tau, gam2, people, CG_max_iter, log=None):
    """denoiserXXT.cpp:57-135; returns (mu, iterations)."""
    N, n4 = ds.N, 4 * ds.mbytes
    mave_p, msig_p, numb_p = people
    with np.errstate(divide="ignore", invalid="ignore"):
        diag = tau * ((numb_p[:N] - 1) / msig_p[:N] / msig_p[:N] + mave_p[:N] ** 2 * numb_p[:N]) / N + gam2
    v = np.asarray(v, float)[:N]
    mu = np.zeros(n4)
    mu[:N] = np.asarray(mu_start, float)[:N]
    r = v - lmmse_mult_aat(ds, mu, tau, gam2)[:N]
    z = r / diag
    p = z.copy()
    its = 0
    for i in range(CG_max_iter):
        its = i + 1
        d = lmmse_mult_aat(ds, p, tau, gam2)[:N]
        alpha = r.dot(z) / d.dot(p)
        mu[:N] += alpha * p
        beta = 1.0 / r.dot(z)
        r = r - alpha * d
        z = r / diag
        beta *= r.dot(z)
        p = z + beta * p
        rel_err = math.sqrt(r.dot(r) / v.dot(v))
        if log is not None:
            log.append(rel_err)
        if rel_err < 1e-4:
            break
    return mu, its


def lmmse_denoiser_aat(ds: "Dataset", r2, mu_last, y, gamw, gam2, people, CG_max_iter):
    """denoiserXXT.cpp:31-55; returns (x2_hat, u)."""
    z2 = ds.Ax(r2)
    v = np.asarray(y, float)[: ds.N] - z2[: ds.N]
    u, _ = cg_solver_aat(ds, v, mu_last, gamw, gam2, people, CG_max_iter)
    return gamw * ds.ATx(u) + np.asarray(r2, float), u
