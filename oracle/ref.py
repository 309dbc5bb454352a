"""oracle/ref.py -- TEST INFRASTRUCTURE: ctypes view of oracle/_ref/libgvamp_ref.so, i.e. the
UNMODIFIED reference classes compiled with the single-rank shims (oracle/Makefile, ref_harness.cpp).
Used to pin oracle.py / gvamp_oracle.c and to generate tests/golden/.  oracle/_ref is git-ignored and
only exists where it was built (the build container) or where the snapshot carried it (GPU box).
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
_LIB = None

c_f64p = ctypes.POINTER(ctypes.c_double)
c_u8p = ctypes.POINTER(ctypes.c_uint8)


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "libgvamp_ref.so"))


def exe(name: str) -> str:
    return os.path.join(REF_DIR, name)


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(os.path.join(REF_DIR, "libgvamp_ref.so"))
        vp, ci, cd, cs = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_char_p
        L.ref_data_create.restype = vp
        L.ref_data_create.argtypes = [cs, cs, ci, ci, ci, ci, cd]
        L.ref_data_create_y.restype = vp
        L.ref_data_create_y.argtypes = [c_f64p, cs, ci, ci, ci, ci, cd]
        L.ref_data_create_bim.restype = vp
        L.ref_data_create_bim.argtypes = [cs, cs, ci, ci, ci, ci, cd, cs]
        L.ref_data_pvals.argtypes = [vp, ci, c_f64p, c_f64p, c_f64p, cs, c_f64p]
        L.ref_data_people_stats.argtypes = [vp, c_f64p, c_f64p, c_f64p]
        L.ref_vamp_cg_aat.argtypes = [vp, vp, c_f64p, c_f64p, ci, cd, c_f64p]
        L.ref_data_destroy.argtypes = [vp]
        L.ref_data_mbytes.restype = ctypes.c_long
        L.ref_data_mbytes.argtypes = [vp]
        L.ref_data_nonas.argtypes = [vp]
        L.ref_data_intercept.restype = cd
        L.ref_data_intercept.argtypes = [vp]
        L.ref_data_scale.restype = cd
        L.ref_data_scale.argtypes = [vp]
        L.ref_data_mask4.argtypes = [vp, c_u8p]
        L.ref_data_phen.argtypes = [vp, c_f64p]
        L.ref_data_filter_pheno.argtypes = [vp, c_f64p]
        L.ref_data_stats.argtypes = [vp, c_f64p, c_f64p]
        L.ref_data_Ax.argtypes = [vp, c_f64p, ci, ci, c_f64p]
        L.ref_data_ATx.argtypes = [vp, c_f64p, ci, ci, c_f64p]
        L.ref_data_read_covariates.argtypes = [vp, cs, ci]
        L.ref_vamp_create.restype = vp
        L.ref_vamp_create.argtypes = [ci, ci, ci, cd, cd, ci, cd, c_f64p, c_f64p, ci, cs, cs, cs, ci, cd, ci, ci,
                                      ctypes.c_ulong, cd]
        L.ref_vamp_destroy.argtypes = [vp]
        L.ref_vamp_set_prior.argtypes = [vp, c_f64p, c_f64p, ci]
        L.ref_vamp_get_prior.argtypes = [vp, c_f64p, c_f64p]
        L.ref_vamp_set_state.argtypes = [vp, cd, cd, cd, cd]
        L.ref_vamp_g1.argtypes = [vp, c_f64p, ci, cd, c_f64p, c_f64p]
        L.ref_vamp_update_prior.argtypes = [vp, c_f64p, ci, cd]
        L.ref_vamp_lmmse_mult.argtypes = [vp, vp, c_f64p, ci, cd, c_f64p]
        L.ref_vamp_cg.argtypes = [vp, vp, c_f64p, c_f64p, ci, cd, ci, c_f64p]
        L.ref_vamp_onsager.restype = cd
        L.ref_vamp_onsager.argtypes = [vp, vp, cd, cd, ci, c_f64p, c_f64p]
        L.ref_vamp_infere.argtypes = [vp, vp, ci, c_f64p]
        L.ref_vamp_gamw.restype = cd
        L.ref_vamp_gamw.argtypes = [vp]
        L.ref_vamp_gam1.restype = cd
        L.ref_vamp_gam1.argtypes = [vp]
        L.ref_vamp_g1_bin_class.argtypes = [vp, c_f64p, c_f64p, c_f64p, ci, cd, c_f64p, c_f64p]
        L.ref_erfcx.restype = cd
        L.ref_erfcx.argtypes = [cd]
        L.ref_simulate.argtypes = [ci, c_f64p, c_f64p, ci, ctypes.c_ulong, c_f64p]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(c_f64p)


class RefData:
    """class data of the reference (data.hpp:93-94), one rank, shard [S, S+M)."""

    def __init__(self, bed_path, N, M, Mt=None, S=0, phen_path=None, y=None, alpha_scale=1.0, bim_path=None):
        L = lib()
        self.N, self.M = N, M
        Mt = M if Mt is None else Mt
        if bim_path is not None:
            self.h = L.ref_data_create_bim(phen_path.encode(), bed_path.encode(), N, M, Mt, S, alpha_scale, bim_path.encode())
        elif phen_path is not None:
            self.h = L.ref_data_create(phen_path.encode(), bed_path.encode(), N, M, Mt, S, alpha_scale)
        else:
            yy = np.ascontiguousarray(np.zeros(N) if y is None else y, dtype=np.float64)
            self.h = L.ref_data_create_y(_p(yy), bed_path.encode(), N, M, Mt, S, alpha_scale)
        self.mbytes = L.ref_data_mbytes(self.h)

    def stats(self):
        a, s = np.empty(self.M), np.empty(self.M)
        lib().ref_data_stats(self.h, _p(a), _p(s))
        return a, s

    def mask4(self):
        m = np.empty(self.mbytes, dtype=np.uint8)
        lib().ref_data_mask4(self.h, m.ctypes.data_as(c_u8p))
        return m

    def phen(self):
        p = np.empty(self.N)
        lib().ref_data_phen(self.h, _p(p))
        return p

    def filter_pheno(self):
        p = np.empty(self.N)
        lib().ref_data_filter_pheno(self.h, _p(p))
        return p

    def nonas(self):
        return lib().ref_data_nonas(self.h)

    def intercept_scale(self):
        return lib().ref_data_intercept(self.h), lib().ref_data_scale(self.h)

    def Ax(self, v, SB=0, LB=None):
        LB = self.mbytes if LB is None else LB
        v = np.ascontiguousarray(v, dtype=np.float64)
        out = np.empty(4 * LB)
        lib().ref_data_Ax(self.h, _p(v), SB, LB, _p(out))
        return out

    def ATx(self, u, SB=0, LB=None):
        LB = self.mbytes if LB is None else LB
        uu = np.zeros(4 * LB)
        u = np.asarray(u, dtype=np.float64)
        uu[: min(len(u), 4 * LB)] = u[: 4 * LB]
        out = np.empty(self.M)
        lib().ref_data_ATx(self.h, _p(uu), SB, LB, _p(out))
        return out

    def people_stats(self):
        a, s, n = np.empty(4 * self.mbytes), np.empty(4 * self.mbytes), np.empty(4 * self.mbytes)
        lib().ref_data_people_stats(self.h, _p(a), _p(s), _p(n))
        return a, s, n

    def pvals(self, z1, y, x1_hat, out_path, loco=False):
        """data::pvals_calc / pvals_calc_LOCO (data.cpp:1108-1353) for one estimator."""
        z1 = np.ascontiguousarray(z1[: self.N], dtype=np.float64)
        y = np.ascontiguousarray(y[: self.N], dtype=np.float64)
        x = np.ascontiguousarray(x1_hat, dtype=np.float64)
        out = np.empty(self.M)
        lib().ref_data_pvals(self.h, int(loco), _p(z1), _p(y), _p(x), out_path.encode(), _p(out))
        return out

    def close(self):
        if self.h:
            lib().ref_data_destroy(self.h)
            self.h = None


class RefVamp:
    """class vamp of the reference (vamp.hpp:78), built with the all-manual constructor."""

    def __init__(self, N, M, Mt, probs, vars_, gam1=1e-6, gamw=2.0, iterations=1, rho=0.15, out_dir="/tmp/",
                 out_name="ref", model="linear", EM_max_iter=2, EM_err_thr=1e-2, CG_max_iter=60, learn_vars=1, seed=1,
                 stop_thr=1e-4):
        pv = np.ascontiguousarray(vars_, dtype=np.float64)
        pp = np.ascontiguousarray(probs, dtype=np.float64)
        self.M, self.N = M, N
        self.h = lib().ref_vamp_create(N, M, Mt, gam1, gamw, iterations, rho, _p(pv), _p(pp), len(pp), out_dir.encode(),
                                       out_name.encode(), model.encode(), EM_max_iter, EM_err_thr, CG_max_iter,
                                       learn_vars, seed, stop_thr)

    def set_prior(self, probs, vars_):
        pv = np.ascontiguousarray(vars_, dtype=np.float64)
        pp = np.ascontiguousarray(probs, dtype=np.float64)
        lib().ref_vamp_set_prior(self.h, _p(pv), _p(pp), len(pp))

    def get_prior(self):
        v, p = np.empty(64), np.empty(64)
        L = lib().ref_vamp_get_prior(self.h, _p(v), _p(p))
        return p[:L].copy(), v[:L].copy()

    def set_state(self, gam1, gam2, gamw, probit_var=1.0):
        lib().ref_vamp_set_state(self.h, gam1, gam2, gamw, probit_var)

    def g1(self, r, gam1):
        r = np.ascontiguousarray(r, dtype=np.float64)
        o, od = np.empty_like(r), np.empty_like(r)
        lib().ref_vamp_g1(self.h, _p(r), len(r), gam1, _p(o), _p(od))
        return o, od

    def update_prior(self, r1, gam1):
        r1 = np.ascontiguousarray(r1, dtype=np.float64)
        lib().ref_vamp_update_prior(self.h, _p(r1), len(r1), gam1)
        return self.get_prior()

    def lmmse_mult(self, data: RefData, v, tau):
        v = np.ascontiguousarray(v, dtype=np.float64)
        out = np.empty(self.M)
        lib().ref_vamp_lmmse_mult(self.h, data.h, _p(v), self.M, tau, _p(out))
        return out

    def cg(self, data: RefData, rhs, mu0, tau, denoiser):
        rhs = np.ascontiguousarray(rhs, dtype=np.float64)
        mu0 = np.ascontiguousarray(mu0, dtype=np.float64)
        out = np.empty(self.M)
        lib().ref_vamp_cg(self.h, data.h, _p(rhs), _p(mu0), self.M, tau, denoiser, _p(out))
        return out

    def cg_aat(self, data: RefData, rhs, mu0, tau):
        n4 = 4 * data.mbytes
        rhs = np.ascontiguousarray(rhs, dtype=np.float64)
        mu0 = np.ascontiguousarray(mu0, dtype=np.float64)
        out = np.empty(n4)
        lib().ref_vamp_cg_aat(self.h, data.h, _p(rhs), _p(mu0), n4, tau, _p(out))
        return out

    def onsager(self, data: RefData, gam2, tau):
        b, q = np.empty(self.M), np.empty(self.M)
        a = lib().ref_vamp_onsager(self.h, data.h, gam2, tau, self.M, _p(b), _p(q))
        return a, b, q

    def infere(self, data: RefData):
        out = np.empty(self.M)
        lib().ref_vamp_infere(self.h, data.h, self.M, _p(out))
        return out

    def gamw(self):
        return lib().ref_vamp_gamw(self.h)

    def g1_bin_class(self, p, tau1, y, mcov):
        p = np.ascontiguousarray(p, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        mc = np.ascontiguousarray(np.broadcast_to(np.asarray(mcov, dtype=np.float64), p.shape))
        o, od = np.empty_like(p), np.empty_like(p)
        lib().ref_vamp_g1_bin_class(self.h, _p(p), _p(y), _p(mc), len(p), tau1, _p(o), _p(od))
        return o, od

    def close(self):
        if self.h:
            lib().ref_vamp_destroy(self.h)
            self.h = None


def erfcx(x):
    return np.array([lib().ref_erfcx(float(v)) for v in np.atleast_1d(x)])
