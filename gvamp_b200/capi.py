"""ctypes binding of libgvamp_b200.so (include/gvamp_b200.h) for the tests and bench.py.

Python is plumbing here: the product is the CUDA library plus the C++ host programs in
gvamp_b200/host.  There is no CPU fallback -- `load()` raises if the library was not built and every
call raises GvbError when the CUDA path fails.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libgvamp_b200.so")

c_f64p = ctypes.POINTER(ctypes.c_double)
c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_i64p = ctypes.POINTER(ctypes.c_int64)
vp, ci, cl, cd = ctypes.c_void_p, ctypes.c_int, ctypes.c_long, ctypes.c_double

NCCL_ID_BYTES = 128

# name -> (restype, argtypes); the non-gpu tests check that every symbol of the header is exported
SIGNATURES = {
    "gvb_last_error": (ctypes.c_char_p, []),
    "gvb_version": (ctypes.c_char_p, []),
    "gvb_device_count": (ci, [ctypes.POINTER(ci)]),
    "gvb_nccl_unique_id": (ci, [vp]),
    "gvb_ctx_create": (ci, [ctypes.POINTER(vp), ci, ci, ci, vp]),
    "gvb_ctx_create_shared": (ci, [ctypes.POINTER(vp), vp]),
    "gvb_ctx_destroy": (None, [vp]),
    "gvb_ctx_sync": (ci, [vp]),
    "gvb_ctx_stream": (vp, [vp]),
    "gvb_ctx_info": (ci, [vp] + [ctypes.POINTER(cl)] * 5),
    "gvb_timer_start": (ci, [vp, ci]),
    "gvb_timer_stop": (ci, [vp, ci]),
    "gvb_timer_elapsed_ms": (ci, [vp, ci, ctypes.POINTER(ctypes.c_float)]),
    "gvb_launch_count": (cl, [vp]),
    "gvb_sweep_count": (cl, [vp]),
    "gvb_host_sync_count": (cl, [vp]),
    "gvb_layout_generation": (cl, [vp]),
    "gvb_vec_reduce_batch": (ci, [vp, ci, vp, c_f64p]),
    "gvb_profile_enable": (ci, [vp, ci]),
    "gvb_profile_read": (ci, [vp, c_f64p]),
    "gvb_profile_read_dual": (ci, [vp, c_f64p]),
    "gvb_dual_sweep_count": (cl, [vp]),
    "gvb_dAx2": (ci, [vp, vp, vp, vp, vp]),
    "gvb_dATx2": (ci, [vp, vp, vp, vp, vp]),
    "gvb_divide_work": (None, [cl, ci, ci, ctypes.POINTER(cl), ctypes.POINTER(cl)]),
    "gvb_bed_load_file": (ci, [vp, ctypes.c_char_p, cl, cl, cl, cl]),
    "gvb_bed_load_host": (ci, [vp, c_u8p, cl, cl, cl, cl]),
    "gvb_bed_synth": (ci, [vp, ctypes.c_uint64, cl, cl, cl, cl, cd]),
    "gvb_bed_decode": (ci, [vp, cl, cl, c_u8p]),
    "gvb_set_mask": (ci, [vp, c_u8p, ci]),
    "gvb_compute_stats": (ci, [vp, cd]),
    "gvb_get_stats": (ci, [vp, c_f64p, c_f64p]),
    "gvb_get_counts": (ci, [vp, c_i64p]),
    "gvb_Ax": (ci, [vp, c_f64p, c_f64p, cl, cl]),
    "gvb_ATx": (ci, [vp, c_f64p, c_f64p, cl, cl]),
    "gvb_marker_dot": (ci, [vp, cl, c_f64p, cl, cl, c_f64p, c_f64p]),
    "gvb_allreduce_host": (ci, [vp, c_f64p, ci]),
    "gvb_vec_alloc": (ci, [vp, cl, ctypes.POINTER(vp)]),
    "gvb_vec_alloc_M": (ci, [vp, ctypes.POINTER(vp)]),
    "gvb_vec_alloc_N": (ci, [vp, ctypes.POINTER(vp)]),
    "gvb_vec_free": (None, [vp, vp]),
    "gvb_vec_len": (cl, [vp]),
    "gvb_vec_ptr": (vp, [vp]),
    "gvb_vec_upload": (ci, [vp, vp, c_f64p, cl]),
    "gvb_vec_upload_changed": (ci, [vp, vp, vp, c_f64p, cl, ctypes.POINTER(ci)]),
    "gvb_vec_download": (ci, [vp, vp, c_f64p, cl]),
    "gvb_vec_copy": (ci, [vp, vp, vp]),
    "gvb_vec_fill": (ci, [vp, vp, cd]),
    "gvb_vec_axpby": (ci, [vp, vp, cd, vp, cd, vp]),
    "gvb_vec_axpby_div": (ci, [vp, vp, cd, vp, cd, vp, cd]),
    "gvb_vec_dots": (ci, [vp, ci, ctypes.POINTER(vp), ctypes.POINTER(vp), ci, c_f64p]),
    "gvb_vec_dist2": (ci, [vp, vp, vp, ci, c_f64p]),
    "gvb_dAx": (ci, [vp, vp, vp]),
    "gvb_dATx": (ci, [vp, vp, vp]),
    "gvb_denoise": (ci, [vp, vp, cd, c_f64p, c_f64p, ci, vp, c_f64p]),
    "gvb_em_stats": (ci, [vp, vp, cd, cd, c_f64p, c_f64p, ci, c_f64p]),
    "gvb_lmmse_mult": (ci, [vp, vp, cd, cd, vp]),
    "gvb_cg_solve": (ci, [vp, vp, vp, cd, cd, ci, ci, ctypes.POINTER(ci), c_f64p]),
    "gvb_cg_solve_ex": (ci, [vp, vp, vp, cd, cd, ci, ci, ctypes.POINTER(ci), c_f64p, vp, c_f64p]),
    "gvb_cg_solve_warm": (ci, [vp, vp, vp, cd, cd, ci, ci, ctypes.POINTER(ci), c_f64p, vp, vp, ci, c_f64p]),
    "gvb_cg_set_companion": (ci, [vp, vp, vp]),
    "gvb_cg_prepare": (ci, [vp, vp, vp, cd, cd, ci, vp, vp, ci, vp, vp]),
    "gvb_cg_solve_prepared": (ci, [vp, vp, vp, cd, cd, ci, ci, ctypes.POINTER(ci), c_f64p, vp, vp, ci, c_f64p]),
    "gvb_cg_solve_cached": (ci, [vp, vp, vp, cd, cd, ci, ci, ctypes.POINTER(ci), c_f64p, vp, ctypes.POINTER(ci), c_f64p]),
    "gvb_people_stats": (ci, [vp, vp, vp, vp]),
    "gvb_cg_solve_aat": (ci, [vp, vp, vp, cd, cd, vp, vp, vp, ci, ctypes.POINTER(ci), c_f64p]),
    "gvb_probit_denoise": (ci, [vp, vp, vp, vp, cd, cd, vp, c_f64p]),
    "gvb_missing_list_entries": (cl, [vp]),
    "gvb_twin_state": (ci, [vp]),
    "gvb_twin_stripes": (cl, [vp]),
    "gvb_twin_release": (ci, [vp]),
    "gvb_snapshot_begin": (ci, [vp, vp, cl, ci]),
    "gvb_snapshot_wait": (ci, [vp, ci, ctypes.POINTER(c_f64p), ctypes.POINTER(cl)]),
    "gvb_assoc_pvals": (ci, [vp, vp, vp, vp, vp]),
    "gvb_probit_cov_pass": (ci, [vp, vp, vp, vp, ci, c_f64p, cd, ci, c_f64p]),
    "gvb_probit_cov_apply": (ci, [vp, vp, ci, c_f64p, vp]),
}



class RedOp(ctypes.Structure):
    """gvb_red_op of the header"""
    _fields_ = [("x", vp), ("y", vp), ("a", cd), ("b", cd), ("kind", ci), ("sync", ci)]


RED_DOT, RED_SQ = 0, 1

XCHECK_LIB_PATH = os.path.join(HERE, "lib", "xcheck", "libgvamp_b200.so")
_LIB = None
_XLIB = None


class GvbError(RuntimeError):
    pass


def load():
    """dlopen the CUDA library; raises (never falls back) when it has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise GvbError(f"{LIB_PATH} is missing: run `python -m gvamp_b200.build` (there is no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def load_xcheck():
    """The TEST build of the library (gvamp_b200/lib/xcheck/): the product library plus the two earlier kernel generations that the
    parity tests use as on-device cross-checks (env GVB_KERNELS=simple|lut1).  Never loaded by the product path or by bench.py."""
    global _XLIB
    if _XLIB is None:
        if not os.path.exists(XCHECK_LIB_PATH):
            raise GvbError(f"{XCHECK_LIB_PATH} is missing: run `python -m gvamp_b200.build`")
        L = ctypes.CDLL(XCHECK_LIB_PATH, mode=ctypes.RTLD_LOCAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _XLIB = L
    return _XLIB


def _chk(rc, L=None):
    if rc != 0:
        raise GvbError(f"gvamp_b200 error {rc}: {(L or load()).gvb_last_error().decode()}")


def _f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(c_f64p)


def device_count() -> int:
    n = ci(0)
    rc = load().gvb_device_count(ctypes.byref(n))
    return n.value if rc == 0 else 0


def nccl_unique_id() -> bytes:
    buf = ctypes.create_string_buffer(NCCL_ID_BYTES)
    _chk(load().gvb_nccl_unique_id(buf))
    return buf.raw


def divide_work(Mt, nranks, rank):
    M, S = cl(0), cl(0)
    load().gvb_divide_work(Mt, nranks, rank, ctypes.byref(M), ctypes.byref(S))
    return M.value, S.value


class Vec:
    """Device-resident FP64 vector owned by a Context."""

    def __init__(self, ctx, handle):
        self.ctx, self.h = ctx, handle

    def __len__(self):
        return self.ctx.L.gvb_vec_len(self.h)

    def upload(self, a):
        a, p = _f64(a)
        _chk(self.ctx.L.gvb_vec_upload(self.ctx.h, self.h, p, len(a)), self.ctx.L)
        return self

    def upload_changed(self, a, stage):
        """upload + whether any bit of the vector changed (gvb_vec_upload_changed); `stage` is a scratch Vec"""
        a, p = _f64(a)
        changed = ci(-1)
        _chk(self.ctx.L.gvb_vec_upload_changed(self.ctx.h, self.h, stage.h, p, len(a), ctypes.byref(changed)), self.ctx.L)
        return bool(changed.value)

    def download(self, n=None):
        n = len(self) if n is None else n
        out = np.empty(n)
        _chk(self.ctx.L.gvb_vec_download(self.ctx.h, self.h, out.ctypes.data_as(c_f64p), n), self.ctx.L)
        return out

    def fill(self, v):
        _chk(self.ctx.L.gvb_vec_fill(self.ctx.h, self.h, float(v)), self.ctx.L)
        return self

    def copy_from(self, other):
        _chk(self.ctx.L.gvb_vec_copy(self.ctx.h, self.h, other.h), self.ctx.L)
        return self

    def free(self):
        if self.h:
            self.ctx.L.gvb_vec_free(self.ctx.h, self.h)
            self.h = None


class Context:
    """One B200 / one marker shard (include/gvamp_b200.h: gvb_ctx)."""

    def __init__(self, device=0, rank=0, nranks=1, nccl_id: bytes | None = None, xcheck: bool = False):
        self.L = load_xcheck() if xcheck else load()
        h = vp()
        idbuf = ctypes.create_string_buffer(nccl_id, NCCL_ID_BYTES) if nccl_id else None
        _chk(self.L.gvb_ctx_create(ctypes.byref(h), device, rank, nranks, idbuf), self.L)
        self.h = h
        self.rank, self.nranks = rank, nranks

    @classmethod
    def shared(cls, parent: "Context") -> "Context":
        """A second context on the parent's device that shares its NCCL communicator (gvb_ctx_create_shared): another matrix
        (a test set, a small self-check problem) next to the resident one.  Close it before the parent."""
        self = cls.__new__(cls)
        self.L = parent.L
        h = vp()
        _chk(self.L.gvb_ctx_create_shared(ctypes.byref(h), parent.h), self.L)
        self.h = h
        self.rank, self.nranks = parent.rank, parent.nranks
        return self

    def close(self):
        if self.h:
            self.L.gvb_ctx_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- matrix
    def load_host(self, bed, N, Mt=None, S=0):
        bed = np.ascontiguousarray(bed, dtype=np.uint8)
        M = bed.shape[0]
        _chk(self.L.gvb_bed_load_host(self.h, bed.ctypes.data_as(c_u8p), N, Mt or M, S, M), self.L)
        return self

    def load_file(self, path, N, Mt, S, M):
        _chk(self.L.gvb_bed_load_file(self.h, path.encode(), N, Mt, S, M), self.L)
        return self

    def synth(self, seed, N, Mt, S, M, miss_rate=0.0):
        _chk(self.L.gvb_bed_synth(self.h, seed, N, Mt, S, M, miss_rate), self.L)
        return self

    def info(self):
        v = [cl(0) for _ in range(5)]
        _chk(self.L.gvb_ctx_info(self.h, *[ctypes.byref(x) for x in v]), self.L)
        return dict(zip(("N", "Mt", "S", "M", "mbytes"), (x.value for x in v)))

    def decode(self, j0, n):
        mb = self.info()["mbytes"]
        out = np.empty((n, mb), dtype=np.uint8)
        _chk(self.L.gvb_bed_decode(self.h, j0, n, out.ctypes.data_as(c_u8p)), self.L)
        return out

    def set_mask(self, mask4, nonas):
        if mask4 is None:
            _chk(self.L.gvb_set_mask(self.h, None, nonas), self.L)
        else:
            m = np.ascontiguousarray(mask4, dtype=np.uint8)
            _chk(self.L.gvb_set_mask(self.h, m.ctypes.data_as(c_u8p), nonas), self.L)
        return self

    def compute_stats(self, alpha_scale=1.0):
        _chk(self.L.gvb_compute_stats(self.h, alpha_scale), self.L)
        return self

    def stats(self):
        M = self.info()["M"]
        a, s = np.empty(M), np.empty(M)
        _chk(self.L.gvb_get_stats(self.h, a.ctypes.data_as(c_f64p), s.ctypes.data_as(c_f64p)), self.L)
        return a, s

    def counts(self):
        M = self.info()["M"]
        out = np.empty((M, 8), dtype=np.int64)
        _chk(self.L.gvb_get_counts(self.h, out.ctypes.data_as(c_i64p)), self.L)
        return out

    # ---- host-pointer drop-ins
    def Ax(self, v, SB=0, LB=None):
        inf = self.info()
        LB = inf["mbytes"] if LB is None else LB
        v, pv = _f64(v)
        assert len(v) == inf["M"]
        out = np.empty(4 * LB)
        _chk(self.L.gvb_Ax(self.h, pv, out.ctypes.data_as(c_f64p), SB, LB), self.L)
        return out

    def ATx(self, u, SB=0, LB=None):
        inf = self.info()
        LB = inf["mbytes"] if LB is None else LB
        uu = np.zeros(4 * LB)
        u = np.asarray(u, dtype=np.float64)
        uu[: min(len(u), 4 * LB)] = u[: 4 * LB]
        out = np.empty(inf["M"])
        _chk(self.L.gvb_ATx(self.h, uu.ctypes.data_as(c_f64p), out.ctypes.data_as(c_f64p), SB, LB), self.L)
        return out

    # ---- device vectors
    def vec(self, n):
        h = vp()
        _chk(self.L.gvb_vec_alloc(self.h, n, ctypes.byref(h)), self.L)
        return Vec(self, h)

    def vecM(self, init=None):
        h = vp()
        _chk(self.L.gvb_vec_alloc_M(self.h, ctypes.byref(h)), self.L)
        v = Vec(self, h)
        return v.upload(init) if init is not None else v

    def vecN(self, init=None):
        h = vp()
        _chk(self.L.gvb_vec_alloc_N(self.h, ctypes.byref(h)), self.L)
        v = Vec(self, h)
        return v.upload(init) if init is not None else v

    def axpby(self, out, a, x, b=0.0, y=None):
        _chk(self.L.gvb_vec_axpby(self.h, out.h, a, x.h, b, y.h if y is not None else None), self.L)

    def dots(self, xs, ys=None, sync=True):
        n = len(xs)
        X = (vp * n)(*[x.h for x in xs])
        Y = (vp * n)(*[(y.h if y is not None else None) for y in (ys or [None] * n)])
        res = np.empty(n)
        _chk(self.L.gvb_vec_dots(self.h, n, X, Y, int(sync), res.ctypes.data_as(c_f64p)), self.L)
        return res

    def reduce_batch(self, ops):
        """ops: list of (kind, x, y, a, b, sync); see gvb_vec_reduce_batch"""
        n = len(ops)
        arr = (RedOp * n)(*[RedOp(x.h, y.h if y is not None else None, a, b, kind, int(sync)) for kind, x, y, a, b, sync in ops])
        res = np.empty(n)
        _chk(self.L.gvb_vec_reduce_batch(self.h, n, ctypes.cast(arr, vp), res.ctypes.data_as(c_f64p)), self.L)
        return res

    def dist2(self, x, y, sync=True):
        res = np.empty(1)
        _chk(self.L.gvb_vec_dist2(self.h, x.h, y.h, int(sync), res.ctypes.data_as(c_f64p)), self.L)
        return float(res[0])

    def dAx(self, v, out):
        _chk(self.L.gvb_dAx(self.h, v.h, out.h), self.L)

    def dATx(self, u, out):
        _chk(self.L.gvb_dATx(self.h, u.h, out.h), self.L)

    def dAx2(self, v0, v1, out0, out1):
        """out0 = X.v0, out1 = X.v1 from one pass over the bed (gvb_dAx2)"""
        _chk(self.L.gvb_dAx2(self.h, v0.h, v1.h, out0.h, out1.h), self.L)

    def denoise(self, r1, gam1, probs, vars_, x1_hat):
        p, pp = _f64(probs)
        v, pv = _f64(vars_)
        sums = np.empty(2)
        _chk(self.L.gvb_denoise(self.h, r1.h, gam1, pp, pv, len(p), x1_hat.h, sums.ctypes.data_as(c_f64p)), self.L)
        return sums

    def em_stats(self, r1, gam1, lam, omegas, vars_):
        o, po = _f64(omegas)
        v, pv = _f64(vars_)
        L = len(o)
        sums = np.empty(2 * L - 1)
        _chk(self.L.gvb_em_stats(self.h, r1.h, gam1, lam, po, pv, L, sums.ctypes.data_as(c_f64p)), self.L)
        return sums

    def lmmse_mult(self, v, tau, gam2, out):
        _chk(self.L.gvb_lmmse_mult(self.h, v.h, tau, gam2, out.h), self.L)

    def cg_solve(self, rhs, mu, tau, gam2, max_iter, denoiser):
        it = ci(0)
        log = np.zeros(4 * max_iter)
        _chk(self.L.gvb_cg_solve(self.h, rhs.h, mu.h, tau, gam2, max_iter, denoiser, ctypes.byref(it), log.ctypes.data_as(c_f64p)), self.L)
        return it.value, log.reshape(max_iter, 4)[: it.value]

    def cg_solve_ex(self, rhs, mu, tau, gam2, max_iter, denoiser, ax_mu=None):
        it = ci(0)
        log = np.zeros(4 * max_iter)
        dots3 = np.zeros(3)
        _chk(self.L.gvb_cg_solve_ex(self.h, rhs.h, mu.h, tau, gam2, max_iter, denoiser, ctypes.byref(it), log.ctypes.data_as(c_f64p),
                                    ax_mu.h if ax_mu is not None else None, dots3.ctypes.data_as(c_f64p)))
        return it.value, log.reshape(max_iter, 4)[: it.value], dots3

    def cg_solve_warm(self, rhs, mu, tau, gam2, max_iter, denoiser, ax_mu, ata_mu, have_start):
        it = ci(0)
        log = np.zeros(4 * max_iter)
        dots3 = np.zeros(3)
        _chk(self.L.gvb_cg_solve_warm(self.h, rhs.h, mu.h, tau, gam2, max_iter, denoiser, ctypes.byref(it), log.ctypes.data_as(c_f64p),
                                      ax_mu.h, ata_mu.h, int(have_start), dots3.ctypes.data_as(c_f64p)))
        return it.value, log.reshape(max_iter, 4)[: it.value], dots3

    def cg_prepare(self, rhs, mu, tau, gam2, max_iter, ax_mu, ata_mu, have_start, extra_v, extra_out):
        """start of a solve + one dual sweep {A p0, extra_out = A extra_v} (gvb_cg_prepare)"""
        _chk(self.L.gvb_cg_prepare(self.h, rhs.h, mu.h, tau, gam2, max_iter, ax_mu.h if ax_mu is not None else None,
                                   ata_mu.h if ata_mu is not None else None, int(have_start), extra_v.h, extra_out.h), self.L)

    def cg_solve_prepared(self, rhs, mu, tau, gam2, max_iter, denoiser, ax_mu, ata_mu, have_start):
        it = ci(0)
        log = np.zeros(4 * max_iter)
        dots3 = np.zeros(3)
        _chk(self.L.gvb_cg_solve_prepared(self.h, rhs.h, mu.h, tau, gam2, max_iter, denoiser, ctypes.byref(it), log.ctypes.data_as(c_f64p),
                                          ax_mu.h if ax_mu is not None else None, ata_mu.h if ata_mu is not None else None, int(have_start),
                                          dots3.ctypes.data_as(c_f64p)), self.L)
        return it.value, log.reshape(max_iter, 4)[: it.value], dots3

    def cg_solve_cached(self, rhs, mu, tau, gam2, max_iter, denoiser, ata_rhs, state):
        """state: a one-element list holding 0 (cache empty) or 1 (filled); updated in place"""
        it, st = ci(0), ci(int(state[0]))
        log = np.zeros(4 * max_iter)
        dots3 = np.zeros(3)
        _chk(self.L.gvb_cg_solve_cached(self.h, rhs.h, mu.h, tau, gam2, max_iter, denoiser, ctypes.byref(it), log.ctypes.data_as(c_f64p),
                                        ata_rhs.h, ctypes.byref(st), dots3.ctypes.data_as(c_f64p)), self.L)
        state[0] = st.value
        return it.value, log.reshape(max_iter, 4)[: it.value], dots3

    def people_stats(self):
        a, s, n = self.vecN(), self.vecN(), self.vecN()
        _chk(self.L.gvb_people_stats(self.h, a.h, s.h, n.h), self.L)
        return a, s, n

    def cg_solve_aat(self, rhs, mu, tau, gam2, people, max_iter):
        it = ci(0)
        log = np.zeros(3 * max_iter)
        _chk(self.L.gvb_cg_solve_aat(self.h, rhs.h, mu.h, tau, gam2, people[0].h, people[1].h, people[2].h, max_iter, ctypes.byref(it),
                                     log.ctypes.data_as(c_f64p)))
        return it.value, log.reshape(max_iter, 3)[: it.value]

    def probit_denoise(self, p1, y, mcov, tau1, probit_var, z1_hat):
        sums = np.empty(2)
        _chk(self.L.gvb_probit_denoise(self.h, p1.h, y.h, mcov.h if mcov is not None else None, tau1, probit_var, z1_hat.h,
                                       sums.ctypes.data_as(c_f64p)))
        return sums

    # ---- timing / counters
    def sync(self):
        _chk(self.L.gvb_ctx_sync(self.h), self.L)

    def timer_start(self, slot=0):
        _chk(self.L.gvb_timer_start(self.h, slot), self.L)

    def timer_stop(self, slot=0):
        _chk(self.L.gvb_timer_stop(self.h, slot), self.L)

    def timer_ms(self, slot=0) -> float:
        ms = ctypes.c_float(0)
        _chk(self.L.gvb_timer_elapsed_ms(self.h, slot, ctypes.byref(ms)), self.L)
        return ms.value

    def profile(self, on=True):
        _chk(self.L.gvb_profile_enable(self.h, int(on)), self.L)

    def profile_read(self):
        out = np.zeros(4)
        _chk(self.L.gvb_profile_read(self.h, out.ctypes.data_as(c_f64p)), self.L)
        return dict(ax_ms=out[0], ax_n=int(out[1]), atx_ms=out[2], atx_n=int(out[3]))

    def dATx2(self, u0, u1, out0, out1):
        """out0 = X^T.u0, out1 = X^T.u1 from one pass over the bed (gvb_dATx2)"""
        _chk(self.L.gvb_dATx2(self.h, u0.h, u1.h, out0.h, out1.h), self.L)

    def profile_read_dual(self):
        out = np.zeros(2)
        _chk(self.L.gvb_profile_read_dual(self.h, out.ctypes.data_as(c_f64p)), self.L)
        return dict(dual_ms=out[0], dual_n=int(out[1]))

    def dual_sweeps(self) -> int:
        return self.L.gvb_dual_sweep_count(self.h)

    def launches(self) -> int:
        return self.L.gvb_launch_count(self.h)

    def sweeps(self) -> int:
        return self.L.gvb_sweep_count(self.h)

    def host_syncs(self) -> int:
        return self.L.gvb_host_sync_count(self.h)

    def layout_generation(self) -> int:
        return self.L.gvb_layout_generation(self.h)

    def probit_cov_pass(self, y, gg, Z, C, eta, probit_var=1.0, what=7):
        e, pe = _f64(eta)
        out = np.empty(1 + 2 * C + C * C)
        _chk(self.L.gvb_probit_cov_pass(self.h, y.h, gg.h if gg is not None else None, Z.h, C, pe, probit_var, what, out.ctypes.data_as(c_f64p)), self.L)
        return out[0], out[1:1 + C].copy(), out[1 + C:1 + 2 * C].copy(), out[1 + 2 * C:].reshape(C, C).copy()

    def probit_cov_apply(self, Z, C, eta, mcov):
        e, pe = _f64(eta)
        _chk(self.L.gvb_probit_cov_apply(self.h, Z.h, C, pe, mcov.h), self.L)

    def assoc_pvals(self, yres, coef, select, pvals):
        _chk(self.L.gvb_assoc_pvals(self.h, yres.h, coef.h if coef is not None else None, select.h if select is not None else None, pvals.h), self.L)

    def missing_list_entries(self) -> int:
        return self.L.gvb_missing_list_entries(self.h)

    def snapshot_begin(self, vec, n, slot):
        _chk(self.L.gvb_snapshot_begin(self.h, vec.h, n, slot), self.L)

    def snapshot_wait(self, slot):
        """copy of the pinned host buffer of a snapshot (blocks until the device -> host copy has landed)"""
        ptr, n = c_f64p(), cl()
        _chk(self.L.gvb_snapshot_wait(self.h, slot, ctypes.byref(ptr), ctypes.byref(n)), self.L)
        return np.ctypeslib.as_array(ptr, shape=(n.value,)).copy()

    def twin_state(self) -> int:
        """1: X.v walks the individual-major twin of the matrix, -1: no twin (too large / GVB_TWIN=0), 0: not decided yet"""
        return self.L.gvb_twin_state(self.h)

    def twin_release(self):
        _chk(self.L.gvb_twin_release(self.h), self.L)

    def twin_stripes(self) -> int:
        return self.L.gvb_twin_stripes(self.h)
