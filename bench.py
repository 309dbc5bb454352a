#!/usr/bin/env python
"""bench.py -- gVAMP iterations on synthetic PLINK-packed genotypes, one process per B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU implementation

A "step" is ONE gVAMP iteration (denoiser + EM prior update, z1 = X.x1, LMMSE by preconditioned CG, Onsager trace
estimate by a second CG, noise-precision update; every output of the reference's iteration is produced) over a
synthetic genotype matrix that lives in HBM.

Workloads (--workload):
  c4          (default) BASELINE.json config 4, N = 400,000 x Mt = 2,200,000 (220 GB packed), STRONG scaling: the whole
              matrix marker-sharded over the GPUs at 2 / 4 / 8 GPUs (110 / 55 / 27.5 GB per GPU).  220 GB does not fit one
              180 GB part, so at 1 GPU the matrix is cut to the largest marker count that fits (<= 1,600,000 markers =
              160 GB, decided from the free HBM and named in config.workload); `value` is iterations/s scaled by
              Mt/2.2M, i.e. in config-4 iterations per second (at >= 2 GPUs the scale factor is exactly 1).
  c4shard     the same matrix in WEAK scaling: 275,000 markers (27.5 GB) per GPU (= config 4 itself at 8 GPUs)
  config2     BASELINE config 2: N = 100,000 x Mt = 500,000 linear (12.5 GB), strong scaling
  config3     BASELINE config 3: N = 200,000 x Mt = 1,000,000 probit (bin_class) with C = 20 covariates, 1 or 2 GPUs
  config5     BASELINE config 5's shard: N = 400,000 x 1,050,000 markers per GPU (105 GB), 1 % missing genotypes; a step is one
              iteration of the LMMSE conjugate-gradient solve (X.v + X^T.u sweep pair + the fused updates)
  config1     BASELINE config 1 (10,000 x 20,000)          tiny: harness smoke test

Protocol: W warm-up iterations of a throw-away run (they also build the lazily created layouts), then a FRESH run timed
from its iteration 1 for exactly K iterations (CUDA events on the library's stream, barrier + synchronize on both sides,
max over ranks), then a second fresh run of the SAME K iterations end to end: y uploaded from pinned host memory every
step, every output vector read back and the reference's per-iteration files written.  Before anything is timed every rank
checks its shard against the oracle (`parity_check`, see below); a failed check fails the bench.  Inputs (>= 12 GB of
packed bed per sweep) are far larger than the 126 MB L2.  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import math
import os
import re
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIG4_MT = 2_200_000
CONFIG4_BYTES = 2_200_000 * 100_000   # packed bed bytes of config 4: Mt * ceil(N/4)
# name: (N, total markers (strong) or markers per GPU (weak), scaling, model, missing rate, description)
WORKLOADS = {
    "c4": (400_000, 2_200_000, "strong", "linear", 0.0, "config 4 (N=400000 x Mt=2200000, 220 GB packed) marker-sharded over the GPUs"),
    "c4shard": (400_000, 275_000, "weak", "linear", 0.0, "config 4 in weak scaling: 275000 markers (27.5 GB packed) per GPU"),
    "config2": (100_000, 500_000, "strong", "linear", 0.0, "config 2 (N=100000 x Mt=500000, 12.5 GB packed)"),
    "config3": (200_000, 1_000_000, "strong", "probit", 0.0, "config 3 (N=200000 x Mt=1000000 probit / bin_class, C=20 covariates, 50 GB packed)"),
    "config5": (400_000, 1_050_000, "weak", "sweeps", 0.01, "config 5 shard (N=400000 x 1050000 markers = 105 GB packed per GPU, 1% missing): CG iterations"),
    "config1": (10_000, 20_000, "strong", "linear", 0.0, "config 1 (N=10000 x Mt=20000, 50 MB packed)"),
    "tiny": (4_000, 8_000, "strong", "linear", 0.0, "smoke-sized workload for testing the harness"),
}
SINGLE_GPU_MAX_MARKERS = 1_600_000    # c4 at 1 GPU: upper bound of the cut (160 GB of the 180 GB part)
H2 = 0.5
RHO = 0.5
SEED = 20261017
METRIC = "gVAMP iter/s (N=400k,M=2.2M)"
UNIT = "iter/s (2.2M-marker equivalent)"


def default_prior(Mt):
    """The reference's 23-component default prior (utilities.cpp:91-140) written out explicitly so that the
    GPU arm and the CPU reference arm (whose sample has Mt < 50000) use the identical prior."""
    p = min(50000.0 / CONFIG4_MT, 1.0) / (2 - 1.0 / 2 ** 21)
    probs = [1 - 50000.0 / CONFIG4_MT]
    for _ in range(22):
        probs.append(p)
        p /= 2
    step = 10 ** (math.log10(1e2 / 1e-5) / 21)
    vars_ = [0.0]
    v = 1e-5
    for _ in range(22):
        vars_.append(v)
        v *= step
    return probs, vars_


def synth_truth(Mt, N, seed):
    """beta: Mt/110 causal markers ~ N(0, h2/CV); noise ~ N(0, 1-h2).  Same on every rank."""
    CV = max(1, Mt // 110)
    rng = np.random.Generator(np.random.Philox(key=seed))
    idx = rng.choice(Mt, size=CV, replace=False)
    beta = np.zeros(Mt)
    beta[idx] = rng.normal(0.0, math.sqrt(H2 / CV), size=CV)
    noise = rng.normal(0.0, math.sqrt(1.0 - H2), size=N)
    return beta, noise


def scale_like_read_phen(y):
    """data::read_phen scales y by 1/sd and does not centre it (data.cpp:171-186)."""
    n = len(y)
    avg = y.sum() / n
    return y * math.sqrt((n - 1) / ((y - avg) ** 2).sum())


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.gpu), "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, reasons, smax, power = [], set(), 0.0, []
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax = max(smax, float(r[2]))
                power.append(float(r[3]))
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax or None, "reasons": sorted(reasons), "samples": len(sm),
                "power_w_median": float(np.median(power)) if power else None}


class QuietStdout:
    """The C++ host classes print the reference's progress lines; keep them out of the JSON stream."""

    def __init__(self, path):
        self.path = path

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        fd = os.open(self.path, os.O_WRONLY | os.O_CREAT | os.O_APPEND, 0o644)
        os.dup2(fd, 1)
        os.close(fd)

    def __exit__(self, *a):
        ctypes.CDLL(None).fflush(None)
        os.dup2(self.saved, 1)
        os.close(self.saved)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel_key):
    """dram bytes per launch of the dominant kernel from a committed `ncu --set full` capture of exactly this kernel and shard
    size (profiles/traffic.json: {key: {"bytes": ..., "source": ...}}); None when no capture covers it."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        ent = json.load(open(p)).get(kernel_key)
        return (int(ent["bytes"]), ent.get("source")) if ent else (None, None)
    except Exception:
        return None, None


# ---------------------------------------------------------------------------------------------------
# the reference's CPU implementation on a bounded sample of the workload
# ---------------------------------------------------------------------------------------------------
def reference_exe():
    ref = os.path.join(ROOT, "oracle", "_ref")
    flags = open("/proc/cpuinfo").read() if os.path.exists("/proc/cpuinfo") else ""
    manvect = os.path.join(ref, "main_real_manvect.exe")
    scalar = os.path.join(ref, "main_real_scalar.exe")
    if os.path.exists(manvect) and "avx512f" in flags:
        return manvect, "reference main_real.exe (-DMANVECT -Ofast, MPI shim: 1 rank)"
    if os.path.exists(scalar):
        return scalar, "reference main_real.exe (scalar build -O2, MPI shim: 1 rank)"
    return None, None


def run_reference_sample(N, M_sample, iterations, cg_max_iter, threads, workdir, log=None):
    """Runs the UNMODIFIED reference (oracle/_ref) on N x M_sample synthetic data from its iteration 1; returns per-iteration seconds."""
    from oracle import oracle as O
    exe, kind = reference_exe()
    if exe is None:
        return None
    bed = O.synth_bed(SEED, 0, M_sample, N)
    bedp, phenp = os.path.join(workdir, "ref.bed"), os.path.join(workdir, "ref.phen")
    O.write_bed(bedp, bed)
    ds = O.Dataset(bed, N)
    beta, noise = synth_truth(M_sample, N, SEED)
    y = ds.Ax(beta * math.sqrt(N))[:N] + noise
    with open(phenp, "w") as fh:
        fh.write("".join(f"{i} {i} {v!r}\n" for i, v in enumerate(y.tolist())))
    probs, vars_ = default_prior(M_sample)
    outd = os.path.join(workdir, "refout") + "/"
    args = [exe, "--run-mode", "infere", "--model", "linear", "--bed-file", bedp, "--phen-files", phenp, "--N", str(N), "--Mt", str(M_sample),
            "--out-dir", outd, "--out-name", "ref", "--iterations", str(iterations), "--CG-max-iter", str(cg_max_iter), "--rho", str(RHO),
            "--probs", ",".join(repr(p) for p in probs), "--vars", ",".join(repr(v) for v in vars_), "--h2", str(H2), "--stop-criteria-thr", "1e-12"]
    env = dict(os.environ, OMP_NUM_THREADS=str(threads))
    t0 = time.time()
    r = subprocess.run(args, capture_output=True, text=True, env=env)
    wall = time.time() - t0
    if log:
        open(log, "w").write(r.stdout[-200000:])
    if r.returncode != 0:
        return None
    times = [float(m) for m in re.findall(r"total iteration time = ([0-9.eE+-]+)", r.stdout)]
    cg = len(re.findall(r"^\[CG\] it = ", r.stdout, flags=re.M))
    return {"iter_s": times, "wall_s": wall, "kind": kind, "cg_lines": cg, "bed_bytes": int(M_sample) * ((N + 3) // 4)}


def reference_arm(args):
    """bench.py --impl reference: the reference's own CPU implementation (main_real.exe built from /root/reference, oracle/_ref),
    all host threads, on a bounded marker sample of the workload; like this repo's arm it is timed from iteration 1 of a fresh
    run, W untimed iterations of a throw-away run first.  The sample's iterations/s are scaled by its packed-bed bytes to
    config-4 iterations/s (the reference's iteration time is linear in the marker count)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    N, _, scaling, model, _, desc = WORKLOADS[args.workload]
    threads = os.cpu_count() or 1
    M_sample = args.ref_markers
    with tempfile.TemporaryDirectory() as tmp:
        if args.warmup > 0:
            run_reference_sample(N, M_sample, min(args.warmup, 2), args.cg_max_iter, threads, tmp)
        res = run_reference_sample(N, M_sample, args.steps, args.cg_max_iter, threads, tmp)
    if res is None or len(res["iter_s"]) < args.steps:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref executables missing or the run failed on this host"}))
        return 0
    timed = res["iter_s"][:args.steps]
    s_per_step = float(np.mean(timed))
    value = (1.0 / s_per_step) * (res["bed_bytes"] / CONFIG4_BYTES)
    sample = (f"N={N} x M={M_sample} markers ({res['bed_bytes'] / 1e6:.0f} MB packed = {100.0 * res['bed_bytes'] / CONFIG4_BYTES:.3f}% of config 4), "
              f"iterations 1..{args.steps} of a fresh run")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * s_per_step, "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}", "sample": sample, "cg_max_iter": args.cg_max_iter, "h2": H2, "rho": RHO,
                   "cg_iterations_total": res["cg_lines"]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample, "build": res["kind"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------------------------
def load_host_lib():
    from gvamp_b200 import capi
    capi.load()
    path = os.path.join(ROOT, "gvamp_b200", "lib", "libgvamp_host.so")
    if not os.path.exists(path):
        raise RuntimeError(f"{path} missing: run python -m gvamp_b200.build")
    H = ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)
    vp, ci, cd = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
    f64p = ctypes.POINTER(ctypes.c_double)
    H.gvbh_options_create.restype = vp
    H.gvbh_options_create.argtypes = [ci, ctypes.POINTER(ctypes.c_char_p)]
    H.gvbh_data_create_resident.restype = vp
    H.gvbh_data_create_resident.argtypes = [vp, f64p, ci, ci, ci, ci, cd]
    H.gvbh_data_destroy.argtypes = [vp]
    H.gvbh_data_set_covs.argtypes = [vp, f64p, ci, ci]
    H.gvbh_vamp_create.restype = vp
    H.gvbh_vamp_create.argtypes = [vp, ci, cd, cd]
    H.gvbh_vamp_destroy.argtypes = [vp]
    H.gvbh_vamp_infere.argtypes = [vp, vp, f64p]
    H.gvbh_vamp_linear_begin.argtypes = [vp, vp]
    H.gvbh_vamp_linear_iteration.restype = ci
    H.gvbh_vamp_linear_iteration.argtypes = [vp, vp, ci, f64p, f64p]
    H.gvbh_vamp_linear_end.argtypes = [vp, f64p, ci]
    H.gvbh_vamp_cg_iters.argtypes = [vp, ctypes.POINTER(ci)]
    H.gvbh_vamp_gamw.restype = cd
    H.gvbh_vamp_gamw.argtypes = [vp]
    return H


class Dist:
    """torch.distributed plumbing of one rank: rendezvous, the NCCL id of the library's own communicator, tiny collectives."""

    def __init__(self, args):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        assert self.world == args.gpus or self.world == 1, f"--gpus {args.gpus} but WORLD_SIZE={self.world} (launch with torch.distributed.run)"
        # the C++ host layer learns its rank from the same environment (gvamp_b200/host/comm.cpp)
        os.environ["GVB_RANK"], os.environ["GVB_NRANKS"], os.environ["GVB_LOCAL_RANK"] = str(self.rank), str(self.world), str(self.local)
        import torch
        self.torch = torch
        torch.cuda.set_device(self.local)
        self.dist = None
        self.nccl_id = None
        if self.world > 1:
            import torch.distributed as dist
            from gvamp_b200 import capi
            self.dist = dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            idt = torch.zeros(capi.NCCL_ID_BYTES, dtype=torch.uint8, device="cuda")
            if self.rank == 0:
                idt.copy_(torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8))
            dist.broadcast(idt, 0)
            self.nccl_id = bytes(idt.cpu().numpy().tobytes())

    def barrier(self):
        if self.dist:
            self.dist.barrier()

    def reduce(self, values, op="max"):
        if not self.dist:
            return [float(v) for v in values]
        t = self.torch.tensor(list(values), device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op={"max": self.dist.ReduceOp.MAX, "sum": self.dist.ReduceOp.SUM, "min": self.dist.ReduceOp.MIN}[op])
        return [float(x) for x in t.cpu()]

    def close(self):
        if self.dist:
            self.dist.destroy_process_group()


def parity_check(D, ctx, N, Mt, S, M, miss):
    """Every rank's shard against the oracle before anything is timed (the driver's GPU test box has ONE GPU, so this is where the
    marker-sharded path is checked at 2 / 4 / 8 GPUs).  The oracle is the checker only; nothing here is timed.
      small:    a 3000 x 2501 matrix (1 % missing) sharded by divide_work over the same communicator: shard bytes bit-exact, X.v
                (all-reduced) and X^T.u within 1e-6, CG solution within 1e-5 with the oracle's iteration count, denoiser sums;
      fullsize: on the resident workload matrix: three sampled markers per rank re-generated by the oracle (bytes bit-exact,
                mu / 1/sigma 1e-10, their X^T.u entries and the all-reduced X.v of a vector supported on all ranks' samples 1e-6),
                adjointness <X v, u> = sum_shards <v, X^T u> with dense vectors, bit-reproducibility of both sweeps."""
    from gvamp_b200 import capi
    from oracle import oracle as O
    rel = lambda a, b: float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))
    out = {"small": {}, "fullsize": {}}
    # ---- small
    n, mt, seed = 3000, 2501, 77
    m, s = capi.divide_work(mt, D.world, D.rank)
    sm = capi.Context.shared(ctx)
    try:
        sm.synth(seed, n, mt, s, m, 0.01)
        sm.compute_stats(1.0)
        bed = O.synth_bed(seed, 0, mt, n, miss_rate=0.01)
        ds = O.Dataset(bed, n)
        rng = np.random.default_rng(5)
        v, u = rng.normal(size=mt), rng.normal(size=n)
        sml = out["small"]
        sml["bytes_equal"] = bool(np.array_equal(sm.decode(0, m), bed[s:s + m]))
        sml["counts_equal"] = bool(np.array_equal(sm.counts(), ds.counts()[s:s + m]))
        sml["ax_rel"] = rel(sm.Ax(v[s:s + m]), ds.Ax(v))
        sml["atx_rel"] = rel(sm.ATx(u), ds.ATx(u)[s:s + m])
        rhs, mu = sm.vecM(v[s:s + m]), sm.vecM()
        its, _ = sm.cg_solve(rhs, mu, 2.0, 0.7, 25, 1)
        mu_ref, its_ref = O.precond_cg(ds, v, np.zeros(mt), 2.0, 0.7, 25, 1)
        sml["cg_iters"], sml["cg_iters_oracle"] = int(its), int(its_ref)
        sml["cg_rel"] = rel(mu.download(), mu_ref[s:s + m])
        x1 = sm.vecM()
        probs, vars_ = [0.9, 0.06, 0.04], [0.0, 0.1, 1.0]
        sums = sm.denoise(rhs, 3.0, probs, vars_, x1)
        sml["denoise_rel"] = abs(sums[0] - O.g1d(v, 3.0, probs, vars_).sum()) / mt
        ok_small = (sml["bytes_equal"] and sml["counts_equal"] and sml["ax_rel"] < 1e-6 and sml["atx_rel"] < 1e-6 and its == its_ref
                    and sml["cg_rel"] < 1e-5 and sml["denoise_rel"] < 1e-8)
    finally:
        sm.close()
    # ---- full size, on the resident matrix
    fs = out["fullsize"]
    picks = {r: sorted({0, capi.divide_work(Mt, D.world, r)[0] // 2, capi.divide_work(Mt, D.world, r)[0] - 1}) for r in range(D.world)}
    glob = [(r, capi.divide_work(Mt, D.world, r)[1] + j) for r in range(D.world) for j in picks[r]]      # (owner, global marker)
    rows = np.concatenate([O.synth_bed(SEED, gj, 1, N, miss_rate=miss) for _, gj in glob])
    dss = O.Dataset(rows, N)
    mine = [k for k, (r, _) in enumerate(glob) if r == D.rank]
    loc = [glob[k][1] - S for k in mine]
    fs["bytes_equal"] = bool(all(np.array_equal(ctx.decode(j, 1)[0], rows[k]) for k, j in zip(mine, loc)))
    mave, msig = ctx.stats()
    fs["stats_rel"] = max(rel(mave[loc], dss.mave[mine]), rel(msig[loc], dss.msig[mine]))
    rng = np.random.default_rng(11)
    u = rng.normal(size=N)                                                    # same on every rank
    vs = np.random.default_rng(12).normal(size=len(glob))
    atx = ctx.ATx(u)
    fs["atx_sampled_rel"] = rel(atx[loc], dss.ATx(u)[mine])
    vloc = np.zeros(M)
    vloc[loc] = vs[mine]
    ax_sparse = ctx.Ax(vloc)                                                  # all-reduced over the shards
    fs["ax_sampled_rel"] = rel(ax_sparse[:N], dss.Ax(vs)[:N])
    vd = np.random.default_rng(1000 + D.rank).normal(size=M)
    ax = ctx.Ax(vd)
    lhs = float(np.dot(ax[:N], u))
    rhs_sum = D.reduce([float(np.dot(vd, atx))], "sum")[0]
    fs["adjointness_rel"] = abs(lhs - rhs_sum) / float(np.linalg.norm(ax[:N]) * np.linalg.norm(u))
    fs["bit_reproducible"] = bool(np.array_equal(ctx.Ax(vd), ax) and np.array_equal(ctx.ATx(u), atx))
    ok_full = (fs["bytes_equal"] and fs["stats_rel"] < 1e-10 and fs["atx_sampled_rel"] < 1e-6 and fs["ax_sampled_rel"] < 1e-6
               and fs["adjointness_rel"] < 1e-6 and fs["bit_reproducible"])
    ok_all = D.reduce([1.0 if (ok_small and ok_full) else 0.0], "min")[0] == 1.0
    # the worst rank's figures travel to rank 0
    for grp in ("small", "fullsize"):
        keys = [k for k, v in out[grp].items() if isinstance(v, float)]
        worst = D.reduce([out[grp][k] for k in keys], "max")
        for k, w in zip(keys, worst):
            out[grp][k] = w
    out["ok"] = bool(ok_all)
    out["ranks"] = D.world
    out["tolerances"] = {"bytes / counts": "bit-exact", "stats": 1e-12, "X.v, X^T.u, adjointness": 1e-6, "CG": 1e-5}
    return out


def pick_single_gpu_markers(mbytes):
    """c4 on one GPU: the largest marker count (multiple of 50,000, <= 1.6M) whose packed matrix leaves 6 GB of the free HBM."""
    import torch
    free, _total = torch.cuda.mem_get_info()
    fit = int((free - (6 << 30)) // mbytes)
    return max(50_000, min(SINGLE_GPU_MAX_MARKERS, fit // 50_000 * 50_000))


def ours_arm(args):
    from gvamp_b200 import capi
    D = Dist(args)
    rank, world, local = D.rank, D.world, D.local
    H = load_host_lib()
    N, m_cfg, scaling, model, miss, desc = WORKLOADS[args.workload]
    mbytes = (N + 3) // 4
    if args.markers_per_gpu:
        Mt, scaling = args.markers_per_gpu * world, "weak"
    elif scaling == "weak":
        Mt = m_cfg * world
    else:
        Mt = m_cfg
        if args.workload == "c4" and world == 1:
            Mt = pick_single_gpu_markers(mbytes)
            desc += f"; 220 GB does not fit one 180 GB GPU: cut to the largest single-GPU fit, {Mt} markers ({Mt * mbytes / 1e9:.0f} GB packed)"
    M, S = capi.divide_work(Mt, world, rank)
    bed_bytes_local = M * mbytes
    work = Mt * mbytes / CONFIG4_BYTES   # this job's packed-bed size in units of config 4's

    logf = os.path.join(tempfile.gettempdir(), f"gvamp_bench_rank{rank}.log")
    if os.path.exists(logf):
        os.remove(logf)
    ctx = capi.Context(local, rank, world, D.nccl_id)
    t_setup = time.time()
    ctx.synth(SEED, N, Mt, S, M, miss)
    ctx.compute_stats(1.0)
    t_setup = time.time() - t_setup
    parity = None
    if not args.no_parity_check:
        t_par = time.time()
        parity = parity_check(D, ctx, N, Mt, S, M, miss)
        parity["seconds"] = time.time() - t_par
        if not parity["ok"]:
            if rank == 0:
                print("bench.py: parity_check FAILED, nothing was timed: " + json.dumps(parity), file=sys.stderr)
            ctx.close()
            D.close()
            return 3
    runner = {"linear": run_linear, "probit": run_probit, "sweeps": run_sweeps}[model]
    res = runner(args, D, ctx, H, N, Mt, S, M, logf)

    K, W = args.steps, args.warmup
    ms_dev, ms_e2e = D.reduce([res["ms_dev"], res["ms_e2e"]], "max")
    value = K / (ms_dev / 1e3) * work
    e2e_value = K / (ms_e2e / 1e3) * work
    peak, peak_src = measured_peak()
    prof = res["prof"]
    # dominant kernel = the sweep kind with the larger total time; algorithmic bytes per launch = M_local * ceil(N/4)
    per = {"X.v": (prof["ax_ms"], prof["ax_n"]), "X^T.u": (prof["atx_ms"], prof["atx_n"])}
    dom = max(per, key=lambda k: per[k][0])
    gbs = {k: (bed_bytes_local / 1e9) / (ms / n / 1e3) if n else None for k, (ms, n) in per.items()}
    twin = ctx.twin_state()
    kname = {"X.v": "ax_pair_kernel<twin>" if twin == 1 else ("ax_pair_kernel<twin>+ax_tile_kernel<gather>" if twin == 2 else "ax_tile_kernel<gather>"),
             "X^T.u": "atx_pair_kernel" + ("+miss_sum_kernel" if miss > 0 else "")}[dom]
    traffic, traffic_src = ncu_traffic(f"{kname}@{bed_bytes_local}")
    dual_ms, dual_n = prof.get("dual_ms", 0.0), prof.get("dual_n", 0)
    sweep_ms_all = prof["ax_ms"] + prof["atx_ms"] + dual_ms
    n_sw = prof["ax_n"] + prof["atx_n"] + dual_n
    roofline = {"bound": "hbm", "kernel": f"{dom}: {kname}", "achieved": gbs[dom], "peak": peak, "unit": "GB/s",
                "frac": gbs[dom] / peak if gbs[dom] else None, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bed_bytes_local, "per_kernel_GBps": gbs,
                "sweep_ms": {k: (ms / n if n else None) for k, (ms, n) in per.items()},
                "sweep_share_of_step": sweep_ms_all / res["ms_dev"],
                "dual_sweeps": {"n": dual_n, "ms_each": dual_ms / dual_n if dual_n else None,
                                "bed_GBps": (bed_bytes_local / 1e9) / (dual_ms / dual_n / 1e3) if dual_n else None,
                                "note": "X.v for two right-hand sides from ONE bed read (z1 = A x1_hat + the first product of the LMMSE solve); "
                                        "not part of `achieved`, which is the single-product X.v / X^T.u sweeps"},
                "note": "achieved = packed bed bytes of the local shard / mean CUDA-event time of all launches of one sweep (scale, table build, main kernel, "
                        "finish), measured inside the timed region on this rank"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}", "N": N, "Mt": Mt, "markers_this_gpu": M, "packed_GB_this_gpu": bed_bytes_local / 1e9,
                   "cg_max_iter": args.cg_max_iter, "h2": H2, "rho": RHO, "prior": "reference default 23-component",
                   "protocol": f"{W} warm-up iterations of a throw-away run, then iterations 1..{K} of a fresh run (device-timed), then the same "
                               f"iterations 1..{K} of another fresh run end to end (host y upload, output read-back, files written)",
                   "iter_per_s_unscaled": K / (ms_dev / 1e3), "config4_scale_factor": work,
                   "sweeps_per_step": res["sweeps"], "cg_iters_per_step": res["cg"], "ms_per_sweep": sweep_ms_all / max(n_sw, 1),
                   "non_sweep_ms_per_step": (res["ms_dev"] - sweep_ms_all) / K, "host_syncs_per_step": res["host_syncs"] / K,
                   "l2_policy": "inputs (>= 12 GB packed bed per sweep at the default workload) larger than the 126 MB L2",
                   "kernels": os.environ.get("GVB_KERNELS", "tile (gen 2)"), "twin_layout": twin, "setup_s": t_setup,
                   "onsager_solve": "by bed sweeps" if (os.environ.get("GVB_ONSAGER_LANCZOS") == "0" or os.environ.get("GVB_REFERENCE_SWEEPS") == "1")
                   else "the reference's CG on the cached Lanczos projection of A^T A (no bed sweep after the first iteration; DESIGN.md section 7)",
                   **res.get("extra", {})},
        "roofline": roofline,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": res["h2d"], "d2h_bytes_per_step": res["d2h"],
                "ms_per_step": ms_e2e / K, "sweeps_per_step": res["e2e_sweeps"], "files_written_per_step": res.get("files", 0), "note": res["e2e_note"]},
        "gpu_launches": res["launches"],
        "clocks": res["clocks"],
        "parity_check": parity,
    }
    # ---- CPU baseline: the reference itself on a bounded sample, rank 0 at N=1 only
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        with tempfile.TemporaryDirectory() as tmp:
            cres = run_reference_sample(N, args.ref_markers, 2, args.cg_max_iter, threads, tmp)
        if cres and cres["iter_s"]:
            s = float(np.mean(cres["iter_s"]))
            line["cpu_baseline"] = {"value": (1.0 / s) * (cres["bed_bytes"] / CONFIG4_BYTES), "unit": UNIT, "cores": threads, "kind": "reference",
                                    "sample": f"N={N} x M={args.ref_markers} markers ({cres['bed_bytes'] / 1e6:.0f} MB packed), iterations 1..2 of a linear-model run",
                                    "s_per_iteration_on_sample": s, "build": cres["kind"]}
        else:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": threads, "kind": "reference", "sample": "oracle/_ref unavailable on this host"}
    ctx.close()
    D.close()
    if rank == 0:
        keep = os.environ.get("GVB_BENCH_KEEP_LOG")     # the host layer's progress lines (per-phase wall times) of rank 0
        if keep:
            shutil.copyfile(logf, keep)
        print(json.dumps(line))
    return 0


def make_phenotype(ctx, N, Mt, S, M):
    beta, noise = synth_truth(Mt, N, SEED)
    g = ctx.Ax(beta[S:S + M] * math.sqrt(N))[:N]          # X_std beta over all shards (NCCL allreduce inside)
    return scale_like_read_phen(g + noise)


def run_linear(args, D, ctx, H, N, Mt, S, M, logf):
    """W iterations of a throw-away run; iterations 1..K of a fresh run on device-resident state; iterations 1..K of another
    fresh run end to end."""
    torch = D.torch
    K, W = args.steps, args.warmup
    mbytes = (N + 3) // 4
    y = make_phenotype(ctx, N, Mt, S, M)
    probs, vars_ = default_prior(Mt)
    outdir = os.path.join(tempfile.gettempdir(), "gvamp_bench_out") + "/"
    if D.rank == 0:
        shutil.rmtree(outdir, ignore_errors=True)
        os.makedirs(outdir, exist_ok=True)
    D.barrier()
    argv = ["bench", "--bed-file", "synthetic-in-hbm", "--N", str(N), "--Mt", str(Mt), "--iterations", str(max(W, K)), "--CG-max-iter",
            str(args.cg_max_iter), "--rho", str(RHO), "--probs", ",".join(repr(p) for p in probs), "--vars", ",".join(repr(v) for v in vars_),
            "--h2", str(H2), "--stop-criteria-thr", "1e-12", "--out-dir", outdir, "--out-name", "gvamp_bench", "--model", "linear", "--run-mode", "infere"]
    carr = (ctypes.c_char_p * len(argv))(*[a.encode() for a in argv])
    f64p = ctypes.POINTER(ctypes.c_double)
    ypin = torch.from_numpy(np.ascontiguousarray(y)).pin_memory()
    yptr = ctypes.cast(ypin.data_ptr(), f64p)
    cg_iters = (ctypes.c_int * 2)()
    out = {}

    with QuietStdout(logf):
        opt = H.gvbh_options_create(len(argv), carr)
        dat = H.gvbh_data_create_resident(ctx.h, yptr, N, M, Mt, S, 1.0)

        def fresh_run(steps, upload, files, timer):
            """One run from iteration 1; returns per-step sweeps / CG iterations and the device time of all its steps."""
            os.environ["GVB_NO_FILES"] = "0" if files else "1"
            vmp = H.gvbh_vamp_create(opt, M, 1e-6, 1.0 / (1.0 - H2))
            H.gvbh_vamp_linear_begin(vmp, dat)
            sweeps, cg, wall = [], [], []
            D.barrier()
            ctx.sync()
            ctx.timer_start(timer)
            for it in range(1, steps + 1):
                s0, t0 = ctx.sweeps(), time.perf_counter()
                H.gvbh_vamp_linear_iteration(vmp, dat, it, yptr if upload else None, None)
                H.gvbh_vamp_cg_iters(vmp, cg_iters)
                wall.append(1e3 * (time.perf_counter() - t0))   # an iteration ends with a host-visible reduction: wall clock ~ device time
                sweeps.append(ctx.sweeps() - s0)
                cg.append((cg_iters[0], cg_iters[1]))
            H.gvbh_vamp_linear_end(vmp, None, 0)       # waits for the last iteration's files
            ctx.timer_stop(timer)
            ctx.sync()
            D.barrier()
            ms = ctx.timer_ms(timer)
            gamw = H.gvbh_vamp_gamw(vmp)
            H.gvbh_vamp_destroy(vmp)
            return sweeps, cg, ms, gamw, wall

        if W > 0:
            fresh_run(W, False, False, 2)
        # ---- timed region 1: device-resident state, iterations 1..K
        sampler = ClockSampler(D.local) if D.rank == 0 else None
        if sampler:
            sampler.start()
            time.sleep(0.3)
        ctx.profile(True)
        launches0, syncs0 = ctx.launches(), ctx.host_syncs()
        out["sweeps"], out["cg"], out["ms_dev"], gamw, wall = fresh_run(K, False, False, 0)
        out["launches"] = ctx.launches() - launches0
        out["host_syncs"] = ctx.host_syncs() - syncs0
        out["prof"] = {**ctx.profile_read(), **ctx.profile_read_dual()}
        ctx.profile(False)
        out["clocks"] = sampler.stop() if sampler else None
        # ---- timed region 2: the same iterations end to end
        out["e2e_sweeps"], _, out["ms_e2e"], gamw2, _ = fresh_run(K, True, True, 1)
        H.gvbh_data_destroy(dat)
    D.barrier()
    nfiles = len([f for f in os.listdir(outdir) if "_it_" in f]) if D.rank == 0 else 0
    if D.rank == 0:
        shutil.rmtree(outdir, ignore_errors=True)     # ~1 GB of per-iteration files at the default workload
    out["files"] = nfiles / K
    out["h2d"] = 8 * N
    out["d2h"] = 8 * (4 * M + 4 * mbytes)
    out["extra"] = {"final_gamw": gamw, "final_gamw_e2e": gamw2, "ms_per_step_each_host_clock": [round(w, 2) for w in wall],
                    "ms_per_step_median_host_clock": float(np.median(wall)),
                    "note_on_K": "iterations/s of a run timed from its iteration 1 depends on K by construction (iteration 1 solves from zero: 4x the sweeps of a "
                                 "late iteration); ms_per_sweep and the per-step list are the K-independent figures"}
    out["e2e_note"] = ("a fresh run of the same iterations: y re-uploaded from pinned host memory every step (the library compares it bit for bit with "
                       "the resident copy on the device and keeps A^T y while it is unchanged, as in a real run, whose y never changes; "
                       "GVB_ATY_CACHE=0 recomputes A^T y after every upload: one more sweep per step), "
                       "x1_hat / r1 / r2 / x2_hat / z1 read back every step into pinned host memory, scaled and written to the reference's "
                       "per-iteration files by a background task; the timed region ends when the last file is closed")
    return out


def run_probit(args, D, ctx, H, N, Mt, S, M, logf):
    """config 3: vamp::infere_bin_class (vamp_probit.cpp:20-658) with C = 20 covariates, K iterations of a fresh run (the covariate
    Newton solve of iteration 1 included), device-timed; then the same run end to end with its files."""
    from scipy.special import ndtr
    K, W, Cc = args.steps, args.warmup, 20
    rng = np.random.Generator(np.random.Philox(key=33))      # identical on every rank
    CV = Mt // 200
    beta = np.zeros(Mt)
    beta[rng.choice(Mt, size=CV, replace=False)] = rng.normal(0.0, math.sqrt(0.5 / CV), size=CV)
    Z = rng.normal(size=(N, Cc))
    eta = 0.25 * (1 - 2 * (np.arange(Cc) % 2))
    g = ctx.Ax(beta[S:S + M] * math.sqrt(N))[:N] + Z @ eta
    y = (rng.random(N) <= ndtr(g)).astype(np.float64)
    # the 3-component prior of the golden probit case: the reference's built-in 23-component prior drives its own probit recursion
    # out of its domain on this synthetic at Mt/N = 5
    probs, vars_ = [0.9, 0.06, 0.04], [0.0, 1e-4, 1e-3]
    outdir = os.path.join(tempfile.gettempdir(), "gvamp_bench_out") + "/"
    if D.rank == 0:
        shutil.rmtree(outdir, ignore_errors=True)
        os.makedirs(outdir, exist_ok=True)
    D.barrier()
    f64p = ctypes.POINTER(ctypes.c_double)
    yc, Zc = np.ascontiguousarray(y), np.ascontiguousarray(Z)
    xout = np.zeros(M)
    out = {}

    def run(iters, files, timer):
        os.environ["GVB_NO_FILES"] = "0" if files else "1"
        argv = ["probit", "--bed-file", "synthetic-in-hbm", "--N", str(N), "--Mt", str(Mt), "--iterations", str(iters), "--CG-max-iter",
                str(args.cg_max_iter), "--rho", "0.5", "--probs", ",".join(repr(p) for p in probs), "--vars", ",".join(repr(v) for v in vars_),
                "--stop-criteria-thr", "1e-12", "--out-dir", outdir, "--out-name", "gvamp_bench", "--model", "bin_class", "--run-mode", "infere",
                "--C", str(Cc)]
        carr = (ctypes.c_char_p * len(argv))(*[x.encode() for x in argv])
        opt = H.gvbh_options_create(len(argv), carr)
        dat = H.gvbh_data_create_resident(ctx.h, yc.ctypes.data_as(f64p), N, M, Mt, S, 1.0)
        H.gvbh_data_set_covs(dat, Zc.ctypes.data_as(f64p), N, Cc)
        vmp = H.gvbh_vamp_create(opt, M, 1e-8, 1.0)
        s0 = ctx.sweeps()
        D.barrier()
        ctx.sync()
        ctx.timer_start(timer)
        H.gvbh_vamp_infere(vmp, dat, xout.ctypes.data_as(f64p))
        ctx.timer_stop(timer)
        ctx.sync()
        D.barrier()
        ms = ctx.timer_ms(timer)
        sw = ctx.sweeps() - s0
        H.gvbh_vamp_destroy(vmp)
        H.gvbh_data_destroy(dat)
        return ms, sw

    with QuietStdout(logf):
        if W > 0:
            run(min(W, 2), False, 2)
        sampler = ClockSampler(D.local) if D.rank == 0 else None
        if sampler:
            sampler.start()
            time.sleep(0.3)
        ctx.profile(True)
        launches0, syncs0 = ctx.launches(), ctx.host_syncs()
        out["ms_dev"], sw = run(K, False, 0)
        out["launches"] = ctx.launches() - launches0
        out["host_syncs"] = ctx.host_syncs() - syncs0
        out["prof"] = {**ctx.profile_read(), **ctx.profile_read_dual()}
        ctx.profile(False)
        out["clocks"] = sampler.stop() if sampler else None
        out["ms_e2e"], sw2 = run(K, True, 1)
    log = open(logf).read()
    cov = [float(tok.split("=")[1]) for l in log.splitlines() if l.startswith("cov_eff[") for tok in l.split(",") if "=" in tok][-Cc:]
    corr = [float(x) for x in re.findall(r"correlation x1_hat = ([0-9.eE+-]+)", log)]
    out["sweeps"], out["cg"], out["e2e_sweeps"] = [sw], [], [sw2]
    out["h2d"] = 8 * N * (1 + Cc) // max(K, 1)
    out["d2h"] = 8 * 2 * M
    out["extra"] = {"model": "bin_class", "C": Cc, "prior": "3-component (golden probit case)", "cov_eff_estimated": cov, "cov_eff_true": eta.tolist(),
                    "corr_x1_truth_last": corr[-1] if corr else None, "sweeps_per_step": [sw / K]}
    out["e2e_note"] = ("vamp::infere through the host class: y and the N x C covariate matrix uploaded once per run (the probit model keeps them resident), "
                       "x1_hat / r1 read back and written every iteration")
    return out


def run_sweeps(args, D, ctx, H, N, Mt, S, M, logf):
    """config 5 shard: a step is one iteration of the LMMSE conjugate-gradient solve (tau X^T X + gam2 I) mu = rhs on the 105 GB
    shard with 1 % missing genotypes: one X.v sweep (all-reduced), one X^T.u sweep incl. the missing-genotype gather, the fused
    updates; the solver runs exactly K iterations (its residual exit cannot fire: the right-hand side is dense noise)."""
    K, W = args.steps, args.warmup
    rng = np.random.default_rng(1000 + D.rank)
    v_h = rng.normal(size=M)
    rhs, mu = ctx.vecM(v_h), ctx.vecM()
    out = {}
    if W > 0:
        ctx.cg_solve(rhs, mu, 2.0, 1e-3, W, 1)
    sampler = ClockSampler(D.local) if D.rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    mu.fill(0.0)
    ctx.profile(True)
    launches0, syncs0, s0 = ctx.launches(), ctx.host_syncs(), ctx.sweeps()
    D.barrier()
    ctx.sync()
    ctx.timer_start(0)
    its, log = ctx.cg_solve(rhs, mu, 2.0, 1e-3, K, 1)
    ctx.timer_stop(0)
    ctx.sync()
    D.barrier()
    out["ms_dev"] = ctx.timer_ms(0) * (K / max(its, 1))
    out["launches"] = ctx.launches() - launches0
    out["host_syncs"] = ctx.host_syncs() - syncs0
    out["prof"] = {**ctx.profile_read(), **ctx.profile_read_dual()}
    ctx.profile(False)
    out["clocks"] = sampler.stop() if sampler else None
    out["sweeps"] = [ctx.sweeps() - s0]
    out["cg"] = [(int(its), 0)]
    # end to end: the right-hand side comes from host memory and the solution goes back
    D.barrier()
    ctx.sync()
    s0 = ctx.sweeps()
    ctx.timer_start(1)
    rhs.upload(v_h)
    mu.fill(0.0)
    its2, _ = ctx.cg_solve(rhs, mu, 2.0, 1e-3, K, 1)
    sol = mu.download()
    ctx.timer_stop(1)
    ctx.sync()
    D.barrier()
    out["ms_e2e"] = ctx.timer_ms(1) * (K / max(its2, 1))
    out["e2e_sweeps"] = [ctx.sweeps() - s0]
    out["h2d"], out["d2h"] = 8 * M // max(K, 1), 8 * M // max(K, 1)
    out["extra"] = {"step": "one CG iteration (X.v + X^T.u sweep pair + fused updates)", "cg_residuals": np.asarray(log)[:, 0].tolist(),
                    "missing_list_entries": ctx.missing_list_entries(), "solution_finite": bool(np.all(np.isfinite(sol)))}
    out["e2e_note"] = "right-hand side uploaded from host memory, K CG iterations, solution read back; per-step bytes = vector bytes / K"
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--markers-per-gpu", type=int, default=0, help="override: this many markers on every GPU (weak scaling)")
    ap.add_argument("--cg-max-iter", type=int, default=20)
    ap.add_argument("--ref-markers", type=int, default=1536, help="markers in the CPU reference's bounded sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    args.steps = max(args.steps, 1)
    if args.impl == "reference":
        return reference_arm(args)
    return ours_arm(args)


if __name__ == "__main__":
    sys.exit(main())
