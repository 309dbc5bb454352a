"""Marker-sharded multi-GPU parity (NCCL allreduce of X.v, fused scalar allreduces); needs >= 2 GPUs."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_two_rank_parity():
    from gvamp_b200 import capi
    if capi.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port",
           "29511", os.path.join(ROOT, "tests", "mgpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "mgpu_check ok" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_main_real_two_ranks(tmp_path):
    """The product executable itself, two ranks started by `torchrun --no-python`: its own NCCL rendezvous
    (host/comm.cpp), marker sharding by divide_work, every rank writing its slice of the output files at byte offset S*8.
    A multi-rank run of the reference differs from its single-rank run only through the Onsager probe - shard S draws it from
    mt19937{seed+S} (vamp.cpp:875-882) - and through summation order.  oracle.infere_linear(nranks=2) emulates exactly that
    (the per-shard probes laid end to end, pinned shard by shard against the live reference in
    tests/test_oracle_vs_reference_live_cpu.py::test_sharded_probe_against_the_live_reference), so the two-rank run is held
    to the north_star's 1e-4 like the single-rank runs."""
    import math
    import numpy as np
    from gvamp_b200 import capi
    from oracle import oracle
    from conftest import golden, relerr
    if capi.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    g = golden("vamp_linear.npz")
    N, M, iters = int(g["N"]), int(g["M"]), int(g["iterations"])
    bed = oracle.synth_bed(int(g["seed"]), 0, M, N)
    bedp, phenp = str(tmp_path / "v.bed"), str(tmp_path / "v.phen")
    oracle.write_bed(bedp, bed)
    oracle.write_phen(phenp, g["y"])
    exe = os.path.join(ROOT, "gvamp_b200", "bin", "main_real")
    extra = [str(a) for a in g["args"]]
    outd = str(tmp_path / "out2") + "/"
    args = ["--run-mode", "infere", "--model", "linear", "--bed-file", bedp, "--phen-files", phenp, "--N", str(N), "--Mt", str(M), "--out-dir", outd,
            "--out-name", "g"]
    for k in range(0, len(extra), 2):
        if extra[k] not in {"--N", "--Mt", "--out-dir", "--out-name"}:
            args += [extra[k], extra[k + 1]]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--no-python", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", exe] + args
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    # the emulated two-rank reference
    y = g["y"]
    avg = float(np.cumsum(y)[-1]) / N
    sqn = math.sqrt((N - 1) / float(np.cumsum((y - avg) * (y - avg))[-1]))
    ds = oracle.Dataset(bed, N, phen=y * sqn)
    cfg = oracle.VampConfig(iterations=iters, rho=0.5, probs=(0.9, 0.06, 0.04), vars=(0, 1e-4, 1e-3), CG_max_iter=20,
                            gamw=1.0 / (1.0 - float(g["h2"])), seed=1, nranks=2)
    tr = oracle.infere_linear(ds, cfg)
    assert relerr(tr.x1_hat[-1], g[f"x1_{iters}"]) > 1e-4     # the two-rank probe does move the result: the check below is not vacuous
    for it in range(1, iters + 1):
        for key, arr in (("", tr.x1_hat), ("_r1", tr.r1), ("_r2", tr.r2)):
            got = np.fromfile(outd + f"g{key}_it_{it}.bin")
            assert got.shape == (M,)
            if np.linalg.norm(arr[it - 1]) > 0:
                assert relerr(got, arr[it - 1]) < 1e-4, (key, it, relerr(got, arr[it - 1]))
            else:
                assert not got.any()
        assert relerr(np.fromfile(outd + f"g_it_{it}_x2_hat.bin"), tr.x2_hat[it - 1]) < 1e-4
    assert np.allclose(np.loadtxt(outd + "g_gam1s.csv"), tr.gam1s, rtol=1e-4)
    assert np.allclose(np.loadtxt(outd + "g_gam2s.csv"), tr.gam2s, rtol=1e-4)
    assert np.allclose(np.loadtxt(outd + "g_R2trains.csv"), tr.R2trains, rtol=1e-4, atol=1e-6)
