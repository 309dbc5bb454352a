"""profiles/run_config5.py -- BASELINE.json config 5: imputed-scale synthetic N=400,000 x Mt=8,400,000 with 1 % missing
genotypes, marker-sharded (1,050,000 markers = 105 GB packed per GPU at 8 GPUs): standalone X.v / X^T.u sweeps and one
LMMSE CG solve (reference data::Ax / ATx, vamp::precondCG_solver), with the size-independent checks that tie the
result to the oracle-checked small cases (adjointness over all shards, bit-reproducibility, CG residual).

    python profiles/run_config5.py                      # one GPU = one shard of the 8-GPU layout
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 profiles/run_config5.py

Rank 0 prints one JSON line.  Times are CUDA events on the library's stream (max over ranks)."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gvamp_b200 import capi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=400_000)
ap.add_argument("--markers-per-gpu", type=int, default=1_050_000)
ap.add_argument("--miss", type=float, default=0.01)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--cg-iters", type=int, default=5)
a = ap.parse_args()

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
nccl_id = None
dist = None
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    idt = torch.zeros(capi.NCCL_ID_BYTES, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    nccl_id = bytes(idt.cpu().numpy().tobytes())


def allmax(x):
    if world == 1:
        return x
    t = torch.tensor([x], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def allsum(x):
    if world == 1:
        return x
    t = torch.tensor([x], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t[0])


N, Mt = a.N, a.markers_per_gpu * world
M, S = capi.divide_work(Mt, world, rank)
bed_local = M * ((N + 3) // 4)
ctx = capi.Context(local, rank, world, nccl_id)
ctx.timer_start(2)
ctx.synth(20261017, N, Mt, S, M, a.miss)
ctx.timer_stop(2)
t_synth = ctx.timer_ms(2)
ctx.timer_start(2)
ctx.compute_stats(1.0)
ctx.timer_stop(2)
t_stats = ctx.timer_ms(2)
rng = np.random.default_rng(7)            # same u on every rank (N-vectors are replicated); v differs per shard
u_h = rng.normal(size=N)
v_h = np.random.default_rng(1000 + rank).normal(size=M)
v, u = ctx.vecM(v_h), ctx.vecN(u_h)
ov, ou = ctx.vecN(), ctx.vecM()
ctx.timer_start(2)
ctx.dATx(u, ou)                           # builds the missing-genotype list on its first call
ctx.timer_stop(2)
t_first_atx = ctx.timer_ms(2)
ax_ms, atx_ms = [], []
for r in range(a.reps):
    ctx.profile(True)
    ctx.dAx(v, ov)
    ctx.dATx(u, ou)
    p = ctx.profile_read()
    ax_ms.append(p["ax_ms"])
    atx_ms.append(p["atx_ms"])
ax1, atx1 = ov.download(), ou.download()
ctx.dAx(v, ov)
ctx.dATx(u, ou)
repro = bool(np.array_equal(ax1, ov.download()) and np.array_equal(atx1, ou.download()))
# <X v, u> over all shards (X.v is all-reduced inside the library) against sum over shards of <v, X^T u>
lhs = float(np.dot(ax1[:N], u_h))
rhs = allsum(float(np.dot(v_h, atx1)))
adj = abs(lhs - rhs) / (np.linalg.norm(ax1) * np.linalg.norm(u_h))
# one LMMSE solve: (tau X^T X + gam2 I) mu = rhs, Jacobi-preconditioned CG with a fixed iteration budget
rhs_v, mu = ctx.vecM(v_h), ctx.vecM()
ctx.timer_start(3)
its, rel = ctx.cg_solve(rhs_v, mu, 2.0, 0.7, a.cg_iters, 1)
ctx.timer_stop(3)
t_cg = ctx.timer_ms(3)
ax_best, atx_best = allmax(min(ax_ms)), allmax(min(atx_ms))
line = {
    "workload": f"config 5 shard: N={N} x {M} markers/GPU x {world} GPU(s), {100 * a.miss:g}% missing, {bed_local / 1e9:.1f} GB packed per GPU",
    "n_gpus": world, "bed_bytes_per_gpu": bed_local, "synth_ms": allmax(t_synth), "stats_ms": allmax(t_stats),
    "first_atx_ms_incl_missing_list_build": allmax(t_first_atx), "missing_list_entries": ctx.missing_list_entries(),
    "missing_list_bytes": 2 * ctx.missing_list_entries(),
    "ax_ms": ax_best, "atx_ms": atx_best, "ax_GBps_per_gpu": bed_local / ax_best / 1e6, "atx_GBps_per_gpu": bed_local / atx_best / 1e6,
    "ax_ms_all": ax_ms, "atx_ms_all": atx_ms,
    "cg": {"iterations": its, "ms": allmax(t_cg), "ms_per_iteration": allmax(t_cg) / max(its, 1), "log": np.asarray(rel).tolist()},
    "adjointness_rel": adj, "bit_reproducible": repro,
}
ok = adj < 1e-6 and repro
if rank == 0:
    print(json.dumps(line))
ctx.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
sys.exit(0 if ok else 1)
