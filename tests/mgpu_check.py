"""tests/mgpu_check.py -- marker-sharded parity check, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py

Every rank owns the shard divide_work() gives it, X.v is combined with the NCCL allreduce inside the
library, X^T.u / CG / dot products are checked against the single-shard oracle on the full matrix.
Used by tests/test_gpu_multi.py (skipped on single-GPU boxes)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gvamp_b200 import capi  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    idt = torch.zeros(capi.NCCL_ID_BYTES, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    ctx = capi.Context(local, rank, world, bytes(idt.cpu().numpy().tobytes()))

    N, Mt, seed = 3000, 2501, 77
    M, S = capi.divide_work(Mt, world, rank)
    ctx.synth(seed, N, Mt, S, M, 0.01)
    ctx.compute_stats(1.0)
    bed = O.synth_bed(seed, 0, Mt, N, miss_rate=0.01)
    ds = O.Dataset(bed, N)
    assert np.array_equal(ctx.decode(0, M), bed[S:S + M])
    rng = np.random.default_rng(5)
    v, u = rng.normal(size=Mt), rng.normal(size=N)
    rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
    ax = ctx.Ax(v[S:S + M])                      # allreduced over shards
    assert rel(ax, ds.Ax(v)) < 1e-6, rel(ax, ds.Ax(v))
    atx = ctx.ATx(u)                             # shard-local
    assert rel(atx, ds.ATx(u)[S:S + M]) < 1e-6
    rhs, mu = ctx.vecM(v[S:S + M]), ctx.vecM()
    d = ctx.dots([rhs], [None])[0]
    assert abs(d - v @ v) < 1e-9 * (v @ v)
    its, _ = ctx.cg_solve(rhs, mu, 2.0, 0.7, 25, 1)
    mu_ref, its_ref = O.precond_cg(ds, v, np.zeros(Mt), 2.0, 0.7, 25, 1)
    assert its == its_ref and rel(mu.download(), mu_ref[S:S + M]) < 1e-5, (its, its_ref)
    x1 = ctx.vecM()
    probs, vars_ = [0.9, 0.06, 0.04], [0.0, 0.1, 1.0]
    sums = ctx.denoise(rhs, 3.0, probs, vars_, x1)
    assert abs(sums[0] - O.g1d(v, 3.0, probs, vars_).sum()) < 1e-8 * Mt
    dist.barrier()
    if rank == 0:
        print(f"mgpu_check ok: world={world}, CG iterations {its}")
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
