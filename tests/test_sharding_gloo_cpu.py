"""world_size-2 CPU check (gloo) of the marker-sharding contract the N>1 path relies on: shards from
divide_work, X.v = allreduce(sum) of the shard products, X^T.u shard-local, EM / dot scalars summed once."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    N, Mt, seed = 501, 333, 9
    M, S = O.divide_work(Mt, world, rank)
    shard = O.Dataset(O.synth_bed(seed, S, M, N, miss_rate=0.02), N, Mt=Mt, S=S)
    full = O.Dataset(O.synth_bed(seed, 0, Mt, N, miss_rate=0.02), N)
    rng = np.random.default_rng(1)
    v, u = rng.normal(size=Mt), rng.normal(size=N)

    def allreduce(x):
        t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64).copy())
        dist.all_reduce(t)
        return t.numpy()

    # the reference's Ax divides by sqrt(N) after the allreduce; linear, so summing shard results is the same
    ax = allreduce(shard.Ax(v[S:S + M]))
    ok = np.allclose(ax, full.Ax(v), rtol=1e-12, atol=1e-13)
    ok &= np.allclose(shard.ATx(u), full.ATx(u)[S:S + M], rtol=1e-12, atol=1e-13)
    ok &= np.array_equal(shard.mave, full.mave[S:S + M]) and np.array_equal(shard.msig, full.msig[S:S + M])
    # EM prior update with the sufficient statistics summed over shards == single-shard update
    probs, vars_ = [0.9, 0.06, 0.04], [0.0, 0.1, 1.0]
    r1 = rng.normal(size=Mt)
    p_s, v_s = O.update_prior(r1[S:S + M], 3.0, probs, vars_, Mt, 3, 1e-6, allreduce=allreduce)
    p_f, v_f = O.update_prior(r1, 3.0, probs, vars_, Mt, 3, 1e-6)
    ok &= np.allclose(p_s, p_f, rtol=1e-12) and np.allclose(v_s, v_f, rtol=1e-12)
    # Onsager probe: every shard seeds mt19937 with seed + S (vamp.cpp:875)
    b = O.bernoulli_probe(1, S, M, Mt)
    ok &= abs(abs(b[0]) - 1 / np.sqrt(Mt)) < 1e-15 and len(b) == M
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_marker_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 400
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
