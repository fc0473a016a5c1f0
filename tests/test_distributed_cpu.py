"""world_size-2 gloo test of the multi-GPU host logic (no GPU): strided mode ownership, the disjoint-support sum used as an
all-gather of the sources, contiguous fine-k slices and the all-reduce of partial C_l sums (north-star item 4)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import symboltz.jl_b200 as sb
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)  # same data on every rank
    nk, nc, nt, nl = 203, 61, 30, 7
    ks = np.linspace(1e-2, 2e3, nk)
    S_all = rng.standard_normal((nc, 2, nt))
    mine = np.arange(rank, nc, world)
    full = torch.zeros((nc, 2, nt), dtype=torch.float64)
    full[torch.from_numpy(mine)] = torch.from_numpy(S_all[mine])
    dist.all_reduce(full)
    assert np.array_equal(full.numpy(), S_all)  # disjoint supports: sum == gather, bit-exact
    theta = rng.standard_normal((2, nl, nk))
    P0 = 1.0 / ks**3
    w = sb.natural_spline_weights(np.concatenate([[0.0], ks]))[1:]
    ck = w * (2 / np.pi) * ks**2 * P0
    lo, hi = (nk * rank) // world, (nk * (rank + 1)) // world
    mask = np.zeros(nk)
    mask[lo:hi] = 1
    part = torch.from_numpy(np.einsum("k,lk,lk->l", ck * mask, theta[0], theta[1]))
    dist.all_reduce(part)
    ref = np.einsum("k,lk,lk->l", ck, theta[0], theta[1])
    assert np.allclose(part.numpy(), ref, rtol=1e-13)
    # cosmology sharding of a parameter sweep (config 4): strided ownership, exact gather of the per-rank P(k) rows, NaN rows kept
    n = 11
    mine = sb.shard_rows(n, rank, world)
    assert np.array_equal(mine, np.arange(rank, n, world))
    P_all = rng.standard_normal((n, 5))
    P_all[3] = np.nan  # a cosmology whose background failed
    got = sb.gather_rows(P_all[mine], mine, n)
    assert np.array_equal(got, P_all, equal_nan=True)
    cnt = sb.gather_rows(np.array([[rank + 1.0, 10.0 * (rank + 1)]]), [rank], world).sum(axis=0)
    assert cnt.tolist() == [3.0, 30.0]
    if rank == 0:
        out.put("ok")
    dist.destroy_process_group()


def test_sharding_and_allreduce_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) == "ok"
