"""Host-side logic of the row-sharded path, on CPU: shard planning, vector slicing, merging of the
bias statistics, and the torch.distributed plumbing over gloo with world_size 2."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import smo_oracle as O
from pysvm_b200 import sharded as SH


@pytest.mark.parametrize("n,world", [(200000, 8), (200000, 2), (569, 2), (1000, 4), (33, 2)])
def test_plan_covers_every_row_once(n, world):
    seen = np.zeros(n, dtype=int)
    for r in range(world):
        try:
            s = SH.plan(n, world, r)
        except ValueError:
            continue
        assert s.rows_per_rank % 32 == 0 and s.row_offset == r * s.rows_per_rank
        seen[s.rows()] += 1
    assert (seen == 1).all()


def test_local_vector_mirror_layout():
    n = 100
    v = np.arange(2 * n, dtype=float)
    s = SH.plan(n, 2, 1)
    loc = SH.local_vector(v, s, mirror=True)
    assert loc.tolist() == list(range(64, 100)) + list(range(164, 200))
    assert SH.local_vector(v[:n], s, mirror=False).tolist() == list(range(64, 100))


def numpy_partials(alpha, g, y, C):
    """What b2svm_bias_partials returns, restated in NumPy (for the merge test)."""
    out = np.zeros(18)
    free = (alpha > 0) & (alpha < C)
    out[0], out[1] = g[free].sum(), free.sum()
    lb = ((alpha == 0) & (y > 0)) | ((alpha == C) & (y < 0))
    ub = ((alpha == 0) & (y < 0)) | ((alpha == C) & (y > 0))
    out[2], out[3] = (g[lb].min() if lb.any() else np.inf), lb.sum()
    out[4], out[5] = (g[ub].max() if ub.any() else -np.inf), ub.sum()
    grad = -y * g
    for c, s in enumerate((1, -1)):
        cls = y == s
        fr = (alpha > 0) & (alpha < 1) & cls
        hi, lo = (alpha == 1) & cls, (alpha == 0) & cls
        o = 6 + 6 * c
        out[o], out[o + 1] = grad[fr].sum(), fr.sum()
        out[o + 2], out[o + 3] = (grad[hi].max() if hi.any() else -np.inf), hi.sum()
        out[o + 4], out[o + 5] = (grad[lo].min() if lo.any() else np.inf), lo.sum()
    return out


def test_merged_partials_reproduce_oracle_bias(breast_cancer):
    X, y = breast_cancer
    fit = O.fit_csvc(X, y, max_iter=10 ** 6)
    st = O.DualState(y=fit.y, p=-np.ones(len(y)), C=1.0, tol=1e-5, alpha=fit.alpha, g=fit.g)
    for world in (1, 2, 4):
        parts = []
        for r in range(world):
            sl = SH.plan(len(y), world, r).rows()
            parts.append(numpy_partials(fit.alpha[sl], fit.g[sl], fit.y[sl], 1.0))
        rho = SH.rho_from_partials(SH.combine_bias_partials(np.array(parts)))
        assert abs(rho - O.bias(st)) < 1e-12
    nu = O.fit_nusvc(X, y, max_iter=300)
    stn = O.DualState(y=nu.y, p=np.zeros(len(y)), C=1.0, tol=1e-5, alpha=nu.alpha, g=nu.g)
    parts = [numpy_partials(nu.alpha[SH.plan(len(y), 2, r).rows()], nu.g[SH.plan(len(y), 2, r).rows()],
                            nu.y[SH.plan(len(y), 2, r).rows()], 1.0) for r in range(2)]
    rho, b = SH.rho_b_from_partials(SH.combine_bias_partials(np.array(parts)))
    want = O.bias_nu(stn)
    assert abs(rho - want[0]) < 1e-12 and abs(b - want[1]) < 1e-12


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # IPC-handle exchange: every rank must end up with every rank's 64 bytes, in rank order
        mine = bytes([rank + 1]) * 64
        allh = SH.all_gather_bytes(mine)
        assert allh == b"".join(bytes([r + 1]) * 64 for r in range(world))
        # reassembly of a sharded per-variable vector (mirror and plain)
        n = 1000
        full = np.arange(2 * n, dtype=float)
        s = SH.plan(n, world, rank)
        got = SH.gather_variables(SH.local_vector(full, s, True), s, True)
        assert np.array_equal(got, full)
        got = SH.gather_variables(SH.local_vector(full[:n], s, False), s, False)
        assert np.array_equal(got, full[:n])
        # bias partials merged over the group
        parts = SH.all_gather_array(np.full(18, float(rank)))
        assert parts.shape == (world, 18) and parts[:, 0].tolist() == list(range(world))
        # results of independent sub-problems (one-vs-one fits) solved on different ranks reach every rank
        sizes = [5, 3, 4, 6, 2]
        owner = SH.schedule_subproblems(sizes, world)
        mine = {k: (np.full(11, 100.0 * k), np.arange(sizes[k]) + 10.0 * k, -np.arange(sizes[k]) - 10.0 * k)
                for k in range(len(sizes)) if owner[k] == rank}
        allres = SH.exchange_subproblem_results(mine, sizes, owner, 11)
        assert len(allres) == len(sizes)
        for k, (st, a, g) in enumerate(allres):
            assert np.array_equal(st, np.full(11, 100.0 * k))
            assert np.array_equal(a, np.arange(sizes[k]) + 10.0 * k) and np.array_equal(g, -a)
        SH.enable()
        assert SH.active()
        SH.disable()
        assert not SH.active()
        q.put((rank, "ok"))
    except Exception as e:      # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_gloo_world_size_2_plumbing():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, "ok"), (1, "ok")], res


def test_subproblem_schedule_is_balanced_and_deterministic():
    """One-vs-one sub-problems are dealt to the GPUs by size (largest first, least loaded rank): every rank derives the
    same owners, every sub-problem has exactly one owner, and the load is within one sub-problem of even."""
    from pysvm_b200.sharded import schedule_subproblems
    rng = np.random.default_rng(0)
    sizes = [int(s) for s in rng.integers(300, 400, size=45)]          # digits: 45 pairs of ~360 rows
    for world in (1, 2, 4, 8):
        owner = schedule_subproblems(sizes, world)
        assert owner == schedule_subproblems(list(sizes), world)
        assert len(owner) == 45 and set(owner) <= set(range(world))
        load = [sum(s for s, o in zip(sizes, owner) if o == r) for r in range(world)]
        assert max(load) - min(load) <= max(sizes)
    assert schedule_subproblems([5, 5, 5], 2) == [0, 1, 0]
    assert schedule_subproblems([], 4) == []


def test_shardable_is_a_pure_predicate():
    """n=200 on 8 GPUs: rpr = 32, rank 7 would be empty -> EVERY rank refuses (none blocks in a collective alone)."""
    assert not SH.shardable(200, 8)
    for r in range(8):
        with pytest.raises(ValueError):
            SH.plan(200, 8, r)
    assert SH.shardable(200, 4) and SH.shardable(33, 2) and SH.shardable(200000, 8)
