"""N > 1 host path on CPU: world_size-2 gloo.  Chains shard by global index, the only
collective is the final all_gather of per-chain summary rows."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nchains, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mcdiff_b200.parallel import all_gather_rows, dist_info, reduce_max, shard_range
    assert dist_info() == (rank, world, True)
    lo, hi = shard_range(nchains, rank, world)
    # a rank's rows carry the GLOBAL chain id, as the device Philox streams do (chain_offset = lo)
    local = np.stack([np.arange(lo, hi, dtype=np.float64), np.full(hi - lo, float(rank))], axis=1)
    allrows = all_gather_rows(local, nchains, device="cpu")
    tmax = reduce_max(1.0 + rank, device="cpu")
    q.put((rank, allrows, tmax))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("nchains", [8, 7])
def test_gloo_world2_gather(nchains):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nchains, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, rows, tmax in got:
        assert rows.shape == (nchains, 2)
        np.testing.assert_array_equal(rows[:, 0], np.arange(nchains))      # global chain order, no gaps
        first = (nchains + 1) // 2
        np.testing.assert_array_equal(rows[:, 1], [0.0] * first + [1.0] * (nchains - first))
        assert tmax == 2.0                                                 # max over ranks


# ---- pooled posterior: the Logger rows of every rank end up on rank 0 in global chain order -------------------
class _FakeMC:
    """What Logger needs from an MCState (no device)."""

    def __init__(self, model, nmc):
        self.model, self.nmc, self.do_radial = model, nmc, False


def _fake_model(ncosF, ncosD, n=12):
    from types import SimpleNamespace
    from oracle import mcdiff_oracle as O
    return SimpleNamespace(ncosF=ncosF, ncosD=ncosD, dim_v=n, dim_w=n, pbc=True, wunit=-1.3, vunit=1.0,
                           edges=np.linspace(-3.0, 3.0, n + 1), v_basis=O.cos_basis_center(n, max(ncosF, 1)),
                           w_basis=O.cos_basis_border(n, n, max(ncosD, 1)))


def _fill_logger(logger, model, chain_lo):
    """Deterministic rows keyed by the GLOBAL chain index (what the device Philox streams guarantee)."""
    nv = model.ncosF if model.ncosF > 0 else model.dim_v
    nw = model.ncosD if model.ncosD > 0 else model.dim_w
    for c in range(logger.nchains):
        rng = np.random.default_rng(1000 + chain_lo + c)
        logger.fill_rows(c, 0, rng.standard_normal((logger.nf_chain, 6 + nv + nw)), model)


def _pool_worker(rank, world, port, nchains, ncosF, ncosD, outdir, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    from mcdiff_b200.log import Logger, write_average_from_pic
    from mcdiff_b200.mc import _pool_logger
    from mcdiff_b200.parallel import shard_range
    model = _fake_model(ncosF, ncosD)
    MC = _FakeMC(model, nmc=300)
    lo, hi = shard_range(nchains, rank, world)
    MC.nchains_total = nchains
    logger = Logger(MC, nchains=hi - lo)
    _fill_logger(logger, model, lo)
    pooled = _pool_logger(MC, logger, rank)
    if rank == 0:
        pooled.model = model
        pic = os.path.join(outdir, "w%d.pic" % world)
        pooled.dump(pic)
        write_average_from_pic(pic, pic + ".ave.dat")
        q.put((world, pooled.nchains, pooled.pack_rows()))
    else:
        assert pooled is None
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("ncosF,ncosD,nchains", [(4, 3, 5), (0, 0, 4), (4, 3, 1)])
def test_pooled_posterior_two_ranks_equal_one_rank(tmp_path, ncosF, ncosD, nchains):
    """`.pic.ave.dat` (pooled posterior mean / std, lib/log.py:90-106, lib/outreading.py:281-301) is byte-identical
    whether one rank holds all chains or two ranks hold a shard each -- also when a rank owns no chain."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    got = {}
    for world in (1, 2):
        port = _free_port()
        procs = [ctx.Process(target=_pool_worker, args=(r, world, port, nchains, ncosF, ncosD, str(tmp_path), q))
                 for r in range(world)]
        for p in procs:
            p.start()
        w, nch, rows = q.get(timeout=180)
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        got[w] = rows
        assert nch == nchains
    np.testing.assert_array_equal(got[1], got[2])
    a = open(os.path.join(str(tmp_path), "w1.pic.ave.dat")).read()
    b = open(os.path.join(str(tmp_path), "w2.pic.ave.dat")).read()
    assert a == b and "===== final F D =====" in a


def test_jitter_depends_on_global_chain_only():
    """Start offsets are keyed by (seed, global chain): shards of two ranks concatenate to the one-rank rows,
    global chain 0 keeps the exact start."""
    from mcdiff_b200.MCState import MCState
    one = MCState(True, nchains=7, seed=3, rank=0, world=1, jitter=0.2)
    parts = [MCState(True, nchains=7, seed=3, rank=r, world=2, jitter=0.2) for r in range(2)]
    full = one._jitter_rows(1, 5)
    np.testing.assert_array_equal(full, np.concatenate([p._jitter_rows(1, 5) for p in parts]))
    assert not full[0].any() and full[1:].all()
    assert not np.array_equal(full[1], full[2])
    assert not np.array_equal(one._jitter_rows(0, 5)[1], full[1])


# ---- lag scan (BASELINE config 4): (problem, chain) pairs block-partitioned over the ranks -------------------
def test_pair_shard_covers_every_pair_once():
    from mcdiff_b200.parallel import pair_shard
    for nprob, nch, world in ((16, 256, 8), (4, 5, 3), (3, 7, 2), (2, 3, 8), (16, 1, 5)):
        seen = np.zeros((nprob, nch), dtype=int)
        sizes = []
        for r in range(world):
            sh = pair_shard(nprob, nch, r, world)
            assert len(sh) == nprob
            sizes.append(sum(hi - lo for lo, hi in sh))
            for k, (lo, hi) in enumerate(sh):
                seen[k, lo:hi] += 1
        assert np.all(seen == 1)
        assert max(sizes) - min(sizes) <= 1                      # balanced to within one pair


def _scan_worker(rank, world, port, nprob, nch, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mcdiff_b200.parallel import gather_ranges_to_root, pair_shard
    shards = [pair_shard(nprob, nch, r, world) for r in range(world)]
    got = []
    for k in range(nprob):
        lo, hi = shards[rank][k]
        local = np.stack([np.full(hi - lo, float(k)), np.arange(lo, hi, dtype=np.float64),
                          np.full(hi - lo, float(rank))], axis=1).reshape(hi - lo, 3)
        got.append(gather_ranges_to_root(local, [shards[r][k] for r in range(world)], device="cpu"))
    q.put((rank, got))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("nprob,nch", [(4, 5), (3, 2)])
def test_gloo_world2_lag_scan_gather(nprob, nch):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_scan_worker, args=(r, world, port, nprob, nch, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(g is None for g in got[1])                         # rank 0 alone holds the gathered rows
    half = (nprob * nch + 1) // 2
    for k, rows in enumerate(got[0]):
        assert rows.shape == (nch, 3)
        np.testing.assert_array_equal(rows[:, 0], k)
        np.testing.assert_array_equal(rows[:, 1], np.arange(nch))          # global chain order within the problem
        np.testing.assert_array_equal(rows[:, 2], [(0.0 if k * nch + c < half else 1.0) for c in range(nch)])
