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
