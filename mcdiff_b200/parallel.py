"""Multi-GPU plumbing: chains shard over ranks, nothing is exchanged while sampling.

One process per GPU (torchrun / torch.distributed, backend "nccl" on GPUs, "gloo" in the CPU
tests).  Every rank owns a contiguous block of the global chain index (the Philox stream of a
chain is keyed by its GLOBAL index, so the trajectories do not depend on the number of GPUs),
and ONE collective at the end collects the per-chain posterior summaries and acceptance
statistics on every rank.
"""
from __future__ import annotations

import numpy as np


def shard_range(nchains: int, rank: int, world: int):
    """Block partition [lo, hi) of the global chain index for `rank` (remainder to the first ranks)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(int(nchains), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pair_shard(nproblems: int, nchains: int, rank: int, world: int):
    """Lag scans (BASELINE config 4): the nproblems x nchains (problem, chain) pairs, numbered
    problem-major, are block-partitioned over the ranks.  Returns [(lo, hi)] per problem: the chains of that
    problem this rank owns (lo == hi: none)."""
    glo, ghi = shard_range(int(nproblems) * int(nchains), rank, world)
    out = []
    for k in range(int(nproblems)):
        lo, hi = max(glo, k * nchains), min(ghi, (k + 1) * nchains)
        out.append((lo - k * nchains, hi - k * nchains) if hi > lo else (0, 0))
    return out


def gather_ranges_to_root(local_rows: np.ndarray, ranges, device=None):
    """Rows [hi_r - lo_r][k] of every rank r (ranges = [(lo_r, hi_r)] for ALL ranks, disjoint, covering
    [0, total)) -> [total][k] on rank 0 in global order; None elsewhere.  One gather of equal padded blocks."""
    import torch
    import torch.distributed as dist
    rank, world, on = dist_info()
    local_rows = np.ascontiguousarray(local_rows, dtype=np.float64)
    if not on or world == 1:
        return local_rows
    k = local_rows.shape[1]
    width = max(1, max(hi - lo for lo, hi in ranges))
    total = sum(hi - lo for lo, hi in ranges)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else "cpu"
    buf = torch.zeros((width, k), dtype=torch.float64, device=device)
    if local_rows.shape[0] > 0:
        buf[: local_rows.shape[0]] = torch.from_numpy(local_rows).to(device)
    parts = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, parts, dst=0)
    if rank != 0:
        return None
    out = np.empty((total, k))
    for r, (lo, hi) in enumerate(ranges):
        if hi > lo:
            out[lo:hi] = parts[r][: hi - lo].cpu().numpy()
    return out


def dist_info():
    """(rank, world, initialised) from torch.distributed, (0, 1, False) when not initialised."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size(), True
    except Exception:
        pass
    return 0, 1, False


def all_gather_rows(local_rows: np.ndarray, nchains_total: int, device=None) -> np.ndarray:
    """Gather per-chain rows [n_local][k] of every rank into [nchains_total][k], chain order =
    global chain index.  Uses all_gather_into_tensor on equal-sized padded blocks."""
    import torch
    import torch.distributed as dist
    rank, world, on = dist_info()
    local_rows = np.ascontiguousarray(local_rows, dtype=np.float64)
    if not on or world == 1:
        return local_rows
    k = local_rows.shape[1]
    sizes = [shard_range(nchains_total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else "cpu"
    buf = torch.zeros((width, k), dtype=torch.float64, device=device)
    buf[: local_rows.shape[0]] = torch.from_numpy(local_rows).to(device)
    out = torch.empty((world * width, k), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(out, buf)
    out = out.cpu().numpy().reshape(world, width, k)
    return np.concatenate([out[r, : hi - lo] for r, (lo, hi) in enumerate(sizes)], axis=0)


def gather_rows_to_root(local_rows: np.ndarray, nchains_total: int, device=None, chunk_bytes: int = 256 << 20):
    """Gather per-chain rows [n_local][k] of every rank on rank 0 only ([nchains_total][k], global chain
    order; other ranks get None).  Used for the Logger rows (every stored sample of every chain), which can be
    large: the columns travel in chunks of at most `chunk_bytes` per rank so that the staging buffers stay small."""
    import torch
    import torch.distributed as dist
    rank, world, on = dist_info()
    local_rows = np.ascontiguousarray(local_rows, dtype=np.float64)
    if not on or world == 1:
        return local_rows
    k = local_rows.shape[1]
    sizes = [shard_range(nchains_total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else "cpu"
    out = np.empty((nchains_total, k)) if rank == 0 else None
    step = max(1, int(chunk_bytes // (8 * max(1, width))))
    for c0 in range(0, k, step):
        c1 = min(k, c0 + step)
        buf = torch.zeros((width, c1 - c0), dtype=torch.float64, device=device)
        if local_rows.shape[0] > 0:
            buf[: local_rows.shape[0]] = torch.from_numpy(np.ascontiguousarray(local_rows[:, c0:c1])).to(device)
        parts = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
        dist.gather(buf, parts, dst=0)
        if rank == 0:
            for r, (lo, hi) in enumerate(sizes):
                out[lo:hi, c0:c1] = parts[r][: hi - lo].cpu().numpy()
    return out


def reduce_max(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    rank, world, on = dist_info()
    if not on or world == 1:
        return float(value)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
