"""Multi-GPU sharding of a single large MSM: contiguous base ranges per rank, one all-gather of the
144-byte partial sums, local addition.  NCCL offers no elliptic-curve reduction op, so this is an all-gather + add, not an
all-reduce.  `torch.distributed` is plumbing only (backend "nccl" on GPUs; the CPU tests use "gloo")."""
from __future__ import annotations


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, as-even-as-possible base range [lo, hi) of `rank`."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allgather_partials(partial, world: int):
    """partial: uint8 tensor of 144 bytes (one Jacobian point) on this rank -> uint8 tensor [world, 144] (same on every rank)."""
    import torch
    import torch.distributed as dist
    out = torch.empty((world, 144), dtype=torch.uint8, device=partial.device)
    if world == 1:
        out[0] = partial
        return out
    dist.all_gather_into_tensor(out.view(-1), partial.contiguous().view(-1))
    return out


def sharded_msm_dev(engine, d_pts_shard: int, d_scalars_shard: int, n_local: int, world: int):
    """Each rank: MSM over its shard (device pointers), all-gather, local sum.  Returns a uint8 CUDA tensor of 144 bytes that is
    identical on every rank.  The Engine must have been created on torch's current stream so that the collective is ordered
    after the MSM kernels."""
    import torch
    partial = torch.zeros(144, dtype=torch.uint8, device="cuda")  # Z = 0: infinity for an empty shard
    lib, h = engine.lib, engine.handle
    if n_local > 0:
        rc = lib.cdp_msm_dev(h, d_pts_shard, d_scalars_shard, n_local, partial.data_ptr())
        if rc != 0:
            raise RuntimeError(lib.cdp_last_error(h).decode())
    allp = allgather_partials(partial, world)
    out = torch.empty(144, dtype=torch.uint8, device="cuda")
    rc = lib.cdp_sum_jacobian_dev(h, allp.data_ptr(), world, out.data_ptr())
    if rc != 0:
        raise RuntimeError(lib.cdp_last_error(h).decode())
    return out
