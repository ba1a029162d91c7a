"""Multi-GPU sharding of a single large MSM (`util::msm`, /root/reference/src/util.rs:19-22): contiguous base
ranges per rank, ONE all-gather of the 144-byte partial sums, local addition -- NCCL offers no elliptic-curve reduction op, so the
"allreduce" of partial G1 sums is an all-gather + add.  The collective lives inside the engine (`cdp_comm_*`, `cdp_msm_sharded_dev` in
include/cdp_msm.h: NCCL on the context's own stream, errors as CDP_ERR_NCCL); this module is the ctypes mirror plus the hand-over of
the NCCL unique id, for which any out-of-band channel does (here: `torch.distributed`, backend "nccl" on GPUs, "gloo" in the CPU tests)."""
from __future__ import annotations

import ctypes
from ctypes import c_size_t, c_void_p

import os

CDP_ERR_NCCL = 5
COMM_ID_BYTES = 128


def _prefer_bundled_nccl():
    """The engine loads NCCL at its first communicator call (`libnccl.so.2`, or $CDP_NCCL_LIB).  Inside a Python process that also uses
    PyTorch the copy must be the one PyTorch was built against (the `nvidia-nccl` wheel): whichever `libnccl.so.2` is loaded first serves
    both, and an older system NCCL lacks symbols libtorch_cuda needs.  So point the engine at the wheel's library when there is one."""
    if os.environ.get("CDP_NCCL_LIB"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for base in (list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []):
            cand = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["CDP_NCCL_LIB"] = cand
                return
    except Exception:
        pass


_prefer_bundled_nccl()


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, as-even-as-possible base range [lo, hi) of `rank` (the same split as cdp_shard_range)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def exchange_unique_id(make_id, rank: int, world: int) -> bytes:
    """Rank 0 creates the communicator id (`make_id()` -> 128 bytes), every rank receives it through torch.distributed."""
    import torch
    import torch.distributed as dist
    buf = torch.zeros(COMM_ID_BYTES, dtype=torch.uint8)
    if rank == 0:
        buf = torch.frombuffer(bytearray(make_id()), dtype=torch.uint8).clone()
    if world > 1:
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        buf = buf.to(dev)
        dist.broadcast(buf, src=0)
    return bytes(buf.cpu().numpy().tobytes())


class Comm:
    """`cdp_comm`: one rank of an n-rank NCCL communicator bound to an Engine (one process per GPU)."""

    def __init__(self, engine, unique_id: bytes, world: int, rank: int):
        from .engine import CdpError
        self.engine = engine
        self._lib = engine.lib
        h = c_void_p()
        idb = (ctypes.c_uint8 * COMM_ID_BYTES).from_buffer_copy(unique_id)
        rc = self._lib.cdp_comm_create(ctypes.byref(h), engine.handle, idb, world, rank)
        if rc != 0:
            raise CdpError(f"cdp_comm_create failed (code {rc}): {self._lib.cdp_last_error(engine.handle).decode()}")
        self._h = h
        self.world, self.rank = world, rank

    @staticmethod
    def unique_id(engine) -> bytes:
        from .engine import CdpError
        out = (ctypes.c_uint8 * COMM_ID_BYTES)()
        rc = engine.lib.cdp_comm_unique_id(out)
        if rc != 0:
            raise CdpError(f"cdp_comm_unique_id failed (code {rc}): NCCL not available")
        return bytes(out)

    @classmethod
    def from_torch_distributed(cls, engine, rank: int, world: int) -> "Comm":
        return cls(engine, exchange_unique_id(lambda: cls.unique_id(engine), rank, world), world, rank)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.cdp_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        from .engine import CdpError
        if rc != 0:
            raise CdpError(f"{what} failed (code {rc}): {self._lib.cdp_comm_last_error(self._h).decode()} / {self._lib.cdp_last_error(self.engine.handle).decode()}")

    def msm_sharded_dev(self, d_pts_shard: int, d_scalars_shard: int, n_local: int, d_out_jac: int):
        """Every rank: MSM over its shard (device pointers) -> all-gather -> sum, all on the Engine's stream; d_out_jac (144 B, device)
        holds the full sum on every rank after the stream is synchronised."""
        self._check(self._lib.cdp_msm_sharded_dev(self._h, c_void_p(d_pts_shard), c_void_p(d_scalars_shard), c_size_t(n_local), c_void_p(d_out_jac)),
                    "cdp_msm_sharded_dev")

    def allreduce_jacobian_dev(self, d_partial: int, d_out: int):
        self._check(self._lib.cdp_allreduce_jacobian_dev(self._h, c_void_p(d_partial), c_void_p(d_out)), "cdp_allreduce_jacobian_dev")

    def msm_sharded(self, pts_shard: bytes, scalars_shard: bytes) -> bytes:
        """Host-buffer convenience: this rank's shard in, the full sum (144-byte Jacobian point, identical on every rank) out."""
        eng, lib = self.engine, self._lib
        n_local = len(scalars_shard) // 32
        d_out = lib.cdp_dev_alloc(eng.handle, 144)
        d_p = d_s = None
        try:
            if n_local:
                d_p, d_s = lib.cdp_dev_alloc(eng.handle, len(pts_shard)), lib.cdp_dev_alloc(eng.handle, len(scalars_shard))
                hp = (ctypes.c_uint8 * len(pts_shard)).from_buffer_copy(pts_shard)
                hs = (ctypes.c_uint8 * len(scalars_shard)).from_buffer_copy(scalars_shard)
                lib.cdp_h2d(eng.handle, d_p, hp, len(pts_shard))
                lib.cdp_h2d(eng.handle, d_s, hs, len(scalars_shard))
            self.msm_sharded_dev(d_p or 0, d_s or 0, n_local, d_out)
            out = (ctypes.c_uint8 * 144)()
            lib.cdp_d2h(eng.handle, out, d_out, 144)
            eng.sync()
            return bytes(out)
        finally:
            for d in (d_out, d_p, d_s):
                if d:
                    lib.cdp_dev_free(eng.handle, d)


def sum_partials_cpu(partials: list, add_jacobian) -> bytes:
    """The combine step on host values (used by the gloo CPU tests with the oracle's group law as `add_jacobian`)."""
    acc = partials[0]
    for p in partials[1:]:
        acc = add_jacobian(acc, p)
    return acc
