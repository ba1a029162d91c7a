"""CPU, world_size = 2, gloo: the N > 1 host logic -- base-range sharding of one MSM, the hand-over of the communicator id from rank 0
(curdleproofs_b200.sharded.exchange_unique_id), all-gather of the 144-byte partial sums, local addition -- with the oracle standing in
for the GPU kernel and a gloo all-gather for the engine's NCCL one (no GPU in this container; the real collective is cdp_msm_sharded_dev,
tests/test_gpu_sharded.py)."""
import os
import random
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    import oracle_lib
    import py_ref as pr
    from curdleproofs_b200.sharded import COMM_ID_BYTES, exchange_unique_id, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # the communicator id travels from rank 0 to everyone (a 128-byte opaque blob, like ncclGetUniqueId's)
    uid = exchange_unique_id(lambda: bytes((7 * i + 3) % 256 for i in range(COMM_ID_BYTES)), rank, world)
    assert uid == bytes((7 * i + 3) % 256 for i in range(COMM_ID_BYTES))
    o = oracle_lib.Oracle()
    rnd = random.Random(42)  # identical inputs on every rank
    sc0 = b"".join(pr.fr_to_bytes(rnd.randrange(pr.R_ORDER)) for _ in range(n))
    pts = o.scalar_mul_batch(o.generator() * n, sc0)
    sc = b"".join(pr.fr_to_bytes(rnd.randrange(pr.R_ORDER)) for _ in range(n))
    lo, hi = shard_range(n, rank, world)
    partial = o.msm(pts[96 * lo:96 * hi], sc[32 * lo:32 * hi])          # this rank's base range
    t = torch.frombuffer(bytearray(partial), dtype=torch.uint8)
    allp = torch.empty((world, 144), dtype=torch.uint8)
    dist.all_gather_into_tensor(allp.view(-1), t)                        # [world, 144], same on every rank
    acc = pr.INF
    for r in range(world):
        acc = pr.add(acc, pr.jacobian_from_bytes(bytes(allp[r].numpy().tobytes())))
    full = pr.jacobian_from_bytes(o.msm(pts, sc))
    q.put((rank, lo, hi, acc == full, pr.compress(acc).hex()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [37, 64])
def test_sharded_msm_allgather_gloo(n):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == n      # contiguous cover
    assert all(r[3] for r in res)                                            # sum of partials == full MSM, on every rank
    assert res[0][4] == res[1][4]                                            # every rank ends with the same point


def test_shard_range_properties():
    from curdleproofs_b200.sharded import shard_range
    for n in (0, 1, 7, 8, 1000, (1 << 22) + 3):
        for world in (1, 2, 4, 8):
            rs = [shard_range(n, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in rs]
            assert max(sizes) - min(sizes) <= 1


def test_prover_batch_split_matches_lane_layout():
    """The prover splits a batch over lanes exactly like shard_range splits bases over ranks (contiguous, sizes differ by <= 1):
    bench.py relies on this when it sizes per-rank batches."""
    from curdleproofs_b200.sharded import shard_range
    B, L = 1023, 8
    off = [0]
    for i in range(L):
        off.append(off[-1] + B // L + (1 if i < B % L else 0))
    assert [(off[i], off[i + 1]) for i in range(L)] == [shard_range(B, i, L) for i in range(L)]
