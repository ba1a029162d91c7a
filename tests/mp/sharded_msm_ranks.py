"""Rank program of tests/test_gpu_sharded.py::test_sharded_msm_two_ranks_one_process_per_gpu, launched by torch.distributed.run (one process
per GPU): every rank holds a contiguous base range, calls cdp_msm_sharded_dev through the engine's own NCCL communicator and compares the
result -- which must be the same on every rank -- with the oracle's MSM over ALL bases.  Prints `RANK r OK` per rank."""
import os
import sys

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
ROOT = os.path.dirname(TESTS)
for p in (ROOT, TESTS, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)


def main():
    from curdleproofs_b200 import Engine
    from curdleproofs_b200.sharded import Comm, shard_range
    from test_gpu_sharded import _inputs
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")  # only carries the 128-byte communicator id
    eng = Engine(local)
    comm = Comm.from_torch_distributed(eng, rank, world)
    ok = True
    for n in (1, 37, 9001, 70000):  # n = 1: every rank but the first holds an empty shard; 70000: the sort-based large path on each shard
        o, pts, sc = _inputs(n)
        lo, hi = shard_range(n, rank, world)
        got = comm.msm_sharded(pts[96 * lo:96 * hi], sc[32 * lo:32 * hi])
        good = o.compress_jac(got) == o.compress_jac(o.msm(pts, sc, threads=4))
        print(f"rank {rank} n={n} shard=[{lo},{hi}) {'ok' if good else 'MISMATCH'}", flush=True)
        ok = ok and good
    dist.barrier()
    comm.close()
    eng.close()
    dist.destroy_process_group()
    if ok:
        print(f"RANK {rank} OK", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
