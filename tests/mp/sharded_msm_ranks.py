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
    # the accumulated check of a batch spread over the ranks (cdp_verify_batch_sharded): all valid -> the cross-rank sum accepts everything;
    # one invalid proof on the last rank -> every rank decides locally, with the oracle's verdicts
    from curdleproofs_b200 import BatchVerifier
    from test_gpu_sharded import _proved_batch
    import oracle_lib
    o = oracle_lib.Oracle()
    ell, batch = 12, 4
    crs, insts, proofs = _proved_batch(o, eng, ell, batch, 500 + 10 * rank)
    bv = BatchVerifier(eng, ell, crs, max_batch=batch, lanes=2)
    r1 = bv.verify_batch(insts, proofs, comm=comm)
    s1 = bv.global_stats()
    bad = list(proofs)
    if rank == world - 1:
        bad[1] = bad[1][:-1] + bytes([bad[1][-1] ^ 1])
    r2 = bv.verify_batch(insts, bad, comm=comm)
    s2 = bv.global_stats()
    want2 = [1, 0, 1, 1] if rank == world - 1 else [1] * batch
    good = r1 == [1] * batch and s1 == {"accepted": 1, "local": 0} and r2 == want2 and s2 == {"accepted": 1, "local": 1}
    print(f"rank {rank} sharded verify: {r1} {s1} {r2} {s2} {'ok' if good else 'MISMATCH'}", flush=True)
    ok = ok and good
    bv.close()
    dist.barrier()
    comm.close()
    eng.close()
    dist.destroy_process_group()
    if ok:
        print(f"RANK {rank} OK", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
