#!/usr/bin/env python3
"""Quick throughput probe of the batched prover (run on the GPU box)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib
from curdleproofs_b200 import Engine, BatchProver

ell = int(sys.argv[1]) if len(sys.argv) > 1 else 252
batches = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [16, 64, 256]
lanes = int(sys.argv[3]) if len(sys.argv) > 3 else 0
host_threads = int(sys.argv[4]) if len(sys.argv) > 4 else 0
oracle = oracle_lib.Oracle()
crs = oracle.crs_points(ell)
inst = oracle.random_instance(ell, crs, seed=1, threads=8)
eng = Engine(0)
for B in batches:
    bp = BatchProver(eng, ell, crs, max_batch=B, lanes=lanes, host_threads=host_threads)
    insts = [inst] * B
    seeds = list(range(B))
    proofs = bp.prove_batch(insts, seeds)  # warm-up
    t = time.time(); proofs = bp.prove_batch(insts, seeds); dt = time.time() - t
    print(f"ell={ell} B={B}: {dt*1e3:.1f} ms -> {B/dt:.1f} proofs/s ; timing {bp.last_timing()} ; lanes {bp.lanes} launches {bp.launch_count}", flush=True)
    if B == batches[0]:
        want = oracle.prove(inst, rng_seed=0, threads=8)
        print("parity vs oracle:", proofs[0] == want, flush=True)
    bp.close()
