#!/usr/bin/env python3
"""Per-kernel device time of one batched-prover step with a given lane count (lanes=1: kernels run alone, nothing overlaps)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib
from curdleproofs_b200 import Engine, BatchProver

ell = int(sys.argv[1]) if len(sys.argv) > 1 else 252
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
lanes = int(sys.argv[3]) if len(sys.argv) > 3 else 1
host_threads = int(sys.argv[4]) if len(sys.argv) > 4 else 0
oracle = oracle_lib.Oracle()
crs = oracle.crs_points(ell)
inst = oracle.random_instance(ell, crs, seed=1, threads=8)
eng = Engine(0)
bp = BatchProver(eng, ell, crs, max_batch=B, lanes=lanes, host_threads=host_threads)
insts = [inst] * B
seeds = list(range(B))
bp.prove_batch(insts, seeds)
bp.prove_batch(insts, seeds)
bp.profile_reset(); bp.profile_enable(True)
t = time.time(); bp.prove_batch(insts, seeds); dt = time.time() - t
prof = bp.profile_read()
bp.profile_enable(False)
print(f"ell={ell} B={B} lanes={bp.lanes}: {dt*1e3:.1f} ms -> {B/dt:.1f} proofs/s ; timing(total,host,gpu_wait,copy) {bp.last_timing()}")
tot = sum(v["ms"] for v in prof.values())
for k, v in prof.items():
    print(f"  {k:12s} {v['ms']:9.2f} ms  {v['launches']:5d} launches  {v['units']:12d} units  {100*v['ms']/tot:5.1f}%")
print(f"  sum {tot:.2f} ms")
