#!/usr/bin/env python3
"""Per-kernel device time of one batched-verifier step (lanes=1: everything runs on the engine's own context, nothing overlaps)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib
from curdleproofs_b200 import Engine, BatchProver, BatchVerifier

ell = int(sys.argv[1]) if len(sys.argv) > 1 else 252
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
lanes = int(sys.argv[3]) if len(sys.argv) > 3 else 1
oracle = oracle_lib.Oracle()
crs = oracle.crs_points(ell)
inst = oracle.random_instance(ell, crs, seed=1, threads=8)
eng = Engine(0)
bp = BatchProver(eng, ell, crs, max_batch=64)
proofs = bp.prove_batch([inst] * 64, list(range(64)))
bp.close()
proofs = (proofs * ((B + 63) // 64))[:B]
bv = BatchVerifier(eng, ell, crs, max_batch=B, lanes=lanes)
insts = [inst] * B
res = bv.verify_batch(insts, proofs)
assert all(r == 1 for r in res), res[:8]
bv.verify_batch(insts, proofs)
eng.profile_reset(); eng.profile_enable(True)
t = time.time(); bv.verify_batch(insts, proofs); dt = time.time() - t
prof = eng.profile_read()
eng.profile_enable(False)
print(f"ell={ell} B={B} lanes={lanes}: {dt*1e3:.1f} ms (incl. python marshalling) -> {B/dt:.1f} verifies/s ; C call {bv.last_timing()}", flush=True)
tot = sum(v["ms"] for v in prof.values())
for k, v in prof.items():
    print(f"  {k:12s} {v['ms']:9.2f} ms  {v['launches']:5d} launches  {v['units']:12d} units")
print(f"  sum {tot:.2f} ms (lane 0 only)")
bad = bytearray(proofs[3]); bad[100] ^= 1
res = bv.verify_batch(insts[:8], proofs[:3] + [bytes(bad)] + proofs[4:8])
print("verdicts with proof 3 corrupted:", res)
