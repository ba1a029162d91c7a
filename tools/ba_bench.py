#!/usr/bin/env python3
"""Batched affine addition micro-benchmark (k_ba_bench, cdp_bench_kernel 9 + INL + 8 * OCC): T threads x K additions sharing one inversion,
operands streamed from an array; INL = which field products are expanded in place (batch_affine.cuh), OCC = 0 / 1 / 2 for 3 / 4 / 5 CTAs of 128
threads per SM.  Prints G additions/s; compare with tools/madd_bench.py (XYZZ chain: 2.95 G/s)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from curdleproofs_b200 import Engine  # noqa: E402

eng = Engine(0)
for occ, ctas in ((0, 3), (1, 4), (2, 5)):
    for inl in (0, 3):
        for waves in (2, 4):
            for K in (16, 32, 64, 128):
                blocks = 148 * ctas * waves
                ms = min(eng.bench_kernel(9 + inl + 8 * occ, blocks, 128, K) for _ in range(3))
                print(f"CTAs/SM {ctas} INL {inl} waves {waves} K {K:4d}: {ms:8.3f} ms  {blocks * 128 * K / ms / 1e6:7.3f} G adds/s", flush=True)
