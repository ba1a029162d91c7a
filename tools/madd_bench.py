#!/usr/bin/env python3
"""Mixed-addition chain micro-benchmark (k_bench_madd): by-value register ABI of the field products (which = 6) against the pointer /
local-memory ABI (which = 7), at several grid sizes.  Prints G additions/s."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from curdleproofs_b200 import Engine  # noqa: E402

eng = Engine(0)
iters = 300
for which, name in ((6, "by-value ABI"), (7, "pointer ABI"), (8, "XYZZ accum.")):
    for bps in (1, 2, 3, 4, 6, 8):
        blocks = 148 * bps
        ms = min(eng.bench_kernel(which, blocks, 128, iters) for _ in range(3))
        print(f"{name:14s} {bps} CTAs/SM requested: {ms:8.3f} ms  {blocks * 128 * iters / ms / 1e6:7.3f} G adds/s", flush=True)
