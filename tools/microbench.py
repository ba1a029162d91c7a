#!/usr/bin/env python3
"""Integer-pipe micro-benchmarks on cuda:0 (run on the GPU box): IMAD.WIDE peak, Fp mul / sqr throughput vs occupancy.
Writes gpurun_out/microbench.json."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from curdleproofs_b200 import Engine  # noqa: E402

eng = Engine(0)
out = {"imad": [], "fpmul": [], "fpsqr": []}
SM = 148
for threads, bps in [(128, 1), (256, 1), (256, 2), (256, 4), (256, 8)]:
    blocks = SM * bps
    iters = 2000
    ms = min(eng.bench_kernel(0, blocks, threads, iters) for _ in range(3))
    ops = blocks * threads * iters * 128
    out["imad"].append({"threads": threads, "blocks_per_sm": bps, "ms": ms, "imad_wide_per_s": ops / (ms * 1e-3)})
for which, key in [(1, "fpmul"), (2, "fpsqr")]:
    for threads, bps in [(32, 1), (128, 1), (128, 2), (128, 4), (256, 2), (256, 4), (256, 6), (256, 8)]:
        blocks = SM * bps
        iters = 2000
        ms = min(eng.bench_kernel(which, blocks, threads, iters) for _ in range(3))
        ops = blocks * threads * iters
        out[key].append({"threads": threads, "blocks_per_sm": bps, "warps_per_sm": threads * bps // 32, "ms": ms,
                         "fp_ops_per_s": ops / (ms * 1e-3), "ns_per_op_per_thread": ms * 1e6 / iters})
# FP64 pipe: DFMA alone, DFMA + IMAD.WIDE in one loop, DFMA + 64-bit integer adds in one loop (128 of each per iteration)
out["dfma"] = []
for which, key in [(3, "dfma"), (4, "dfma+imad_wide"), (5, "dfma+iadd64")]:
    for threads, bps in [(256, 1), (256, 2), (256, 4)]:
        blocks = SM * bps
        iters = 2000
        ms = min(eng.bench_kernel(which, blocks, threads, iters) for _ in range(3))
        ops = blocks * threads * iters * 128
        out["dfma"].append({"kernel": key, "threads": threads, "blocks_per_sm": bps, "ms": ms, "dfma_per_s": ops / (ms * 1e-3)})
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/microbench.json", "w"), indent=1)
for k, v in out.items():
    for r in v:
        print(k, r)
