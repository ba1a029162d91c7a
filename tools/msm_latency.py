#!/usr/bin/env python3
"""Latency of ONE `util::msm` call (/root/reference/src/util.rs:19-22) by size: device-resident (cdp_msm_dev) and through the host-buffer
drop-in (cdp_msm, copies included), median of several calls on an otherwise idle GPU.  Env knobs under test: CDP_MSM_SINGLE_CHUNK,
CDP_BIG_MIN_LOG2, CDP_BIG_C.  Usage: msm_latency.py [lo_log2 [hi_log2]]"""
import ctypes
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from bench import GX, GY, mont
from curdleproofs_b200 import Engine

lo = int(sys.argv[1]) if len(sys.argv) > 1 else 8
hi = int(sys.argv[2]) if len(sys.argv) > 2 else 20
stream = torch.cuda.Stream()
eng = Engine(0, stream=stream.cuda_stream)
lib, h = eng.lib, eng.handle
N = 1 << hi
rng = np.random.Generator(np.random.Philox(key=99))
t = rng.integers(0, 2 ** 64, size=(N, 4), dtype=np.uint64); t[:, 3] &= np.uint64((1 << 62) - 1)
s = rng.integers(0, 2 ** 64, size=(N, 4), dtype=np.uint64); s[:, 3] &= np.uint64((1 << 62) - 1)
g = mont(GX) + mont(GY)
pts = eng.scalar_mul_batch(g * N, t.tobytes())
sc = s.tobytes()
d_p, d_s, d_o = lib.cdp_dev_alloc(h, 96 * N), lib.cdp_dev_alloc(h, 32 * N), lib.cdp_dev_alloc(h, 144)
hp, hs = (ctypes.c_uint8 * len(pts)).from_buffer_copy(pts), (ctypes.c_uint8 * len(sc)).from_buffer_copy(sc)
lib.cdp_h2d(h, d_p, hp, len(pts)); lib.cdp_h2d(h, d_s, hs, len(sc)); eng.sync()
print(f"env: chunk={os.environ.get('CDP_MSM_SINGLE_CHUNK', '-')} big_min_log2={os.environ.get('CDP_BIG_MIN_LOG2', '-')} big_c={os.environ.get('CDP_BIG_C', '-')}")
for lg in range(lo, hi + 1):
    n = 1 << lg
    for _ in range(2):
        lib.cdp_msm_dev(h, d_p, d_s, n, d_o)
    torch.cuda.synchronize()
    ms = []
    for _ in range(7 if lg < 18 else 3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); lib.cdp_msm_dev(h, d_p, d_s, n, d_o); e1.record(stream)
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    out = (ctypes.c_uint8 * 144)()
    t0 = time.perf_counter(); lib.cdp_msm(h, hp, hs, n, out); t1 = time.perf_counter()
    t0 = time.perf_counter(); lib.cdp_msm(h, hp, hs, n, out); t1 = time.perf_counter()
    m = statistics.median(ms)
    print(f"2^{lg:2d}: resident {m:8.3f} ms  {n / m / 1e3:9.3f} M pairs/s   host-buffer call {1e3 * (t1 - t0):8.3f} ms   result {eng.compress_batch(bytes(out)).hex()[:16]}", flush=True)
