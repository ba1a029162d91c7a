#!/usr/bin/env python3
"""Standalone G1 MSM sweep (BASELINE.json config 5) on cuda:0: device-resident pairs/s for N = 2^10 .. 2^max.
Bases: N random subgroup points made on the GPU; scalars: uniform < r from a seeded PRNG."""
import ctypes, json, os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from curdleproofs_b200 import Engine
from bench import mont, GX, GY, R_MOD

max_log = int(sys.argv[1]) if len(sys.argv) > 1 else 20
stream = torch.cuda.Stream()
eng = Engine(0, stream=stream.cuda_stream)
lib, h = eng.lib, eng.handle
rnd = random.Random(0)
N = 1 << max_log
g = mont(GX) + mont(GY)
# distinct random points: s_i * G for 2^14 scalars, then tile with per-tile scalar multiples (cheap, still distinct points)
base_n = min(N, 1 << 14)
pts = eng.scalar_mul_batch(g * base_n, b"".join((rnd.randrange(1, R_MOD)).to_bytes(32, "little") for _ in range(base_n)))
while len(pts) < 96 * N:
    k = rnd.randrange(1, R_MOD).to_bytes(32, "little")
    pts += eng.scalar_mul_batch(pts[:96 * base_n], k * base_n)
sc = b"".join(rnd.randrange(R_MOD).to_bytes(32, "little") for _ in range(N))
d_pts = lib.cdp_dev_alloc(h, len(pts)); d_sc = lib.cdp_dev_alloc(h, len(sc)); d_out = lib.cdp_dev_alloc(h, 144)
bp = (ctypes.c_uint8 * len(pts)).from_buffer_copy(pts); bs = (ctypes.c_uint8 * len(sc)).from_buffer_copy(sc)
lib.cdp_h2d(h, d_pts, bp, len(pts)); lib.cdp_h2d(h, d_sc, bs, len(sc)); eng.sync()
res = []
for lg in range(10, max_log + 1):
    n = 1 << lg
    for _ in range(2):
        lib.cdp_msm_dev(h, d_pts, d_sc, n, d_out)
    eng.sync()
    reps = 5 if lg >= 18 else 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
    for _ in range(reps):
        lib.cdp_msm_dev(h, d_pts, d_sc, n, d_out)
    with torch.cuda.stream(stream):
        e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    t = time.time(); out = eng.msm(pts[:96 * n], sc[:32 * n]); e2e_ms = (time.time() - t) * 1e3
    res.append({"log2_n": lg, "ms": ms, "pairs_per_s": n / (ms * 1e-3), "GBps_algorithmic": 128 * n / (ms * 1e-3) / 1e9, "e2e_ms": e2e_ms})
    print(res[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/msm_sweep.json", "w"), indent=1)
