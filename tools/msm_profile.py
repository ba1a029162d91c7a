#!/usr/bin/env python3
"""Per-launch timing of one cdp_msm_dev call at N = 2^k (run with CDP_PROFILE_DUMP=file)."""
import ctypes, os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from curdleproofs_b200 import Engine
from bench import mont, GX, GY, R_MOD
lg = int(sys.argv[1]); N = 1 << lg
eng = Engine(0); lib, h = eng.lib, eng.handle
rnd = random.Random(0)
g = mont(GX) + mont(GY)
base_n = min(N, 1 << 14)
pts = eng.scalar_mul_batch(g * base_n, b"".join((rnd.randrange(1, R_MOD)).to_bytes(32, "little") for _ in range(base_n)))
pts = (pts * (N // base_n + 1))[:96 * N]
sc = b"".join(rnd.randrange(R_MOD).to_bytes(32, "little") for _ in range(N))
d_pts = lib.cdp_dev_alloc(h, len(pts)); d_sc = lib.cdp_dev_alloc(h, len(sc)); d_out = lib.cdp_dev_alloc(h, 144)
lib.cdp_h2d(h, d_pts, (ctypes.c_uint8 * len(pts)).from_buffer_copy(pts), len(pts)); lib.cdp_h2d(h, d_sc, (ctypes.c_uint8 * len(sc)).from_buffer_copy(sc), len(sc)); eng.sync()
lib.cdp_msm_dev(h, d_pts, d_sc, N, d_out); eng.sync()
eng.profile_reset(); eng.profile_enable(True)
lib.cdp_msm_dev(h, d_pts, d_sc, N, d_out); eng.sync()
print(lg, eng.profile_read())
