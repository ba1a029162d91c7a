#!/usr/bin/env python
"""Timing of the fixed-base path: table construction for a CRS-sized base set and a prover-round-shaped batch of segments."""
import ctypes
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import py_ref as pr
from curdleproofs_b200 import Engine, FixedSeg

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 259
bits = int(sys.argv[2]) if len(sys.argv) > 2 else 16
eng = Engine(0)
rnd = random.Random(1)
g = pr.affine_to_bytes(pr.G1)
sc = b"".join(pr.fr_to_bytes(rnd.randrange(pr.R_ORDER)) for _ in range(nb))
bases = eng.scalar_mul_batch(g * nb, sc)
t0 = time.time()
tab = eng.fixed_table_create(bases, bits)
print(f"table: {nb} bases, c={bits}, {tab.nbytes / 2**30:.2f} GiB, built in {time.time() - t0:.3f} s")
lib, h = eng.lib, eng.handle
for (B, K, n) in ((1024, 4, 128), (128, 4, 128), (1024, 5, 256), (128, 5, 256), (1024, 4, 1)):
    nsc = B * K * n
    scal = os.urandom(32 * nsc)
    scal = bytes(b & 0x3F if i % 32 == 31 else b for i, b in enumerate(scal)) if nsc < 100000 else bytes(bytearray(scal)[:])  # < 2^254
    if nsc >= 100000:
        ba = bytearray(scal)
        ba[31::32] = bytes(x & 0x3F for x in ba[31::32])
        scal = bytes(ba)
    segs = (FixedSeg * (B * K))()
    for i in range(B * K):
        segs[i].base_off = 0; segs[i].scalars_off = i * n; segs[i].n = n; segs[i].remap_from = 0xFFFFFFFF; segs[i].out_idx = i
    d_sc = lib.cdp_dev_alloc(h, len(scal)); d_sg = lib.cdp_dev_alloc(h, ctypes.sizeof(segs)); d_out = lib.cdp_dev_alloc(h, B * K * 144)
    buf = (ctypes.c_uint8 * len(scal)).from_buffer_copy(scal)
    lib.cdp_h2d(h, d_sc, buf, len(scal)); lib.cdp_h2d(h, d_sg, segs, ctypes.sizeof(segs)); eng.sync()
    for it in range(2):
        lib.cdp_msm_fixed_batch_dev(h, tab.handle, d_sc, d_sg, B * K, nsc, None, d_out)
    eng.sync()
    eng.profile_reset(); eng.profile_enable(True)
    for it in range(5):
        lib.cdp_msm_fixed_batch_dev(h, tab.handle, d_sc, d_sg, B * K, nsc, None, d_out)
    eng.sync()
    p = eng.profile_read()["msm_fixed"]
    eng.profile_enable(False)
    ms = p["ms"] / p["launches"]
    nw = (256 + bits - 1) // bits
    print(f"B={B} K={K} n={n}: {ms:.3f} ms/launch, {nsc / ms / 1e3:.2f} M pairs/s, {nsc * nw / ms / 1e6:.3f} G adds/s")
    # the same batch as a tree of batched affine additions
    for it in range(2):
        lib.cdp_msm_fixed_batch_dev_tree(h, tab.handle, d_sc, d_sg, B * K, nsc, None, d_out, n)
    eng.sync()
    eng.profile_reset(); eng.profile_enable(True)
    for it in range(5):
        lib.cdp_msm_fixed_batch_dev_tree(h, tab.handle, d_sc, d_sg, B * K, nsc, None, d_out, n)
    eng.sync()
    p = eng.profile_read()["msm_fixed"]
    eng.profile_enable(False)
    ms = p["ms"] / p["launches"]
    print(f"   tree: {ms:.3f} ms/launch, {nsc / ms / 1e3:.2f} M pairs/s, {nsc * nw / ms / 1e6:.3f} G adds/s")
    for d in (d_sc, d_sg, d_out):
        lib.cdp_dev_free(h, d)
