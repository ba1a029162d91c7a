#!/usr/bin/env python3
"""BASELINE config 5, the extra scalar / base distributions: uniform, all-equal scalars (the SamePerm case, a11),
25 % zero scalars, 32-bit scalars, 1 % infinity bases -- timing of cdp_msm_dev at N = 2^k on one GPU plus a size-independent
correctness property for each (no CPU oracle involved at this size):

    all-equal   msm(P, s..s)            == s * msm(P, 1..1)                       (scalar-mul kernel on the sum)
    25 % zeros  msm(P, s)               == msm(P', s') with the zero pairs removed
    32-bit      msm(P, s) + msm(P, t)   == msm(P, s + t)                          (linearity)
    1 % inf     msm(P, s)               == msm(P', s') with the infinity pairs removed
    uniform     msm(P, s) + msm(P, r-s) == identity

    python tools/msm_distributions.py 20
"""
import json, os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from curdleproofs_b200 import Engine
from bench import mont, GX, GY, R_MOD

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
N = 1 << lg
torch.cuda.set_device(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
eng = Engine(0, stream=stream.cuda_stream)
lib, h = eng.lib, eng.handle
rnd = random.Random(5)
g = mont(GX) + mont(GY)
base_n = min(N, 1 << 14)
pts = eng.scalar_mul_batch(g * base_n, b"".join(rnd.randrange(1, R_MOD).to_bytes(32, "little") for _ in range(base_n)))
while len(pts) < 96 * N:
    k = rnd.randrange(1, R_MOD).to_bytes(32, "little")
    pts += eng.scalar_mul_batch(pts[:96 * base_n], k * base_n)
pts = pts[:96 * N]
fr = lambda v: v.to_bytes(32, "little")  # noqa: E731


def dev(b):
    return torch.frombuffer(bytearray(b), dtype=torch.uint8).cuda()


def msm_dev(t_pts, t_sc, n, reps=0):
    out = torch.zeros(144, dtype=torch.uint8, device="cuda")
    rc = lib.cdp_msm_dev(h, t_pts.data_ptr(), t_sc.data_ptr(), n, out.data_ptr())
    assert rc == 0, lib.cdp_last_error(h)
    ms = None
    if reps:
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            lib.cdp_msm_dev(h, t_pts.data_ptr(), t_sc.data_ptr(), n, out.data_ptr())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
    torch.cuda.synchronize()
    return out.cpu().numpy().tobytes(), ms


def same(j1, j2):
    return eng.compress_batch(j1) == eng.compress_batch(j2)


def add(j1, j2):  # sum of two Jacobian points through the engine: msm of the two normalised points with unit scalars
    return eng.msm(eng.normalize_batch(j1 + j2), fr(1) * 2)


t_pts = dev(pts)
rows = []
# uniform
sc = [rnd.randrange(R_MOD) for _ in range(N)]
r1, ms = msm_dev(t_pts, dev(b"".join(map(fr, sc))), N, reps=5)
r2, _ = msm_dev(t_pts, dev(b"".join(fr((R_MOD - s) % R_MOD) for s in sc)), N)
rows.append({"distribution": "uniform", "ms": ms, "ok": eng.compress_batch(add(r1, r2)) == bytes([0xC0]) + bytes(47)})
# all-equal scalars
s = rnd.randrange(1, R_MOD)
r1, ms = msm_dev(t_pts, dev(fr(s) * N), N, reps=5)
ones, _ = msm_dev(t_pts, dev(fr(1) * N), N)
want = eng.scalar_mul_batch(eng.normalize_batch(ones), fr(s))
rows.append({"distribution": "all-equal scalars", "ms": ms, "ok": eng.compress_batch(r1) == eng.compress_batch(want + mont(1) if want != bytes(96) else bytes(144))})
# 25 % zero scalars
sc = [0 if rnd.random() < 0.25 else rnd.randrange(R_MOD) for _ in range(N)]
r1, ms = msm_dev(t_pts, dev(b"".join(map(fr, sc))), N, reps=5)
keep = [i for i, v in enumerate(sc) if v]
r2, _ = msm_dev(dev(b"".join(pts[96 * i:96 * i + 96] for i in keep)), dev(b"".join(fr(sc[i]) for i in keep)), len(keep))
rows.append({"distribution": "25% zero scalars", "ms": ms, "ok": same(r1, r2)})
# 32-bit scalars
s32, t32 = [rnd.randrange(1 << 32) for _ in range(N)], [rnd.randrange(1 << 32) for _ in range(N)]
r1, ms = msm_dev(t_pts, dev(b"".join(map(fr, s32))), N, reps=5)
r2, _ = msm_dev(t_pts, dev(b"".join(map(fr, t32))), N)
r3, _ = msm_dev(t_pts, dev(b"".join(fr(a + b) for a, b in zip(s32, t32))), N)
rows.append({"distribution": "32-bit scalars", "ms": ms, "ok": same(add(r1, r2), r3)})
# 1 % infinity bases
inf = set(rnd.sample(range(N), N // 100))
pts_inf = b"".join(bytes(96) if i in inf else pts[96 * i:96 * i + 96] for i in range(N))
sc = [rnd.randrange(R_MOD) for _ in range(N)]
r1, ms = msm_dev(dev(pts_inf), dev(b"".join(map(fr, sc))), N, reps=5)
keep = [i for i in range(N) if i not in inf]
r2, _ = msm_dev(dev(b"".join(pts[96 * i:96 * i + 96] for i in keep)), dev(b"".join(fr(sc[i]) for i in keep)), len(keep))
rows.append({"distribution": "1% infinity bases", "ms": ms, "ok": same(r1, r2)})
for r in rows:
    r.update({"log2_n": lg, "pairs_per_s": N / (r["ms"] * 1e-3)})
    print(r, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open(f"gpurun_out/msm_distributions_2p{lg}.json", "w"), indent=1)
assert all(r["ok"] for r in rows)
