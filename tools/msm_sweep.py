#!/usr/bin/env python3
"""Standalone G1 MSM sweep (BASELINE.json config 5): pairs/s for N = 2^10 .. 2^max at 1..8 GPUs.

    python tools/msm_sweep.py 22
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 tools/msm_sweep.py 22

Bases: N random subgroup points made on the GPU; scalars: uniform < r from a seeded PRNG (seed 0).  With W ranks the bases are
sharded by contiguous range; every rank computes its partial sum, the 144-byte partials are all-gathered and added locally.
Device-resident timing (CUDA events, max over ranks); the 1-GPU run also reports the host-buffer call."""
import json, os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from curdleproofs_b200 import Engine
from curdleproofs_b200.sharded import shard_range, sharded_msm_dev
from bench import mont, GX, GY, R_MOD

max_log = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)  # kernels of the engine and the NCCL collective share this stream
eng = Engine(local, stream=stream.cuda_stream)
rnd = random.Random(0)
N = 1 << max_log
g = mont(GX) + mont(GY)
base_n = min(N, 1 << 14)
pts = eng.scalar_mul_batch(g * base_n, b"".join((rnd.randrange(1, R_MOD)).to_bytes(32, "little") for _ in range(base_n)))
while len(pts) < 96 * N:
    k = rnd.randrange(1, R_MOD).to_bytes(32, "little")
    pts += eng.scalar_mul_batch(pts[:96 * base_n], k * base_n)
sc = b"".join(rnd.randrange(R_MOD).to_bytes(32, "little") for _ in range(N))
t_pts = torch.frombuffer(bytearray(pts), dtype=torch.uint8).cuda()
t_sc = torch.frombuffer(bytearray(sc), dtype=torch.uint8).cuda()
res = []
for lg in range(10, max_log + 1):
    n = 1 << lg
    lo, hi = shard_range(n, rank, world)
    run = lambda: sharded_msm_dev(eng, t_pts.data_ptr() + 96 * lo, t_sc.data_ptr() + 32 * lo, hi - lo, world)  # noqa: E731
    for _ in range(2):
        out = run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    reps = 5 if lg >= 18 else 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    if world > 1:
        t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    comp = eng.compress_batch(out.cpu().numpy().tobytes())
    row = {"log2_n": lg, "n_gpus": world, "ms": ms, "pairs_per_s": n / (ms * 1e-3), "GBps_algorithmic": 128 * n / (ms * 1e-3) / 1e9, "result": comp.hex()[:16]}
    if world == 1:
        t = time.time(); eng.msm(pts[:96 * n], sc[:32 * n]); row["e2e_ms"] = (time.time() - t) * 1e3
    res.append(row)
    if rank == 0:
        print(row, flush=True)
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open(f"gpurun_out/msm_sweep_N{world}.json", "w"), indent=1)
if world > 1:
    dist.destroy_process_group()
