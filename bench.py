#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200 Curdleproofs engine.

Metric (BASELINE.json): shuffle proofs/s at ell = 252 (`CurdleproofsProof::new`, /root/reference/src/curdleproofs.rs:59-184), one batch of
independent proofs per step per GPU; beside it verifies/s and the standalone G1 MSM sweep (pairs/s, sharded by base range over the ranks).
One JSON line on stdout (rank 0).

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's CPU path (oracle port; arkworks is not buildable here)

`value`  : proofs/s with the instance vectors already resident in HBM when the timed region starts.
`e2e`    : the same through the host-buffer C ABI (cdp_prove_batch): instance + witnesses + randomness H2D, proofs D2H inside the timed region.
The whole protocol (Fiat-Shamir transcript and scalar algebra included) runs on the GPU; the host draws the prover's randomness (ChaCha12).
`roofline`: the dominant kernel timed in a SERIALISED pass of the same step (lanes one after the other, so CUDA-event durations are
device time of that kernel alone), against the measured HBM copy bandwidth (MEASURED_PEAKS.json) and the measured IMAD.WIDE rate.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

P_MOD = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
GX = 0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB
GY = 0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1
README_PROOFS_PER_S = 1.0 / 0.560  # reference README.md:49, ell = 252 proving on an i7-8550U
IMAD_PER_MIXED_ADD = 8 * 288 + 2 * 222  # XYZZ accumulator + affine point: 8 products, 2 squarings
IMAD_PER_AFFINE_ADD = 5 * 288 + 222     # batched affine addition: 5 products, 1 squaring (shared inversion not counted)
# a table point summed by the tree path (cdp_msm_fixed_batch_dev_tree): 5 halving rounds of affine additions, the last 1/32 by the lane kernel
IMAD_PER_TREE_POINT = (31 * IMAD_PER_AFFINE_ADD + IMAD_PER_MIXED_ADD) / 32


def mont(v):
    return (v * (1 << 384) % P_MOD).to_bytes(48, "little")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ell", type=int, default=252)
    ap.add_argument("--batch", type=int, default=4096, help="proofs per step per GPU")
    ap.add_argument("--lanes", type=int, default=0, help="concurrent sub-batch pipelines per GPU (0 = default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--msm-sizes", default="14,18,22", help="log2 sizes of the MSM sweep ('' = skip)")
    ap.add_argument("--no-extras", action="store_true", help="skip config 4 (small batches gathered to rank 0) and the worst-case verifier step")
    return ap.parse_args()


def common_config(args):
    """The workload both arms run, key for key (arm-specific remarks live outside `config`)."""
    return {"workload": f"ell={args.ell} CurdleproofsProof::new, {args.batch} independent proofs per step per GPU, bit-exact vs reference CPU path",
            "ell": args.ell, "batch_per_gpu": args.batch, "l2_flush": "256 MiB memset between steps",
            "parallelism": "proofs sharded over the GPUs of the node, no data-path collective"}


# ---------------------------------------------------------------------------------------------- synthetic workload
def make_instances(eng, ell, batch, seed):
    """`batch` random shuffle instances on random subgroup points, built with the GPU engine itself.
    Mirrors the reference's benchmark setup (/root/reference/benches/perf.rs:25-66)."""
    import random
    rnd = random.Random(seed)
    g = mont(GX) + mont(GY)
    fr = lambda v: (v % R_MOD).to_bytes(32, "little")  # noqa: E731
    rs = lambda k: b"".join(fr(rnd.randrange(1, R_MOD)) for _ in range(k))  # noqa: E731
    crs = eng.scalar_mul_batch(g * (ell + 7), rs(ell + 7))  # the same CRS for every rank (seeded before the rank offset)
    return crs, rnd, g, fr, rs


def build_batch(eng, crs, ell, batch, rnd, g, fr, rs):
    RS = eng.scalar_mul_batch(g * (2 * ell * batch), rs(2 * ell * batch))
    insts = []
    ks = [rnd.randrange(1, R_MOD) for _ in range(batch)]
    kRS = eng.scalar_mul_batch(RS, b"".join(fr(k) * (2 * ell) for k in ks))
    msm_items = []
    perms, mbs = [], []
    for b in range(batch):
        perm = list(range(ell))
        rnd.shuffle(perm)
        mb = [rnd.randrange(R_MOD) for _ in range(4)]
        perms.append(perm)
        mbs.append(mb)
        msm_items.append((crs[:96 * (ell + 4)], b"".join(fr(x) for x in perm) + b"".join(fr(x) for x in mb)))
    Ms = eng.msm_batch(msm_items)  # M = msm(vec_G, sigma) + msm(vec_H, r_m)   (src/util.rs:98-103)
    for b in range(batch):
        base = 96 * 2 * ell * b
        R, S = RS[base:base + 96 * ell], RS[base + 96 * ell:base + 192 * ell]
        kR, kS = kRS[base:base + 96 * ell], kRS[base + 96 * ell:base + 192 * ell]
        T = b"".join(kR[96 * i:96 * i + 96] for i in perms[b])
        U = b"".join(kS[96 * i:96 * i + 96] for i in perms[b])
        insts.append(dict(ell=ell, crs=crs, R=R, S=S, T=T, U=U, M=Ms[b], perm=perms[b], k=fr(ks[b]),
                          m_blinders=b"".join(fr(x) for x in mbs[b])))
    return insts


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.p:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name).read().strip().split("\n") if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        return out


# ---------------------------------------------------------------------------------------------- CPU port: unit costs
def port_unit_costs(o):
    """Single-thread unit costs of the CPU port, so that its speed can be judged next to arkworks' published figures."""
    import ctypes
    L = o.L
    L.oracle_time_fp_mul_ns.restype = ctypes.c_double
    L.oracle_time_mixed_add_ns.restype = ctypes.c_double
    g = o.generator()
    k = (R_MOD - 12345).to_bytes(32, "little")
    n = 64
    t = time.perf_counter()
    o.scalar_mul_batch(g * n, k * n)
    smul_us = (time.perf_counter() - t) / n * 1e6
    return {"fp_mul_ns": L.oracle_time_fp_mul_ns(2000000), "mixed_add_ns": L.oracle_time_mixed_add_ns(200000), "scalar_mul_255bit_us": smul_us,
            "note": "one thread; Fp product = six-limb CIOS in C (no assembly), scalar multiplication = double-and-add without GLV, as arkworks' Mul<Fr>"}


# ---------------------------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path on the host cores.  arkworks cannot be built in this image
    (no Rust), so this is the C port under oracle/ (pinned bit-exact to the reference's golden proofs), one proof per
    host thread, all threads busy.  A step = a bounded sample of the workload (one proof per host thread)."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    o = oracle_lib.Oracle()
    cores = os.cpu_count() or 1
    ell = args.ell
    crs = o.crs_points(ell)
    inst = o.random_instance(ell, crs, seed=1, threads=cores)
    import ctypes
    perm = (ctypes.c_uint32 * ell)(*inst["perm"])
    buf = lambda b: (ctypes.c_uint8 * len(b)).from_buffer_copy(b)  # noqa: E731
    a = [buf(inst[k]) for k in ("crs", "R", "S", "T", "U", "M")]
    kk, mb = buf(inst["k"]), buf(inst["m_blinders"])
    count = cores  # bounded sample per step: one proof per host thread

    def step():
        return o.L.oracle_time_prove(ell, *a, perm, kk, mb, count, cores, None)

    for _ in range(args.warmup):
        step()
    times = [step() for _ in range(args.steps)]
    t = sum(times)
    value = count * args.steps / t
    line = {"impl": "reference", "metric": f"shuffle_proofs_per_sec_ell{ell}", "value": value, "unit": "proofs/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": value / README_PROOFS_PER_S if ell == 252 else None, "dtype": "u64", "data": "synthetic",
            "config": common_config(args),
            "reference_arm": "the same workload on the host cores, a bounded sample per step (C port of the reference path under oracle/, pinned to "
                             "the reference's golden proofs; arkworks itself is not buildable here: no Rust toolchain)",
            "cpu_baseline": {"value": value, "unit": "proofs/s", "cores": cores, "kind": "port",
                             "sample": f"{count} proofs per step, one per host thread, {args.steps} steps after {args.warmup} warm-up steps",
                             "unit_costs": port_unit_costs(o)},
            "e2e": {"value": value, "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- MSM sweep (metric 2)
def msm_sweep(args, eng, stream, rank, world, flush_buf, with_cpu):
    """Standalone G1 MSM (`util::msm`, /root/reference/src/util.rs:19-22) of 2^k pairs, device-resident, sharded by base range over the ranks
    through the engine's communicator: the only collective is one all-gather of the 144-byte partial sums.  Inputs depend on the global
    index only, so the result (`checksum`: its 48-byte encoding) is the same for every N."""
    import ctypes

    import numpy as np
    import torch
    import torch.distributed as dist
    from curdleproofs_b200.sharded import Comm, shard_range
    sizes = [int(x) for x in args.msm_sizes.split(",") if x.strip()]
    if not sizes:
        return None
    lib, h = eng.lib, eng.handle
    comm = Comm.from_torch_distributed(eng, rank, world)
    g = mont(GX) + mont(GY)
    out = {"metric": "g1_msm_pairs_per_sec", "unit": "pairs/s", "collective": "one ncclAllGather of the 144-byte partial sums per MSM (cdp_msm_sharded_dev)",
           "scalars": "uniform 254-bit (< r)", "bases": "t_i * G, t_i uniform 254-bit, seeded by the global index", "points": []}
    o = None
    if with_cpu:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib
        o = oracle_lib.Oracle()
        o.L.oracle_msm.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int]
    for lg in sizes:
        n = 1 << lg
        lo, hi = shard_range(n, rank, world)
        nl = hi - lo
        rng = np.random.Generator(np.random.Philox(key=7000 + lg))
        t_all = rng.integers(0, 2 ** 64, size=(n, 4), dtype=np.uint64)
        s_all = rng.integers(0, 2 ** 64, size=(n, 4), dtype=np.uint64)
        t_all[:, 3] &= np.uint64((1 << 62) - 1)
        s_all[:, 3] &= np.uint64((1 << 62) - 1)
        sc = s_all[lo:hi].tobytes()
        pts = eng.scalar_mul_batch(g * nl, t_all[lo:hi].tobytes()) if nl else b""
        d_p, d_s, d_o = lib.cdp_dev_alloc(h, max(96 * nl, 96)), lib.cdp_dev_alloc(h, max(32 * nl, 32)), lib.cdp_dev_alloc(h, 144)
        if nl:
            lib.cdp_h2d(h, d_p, (ctypes.c_uint8 * len(pts)).from_buffer_copy(pts), len(pts))
            lib.cdp_h2d(h, d_s, (ctypes.c_uint8 * len(sc)).from_buffer_copy(sc), len(sc))
        eng.sync()
        reps = 3 if lg >= 20 else 10
        for _ in range(2):
            comm.msm_sharded_dev(d_p, d_s, nl, d_o)
        eng.sync()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms_list = []
        for _ in range(reps):
            with torch.cuda.stream(stream):
                flush_buf.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            comm.msm_sharded_dev(d_p, d_s, nl, d_o)
            e1.record(stream)
            torch.cuda.synchronize()
            ms_list.append(e0.elapsed_time(e1))
        ms = statistics.median(ms_list)
        if world > 1:
            tt = torch.tensor([ms], device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        res = (ctypes.c_uint8 * 144)()
        lib.cdp_d2h(h, res, d_o, 144)
        eng.sync()
        checksum = eng.compress_batch(bytes(res)).hex()
        pt = {"log2_n": lg, "pairs_per_s": n / (ms * 1e-3), "ms": ms, "reps": reps, "checksum": checksum,
              "achieved_GBps_128B_per_pair": 128 * n / (ms * 1e-3) / 1e9}
        if o is not None and world == 1:
            cores = os.cpu_count() or 1
            cj = (ctypes.c_uint8 * 144)()
            t0 = time.perf_counter()
            o.L.oracle_msm(pts, sc, n, cj, cores)
            dt = time.perf_counter() - t0
            pt["cpu"] = {"pairs_per_s": n / dt, "seconds": dt, "cores": cores, "kind": "port",
                         "result_matches_gpu": o.compress_jac(bytes(cj)).hex() == checksum}
        out["points"].append(pt)
        for d in (d_p, d_s, d_o):
            lib.cdp_dev_free(h, d)
        del t_all, s_all, pts, sc
    comm.close()
    return out


# ---------------------------------------------------------------------------------------------- main arm
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import ctypes

    import torch
    import torch.distributed as dist
    from curdleproofs_b200 import BatchProver, BatchVerifier, Engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.Stream()
    eng = Engine(local_rank, stream=stream.cuda_stream)
    ell, B = args.ell, args.batch

    crs, rnd, g, fr, rs = make_instances(eng, ell, B, seed=2024)
    rnd.seed(7777 + rank)  # rank-specific instances, shared CRS
    insts = build_batch(eng, crs, ell, B, rnd, g, fr, rs)
    # host threads: the box's cores are shared by the ranks of this node (one process per GPU); the host only stages buffers and draws randomness
    cores = os.cpu_count() or 8
    host_threads = max(1, cores // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", world))))
    bp = BatchProver(eng, ell, crs, max_batch=B, lanes=args.lanes, host_threads=host_threads)

    cat = lambda key: b"".join(i[key] for i in insts)  # noqa: E731
    arr = lambda b: (ctypes.c_uint8 * len(b)).from_buffer_copy(b)  # noqa: E731
    # the step's inputs live in page-locked host memory (cdp_host_alloc): the e2e copies are DMAs from the caller's own buffers
    R, S, T, U = (eng.pinned_array(cat(k)) for k in ("R", "S", "T", "U"))
    M, K, MB = (arr(cat(k)) for k in ("M", "k", "m_blinders"))
    perm = (ctypes.c_uint32 * (B * ell))(*[x for i in insts for x in i["perm"]])
    seeds = (ctypes.c_uint64 * B)(*range(1000 * rank, 1000 * rank + B))
    out = (ctypes.c_uint8 * (B * bp.proof_size))()
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def step(resident, prover=None, n=B, outbuf=None):
        prover = prover or bp
        with torch.cuda.stream(stream):
            flush_buf.zero_()  # L2 flush between steps (256 MiB > 126 MB L2)
        if resident:
            prover.prove_raw(n, None, None, None, None, None, perm, K, MB, seeds, out=outbuf or out, split=False)
        else:
            prover.prove_raw(n, R, S, T, U, M, perm, K, MB, seeds, out=outbuf or out, split=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = bp.launch_count
        with torch.cuda.stream(stream):
            e0.record(stream)
        for _ in range(steps):
            fn()
        with torch.cuda.stream(stream):
            e1.record(stream)
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)), bp.launch_count - l0

    step(False)  # stages the instance batch in HBM (and is the first warm-up step)
    for _ in range(max(0, args.warmup - 1)):
        step(True)
    # ---- timed region 1: inputs resident in HBM
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_res, launches = timed(lambda: step(True), args.steps)
    clocks = sampler.stop() if sampler else None
    timing = bp.last_timing()
    # ---- timed region 2: end to end through the host-buffer API
    ms_e2e, _ = timed(lambda: step(False), args.steps)
    traffic = bp.last_traffic()
    proof0 = bytes(out[:bp.proof_size])
    # ---- per-kernel device time: one more step with the lanes SERIALISED (nothing overlaps, so event durations are device time of each kernel)
    bp.set_serial(True)
    bp.profile_reset()
    bp.profile_enable(True)
    t_ser = time.perf_counter()
    step(True)
    torch.cuda.synchronize()
    t_ser = (time.perf_counter() - t_ser) * 1e3
    prof = bp.profile_read()
    bp.profile_enable(False)
    bp.set_serial(False)
    table_bytes = bp.table_bytes

    # ---- secondary metric: verifies/s on the proofs just produced (CurdleproofsProof::deserialize + verify through cdp_verify_batch)
    VB = B
    bv = BatchVerifier(eng, ell, crs, max_batch=VB, host_threads=host_threads)  # shares the prover's digit table
    vout = (ctypes.c_uint8 * VB)()

    def vstep(proofs=out):
        bv.verify_raw(VB, R, S, T, U, M, proofs, None, out=vout)

    vstep()
    ms_ver, _ = timed(vstep, args.steps)
    all_ok = all(x == 1 for x in vout)
    verify = {"metric": f"shuffle_verifies_per_sec_ell{ell}", "value": world * VB * args.steps / (ms_ver * 1e-3), "unit": "verifies/s",
              "batch_per_gpu": VB, "ms_per_step": ms_ver / args.steps, "all_accepted": all_ok,
              "note": "host buffers in (instance + serialised proofs), verdicts out; README.md:49 reference: 35 ms/verify on i7-8550U"}
    extras = {}
    if not args.no_extras:
        # worst case for the merged check: one invalid proof in every lane's sub-batch -> every lane falls back to proof-by-proof MSMs
        bad = (ctypes.c_uint8 * len(out)).from_buffer_copy(out)
        nl = max(1, bv.lanes)
        per = (VB + nl - 1) // nl
        bad_idx = list(range(0, VB, per))
        for i in bad_idx:
            bad[i * bp.proof_size + bp.proof_size - 1] ^= 1  # x_final: still canonical with overwhelming probability, no longer valid
        s0 = bv.merge_stats()
        vstep(bad)
        ms_bad, _ = timed(lambda: vstep(bad), 2)
        s1 = bv.merge_stats()
        verdicts_ok = all((vout[i] != 1) == (i in set(bad_idx)) for i in range(VB))
        verify["worst_case"] = {"value": world * VB * 2 / (ms_bad * 1e-3), "unit": "verifies/s", "ms_per_step": ms_bad / 2,
                                "what": f"{len(bad_idx)} invalid proofs per step, one in every lane sub-batch: the merged check rejects and every proof is decided by its own accumulated MSM",
                                "fallback_sub_batches": s1["fallback"] - s0["fallback"], "verdicts_correct": verdicts_ok}
    bv.close()

    if not args.no_extras:
        # config 4 of BASELINE.json: 1024 proofs over 8 GPUs = 128 per GPU per step, proofs gathered to rank 0 (end to end, host buffers)
        B4 = min(128, B)
        bp4 = BatchProver(eng, ell, crs, max_batch=B4, host_threads=host_threads)  # shares the digit table
        out4 = (ctypes.c_uint8 * (B4 * bp.proof_size))()
        d_gather = torch.empty(B4 * bp.proof_size, dtype=torch.uint8, device="cuda")
        gathered = [torch.empty_like(d_gather) for _ in range(world)] if (world > 1 and rank == 0) else None

        def step4():
            step(False, prover=bp4, n=B4, outbuf=out4)
            if world > 1:
                d_gather.copy_(torch.frombuffer(out4, dtype=torch.uint8), non_blocking=False)
                dist.gather(d_gather, gathered, dst=0)

        step4(); step4()
        ms4, _ = timed(step4, args.steps)
        extras["config4"] = {"what": f"{B4} proofs per GPU per step ({world * B4} per step over {world} GPU(s)), end to end from host buffers, proofs gathered to rank 0 (NCCL gather)",
                             "value": world * B4 * args.steps / (ms4 * 1e-3), "unit": "proofs/s", "ms_per_step": ms4 / args.steps, "lanes": bp4.lanes}
        bp4.close()

    sweep = None
    try:
        sweep = msm_sweep(args, eng, stream, rank, world, flush_buf, with_cpu=(world == 1 and not args.no_cpu_baseline))
    except Exception as e:  # never lose the headline line
        sweep = {"error": repr(e)}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    value = world * B * args.steps / (ms_res * 1e-3)
    e2e = world * B * args.steps / (ms_e2e * 1e-3)
    # ---- roofline of the dominant kernel, from the serialised pass
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, peak_src = (peaks["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (burst)") if "hbm_gbs" in peaks else (6650.0, "fallback of B200_PROFILING.md")
    # ncu DRAM bytes per unit of the FINAL kernels (profiles/r02_ncu_traffic.json, written from this round's `ncu --set full` captures)
    ncu_traffic = {}
    try:
        ncu_traffic = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")))
    except Exception:
        pass
    throughput_kinds = ("msm_fixed", "msm_buckets", "smul")
    dom = max(throughput_kinds, key=lambda k: prof[k]["ms"])
    d = prof[dom]
    avg_ms = d["ms"] / max(1, d["launches"])
    units_per_launch = d["units"] / max(1, d["launches"])
    # an MSM pair is 128 algorithmic bytes (32-byte scalar + 96-byte base), a fold element 288
    contract_bytes = {"msm_fixed": 128, "msm_buckets": 128, "smul": 288}[dom]
    achieved = contract_bytes * units_per_launch / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
    imad_ms = min(eng.bench_kernel(0, 148 * 4, 256, 2000) for _ in range(3))
    imad_peak = 148 * 4 * 256 * 2000 * 128 / (imad_ms * 1e-3)
    total_kernel_ms = sum(v["ms"] for v in prof.values())
    tr = ncu_traffic.get(dom, {})
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": tr.get("dram_bytes_per_unit") * units_per_launch if tr.get("dram_bytes_per_unit") else None,
                "traffic_source": tr.get("source"), "peak_source": peak_src,
                "algorithmic_bytes_per_unit": contract_bytes, "units_per_launch": units_per_launch, "avg_launch_ms": avg_ms,
                "timing": "CUDA events around every launch of ONE extra step with the prover's lanes serialised (cdp_prover_set_serial): kernels run alone",
                "share_of_kernel_time": d["ms"] / total_kernel_ms if total_kernel_ms else None,
                "serialised_step_ms": t_ser, "serialised_kernel_ms_sum": total_kernel_ms,
                "kernel_ms": {k: v["ms"] for k, v in prof.items()}, "kernel_launches": {k: v["launches"] for k, v in prof.items()},
                "note": "381-bit modular arithmetic: the binding roofline is the integer multiply pipe, not HBM (int_pipe); fixed-base pairs also gather "
                        "16 table entries of 96 B each (table_gather_*), which the 128-byte contract figure does not count"}
    if dom == "msm_fixed":
        gather_bytes = 32 + 16 * 96
        adds_per_s = 16 * units_per_launch / (avg_ms * 1e-3)
        roofline["table_gather_bytes_per_pair"] = gather_bytes
        roofline["table_gather_achieved_GBps"] = gather_bytes * units_per_launch / (avg_ms * 1e-3) / 1e9
        roofline["table_gather_frac"] = roofline["table_gather_achieved_GBps"] / hbm_peak
        # most table points of a prover step are summed by the tree path (long uniform segments); the rest by the lane kernel's XYZZ accumulators.
        # The fraction is quoted on the tree's count -- the fewest multiply-adds the method in use needs -- and the lane count is given beside it.
        roofline["int_pipe"] = {"peak": imad_peak, "unit": "IMAD.WIDE.U32/s", "peak_source": "measured live (k_bench_imad, register-only)",
                                "achieved": adds_per_s * IMAD_PER_TREE_POINT, "frac": adds_per_s * IMAD_PER_TREE_POINT / imad_peak,
                                "table_points_per_s": adds_per_s, "imad_wide_per_table_point": IMAD_PER_TREE_POINT,
                                "imad_wide_per_affine_add": IMAD_PER_AFFINE_ADD, "imad_wide_per_xyzz_mixed_add": IMAD_PER_MIXED_ADD,
                                "frac_if_all_xyzz": adds_per_s * IMAD_PER_MIXED_ADD / imad_peak}
    elif dom == "msm_buckets":
        # XYZZ mixed additions per pair of the bucket kernel: 2 GLV halves x the windows of the launch's width; over the launches of an
        # ell = 252 step (1792 pairs at c = 5 -> 52 additions, 192 at c = 4 -> 66, 48 at c = 3 -> 88, 15 at c = 2 -> 130) 54.7 on average
        adds_per_pair = 54.7
        adds_per_s = adds_per_pair * units_per_launch / (avg_ms * 1e-3)
        roofline["int_pipe"] = {"peak": imad_peak, "unit": "IMAD.WIDE.U32/s", "peak_source": "measured live (k_bench_imad, register-only)",
                                "achieved": adds_per_s * IMAD_PER_MIXED_ADD, "frac": adds_per_s * IMAD_PER_MIXED_ADD / imad_peak,
                                "mixed_adds_per_s": adds_per_s, "mixed_adds_per_pair": adds_per_pair, "imad_wide_per_xyzz_mixed_add": IMAD_PER_MIXED_ADD,
                                "note": "bucket accumulation only; the window combine is a separate kernel (kernel_ms.msm_combine)"}
    else:
        # fold elements: 128 doublings (2M + 5S) + ~64 mixed additions (7M + 4S) each
        imad_per_elem = 128 * (2 * 288 + 5 * 222) + 64 * (7 * 288 + 4 * 222)
        elems_per_s = units_per_launch / (avg_ms * 1e-3)
        roofline["int_pipe"] = {"peak": imad_peak, "unit": "IMAD.WIDE.U32/s", "peak_source": "measured live (k_bench_imad, register-only)",
                                "achieved": elems_per_s * imad_per_elem, "frac": elems_per_s * imad_per_elem / imad_peak,
                                "elements_per_s": elems_per_s, "imad_wide_per_element": imad_per_elem}
    # ---- the same kernel timed ALONE on synthetic scalars (nothing else on the GPU): one fixed-base launch of the IPA-round shape for the whole batch
    try:
        from curdleproofs_b200 import FixedSeg
        n_ = ell + 4
        tab = eng.fixed_table_create(crs, 16)
        nseg = B * 4
        segs = (FixedSeg * nseg)()
        for i in range(nseg):
            segs[i].base_off = 0; segs[i].scalars_off = (i // 4) * (2 * n_ + 2) + (n_ + 2) * ((i % 4) // 2); segs[i].n = n_ // 2
            segs[i].sel_h = n_ // 2; segs[i].sel_val = (n_ // 2) * (i % 2); segs[i].remap_from = 0xFFFFFFFF
            if i % 4 < 2:  # L_C / R_C carry the extra `+ ip * H` pair (scalars n_, n_ + 1 of the proof's block)
                segs[i].extra_base = n_ + 1; segs[i].extra_scalar = n_ + (i % 2)
            segs[i].out_idx = i
        sc = bytearray(os.urandom(32 * B * (2 * n_ + 2)))
        sc[31::32] = bytes(x & 0x3F for x in sc[31::32])
        lib, h = eng.lib, eng.handle
        d_sc, d_sg, d_out = lib.cdp_dev_alloc(h, len(sc)), lib.cdp_dev_alloc(h, ctypes.sizeof(segs)), lib.cdp_dev_alloc(h, nseg * 144)
        lib.cdp_h2d(h, d_sc, arr(bytes(sc)), len(sc)); lib.cdp_h2d(h, d_sg, segs, ctypes.sizeof(segs)); eng.sync()
        pairs = B * (2 * n_ + 2)

        def isolated(call):
            for _ in range(3):
                call()
            eng.sync(); eng.profile_reset(); eng.profile_enable(True)
            for _ in range(5):
                with torch.cuda.stream(stream):
                    flush_buf.zero_()
                call()
            eng.sync()
            pf = eng.profile_read()["msm_fixed"]
            eng.profile_enable(False)
            return pf["ms"] / pf["launches"]

        iso_tree = isolated(lambda: lib.cdp_msm_fixed_batch_dev_tree(h, tab.handle, d_sc, d_sg, nseg, pairs, None, d_out, n_ // 2 + 1))
        iso_lanes = isolated(lambda: lib.cdp_msm_fixed_batch_dev_lanes(h, tab.handle, d_sc, d_sg, nseg, pairs, None, d_out, 16))
        iso_ms = iso_tree
        madds = pairs * 16 / (iso_ms * 1e-3)
        roofline["isolated"] = {"kernel": "fixed-base MSM as the prover launches it (tree of batched affine additions: k_fixed_ba_first / _next + k_fixed_msm), "
                                          "IPA-round shape, whole batch in one call, nothing else running", "ms": iso_ms,
                                "pairs_per_s": pairs / (iso_ms * 1e-3),
                                "achieved_GBps_128B_per_pair": pairs * 128 / (iso_ms * 1e-3) / 1e9, "hbm_frac_128B_per_pair": pairs * 128 / (iso_ms * 1e-3) / 1e9 / hbm_peak,
                                "achieved_GBps_table_gather": pairs * (32 + 16 * 96) / (iso_ms * 1e-3) / 1e9,
                                "hbm_frac_table_gather": pairs * (32 + 16 * 96) / (iso_ms * 1e-3) / 1e9 / hbm_peak,
                                "table_points_per_s": madds, "imad_wide_per_table_point": IMAD_PER_TREE_POINT,
                                "int_pipe_frac": madds * IMAD_PER_TREE_POINT / imad_peak,
                                "lane_kernel": {"kernel": "k_fixed_msm alone, 16 lanes per segment, XYZZ accumulators", "ms": iso_lanes,
                                                "pairs_per_s": pairs / (iso_lanes * 1e-3), "mixed_adds_per_s": pairs * 16 / (iso_lanes * 1e-3),
                                                "imad_wide_per_mixed_add": IMAD_PER_MIXED_ADD,
                                                "int_pipe_frac": pairs * 16 / (iso_lanes * 1e-3) * IMAD_PER_MIXED_ADD / imad_peak}}
        for dd in (d_sc, d_sg, d_out):
            lib.cdp_dev_free(h, dd)
        tab.close()
    except Exception as e:  # diagnostics only: never lose the headline line
        roofline["isolated"] = {"error": repr(e)}
    line = {"metric": f"shuffle_proofs_per_sec_ell{ell}", "value": value, "unit": "proofs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": value / README_PROOFS_PER_S if ell == 252 else None, "dtype": "u32", "data": "synthetic",
            "config": common_config(args),
            "run": {"host_threads": host_threads, "host_cores": cores, "lanes": bp.lanes, "table_bytes": table_bytes,
                    "baseline": "README.md:49 560 ms/proof on i7-8550U (other hardware)",
                    "transcript": "device (cdp_prove_stage_dev)" if os.environ.get("CDP_PROVE_HOST_TRANSCRIPT", "0") in ("", "0") else "host"},
            "e2e": {"value": e2e, "unit": "proofs/s", "h2d_bytes_per_step": traffic["h2d_bytes"], "d2h_bytes_per_step": traffic["d2h_bytes"],
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "roofline": roofline, "clocks": clocks, "verify": verify, "msm_sweep": sweep, "extra": extras,
            "host_breakdown_last_step_ms": timing}
    bp.close()
    # ---- CPU baseline beside it (rank 0, N = 1 only): the oracle port on a bounded sample, and a parity check of proof 0
    if world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib
        o = oracle_lib.Oracle()
        inst = insts[0]
        pc = (ctypes.c_uint32 * ell)(*inst["perm"])
        a = [arr(inst[k]) for k in ("crs", "R", "S", "T", "U", "M")]
        last = (ctypes.c_uint8 * bp.proof_size)()
        count = 2 * cores
        t = o.L.oracle_time_prove(ell, *a, pc, arr(inst["k"]), arr(inst["m_blinders"]), count, cores, last)
        want0 = o.prove(inst, rng_seed=int(seeds[0]), threads=cores)
        allok = ctypes.c_int(0)
        tv = o.L.oracle_time_verify(ell, *a, arr(want0), 4 * cores, cores, ctypes.byref(allok))
        line["verify"]["cpu_baseline"] = {"value": 4 * cores / tv, "unit": "verifies/s", "cores": cores, "kind": "port", "all_accepted": bool(allok.value)}
        line["cpu_baseline"] = {"value": count / t, "unit": "proofs/s", "cores": cores, "kind": "port",
                                "sample": f"{count} ell={ell} proofs, one per host thread ({cores} threads), oracle C port",
                                "parity_proof0_bit_exact": bool(want0 == proof0), "oracle_verifies_gpu_proof": o.verify(inst, proof0) == 1,
                                "unit_costs": port_unit_costs(o)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
