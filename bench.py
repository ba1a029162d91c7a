#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200 Curdleproofs engine.

Metric (BASELINE.json): shuffle proofs/s at ell = 252 (`CurdleproofsProof::new`, /root/reference/src/curdleproofs.rs:59-184),
one batch of independent proofs per step per GPU.  One JSON line on stdout (rank 0).

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's CPU path (oracle port; arkworks is not buildable here)

`value`  : proofs/s with the instance vectors already resident in HBM when the timed region starts.
`e2e`    : the same through the host-buffer C ABI (cdp_prove_batch): instance H2D + proofs D2H inside the timed region.
Both include the per-round scalar uploads / 48-byte point downloads that the host side of the Fiat-Shamir transcript needs
(its opening -- the bulk of the hashing -- runs on the GPU).  The CRS digit table (fixed-base MSM) is built once at prover creation.
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
    return ap.parse_args()


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


# ---------------------------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path on the host cores.  arkworks cannot be built in this image
    (no Rust), so this is the C port under oracle/ (pinned bit-exact to the reference's golden proofs), one proof per
    host thread, all threads busy."""
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
    for _ in range(min(args.warmup, 1)):
        step()
    times = [step() for _ in range(args.steps)]
    t = sum(times)
    value = count * args.steps / t
    line = {"impl": "reference", "metric": f"shuffle_proofs_per_sec_ell{ell}", "value": value, "unit": "proofs/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": value / README_PROOFS_PER_S if ell == 252 else None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": f"ell={ell} CurdleproofsProof::new, {args.batch} independent proofs per step per GPU, bit-exact vs reference CPU path",
                       "ell": ell, "batch_per_gpu": args.batch, "proofs_per_step": count,
                       "reference_arm": "the same workload on the host cores, bounded sample per step (C port of the reference path under oracle/, "
                                        "pinned to the reference's golden proofs; arkworks itself is not buildable here: no Rust toolchain)"},
            "cpu_baseline": {"value": value, "unit": "proofs/s", "cores": cores, "kind": "port",
                             "sample": f"{count} proofs per step, one per host thread, {args.steps} steps"},
            "e2e": {"value": value, "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- main arm
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

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
    # host threads: the box's cores are shared by the ranks of this node (one process per GPU)
    host_threads = max(16, (os.cpu_count() or 8) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", world))))  # >= 2 per lane
    bp = BatchProver(eng, ell, crs, max_batch=B, lanes=args.lanes, host_threads=host_threads)

    import ctypes
    cat = lambda key: b"".join(i[key] for i in insts)  # noqa: E731
    arr = lambda b: (ctypes.c_uint8 * len(b)).from_buffer_copy(b)  # noqa: E731
    R, S, T, U, M, K, MB = (arr(cat(k)) for k in ("R", "S", "T", "U", "M", "k", "m_blinders"))
    perm = (ctypes.c_uint32 * (B * ell))(*[x for i in insts for x in i["perm"]])
    seeds = (ctypes.c_uint64 * B)(*range(1000 * rank, 1000 * rank + B))
    out = (ctypes.c_uint8 * (B * bp.proof_size))()
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def step(resident):
        with torch.cuda.stream(stream):
            flush_buf.zero_()  # L2 flush between steps (256 MiB > 126 MB L2)
        if resident:
            bp.prove_raw(B, None, None, None, None, None, perm, K, MB, seeds, out=out, split=False)
        else:
            bp.prove_raw(B, R, S, T, U, M, perm, K, MB, seeds, out=out, split=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(resident, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = bp.launch_count
        with torch.cuda.stream(stream):
            e0.record(stream)
        for _ in range(steps):
            step(resident)
        with torch.cuda.stream(stream):
            e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, bp.launch_count - l0

    step(False)  # stages the instance batch in HBM (and is the first warm-up step)
    for _ in range(max(0, args.warmup - 1)):
        step(True)
    # ---- timed region 1: inputs resident in HBM, per-kernel profile on
    bp.profile_reset()
    bp.profile_enable(True)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_res, launches = timed(True, args.steps)
    clocks = sampler.stop() if sampler else None
    prof = bp.profile_read()
    bp.profile_enable(False)
    timing = bp.last_timing()
    # ---- timed region 2: end to end through the host-buffer API
    ms_e2e, _ = timed(False, args.steps)
    traffic = bp.last_traffic()
    proof0 = bytes(out[:bp.proof_size])
    # ---- secondary metric: verifies/s on the proofs just produced (CurdleproofsProof::deserialize + verify through cdp_verify_batch)
    bp.close()
    VB = B
    bv = BatchVerifier(eng, ell, crs, max_batch=VB, host_threads=host_threads)
    vout = (ctypes.c_uint8 * VB)()

    def vstep():
        bv.verify_raw(VB, R, S, T, U, M, out, None, out=vout)

    vstep()
    barrier()
    ve0, ve1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ve0.record(stream)
    for _ in range(args.steps):
        vstep()
    with torch.cuda.stream(stream):
        ve1.record(stream)
    barrier()
    ms_ver = ve0.elapsed_time(ve1)
    if world > 1:
        t = torch.tensor([ms_ver], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_ver = float(t.item())
    all_ok = all(x == 1 for x in vout)

    if rank != 0:
        return
    value = world * B * args.steps / (ms_res * 1e-3)
    e2e = world * B * args.steps / (ms_e2e * 1e-3)
    # ---- roofline of the dominant kernel (by device time inside the timed region)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, peak_src = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")
    # algorithmic bytes per unit: variable-base pair = 32 B scalar + 96 B base; fixed-base pair = 32 B scalar +
    # 16 gathered 96-byte table points (c = 16); fold element = 192 B in + 96 B out; normalised point = 144 B in + 96 B out
    bytes_per_unit = {"msm_buckets": 128, "msm_fixed": 32 + 16 * 96, "smul": 288, "normalize": 240, "msm_combine": 144, "other": 144}
    # measured DRAM bytes per unit of the same kernels from the committed `ncu --set full` captures (dram__bytes_read + write per launch /
    # units of that launch): profiles/r01_ncu_fixed_msm_v1.txt (407.5 MB / 131,584 pairs), profiles/r01_ncu_msm_buckets_v3.txt (157.3 MB / 520,192 pairs)
    ncu_traffic_per_unit = {"msm_fixed": 3097.0, "msm_buckets": 302.0}
    # dominant = the throughput kernel with the most device time; the latency-bound launches (combine: 130 dependent doublings per
    # thread, normalise: one Fp inversion per thread) run a handful of warps each, and with 8 lanes in flight their summed durations
    # mostly measure waiting for SM time, not work
    dom = max(("msm_fixed", "msm_buckets", "smul"), key=lambda k: prof[k]["ms"])
    d = prof[dom]
    avg_ms = d["ms"] / max(1, d["launches"])
    alg_bytes = bytes_per_unit[dom] * d["units"] / max(1, d["launches"])
    achieved = alg_bytes / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
    # integer-pipe view: Fp multiplications are 300 IMAD.WIDE.U32 each; peak measured live with a register-only kernel
    imad_ms = min(eng.bench_kernel(0, 148 * 4, 256, 2000) for _ in range(3))
    imad_peak = 148 * 4 * 256 * 2000 * 128 / (imad_ms * 1e-3)
    total_kernel_ms = sum(v["ms"] for v in prof.values())
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": ncu_traffic_per_unit[dom] * d["units"] / max(1, d["launches"]) if dom in ncu_traffic_per_unit else None,
                "traffic_source": "ncu dram bytes per unit (profiles/r01_ncu_*) x units per launch", "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": avg_ms,
                "share_of_kernel_time": d["ms"] / total_kernel_ms if total_kernel_ms else None,
                "note": "381-bit modular arithmetic: the binding roofline is the integer multiply pipe, not HBM (see int_pipe)",
                "int_pipe": {"peak_imad_wide_per_s": imad_peak, "unit": "IMAD.WIDE.U32/s", "peak_source": "measured live (k_bench_imad)"},
                "kernel_ms": {k: v["ms"] / args.steps for k, v in prof.items()}}
    # ---- the same kernel timed ALONE (nothing else on the GPU): one fixed-base launch of the IPA-round shape for the whole batch
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
        for _ in range(3):
            lib.cdp_msm_fixed_batch_dev(h, tab.handle, d_sc, d_sg, nseg, B * (2 * n_ + 2), None, d_out)
        eng.sync(); eng.profile_reset(); eng.profile_enable(True)
        for _ in range(5):
            with torch.cuda.stream(stream):
                flush_buf.zero_()
            lib.cdp_msm_fixed_batch_dev(h, tab.handle, d_sc, d_sg, nseg, B * (2 * n_ + 2), None, d_out)
        eng.sync()
        pf = eng.profile_read()["msm_fixed"]
        eng.profile_enable(False)
        iso_ms = pf["ms"] / pf["launches"]
        pairs = B * (2 * n_ + 2)
        madds = pairs * 16 / (iso_ms * 1e-3)
        roofline["isolated"] = {"kernel": "k_fixed_msm, IPA-round shape, whole batch in one launch, nothing else running", "ms": iso_ms,
                                "pairs_per_s": pairs / (iso_ms * 1e-3), "achieved_GBps": pairs * (32 + 16 * 96) / (iso_ms * 1e-3) / 1e9,
                                "hbm_frac": pairs * (32 + 16 * 96) / (iso_ms * 1e-3) / 1e9 / hbm_peak,
                                "mixed_adds_per_s": madds, "imad_wide_per_mixed_add": 7 * 288 + 4 * 222,
                                "int_pipe_frac": madds * (7 * 288 + 4 * 222) / imad_peak}
        for dd in (d_sc, d_sg, d_out):
            lib.cdp_dev_free(h, dd)
        tab.close()
    except Exception as e:  # diagnostics only: never lose the headline line
        roofline["isolated"] = {"error": repr(e)}
    line = {"metric": f"shuffle_proofs_per_sec_ell{ell}", "value": value, "unit": "proofs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": value / README_PROOFS_PER_S if ell == 252 else None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": f"ell={ell} CurdleproofsProof::new, {B} independent proofs per step per GPU, bit-exact vs reference CPU path",
                       "ell": ell, "batch_per_gpu": B, "l2_flush": "256 MiB memset between steps", "parallelism": f"proofs sharded over {world} GPU(s), no collective",
                       "baseline": "README.md:49 560 ms/proof on i7-8550U (other hardware)", "host_threads": host_threads, "lanes": bp.lanes},
            "e2e": {"value": e2e, "unit": "proofs/s", "h2d_bytes_per_step": traffic["h2d_bytes"], "d2h_bytes_per_step": traffic["d2h_bytes"],
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "roofline": roofline, "clocks": clocks,
            "verify": {"metric": f"shuffle_verifies_per_sec_ell{ell}", "value": world * VB * args.steps / (ms_ver * 1e-3), "unit": "verifies/s",
                       "batch_per_gpu": VB, "ms_per_step": ms_ver / args.steps, "all_accepted": all_ok,
                       "note": "host buffers in (instance + serialised proofs), verdicts out; README.md:49 reference: 35 ms/verify on i7-8550U"},
            "host_breakdown_last_step_ms": timing}
    # ---- CPU baseline beside it (rank 0, N = 1 only): the oracle port on a bounded sample, and a parity check of proof 0
    if world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib
        o = oracle_lib.Oracle()
        cores = os.cpu_count() or 1
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
                                "parity_proof0_bit_exact": bool(want0 == proof0), "oracle_verifies_gpu_proof": o.verify(inst, proof0) == 1}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
