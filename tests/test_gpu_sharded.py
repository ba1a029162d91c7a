"""GPU: one large MSM sharded by base range across the ranks of a node through the engine's own communicator (`cdp_comm_*`,
`cdp_msm_sharded_dev`, include/cdp_msm.h; SURVEY.md 8(e), `util::msm` /root/reference/src/util.rs:19-22): every rank's result must be the
oracle's MSM over ALL bases.  World size 1 runs on any box (NCCL with one rank); the two-rank forms (one process per GPU, and one
process driving two GPUs) need a 2-GPU lease and are skipped with that reason otherwise."""
import os
import random
import socket
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

pytestmark = pytest.mark.gpu


def _inputs(n, seed=5):
    import oracle_lib
    import py_ref as pr
    o = oracle_lib.Oracle()
    rnd = random.Random(seed)
    k = min(n, 512)  # distinct base points; repeated beyond that (the MSM does not care)
    base = o.scalar_mul_batch(o.generator() * k, b"".join(pr.fr_to_bytes(rnd.randrange(1, pr.R_ORDER)) for _ in range(k)))
    pts = (base * ((n + k - 1) // k))[:96 * n]
    sc = b"".join(pr.fr_to_bytes(rnd.randrange(pr.R_ORDER)) for _ in range(n))
    return o, pts, sc


def _gpu_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("n", [1, 300, 20000])
def test_sharded_msm_world1(n):
    from curdleproofs_b200 import Engine
    from curdleproofs_b200.sharded import Comm
    o, pts, sc = _inputs(n)
    eng = Engine(0)
    comm = Comm(eng, Comm.unique_id(eng), 1, 0)
    got = comm.msm_sharded(pts, sc)
    assert o.compress_jac(got) == o.compress_jac(o.msm(pts, sc))
    comm.close()
    eng.close()


def _proved_batch(oracle, eng, ell, batch, seed0):
    from curdleproofs_b200 import BatchProver
    crs = oracle.crs_points(ell)
    insts = [oracle.random_instance(ell, crs, seed=seed0 + i) for i in range(batch)]
    bp = BatchProver(eng, ell, crs, max_batch=batch, lanes=2)
    proofs = bp.prove_batch(insts, [seed0 + 50 + i for i in range(batch)])
    bp.close()
    return crs, insts, proofs


def test_sharded_accumulated_verify_world1():
    """cdp_verify_batch_sharded with a one-rank communicator: the merged sums of the lanes are added, all-gathered (one rank) and decide the
    whole batch; a batch with one invalid proof is rejected by the global sum and decided locally, with the same verdicts as cdp_verify_batch."""
    import oracle_lib
    from curdleproofs_b200 import BatchVerifier, Engine
    from curdleproofs_b200.sharded import Comm
    o = oracle_lib.Oracle()
    eng = Engine(0)
    comm = Comm(eng, Comm.unique_id(eng), 1, 0)
    ell, batch = 12, 6
    crs, insts, proofs = _proved_batch(o, eng, ell, batch, 300)
    bv = BatchVerifier(eng, ell, crs, max_batch=batch, lanes=2)
    assert bv.verify_batch(insts, proofs, comm=comm) == [1] * batch
    assert bv.global_stats() == {"accepted": 1, "local": 0}
    bad = list(proofs)
    bad[4] = bad[4][:-1] + bytes([bad[4][-1] ^ 1])  # x_final
    want = [o.verify(i, p) for i, p in zip(insts, bad)]
    assert want == [1, 1, 1, 1, 0, 1]
    assert bv.verify_batch(insts, bad, comm=comm) == want
    assert bv.verify_batch(insts, bad) == want
    assert bv.global_stats() == {"accepted": 1, "local": 1}
    bv.close()
    comm.close()
    eng.close()


def test_prover_reads_pinned_inputs_directly():
    """Page-locked caller buffers (Engine.pinned_array) take the strided-DMA path of cdp_prove_batch / cdp_verify_batch: same proofs, same verdicts."""
    import ctypes

    import oracle_lib
    from curdleproofs_b200 import BatchProver, BatchVerifier, Engine
    o = oracle_lib.Oracle()
    eng = Engine(0)
    ell, batch = 12, 5
    crs = o.crs_points(ell)
    insts = [o.random_instance(ell, crs, seed=800 + i) for i in range(batch)]
    seeds = [40 + i for i in range(batch)]
    bp = BatchProver(eng, ell, crs, max_batch=batch, lanes=2)
    want = bp.prove_batch(insts, seeds)
    cat = lambda k: b"".join(i[k] for i in insts)  # noqa: E731
    arr = lambda b: (ctypes.c_uint8 * len(b)).from_buffer_copy(b)  # noqa: E731
    R, S, T, U = (eng.pinned_array(cat(k)) for k in ("R", "S", "T", "U"))
    assert eng.lib.cdp_host_is_pinned(R) == 1 and eng.lib.cdp_host_is_pinned(arr(b"x" * 64)) == 0
    perm = (ctypes.c_uint32 * (batch * ell))(*[x for i in insts for x in i["perm"]])
    got = bp.prove_raw(batch, R, S, T, U, arr(cat("M")), perm, arr(cat("k")), arr(cat("m_blinders")), (ctypes.c_uint64 * batch)(*seeds))
    assert got == want and got[0] == o.prove(insts[0], rng_seed=seeds[0])
    bp.close()
    bv = BatchVerifier(eng, ell, crs, max_batch=batch, lanes=2)
    res = bv.verify_raw(batch, R, S, T, U, arr(cat("M")), arr(b"".join(got)))
    assert list(res) == [1] * batch
    bv.close()
    eng.close()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_sharded_msm_two_ranks_one_process_per_gpu():
    """One process per GPU under torch.distributed.run (tests/mp/sharded_msm_ranks.py): n = 1 (an empty shard), uneven splits, and a size
    that takes the sort-based large path on every shard."""
    if _gpu_count() < 2:
        pytest.skip("needs a 2-GPU lease (this box has %d GPU): run under `gpurun --gpus 2`" % _gpu_count())
    import subprocess
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", str(_free_port()), os.path.join(HERE, "mp", "sharded_msm_ranks.py")],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "RANK 0 OK" in out.stdout and "RANK 1 OK" in out.stdout and "MISMATCH" not in out.stdout


def test_sharded_msm_group_one_process_two_gpus():
    if _gpu_count() < 2:
        pytest.skip("needs a 2-GPU lease (this box has %d GPU): run under `gpurun --gpus 2`" % _gpu_count())
    import ctypes
    from ctypes import c_size_t, c_void_p
    from curdleproofs_b200 import Engine
    from curdleproofs_b200.sharded import shard_range
    n, world = 30011, 2
    o, pts, sc = _inputs(n)
    engs = [Engine(r) for r in range(world)]
    lib = engs[0].lib
    comms = (c_void_p * world)()
    ctxs = (c_void_p * world)(*[e.handle for e in engs])
    assert lib.cdp_comm_create_all(comms, ctxs, world) == 0
    d_p, d_s, d_o, nl = (c_void_p * world)(), (c_void_p * world)(), (c_void_p * world)(), (c_size_t * world)()
    for r in range(world):
        lo, hi = shard_range(n, r, world)
        nl[r] = hi - lo
        h = engs[r].handle
        d_p[r], d_s[r], d_o[r] = lib.cdp_dev_alloc(h, 96 * (hi - lo)), lib.cdp_dev_alloc(h, 32 * (hi - lo)), lib.cdp_dev_alloc(h, 144)
        lib.cdp_h2d(h, d_p[r], (ctypes.c_uint8 * (96 * (hi - lo))).from_buffer_copy(pts[96 * lo:96 * hi]), 96 * (hi - lo))
        lib.cdp_h2d(h, d_s[r], (ctypes.c_uint8 * (32 * (hi - lo))).from_buffer_copy(sc[32 * lo:32 * hi]), 32 * (hi - lo))
        engs[r].sync()
    assert lib.cdp_msm_sharded_group(comms, world, d_p, d_s, nl, d_o) == 0
    want = o.compress_jac(o.msm(pts, sc))
    for r in range(world):
        out = (ctypes.c_uint8 * 144)()
        lib.cdp_d2h(engs[r].handle, out, d_o[r], 144)
        engs[r].sync()
        assert o.compress_jac(bytes(out)) == want
    for r in range(world):
        lib.cdp_comm_destroy(comms[r])
        for d in (d_p[r], d_s[r], d_o[r]):
            lib.cdp_dev_free(engs[r].handle, d)
        engs[r].close()
