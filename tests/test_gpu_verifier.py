"""GPU: decompression + the batched verifier against the CPU oracle's verdicts.  Mirrors the reference's completeness and
negative tests (/root/reference/src/curdleproofs.rs:335-471, src/inner_product_argument.rs:423-446)."""
import os
import random
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
import py_ref as pr  # noqa: E402

pytestmark = pytest.mark.gpu


def test_decompress(engine, oracle):
    rnd = random.Random(11)
    pts = [pr.mul(pr.G1, rnd.randrange(1, pr.R_ORDER)) for _ in range(20)] + [pr.INF]
    comp = b"".join(pr.compress(p) for p in pts)
    aff, st = engine.decompress_batch(comp)
    assert st == [0] * len(pts)
    assert aff == b"".join(pr.affine_to_bytes(p) for p in pts)
    assert aff == oracle.decompress(comp)
    # a curve point outside the prime-order subgroup (cofactor not cleared) must be rejected with status 3
    x = 5
    while True:
        y = pr.sqrt_fp((x ** 3 + 4) % pr.P)
        if y is not None and pr.mul((x, y), pr.R_ORDER) is not pr.INF:
            break
        x += 1
    # x whose x^3 + 4 is a non-residue -> status 2
    xn = 1
    while pr.sqrt_fp((xn ** 3 + 4) % pr.P) is not None:
        xn += 1
    bad_x = bytearray(xn.to_bytes(48, "big")); bad_x[0] |= 0x80
    too_big = bytearray(pr.P.to_bytes(48, "big")); too_big[0] |= 0x80          # x = p: non-canonical
    not_compressed = bytearray(pr.compress(pts[0])); not_compressed[0] &= 0x7F
    bad_inf = bytearray(pr.compress(pr.INF)); bad_inf[47] = 1
    cases = [pr.compress((x, y)), bytes(bad_x), bytes(too_big), bytes(not_compressed), bytes(bad_inf), pr.compress(pts[3])]
    aff, st = engine.decompress_batch(b"".join(cases))
    assert st == [3, 2, 1, 1, 1, 0]
    for c, s in zip(cases, st):
        ok = True
        try:
            oracle.decompress(c)
        except ValueError:
            ok = False
        assert ok == (s == 0)


@pytest.fixture(scope="module")
def proved(engine, oracle):
    from curdleproofs_b200 import BatchProver
    ell = 28
    crs = oracle.crs_points(ell)
    insts = [oracle.random_instance(ell, crs, seed=300 + i) for i in range(4)]
    bp = BatchProver(engine, ell, crs, max_batch=4)
    proofs = bp.prove_batch(insts, [5, 6, 7, 8])
    bp.close()
    return ell, crs, insts, proofs


def test_verifier_accepts_valid(engine, oracle, proved):
    from curdleproofs_b200 import BatchVerifier
    ell, crs, insts, proofs = proved
    bv = BatchVerifier(engine, ell, crs, max_batch=8)
    assert bv.verify_batch(insts, proofs, rng_seeds=[1, 2, 3, 4]) == [1, 1, 1, 1]
    assert [oracle.verify(i, p) for i, p in zip(insts, proofs)] == [1, 1, 1, 1]
    bv.close()


def test_verifier_rejects_like_the_reference(engine, oracle, proved):
    """test_bad_shuffle_arguments (curdleproofs.rs:384-471): swapped R/S, other permutation, wrong M, other randomiser;
    plus tampered proof scalars / points and proofs attached to the wrong instance."""
    from curdleproofs_b200 import BatchVerifier
    ell, crs, insts, proofs = proved
    inst, proof = insts[0], proofs[0]
    other = insts[1]
    rnd = random.Random(3)
    perm2 = list(range(ell)); rnd.shuffle(perm2)
    permute = lambda b: b"".join(b[96 * i:96 * i + 96] for i in perm2)  # noqa: E731
    k2 = pr.fr_to_bytes(rnd.randrange(1, pr.R_ORDER))
    cases = [
        (dict(inst, R=inst["S"], S=inst["R"]), proof),                                     # swapped inputs
        (dict(inst, T=permute(inst["T"]), U=permute(inst["U"])), proof),                   # a different permutation
        (dict(inst, M=other["M"]), proof),                                                 # wrong permutation commitment
        (dict(inst, T=oracle.scalar_mul_batch(inst["T"], k2 * ell), U=oracle.scalar_mul_batch(inst["U"], k2 * ell)), proof),  # other randomiser
        (other, proof),                                                                    # proof of another instance
        (inst, proof[:-32] + pr.fr_to_bytes((int.from_bytes(proof[-32:], "little") + 1) % pr.R_ORDER)),   # x_final + 1
        (inst, proof[:48 * 9] + pr.fr_to_bytes((int.from_bytes(proof[48 * 9:48 * 9 + 32], "little") + 1) % pr.R_ORDER) + proof[48 * 9 + 32:]),  # r_p + 1
        (inst, proofs[1][:48] + proof[48:]),                                               # A replaced by a valid but wrong point
        (dict(inst, T=bytes(96) + inst["T"][96:]), proof),                                 # vec_T[0] = infinity (curdleproofs.rs:218)
        (inst, proof),                                                                     # control: valid
    ]
    bv = BatchVerifier(engine, ell, crs, max_batch=len(cases), lanes=2)
    got = bv.verify_batch([c[0] for c in cases], [c[1] for c in cases])
    want = [oracle.verify(c[0], c[1]) for c in cases]
    assert want == [0] * 9 + [1]
    assert got == want
    # a proof point that is not a valid encoding: deserialisation error (2 here, -1 in the oracle)
    broken = bytearray(proof); broken[48 * 5 + 20] ^= 0x55
    assert oracle.verify(inst, bytes(broken)) in (-1, 0)
    assert bv.verify_batch([inst], [bytes(broken)])[0] in (0, 2)
    bv.close()


@pytest.mark.parametrize("exact", [0, 1])
def test_verifier_same_scalar_checks(engine, oracle, proved, exact, monkeypatch):
    """The SameScalar equalities (same_scalar_argument.rs:127-136): tampered z_k / z_t / z_u and swapped commitment points must be
    rejected, both when the equalities join the accumulated check (default) and as four exact MSMs (CDP_VERIFY_EXACT_EQ=1)."""
    from curdleproofs_b200 import BatchVerifier
    monkeypatch.setenv("CDP_VERIFY_EXACT_EQ", str(exact))
    ell, crs, insts, proofs = proved
    inst, proof = insts[0], proofs[0]
    m = 5
    o_pts = 9 * 48 + 32 + (2 + 4 * m) * 48 + 64        # A1 | A2 | B1 | B2
    o_z = o_pts + 4 * 48                                # z_k | z_t | z_u
    bump = lambda b, o: b[:o] + pr.fr_to_bytes((int.from_bytes(b[o:o + 32], "little") + 1) % pr.R_ORDER) + b[o + 32:]  # noqa: E731
    swap = lambda b, o: b[:o] + proofs[1][o:o + 48] + b[o + 48:]  # noqa: E731
    cases = [bump(proof, o_z), bump(proof, o_z + 32), bump(proof, o_z + 64)] + [swap(proof, o_pts + 48 * i) for i in range(4)] + [proof]
    bv = BatchVerifier(engine, ell, crs, max_batch=len(cases))
    got = bv.verify_batch([inst] * len(cases), cases)
    want = [oracle.verify(inst, c) for c in cases]
    assert want == [0] * 7 + [1]
    assert got == want
    bv.close()


def test_verifier_chunked_accumulated_msm(engine, oracle):
    """ell = 508: 5*ell + 8 + proof points > 2048 bases, so the accumulated MSM runs in chunks (the ell = 1020 config path)."""
    from curdleproofs_b200 import BatchProver, BatchVerifier
    ell = 508
    crs = oracle.crs_points(ell)
    inst = oracle.random_instance(ell, crs, seed=900, threads=8)
    bp = BatchProver(engine, ell, crs, max_batch=2)
    proofs = bp.prove_batch([inst, inst], [1, 2])
    bp.close()
    bv = BatchVerifier(engine, ell, crs, max_batch=2)
    bad = dict(inst, R=inst["S"], S=inst["R"])
    assert bv.verify_batch([inst, bad], proofs) == [1, 0]
    assert oracle.verify(inst, proofs[0], threads=8) == 1
    bv.close()


def test_verifier_merged_check_and_fallback(engine, oracle, proved, monkeypatch):
    """The merged accumulated check of a lane's sub-batch (one large MSM over the per-proof bases of all its proofs + the summed CRS
    parts; SURVEY.md 8(f) rank 3, BASELINE config 3) accepts a batch of valid proofs in one go, and a batch with a bad proof falls back
    to proof-by-proof checks with exactly the oracle's verdicts.  64 proofs x 191 slots is past the large-Pippenger threshold (2^13)."""
    from curdleproofs_b200 import BatchVerifier
    ell, crs, insts, proofs = proved
    B = 64
    bi, bp_ = [insts[i % 4] for i in range(B)], [proofs[i % 4] for i in range(B)]
    bv = BatchVerifier(engine, ell, crs, max_batch=B, lanes=1)
    assert bv.verify_batch(bi, bp_, rng_seeds=list(range(100, 100 + B))) == [1] * B
    assert bv.merge_stats() == {"merged": 1, "fallback": 0}
    # one proof attached to the wrong instance, one with a tampered scalar: the merged check must fail and the fallback must find exactly them
    bad_i, bad_p = list(bi), list(bp_)
    bad_i[17] = insts[(17 + 1) % 4]
    bad_p[40] = bad_p[40][:-32] + pr.fr_to_bytes((int.from_bytes(bad_p[40][-32:], "little") + 1) % pr.R_ORDER)
    want = [1] * B
    want[17] = want[40] = 0
    assert [oracle.verify(bad_i[k], bad_p[k]) for k in (16, 17, 40)] == [1, 0, 0]
    assert bv.verify_batch(bad_i, bad_p) == want
    assert bv.merge_stats() == {"merged": 1, "fallback": 1}
    # a malformed proof keeps the whole sub-batch off the merged path (its points must not enter the sum)
    broken = bytearray(bp_[5]); broken[48 * 5 + 20] ^= 0x55
    mal_p = list(bp_); mal_p[5] = bytes(broken)
    got = bv.verify_batch(bi, mal_p)
    assert got[5] in (0, 2) and got[:5] + got[6:] == [1] * (B - 1)
    assert bv.merge_stats() == {"merged": 1, "fallback": 1}
    # two valid proofs whose errors would cancel in a naive (unweighted) sum cannot exist here, but two INVALID ones must not cancel either:
    # swap the proofs of two different instances
    sw_p = list(bp_); sw_p[0], sw_p[1] = bp_[1], bp_[0]
    got = bv.verify_batch(bi, sw_p)
    assert got[:2] == [0, 0] and got[2:] == [1] * (B - 2)
    bv.close()
    monkeypatch.setenv("CDP_VERIFY_MERGE", "0")
    bv = BatchVerifier(engine, ell, crs, max_batch=B, lanes=2)
    assert bv.verify_batch(bad_i, bad_p) == want
    assert bv.verify_batch(bi, bp_) == [1] * B
    assert bv.merge_stats() == {"merged": 0, "fallback": 0}
    bv.close()


def test_verifier_host_transcript_path_agrees(engine, oracle, proved, monkeypatch):
    """CDP_VERIFY_HOST_TRANSCRIPT=1 keeps the per-round transcript and the challenges on the host (the older path): same verdicts as the
    default device-side transcript on valid, tampered and malformed proofs."""
    from curdleproofs_b200 import BatchVerifier
    ell, crs, insts, proofs = proved
    bad = proofs[1][:-32] + pr.fr_to_bytes((int.from_bytes(proofs[1][-32:], "little") + 1) % pr.R_ORDER)
    too_big = proofs[2][:-32] + (pr.R_ORDER + 5).to_bytes(32, "little")          # x_final >= r: does not deserialise
    broken = bytearray(proofs[3]); broken[48 * 5 + 20] ^= 0x55
    inf_T = dict(insts[0], T=bytes(96) + insts[0]["T"][96:])
    cases_i = [insts[0], insts[1], insts[2], insts[3], inf_T, insts[1]]
    cases_p = [proofs[0], bad, too_big, bytes(broken), proofs[0], proofs[1]]
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("CDP_VERIFY_HOST_TRANSCRIPT", mode)
        bv = BatchVerifier(engine, ell, crs, max_batch=8, lanes=1)
        res[mode] = bv.verify_batch(cases_i, cases_p, rng_seeds=[11, 12, 13, 14, 15, 16])
        assert bv.verify_batch(insts, proofs) == [1, 1, 1, 1]
        bv.close()
    assert res["0"] == res["1"]
    assert res["0"][0] == 1 and res["0"][1] == 0 and res["0"][2] == 2 and res["0"][3] in (0, 2) and res["0"][4] == 0 and res["0"][5] == 1
