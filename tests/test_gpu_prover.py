"""GPU: the batched B200 prover against the CPU oracle (byte-identical proofs for the same inputs and RNG stream) and
against the reference's own golden vector (/root/reference/src/whisk.rs:455)."""
import os

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("ell,batch", [(4, 3), (28, 5), (60, 2)])
def test_batch_prover_matches_oracle(engine, oracle, ell, batch):
    from curdleproofs_b200 import BatchProver
    crs = oracle.crs_points(ell)
    insts = [oracle.random_instance(ell, crs, seed=100 + i) for i in range(batch)]
    seeds = [1000 + i for i in range(batch)]
    bp = BatchProver(engine, ell, crs, max_batch=batch + 1)
    proofs = bp.prove_batch(insts, seeds)
    for inst, seed, proof in zip(insts, seeds, proofs):
        want = oracle.prove(inst, rng_seed=seed, threads=4)
        assert proof == want
        assert oracle.verify(inst, proof) == 1
    # a second batch on the same prover object (state reuse), smaller than max_batch
    proofs2 = bp.prove_batch(insts[:1], [77])
    assert proofs2[0] == oracle.prove(insts[0], rng_seed=77, threads=4)
    bp.close()


def test_whisk_golden_vector_through_gpu(engine, oracle):
    """The seed-0 whisk shuffle proof of the reference's own test, produced by the B200 prover: M || proof must equal the
    4496-byte hex string at /root/reference/src/whisk.rs:455."""
    from curdleproofs_b200 import BatchProver
    golden = bytes.fromhex(open(os.path.join(HERE, "golden", "whisk_shuffle_proof_seed0.hex")).read().strip())
    ell = 124
    ref_proof, ok, inst = oracle.whisk_shuffle_proof_seed0(ell, threads=8, want_instance=True)
    assert ok and ref_proof == golden
    bp = BatchProver(engine, ell, inst["crs"], max_batch=2)
    proofs = bp.prove_batch([inst, inst], [0, 0], rng_skip_words=[inst["rng_words"]] * 2)
    m_comp = engine.compress_batch(inst["M"])
    assert m_comp + proofs[0] == golden
    assert proofs[1] == proofs[0]
    bp.close()


def test_prover_rejects_bad_sizes(engine, oracle):
    from curdleproofs_b200 import BatchProver, CdpError
    crs = oracle.crs_points(4)
    with pytest.raises(CdpError):
        BatchProver(engine, 5, crs + bytes(96), max_batch=1)  # ell + 4 not a power of two (inner_product_argument.rs:116)


def test_full_size_ell252_multi_lane_round_trip(engine, oracle):
    """BASELINE config 2 at its full size: ell = 252 (n = 256), a batch large enough for several prover / verifier lanes.
    Two proofs are compared byte for byte with the oracle; the whole batch goes through prove -> verify (size-independent
    property: everything the prover emits is accepted, the same proofs attached to other instances are rejected)."""
    from curdleproofs_b200 import BatchProver, BatchVerifier
    ell, batch = 252, 72
    crs = oracle.crs_points(ell)
    base = [oracle.random_instance(ell, crs, seed=500 + i, threads=8) for i in range(3)]
    insts = [base[i % 3] for i in range(batch)]
    seeds = [9000 + i for i in range(batch)]
    bp = BatchProver(engine, ell, crs, max_batch=batch, lanes=4)
    assert bp.lanes == 4
    proofs = bp.prove_batch(insts, seeds)
    bp.close()
    for i in (0, batch - 1):
        assert proofs[i] == oracle.prove(insts[i], rng_seed=seeds[i], threads=8)
        assert oracle.verify(insts[i], proofs[i], threads=8) == 1
    assert len(set(proofs)) == batch  # different RNG streams: all proofs distinct
    bv = BatchVerifier(engine, ell, crs, max_batch=batch, lanes=3)
    assert bv.verify_batch(insts, proofs) == [1] * batch
    shifted = insts[1:] + insts[:1]   # every proof now sits on a different instance
    assert bv.verify_batch(shifted, proofs) == [0] * batch
    bv.close()


def test_ell252_sixteen_proofs_across_all_lanes_byte_identical(engine, oracle):
    """BASELINE config 2 (ell = 252): 64 proofs over 8 lanes; the first and the last proof of EVERY lane's sub-batch (16 proofs) are
    compared byte for byte with the oracle's `CurdleproofsProof::new` (/root/reference/src/curdleproofs.rs:59-184), half of them with the
    rng given as a 32-byte key instead of the u64 test seed."""
    from curdleproofs_b200 import BatchProver
    ell, batch, lanes = 252, 64, 8
    crs = oracle.crs_points(ell)
    base = [oracle.random_instance(ell, crs, seed=700 + i, threads=8) for i in range(4)]
    insts = [base[(3 * i) % 4] for i in range(batch)]
    seeds = [31000 + 7 * i for i in range(batch)]
    bp = BatchProver(engine, ell, crs, max_batch=batch, lanes=lanes)
    assert bp.lanes == lanes
    proofs = bp.prove_batch(insts, seeds)
    keys = [oracle.stdrng_seed_bytes(s) for s in seeds]
    proofs_k = bp.prove_batch(insts, rng_keys=keys)
    bp.close()
    per = batch // lanes
    picked = [l * per for l in range(lanes)] + [l * per + per - 1 for l in range(lanes)]
    assert len(picked) == 16
    for j, i in enumerate(picked):
        want = oracle.prove(insts[i], rng_seed=seeds[i], threads=8)
        assert (proofs if j % 2 == 0 else proofs_k)[i] == want, f"proof {i} (lane {i // per}) differs from the oracle"
    assert proofs == proofs_k


def test_ell1020_proof_and_verdict_parity(engine, oracle):
    """BASELINE config 3 (ell = 1020, n = 1024): one proof byte-identical to the oracle's, accepted by the oracle's verifier and by the
    batched GPU verifier; a proof attached to a permuted instance is rejected by both."""
    from curdleproofs_b200 import BatchProver, BatchVerifier
    ell = 1020
    crs = oracle.crs_points(ell)
    inst = oracle.random_instance(ell, crs, seed=4242, threads=8)
    bp = BatchProver(engine, ell, crs, max_batch=2)
    proofs = bp.prove_batch([inst, inst], [5, 6])
    bp.close()
    assert proofs[0] == oracle.prove(inst, rng_seed=5, threads=8)
    assert oracle.verify(inst, proofs[1], threads=8) == 1
    bad = dict(inst, T=inst["U"], U=inst["T"])
    bv = BatchVerifier(engine, ell, crs, max_batch=2)
    assert bv.verify_batch([inst, bad], proofs) == [1, 0]
    assert oracle.verify(bad, proofs[1], threads=8) == 0
    bv.close()


def _chacha12_words(key: bytes, first_word: int, count: int):
    """Words [first_word, first_word + count) of the ChaCha12 keystream of rand 0.8's StdRng (64-bit block counter, stream 0)."""
    import struct
    k = struct.unpack("<8I", key)
    M = 0xFFFFFFFF

    def rotl(v, n):
        return ((v << n) | (v >> (32 - n))) & M

    def block(ctr):
        s = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574, *k, ctr & M, ctr >> 32, 0, 0]
        x = list(s)

        def qr(a, b, c, d):
            x[a] = (x[a] + x[b]) & M; x[d] = rotl(x[d] ^ x[a], 16)
            x[c] = (x[c] + x[d]) & M; x[b] = rotl(x[b] ^ x[c], 12)
            x[a] = (x[a] + x[b]) & M; x[d] = rotl(x[d] ^ x[a], 8)
            x[c] = (x[c] + x[d]) & M; x[b] = rotl(x[b] ^ x[c], 7)
        for _ in range(6):
            qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
            qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
        return [(a + b) & M for a, b in zip(x, s)]
    out = []
    b = first_word // 16
    while len(out) < first_word % 16 + count:
        out += block(b)
        b += 1
    return out[first_word % 16:first_word % 16 + count]


@pytest.mark.parametrize("ell", [4, 28])
def test_prover_randomness_on_the_device(engine, oracle, ell):
    """cdp_prove_random_dev against a plain-Python restatement of `StdRng` (ChaCha12) + ark-ff `Fr::rand` (eight consecutive stream words, top
    bit cleared, rejected when >= r, taken as they are): keys from the u64 test seeds and raw 32-byte keys, stream positions that are aligned,
    odd, and straddle a ChaCha block -- /root/reference/src/curdleproofs.rs:74 takes `rng: &mut impl RngCore` at whatever position the caller left it."""
    import struct
    R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
    n = ell + 4
    nrnd = 3 * n + 11
    keys = [oracle.stdrng_seed_bytes(7), oracle.stdrng_seed_bytes(2 ** 63 + 5), bytes(range(32)), bytes(32), b"\xff" * 32]
    skips = [0, 3, 13, 16 * 5 + 9, 2 ** 33 + 7]
    got = engine.prove_random(b"".join(keys), ell, skips)
    got0 = engine.prove_random(b"".join(keys), ell)      # NULL positions = fresh generators
    for i, (key, skip) in enumerate(zip(keys, skips)):
        for blob, sk in ((got, skip), (got0, 0)):
            words = _chacha12_words(key, sk, 8 * (nrnd + 200))
            draws, a = [], 0
            while len(draws) < 3 * n + 9:
                w = words[8 * a:8 * a + 8]
                a += 1
                v = sum(x << (32 * j) for j, x in enumerate(w)) & ((1 << 255) - 1)
                if v < R_MOD:
                    draws.append(v)
            want = draws[:6 + 2 * n - 2] + [0, 0] + draws[6 + 2 * n - 2:]
            have = [int.from_bytes(blob[32 * (i * nrnd + j):32 * (i * nrnd + j + 1)], "little") for j in range(nrnd)]
            assert have == want, (i, sk)
