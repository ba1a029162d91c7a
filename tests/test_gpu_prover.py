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
    bp = BatchProver(engine, ell, crs, max_batch=batch)
    assert bp.lanes >= 2
    proofs = bp.prove_batch(insts, seeds)
    bp.close()
    for i in (0, batch - 1):
        assert proofs[i] == oracle.prove(insts[i], rng_seed=seeds[i], threads=8)
        assert oracle.verify(insts[i], proofs[i], threads=8) == 1
    assert len(set(proofs)) == batch  # different RNG streams: all proofs distinct
    bv = BatchVerifier(engine, ell, crs, max_batch=batch)
    assert bv.verify_batch(insts, proofs) == [1] * batch
    shifted = insts[1:] + insts[:1]   # every proof now sits on a different instance
    assert bv.verify_batch(shifted, proofs) == [0] * batch
    bv.close()
