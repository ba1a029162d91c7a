"""GPU: the Whisk byte-level API (`generate_whisk_shuffle_proof` / `is_valid_whisk_shuffle_proof`, /root/reference/src/whisk.rs:106-179)
through cdp_whisk_*: the reference's own seed-0 golden proof (src/whisk.rs:455) must come out of the product API byte for byte,
post-trackers included, and the validity check must accept it and reject tampered inputs."""
import os

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def zip_trackers(oracle, a, b, ell):
    ca, cb = oracle.compress(a), oracle.compress(b)
    return b"".join(ca[48 * i:48 * i + 48] + cb[48 * i:48 * i + 48] for i in range(ell))


def test_whisk_api_reproduces_the_golden_vector(engine, oracle):
    from curdleproofs_b200 import BatchProver, BatchVerifier
    golden = bytes.fromhex(open(os.path.join(HERE, "golden", "whisk_shuffle_proof_seed0.hex")).read().strip())
    ell = 124
    ref_proof, ok, inst = oracle.whisk_shuffle_proof_seed0(ell, threads=8, want_instance=True)
    assert ok and ref_proof == golden
    pre = zip_trackers(oracle, inst["R"], inst["S"], ell)
    want_post = zip_trackers(oracle, inst["T"], inst["U"], ell)
    bp = BatchProver(engine, ell, inst["crs"], max_batch=3)
    # the reference's test hands generate_whisk_shuffle_proof an rng that already produced the 124 trackers: same stream position here
    res = bp.whisk_generate_shuffle_proofs([pre, pre], [0, 12345], rng_skip_words=[inst["whisk_entry_words"], 0])
    bp.close()
    post0, proof0 = res[0]
    assert len(proof0) == 4496
    assert proof0 == golden
    assert post0 == want_post
    post1, proof1 = res[1]
    assert proof1 != proof0 and post1 != post0
    bv = BatchVerifier(engine, ell, inst["crs"], max_batch=8)
    swapped = post0[96:192] + post0[:96] + post0[192:]                  # two post-trackers exchanged
    tampered = proof0[:48] + proof1[48:]                                 # M of one shuffle, proof of another
    broken = bytearray(pre); broken[5] ^= 0x40                           # a tracker that does not deserialise
    got = bv.whisk_verify_shuffle_proofs([pre, pre, pre, pre, pre, bytes(broken)], [post0, post1, swapped, post0, post1, post0],
                                         [proof0, proof1, proof0, tampered, proof0, proof0])
    assert got[:5] == [1, 1, 0, 0, 0]
    assert got[5] in (0, 2)
    bv.close()


def test_whisk_tracker_proof_golden(engine, oracle):
    """`whisk_tracker_proof` (src/whisk.rs:381-402): the 128-byte seed-0 proof through the product API, and its validity check."""
    from curdleproofs_b200 import whisk_generate_tracker_proofs, whisk_verify_tracker_proofs
    golden = bytes.fromhex(open(os.path.join(HERE, "golden", "whisk_tracker_proof_seed0.hex")).read().strip())
    t = oracle.whisk_tracker_inputs_seed0()
    proofs = whisk_generate_tracker_proofs(engine, [t["tracker"]] * 2, [t["k"]] * 2, [0, 7], rng_skip_words=[t["rng_words"], 0])
    assert proofs[0] == golden
    assert proofs[1] != golden
    bad_s = proofs[0][:96] + bytes([proofs[0][96] ^ 1]) + proofs[0][97:]
    wrong_commitment = t["tracker"][:48]  # r_G instead of k_G
    got = whisk_verify_tracker_proofs(engine, [t["tracker"]] * 4, [t["k_commitment"]] * 3 + [wrong_commitment], [proofs[0], proofs[1], bad_s, proofs[0]])
    assert got == [1, 1, 0, 0]
