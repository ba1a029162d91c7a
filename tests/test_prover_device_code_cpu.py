"""CPU: the device-side prover (curdleproofs_b200/csrc/k_prove.cu -- the Fiat-Shamir transcript and the Fr algebra of every step of
`CurdleproofsProof::new`, /root/reference/src/curdleproofs.rs:59-184) and the host driver that sequences it
(curdleproofs_b200/host/prover.cpp), without a GPU: the kernel sources are compiled as plain C++ (one emulated CTA per proof) and the
group kernels are replaced by the CPU oracle evaluating the same segment descriptors (tests/host/cpu_engine_mock.cpp -- test
infrastructure, never part of the product).  Whole proofs are compared with the oracle's prover byte for byte."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    d = tmp_path_factory.mktemp("pdc")
    exe = str(d / "prove_dev_check")
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    subprocess.run(["g++", "-O1", "-march=x86-64-v3", "-std=c++17", "-pthread", "-o", exe,
                    os.path.join(ROOT, "tests/host/prove_dev_check.cpp"), os.path.join(ROOT, "tests/host/cpu_engine_mock.cpp"),
                    os.path.join(ROOT, "curdleproofs_b200/host/prover.cpp"),
                    "-L" + os.path.join(ROOT, "oracle"), "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle")], check=True)
    return exe


@pytest.mark.parametrize("ell,batch,lanes", [(4, 3, 1), (12, 4, 2), (28, 2, 1), (60, 2, 2)])
def test_device_prover_code_reproduces_oracle_proofs(harness, ell, batch, lanes):
    """Random instances; the device path (seed and 32-byte-key forms of the rng) and the older host-transcript path all give the
    oracle's proof bytes."""
    out = subprocess.run([harness, "random", str(ell), str(batch), str(lanes)], capture_output=True, text=True)
    assert out.returncode == 0 and " ok " in out.stdout and "MISMATCH" not in out.stdout, out.stdout + out.stderr


@pytest.mark.parametrize("ell,switch_len", [(12, 4), (28, 2), (60, 32), (60, 16), (60, 8)])
def test_device_prover_switch_round_variants(harness, ell, switch_len):
    """The round at which the prover stops writing the round MSMs over the original CRS bases and materialises the folded bases instead
    (cdp_prove_dev::switch_round; CDP_PROVE_SWITCH_LEN, off by default: it measured slower on the B200) must not change a single proof byte."""
    env = dict(os.environ, CDP_PROVE_SWITCH_LEN=str(switch_len))
    out = subprocess.run([harness, "random", str(ell), "2", "1"], capture_output=True, text=True, env=env)
    assert out.returncode == 0 and " ok " in out.stdout and "MISMATCH" not in out.stdout, out.stdout + out.stderr


def test_device_prover_code_reproduces_reference_golden_proof(harness):
    """The seed-0 whisk shuffle of the reference's own test (src/whisk.rs:416-456): the CurdleproofsProof part of the 4496-byte golden
    vector (src/whisk.rs:455; tests/test_oracle_golden.py pins the oracle's copy against the committed hex)."""
    out = subprocess.run([harness, "golden"], capture_output=True, text=True)
    assert out.returncode == 0 and "golden ok" in out.stdout, out.stdout + out.stderr


def test_prover_rejects_malformed_witnesses(harness):
    """cdp_prove_batch validates the witnesses before any work (the reference panics on an out-of-range permutation index)."""
    out = subprocess.run([harness, "badinput"], capture_output=True, text=True)
    assert out.returncode == 0 and "badinput ok" in out.stdout, out.stdout + out.stderr


def test_device_fr_inversion_euclid_matches_fermat(harness):
    """csrc/fr256.cuh fr_inverse_euclid (one thread inverts each round challenge on the prover's critical path) against the Fermat ladder."""
    out = subprocess.run([harness, "inverse"], capture_output=True, text=True)
    assert out.returncode == 0 and "inverse ok" in out.stdout, out.stdout + out.stderr
