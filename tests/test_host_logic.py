"""CPU: the product's host-side Fiat-Shamir transcript (merlin / STROBE-128 / unrolled Keccak-f) and Fr arithmetic
(curdleproofs_b200/host/merlin.hpp, fr.hpp) against the oracle's C implementation on a fixed script: message lengths around
the STROBE rate (166), long squeezes, 300 `get_and_append_challenge` draws (/root/reference/src/transcript.rs:41-54), products,
an inversion.  Byte-for-byte."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_transcript_and_fr_match_oracle():
    with tempfile.TemporaryDirectory() as d:
        prod, orc = os.path.join(d, "prod"), os.path.join(d, "orc")
        subprocess.run(["g++", "-O2", "-march=x86-64-v3", "-std=c++17", "-o", prod, os.path.join(ROOT, "tests/host/host_script_product.cpp")], check=True)
        subprocess.run(["gcc", "-O2", "-o", orc, os.path.join(ROOT, "tests/host/host_script_oracle.c")], check=True)
        a = subprocess.run([prod], check=True, capture_output=True).stdout
        b = subprocess.run([orc], check=True, capture_output=True).stdout
    assert len(a.splitlines()) == 330
    assert a == b


def _build_transcript_harness(d):
    exe = os.path.join(d, "tdc")
    subprocess.run(["g++", "-O1", "-march=x86-64-v3", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests/host/transcript_dev_check.cpp")], check=True)
    return exe


def test_device_transcript_code_matches_host_transcript_on_cpu():
    """The device transcript-opening code (csrc/k_transcript.cu, compiled as plain C++) against the host merlin implementation:
    vec_a and the continued transcript agree for ell = 4 ... 252, aligned and misaligned inputs."""
    with tempfile.TemporaryDirectory() as d:
        exe = _build_transcript_harness(d)
        out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert "MISMATCH" not in out.stdout
