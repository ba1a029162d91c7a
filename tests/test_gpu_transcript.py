"""GPU: cdp_transcript_open_dev (merlin / STROBE-128 / Keccak-f on the device, one thread per proof) against the same code run on the
CPU (tests/host/transcript_dev_check.cpp, itself checked against the host transcript and, through it, the oracle).  Byte for byte."""
import os
import subprocess
import tempfile

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def lcg_bytes(x, n):
    out = bytearray(n)
    for i in range(n):
        x = (x * 1664525 + 1013904223) & 0xFFFFFFFF
        out[i] = x >> 24
    return x, bytes(out)


@pytest.mark.parametrize("ell,B", [(4, 3), (12, 2), (28, 5), (124, 33), (252, 4)])
def test_transcript_open_matches_cpu_run(engine, ell, B):
    x = (ell * 7919 + B) & 0xFFFFFFFF
    x, vecs = lcg_bytes(x, B * 4 * ell * 48)
    x, M = lcg_bytes(x, B * 48)
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "tdc")
        subprocess.run(["g++", "-O1", "-march=x86-64-v3", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests/host/transcript_dev_check.cpp")], check=True)
        want = subprocess.run([exe, "dump", str(ell), str(B)], check=True, capture_output=True, text=True).stdout.split()
    va, st = engine.transcript_open(vecs, M, ell)
    assert va.hex() == want[0]
    assert st.hex() == want[1]
