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


def test_generated_squaring_schedule_is_model_checked_and_current():
    """curdleproofs_b200/csrc/fp_sqr_rows.inc is generated: tools/gen_fp_sqr.py runs the exact instruction sequence of the row-wise
    Montgomery squaring on a 32-bit register / carry-flag model against big integers (any dropped carry asserts) before it writes the
    file.  Re-run the generator and require the committed file to be what it produces."""
    path = os.path.join(ROOT, "curdleproofs_b200", "csrc", "fp_sqr_rows.inc")
    before = open(path).read()
    out = subprocess.run(["python", os.path.join(ROOT, "tools", "gen_fp_sqr.py")], check=True, capture_output=True, text=True).stdout
    assert "model ok" in out and "222 wide multiply-adds" in out
    assert open(path).read() == before


def test_crs_text_format_errors_need_no_device():
    """`from_hex_g1affine` (/root/reference/src/crs.rs:128-139): prefix, hex digits and length are checked before any point decoding;
    `from_points` refuses too few points (src/crs.rs:40-42)."""
    import pytest
    from curdleproofs_b200.crs import CrsError, CurdleproofsCrs
    for bad in ("97f1" * 24, "0x" + "zz" * 48, "0x" + "00" * 47, "0x" + "00" * 49, 7, None):
        with pytest.raises(CrsError):
            CurdleproofsCrs._parse_hex_point(bad)
    assert CurdleproofsCrs._parse_hex_point("0x" + "c0" + "00" * 47) == bytes([0xC0]) + bytes(47)
    with pytest.raises(CrsError):
        CurdleproofsCrs.from_points(None, 4, bytes(96 * 10))  # needs 4 + 4 + 3 points
    with pytest.raises(CrsError):
        CurdleproofsCrs.from_json(None, "[1, 2]")
    with pytest.raises(CrsError):
        CurdleproofsCrs.from_hex(None, {"vec_G": []})


def test_device_verifier_scalar_code_matches_host_restatement_on_cpu():
    """The device-side verifier scalar preparation (csrc/k_vcoeffs.cu, compiled as plain C++): every coefficient of the accumulated check
    -- verification scalars s_i / 1/s_i (/root/reference/src/util.rs:40-64), beta^-(i+1) rescaling, the `a * x_i` products of the eight
    accumulate_check calls (src/msm_accumulator.rs:37-52) and the four SameScalar equalities -- against a vector-at-a-time host restatement
    with the product's host Fr, on random challenges, ell = 4 / 12 / 124 / 252, both SameScalar modes, three thread counts.  The same
    harness checks the prover's device-side round expansion (k_round_expand: prefix weights x folded vector) against the per-index host loops
    it replaces (/root/reference/src/inner_product_argument.rs:158-179, src/same_multiscalar_argument.rs:107-131)."""
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "vcc")
        subprocess.run(["g++", "-O1", "-march=x86-64-v3", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests/host/vcoeffs_check.cpp")], check=True)
        out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert out.stdout.count(" ok ") == 8 + 3  # 8 coefficient cases + the prover's round expansion for n = 8 / 16 / 256 (all rounds, both arguments)


def test_device_verifier_transcript_code_matches_host_transcript_on_cpu():
    """The device verifier-transcript kernels (k_verify_transcript_a / _b in csrc/k_transcript.cu, compiled as plain C++) against the
    product's host merlin / Fr following `CurdleproofsProof::verify` (/root/reference/src/curdleproofs.rs:226-296): every challenge, the
    inverted round challenges, z, the stage scalars, the identity flag of vec_T[0] and the final STROBE state; ell = 12 / 124 / 252."""
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "vtc")
        subprocess.run(["g++", "-O1", "-march=x86-64-v3", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests/host/vtranscript_check.cpp")], check=True)
        out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert out.stdout.count(" ok ") == 3


def test_device_fp_inversion_on_cpu():
    """csrc/fp_inv_safegcd.cuh (Bernstein-Yang division steps: the inversion of `into_affine` / `normalize_batch` and of the batched affine
    additions on the device) and csrc/fp_inv_euclid.cuh (binary Euclid), both plain C++, against each other and the oracle's field:
    a * inverse(a) == 1 mod p for edge values and 20 000 pseudo-random ones."""
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "fpinv")
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
        subprocess.run(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests/host/fp_inv_check.cpp"), "-L" + os.path.join(ROOT, "oracle"),
                        "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle")], check=True)
        out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "fp inverse ok" in out.stdout, out.stdout


def test_device_glv_split_on_cpu():
    """csrc/glv_split.cuh (plain C++; Barrett division by lambda in front of every device MSM / scalar multiplication): k = k2 lambda + k1 with
    0 <= k1 < lambda, checked by multiplying back, for edge values and 180 000 pseudo-random scalars below r."""
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "glv")
        subprocess.run(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests/host/glv_check.cpp")], check=True)
        out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "glv split ok" in out.stdout, out.stdout
