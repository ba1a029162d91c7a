"""GPU: `CurdleproofsCrs::from_points` and the JSON/hex form (/root/reference/src/crs.rs:37-58, :77-142; its own test
`serde_crs_json`, :148-160) through the product: sums, encodings and decoding run on the device and are compared with the oracle."""
import json

import pytest

pytestmark = pytest.mark.gpu


def test_from_points_matches_oracle(engine, oracle):
    from curdleproofs_b200 import CurdleproofsCrs
    ell = 60  # the reference's serde test: generate_crs(64 - N_BLINDERS)
    pts = oracle.crs_points(ell)
    crs = CurdleproofsCrs.from_points(engine, ell, pts)
    ones = (1).to_bytes(32, "little")
    assert crs.G_sum == oracle.normalize_batch(oracle.msm(pts[:96 * ell], ones * ell))
    assert crs.H_sum == oracle.normalize_batch(oracle.msm(pts[96 * ell:96 * (ell + 4)], ones * 4))
    assert crs.points() == pts and crs.ell == ell and crs.log2_n() == 6
    # more points than needed are allowed, fewer are an error (src/crs.rs:40-42)
    assert CurdleproofsCrs.from_points(engine, ell - 1, pts).ell == ell - 1
    from curdleproofs_b200 import CrsError
    with pytest.raises(CrsError):
        CurdleproofsCrs.from_points(engine, ell + 1, pts)


def test_serde_crs_json(engine, oracle):
    from curdleproofs_b200 import CrsError, CurdleproofsCrs
    ell = 60
    pts = oracle.crs_points(ell)
    crs = CurdleproofsCrs.from_points(engine, ell, pts)
    hexed = crs.to_hex(engine)
    want = oracle.compress(pts)
    assert hexed["vec_G"] == ["0x" + want[48 * i:48 * i + 48].hex() for i in range(ell)]
    assert hexed["vec_H"] == ["0x" + want[48 * i:48 * i + 48].hex() for i in range(ell, ell + 4)]
    assert hexed["H"] == "0x" + want[48 * (ell + 4):48 * (ell + 5)].hex()
    assert hexed["G_sum"] == "0x" + oracle.compress(crs.G_sum).hex()
    text = crs.to_json(engine)
    assert json.loads(text) == hexed
    back = CurdleproofsCrs.from_json(engine, text)
    assert back == crs and back.H_sum == crs.H_sum
    # InvalidData cases of from_hex_g1affine (src/crs.rs:128-139) and of the point decoder behind it
    for bad in (hexed["H"][2:], "0x" + hexed["H"][4:], hexed["H"][:-2] + "zz", "0x" + "00" * 48, "0x" + "ff" * 48):
        d = dict(hexed); d["G_t"] = bad
        with pytest.raises(CrsError):
            CurdleproofsCrs.from_hex(engine, d)
    x_only = bytearray(bytes.fromhex(hexed["H"][2:])); x_only[47] ^= 1   # almost surely no longer a subgroup point / a square
    d = dict(hexed); d["vec_G"] = list(hexed["vec_G"]); d["vec_G"][3] = "0x" + bytes(x_only).hex()
    try:
        got = CurdleproofsCrs.from_hex(engine, d)
        oracle.decompress(bytes(x_only))  # the oracle must agree that it decodes
        assert got.vec_G[3 * 96:4 * 96] == oracle.decompress(bytes(x_only))
    except CrsError:
        with pytest.raises(ValueError):
            oracle.decompress(bytes(x_only))
    inf = CurdleproofsCrs(b"", b"", bytes(96), crs.G_t, crs.G_u, bytes(96), bytes(96))  # the identity encodes as c0 00 ..
    assert inf.to_hex(engine)["H"] == "0xc0" + "00" * 47
    assert CurdleproofsCrs.from_hex(engine, inf.to_hex(engine)) == inf
