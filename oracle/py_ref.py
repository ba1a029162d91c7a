"""Pure-Python big-int BLS12-381 G1 -- the *mathematical* oracle.

TEST INFRASTRUCTURE ONLY.  Nothing under ``curdleproofs_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` leg may.

This file follows no arkworks code: it restates the *group law* that every
observable value on the reference's MSM / fold path is defined by
(`/root/reference/src/util.rs:19-29` -> ark-ec ``VariableBaseMSM``; every value
that leaves that boundary is a canonical group element).
It is deliberately naive (affine formulas, ``pow(x, -1, p)``) so that it shares
no structure -- and hence no bugs -- with either the C oracle (``oracle/*.c``)
or the CUDA kernels.

Byte layouts used by the helper converters are those of ``include/cdp_msm.h``:
  * Fp  : 48 bytes, 6 x u64 little-endian limbs, **Montgomery form** (R = 2^384)
  * Fr  : 32 bytes, 4 x u64 little-endian limbs, canonical integer
  * affine  : x || y (96 B), infinity = all zero bytes
  * jacobian: X || Y || Z (144 B), infinity <=> Z == 0
  * compressed: 48 B big-endian x with ZCash flag bits
    (pinned by the KAT at `/root/reference/src/whisk.rs:363-368`).
"""
from __future__ import annotations

P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
R_ORDER = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
B_COEFF = 4
COFACTOR = 0x396C8C005555E1568C00AAAB0000AAAB
GX = 0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB
GY = 0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1
MONT_R = 1 << 384          # Fp Montgomery radix
MONT_R_INV = pow(MONT_R, -1, P)
FR_MONT_R = 1 << 256
FR_MONT_R_INV = pow(FR_MONT_R, -1, R_ORDER)

INF = None                 # point at infinity
G1 = (GX, GY)


def is_on_curve(pt) -> bool:
    if pt is INF:
        return True
    x, y = pt
    return (y * y - x * x * x - B_COEFF) % P == 0


def neg(pt):
    if pt is INF:
        return INF
    return (pt[0], (-pt[1]) % P)


def add(a, b):
    if a is INF:
        return b
    if b is INF:
        return a
    x1, y1 = a
    x2, y2 = b
    if x1 == x2:
        if (y1 + y2) % P == 0:
            return INF
        lam = 3 * x1 * x1 * pow(2 * y1, -1, P) % P
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, P) % P
    x3 = (lam * lam - x1 - x2) % P
    y3 = (lam * (x1 - x3) - y1) % P
    return (x3, y3)


def mul(pt, k: int):
    """Plain double-and-add; k is any non-negative integer (not reduced)."""
    acc = INF
    addend = pt
    while k:
        if k & 1:
            acc = add(acc, addend)
        addend = add(addend, addend)
        k >>= 1
    return acc


def msm(points, scalars):
    """sum_i scalars[i] * points[i]  (what `util::msm` returns, src/util.rs:19-22)."""
    assert len(points) == len(scalars)
    acc = INF
    for pt, s in zip(points, scalars):
        acc = add(acc, mul(pt, s % R_ORDER))
    return acc


def sqrt_fp(a: int):
    """p = 3 mod 4  =>  sqrt = a^((p+1)/4) when it exists."""
    s = pow(a, (P + 1) // 4, P)
    return s if s * s % P == a % P else None


# --------------------------------------------------------------------------- encodings
def compress(pt) -> bytes:
    if pt is INF:
        return bytes([0xC0]) + bytes(47)
    x, y = pt
    out = bytearray(x.to_bytes(48, "big"))
    out[0] |= 0x80
    if y > P - y:
        out[0] |= 0x20
    return bytes(out)


def decompress(buf: bytes, check_subgroup: bool = True):
    assert len(buf) == 48
    flags = buf[0] >> 5
    if not flags & 0b100:
        raise ValueError("uncompressed encoding not supported")
    if flags & 0b010:
        if any(buf[1:]) or buf[0] & 0x3F:
            raise ValueError("non-canonical infinity")
        return INF
    x = int.from_bytes(bytes([buf[0] & 0x1F]) + buf[1:], "big")
    if x >= P:
        raise ValueError("x >= p")
    y = sqrt_fp((x * x * x + B_COEFF) % P)
    if y is None:
        raise ValueError("not on curve")
    if (y > P - y) != bool(flags & 0b001):
        y = P - y
    pt = (x, y)
    if check_subgroup and mul(pt, R_ORDER) is not INF:
        raise ValueError("not in subgroup")
    return pt


def fp_to_mont_bytes(a: int) -> bytes:
    return (a * MONT_R % P).to_bytes(48, "little")


def fp_from_mont_bytes(b: bytes) -> int:
    return int.from_bytes(b, "little") * MONT_R_INV % P


def affine_to_bytes(pt) -> bytes:
    if pt is INF:
        return bytes(96)
    return fp_to_mont_bytes(pt[0]) + fp_to_mont_bytes(pt[1])


def affine_from_bytes(b: bytes):
    assert len(b) == 96
    if not any(b):
        return INF
    return (fp_from_mont_bytes(b[:48]), fp_from_mont_bytes(b[48:]))


def jacobian_from_bytes(b: bytes):
    assert len(b) == 144
    X, Y, Z = (fp_from_mont_bytes(b[i * 48:(i + 1) * 48]) for i in range(3))
    if Z == 0:
        return INF
    zi = pow(Z, -1, P)
    return (X * zi * zi % P, Y * zi * zi * zi % P)


def jacobian_to_bytes(pt, z: int = 1) -> bytes:
    """Encode with an arbitrary non-zero Z (tests use z != 1 to exercise projective inputs)."""
    if pt is INF:
        return fp_to_mont_bytes(1) + fp_to_mont_bytes(1) + bytes(48)
    x, y = pt
    return fp_to_mont_bytes(x * z * z % P) + fp_to_mont_bytes(y * z * z * z % P) + fp_to_mont_bytes(z % P)


def fr_to_bytes(s: int) -> bytes:
    return (s % R_ORDER).to_bytes(32, "little")


def fr_from_bytes(b: bytes) -> int:
    return int.from_bytes(b, "little")


if __name__ == "__main__":  # tiny self-check
    assert is_on_curve(G1)
    assert mul(G1, R_ORDER) is INF
    assert compress(G1).hex().startswith("97f1d3a7")
    assert decompress(compress(G1)) == G1
    print("py_ref ok")
