"""`CurdleproofsCrs` and its JSON/hex form (/root/reference/src/crs.rs:17-142) on top of the GPU engine.

The group work -- the two sums of `from_points` (`sum_affine_points`, src/crs.rs:46-47), the 48-byte encodings of `to_hex`
(`to_bytes_g1affine`) and the decoding + subgroup check of `from_hex` (`from_bytes_g1affine` = `deserialize_compressed`) -- runs
on the device through the C ABI; this file only holds the container and the text format.  Points are the ABI's affine form
(96 bytes, x || y in arkworks' Montgomery limbs, infinity = all zero).  There is no CPU path."""
from __future__ import annotations

import json

from .engine import AFFINE_BYTES, COMPRESSED_BYTES, Engine

N_BLINDERS = 4          # src/lib.rs
CRS_EXTRA_POINTS = 3    # crs_H, crs_G_t, crs_G_u (src/crs.rs:14)
_P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
_MONT_ONE = ((1 << 384) % _P).to_bytes(48, "little")
_FR_ONE = (1).to_bytes(32, "little")


class CrsError(ValueError):
    """The reference's `SerializationError::InvalidData` / `Err("not enough points")`."""


def _affine_to_jacobian(points: bytes) -> bytes:
    out = bytearray()
    zero = bytes(AFFINE_BYTES)
    for i in range(0, len(points), AFFINE_BYTES):
        p = points[i:i + AFFINE_BYTES]
        out += p + (bytes(48) if p == zero else _MONT_ONE)
    return bytes(out)


def _sum_affine_points(engine: Engine, points: bytes) -> bytes:
    """`sum_affine_points` (src/util.rs:108-113): the plain sum, as one MSM with unit scalars, back in affine form."""
    n = len(points) // AFFINE_BYTES
    if n == 0:
        return bytes(AFFINE_BYTES)
    return engine.normalize_batch(engine.msm(points, _FR_ONE * n))


class CurdleproofsCrs:
    """Fields as in src/crs.rs:17-34 (H, G_t, G_u are kept affine: a G1Projective only ever leaves the reference normalised)."""

    def __init__(self, vec_G: bytes, vec_H: bytes, H: bytes, G_t: bytes, G_u: bytes, G_sum: bytes, H_sum: bytes):
        self.vec_G, self.vec_H, self.H, self.G_t, self.G_u, self.G_sum, self.H_sum = vec_G, vec_H, H, G_t, G_u, G_sum, H_sum

    # ---- construction (src/crs.rs:37-58)
    @classmethod
    def from_points(cls, engine: Engine, ell: int, points: bytes) -> "CurdleproofsCrs":
        n = ell + N_BLINDERS
        if len(points) % AFFINE_BYTES:
            raise CrsError("malformed point buffer")
        if len(points) // AFFINE_BYTES < n + CRS_EXTRA_POINTS:
            raise CrsError("not enough points")
        pt = lambda i, j=None: points[i * AFFINE_BYTES:(i + 1 if j is None else j) * AFFINE_BYTES]  # noqa: E731
        vec_G, vec_H = pt(0, ell), pt(ell, n)
        return cls(vec_G, vec_H, pt(n), pt(n + 1), pt(n + 2), _sum_affine_points(engine, vec_G), _sum_affine_points(engine, vec_H))

    @property
    def ell(self) -> int:
        return len(self.vec_G) // AFFINE_BYTES

    def log2_n(self) -> int:
        """src/crs.rs:72-75"""
        n = (len(self.vec_G) + len(self.vec_H)) // AFFINE_BYTES
        return max(0, (n - 1).bit_length())

    def points(self) -> bytes:
        """The ell + 7 points in `from_points` order: what `BatchProver` / `BatchVerifier` take as `crs_points`."""
        return self.vec_G + self.vec_H + self.H + self.G_t + self.G_u

    def __eq__(self, other):
        return isinstance(other, CurdleproofsCrs) and self.__dict__ == other.__dict__

    # ---- CurdleproofsCrsHex (src/crs.rs:77-142): "0x" + hex of the 48-byte compressed encoding, one string per point
    def to_hex(self, engine: Engine) -> dict:
        flat = self.vec_G + self.vec_H + self.H + self.G_t + self.G_u + self.G_sum + self.H_sum
        comp = engine.compress_batch(_affine_to_jacobian(flat))
        hexes = ["0x" + comp[i:i + COMPRESSED_BYTES].hex() for i in range(0, len(comp), COMPRESSED_BYTES)]
        ng, nh = len(self.vec_G) // AFFINE_BYTES, len(self.vec_H) // AFFINE_BYTES
        tail = hexes[ng + nh:]
        return {"vec_G": hexes[:ng], "vec_H": hexes[ng:ng + nh], "H": tail[0], "G_t": tail[1], "G_u": tail[2], "G_sum": tail[3], "H_sum": tail[4]}

    def to_json(self, engine: Engine) -> str:
        return json.dumps(self.to_hex(engine))

    @staticmethod
    def _parse_hex_point(s) -> bytes:
        """The text part of `from_hex_g1affine` (src/crs.rs:128-139): prefix, hex digits, exactly 48 bytes."""
        if not isinstance(s, str) or not s.startswith("0x"):
            raise CrsError("InvalidData: missing 0x prefix")
        try:
            b = bytes.fromhex(s[2:])
        except ValueError:
            raise CrsError("InvalidData: not hex") from None
        if len(b) != COMPRESSED_BYTES or len(s) != 2 + 2 * COMPRESSED_BYTES:
            raise CrsError("InvalidData: a G1 encoding has 48 bytes")
        return b

    @classmethod
    def from_hex(cls, engine: Engine, d: dict) -> "CurdleproofsCrs":
        try:
            vg, vh = list(d["vec_G"]), list(d["vec_H"])
            order = vg + vh + [d["H"], d["G_t"], d["G_u"], d["G_sum"], d["H_sum"]]
        except (KeyError, TypeError):
            raise CrsError("InvalidData: missing field") from None
        comp = b"".join(cls._parse_hex_point(s) for s in order)
        aff, status = engine.decompress_batch(comp)  # on the device: sqrt, sign, curve and subgroup checks
        if any(status):
            i = next(k for k, st in enumerate(status) if st)
            raise CrsError(f"InvalidData: point {i} does not deserialise (status {status[i]})")
        cut = lambda i, j: aff[i * AFFINE_BYTES:j * AFFINE_BYTES]  # noqa: E731
        ng, n = len(vg), len(vg) + len(vh)
        return cls(cut(0, ng), cut(ng, n), cut(n, n + 1), cut(n + 1, n + 2), cut(n + 2, n + 3), cut(n + 3, n + 4), cut(n + 4, n + 5))

    @classmethod
    def from_json(cls, engine: Engine, text: str) -> "CurdleproofsCrs":
        try:
            d = json.loads(text)
        except json.JSONDecodeError:
            raise CrsError("InvalidData: not JSON") from None
        if not isinstance(d, dict):
            raise CrsError("InvalidData: not an object")
        return cls.from_hex(engine, d)
