#!/usr/bin/env python3
"""Extract the known-answer hex strings held by the reference's own tests into tests/golden/*.hex.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/extract_reference_vectors.py
Sources:
  whisk_tracker_proof_seed0.hex  <- /root/reference/src/whisk.rs:401   (128-byte tracker proof, StdRng seed 0)
  whisk_shuffle_proof_seed0.hex  <- /root/reference/src/whisk.rs:455   (4496-byte shuffle proof, N=128, StdRng seed 0)
  g1_generator_compressed.hex    <- /root/reference/src/whisk.rs:365
  fr_roundtrip.hex               <- /root/reference/src/whisk.rs:357
"""
import os
import re

SRC = "/root/reference/src/whisk.rs"
OUT = os.path.dirname(os.path.abspath(__file__))
lines = open(SRC).read().split("\n")


def hex_on_line(lineno, minlen):
    m = re.findall(r'"([0-9a-f]{%d,})"' % minlen, lines[lineno - 1])
    assert len(m) == 1, (lineno, m)
    return m[0]


vectors = {
    "whisk_tracker_proof_seed0.hex": hex_on_line(401, 256),
    "whisk_shuffle_proof_seed0.hex": hex_on_line(455, 8992),
    "g1_generator_compressed.hex": hex_on_line(365, 96),
    "fr_roundtrip.hex": hex_on_line(357, 64),
}
assert len(vectors["whisk_tracker_proof_seed0.hex"]) == 256
assert len(vectors["whisk_shuffle_proof_seed0.hex"]) == 8992
for name, val in vectors.items():
    with open(os.path.join(OUT, name), "w") as f:
        f.write(val + "\n")
    print(name, len(val) // 2, "bytes")
