"""GPU: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded inputs.
Bit-exact: every comparison is on canonical encodings (affine Montgomery bytes or 48-byte compressed points)."""
import os
import random
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
import py_ref as pr  # noqa: E402

pytestmark = pytest.mark.gpu


def rand_scalars(rnd, n):
    return b"".join(pr.fr_to_bytes(rnd.randrange(pr.R_ORDER)) for _ in range(n))


def rand_points(oracle, rnd, n):
    """n subgroup points s_i * G (affine, 96 B each), made by the oracle."""
    return oracle.scalar_mul_batch(oracle.generator() * n, rand_scalars(rnd, n))


@pytest.fixture(scope="module")
def pool(oracle):
    rnd = random.Random(2024)
    return rand_points(oracle, rnd, 2100)


def test_smoke_one_point(engine, oracle):
    g = oracle.generator()
    for k in (0, 1, 2, 3, 5, pr.R_ORDER - 1, 2**128 - 1, 2**128, int(pr.R_ORDER // 3)):
        got = engine.msm(g, pr.fr_to_bytes(k))
        assert engine.compress_batch(got) == pr.compress(pr.mul(pr.G1, k)), k


@pytest.mark.parametrize("n", [1, 2, 3, 4, 7, 11, 12, 16, 32, 39, 40, 64, 95, 96, 128, 252, 256, 1024, 1268, 2048])
def test_msm_sizes(engine, oracle, pool, n):
    rnd = random.Random(n)
    pts, sc = pool[:96 * n], rand_scalars(rnd, n)
    got = engine.msm(pts, sc)
    assert oracle.compress_jac(got) == oracle.compress_jac(oracle.msm(pts, sc, threads=4))


def test_msm_chunked_large(engine, oracle, pool):
    n = 5108  # accumulated-verify size at ell = 1020: goes through the chunked path
    rnd = random.Random(5108)
    pts = (pool * 3)[:96 * n]
    sc = rand_scalars(rnd, n)
    assert oracle.compress_jac(engine.msm(pts, sc)) == oracle.compress_jac(oracle.msm(pts, sc, threads=8))


def test_msm_empty_and_mismatch(engine):
    out = engine.msm(b"", b"")
    assert pr.jacobian_from_bytes(out) is pr.INF
    with pytest.raises(ValueError):   # reference: assert_eq!(points.len(), scalars.len())  src/util.rs:20
        engine.msm(bytes(96), bytes(64))


def test_msm_edge_cases(engine, oracle, pool):
    """infinity bases, zero scalars, all-equal scalars, repeated and opposite points."""
    rnd = random.Random(9)
    n = 64
    pts = bytearray(pool[:96 * n])
    sc = bytearray(rand_scalars(rnd, n))
    for i in (0, 5, 63):
        pts[96 * i:96 * i + 96] = bytes(96)           # infinity bases (vec_T/vec_U blinders, curdleproofs.rs:142-155)
    for i in (1, 6, 62):
        sc[32 * i:32 * i + 32] = bytes(32)            # zero scalars (vec_r_a_prime, curdleproofs.rs:88-89)
    pts[96 * 10:96 * 11] = pts[96 * 11:96 * 12]       # P, P with ...
    sc[32 * 10:32 * 11] = sc[32 * 11:32 * 12]         # ... equal scalars (bucket P + P -> doubling path)
    p12 = pr.affine_from_bytes(bytes(pts[96 * 12:96 * 13]))
    pts[96 * 13:96 * 14] = pr.affine_to_bytes(pr.neg(p12))
    sc[32 * 13:32 * 14] = sc[32 * 12:32 * 13]         # P, -P with equal scalars (bucket cancels to infinity)
    pts, sc = bytes(pts), bytes(sc)
    assert oracle.compress_jac(engine.msm(pts, sc)) == oracle.compress_jac(oracle.msm_naive(pts, sc))
    # all-equal scalars: vec_beta_repeated (same_permutation_argument.rs:75-76)
    beta = pr.fr_to_bytes(rnd.randrange(pr.R_ORDER))
    for m in (4, 64, 252):
        assert oracle.compress_jac(engine.msm(pool[:96 * m], beta * m)) == oracle.compress_jac(oracle.msm(pool[:96 * m], beta * m))
    # all scalars zero, all bases infinity
    assert pr.jacobian_from_bytes(engine.msm(pool[:96 * 8], bytes(32 * 8))) is pr.INF
    assert pr.jacobian_from_bytes(engine.msm(bytes(96 * 8), rand_scalars(rnd, 8))) is pr.INF
    # small scalars
    small = b"".join(pr.fr_to_bytes(rnd.randrange(1 << 32)) for _ in range(100))
    assert oracle.compress_jac(engine.msm(pool[:9600], small)) == oracle.compress_jac(oracle.msm(pool[:9600], small))


def test_msm_linearity_full_size(engine, oracle, pool):
    """Size-independent property at a size the oracle would take long on: msm(P, a) + msm(P, b) == msm(P, a + b)."""
    rnd = random.Random(77)
    n = 2048 * 8
    pts = (pool * 8)[:96 * n]
    a = [rnd.randrange(pr.R_ORDER) for _ in range(n)]
    b = [rnd.randrange(pr.R_ORDER) for _ in range(n)]
    ab = b"".join(pr.fr_to_bytes((x + y) % pr.R_ORDER) for x, y in zip(a, b))
    ja = engine.msm(pts, b"".join(map(pr.fr_to_bytes, a)))
    jb = engine.msm(pts, b"".join(map(pr.fr_to_bytes, b)))
    jab = engine.msm(pts, ab)
    s = pr.add(pr.jacobian_from_bytes(ja), pr.jacobian_from_bytes(jb))
    assert s == pr.jacobian_from_bytes(jab)


def test_msm_batch_mixed_sizes(engine, oracle, pool):
    rnd = random.Random(5)
    sizes = [256, 128, 1, 64, 0, 32, 16, 2, 8, 4, 252, 3, 100, 40, 12]
    items = []
    off = 0
    for n in sizes:
        items.append((pool[96 * off:96 * (off + n)], rand_scalars(rnd, n)))
        off += n
    got = engine.msm_batch(items)
    for (p, s), g in zip(items, got):
        assert oracle.compress_jac(g) == oracle.compress_jac(oracle.msm(p, s))


def test_msm_from_projective(engine, oracle, pool):
    rnd = random.Random(6)
    m = 8  # the verifier's size-m MSMs (inner_product_argument.rs:305-321)
    aff = [pr.affine_from_bytes(pool[96 * i:96 * i + 96]) for i in range(m)]
    aff[2] = pr.INF
    jac = b"".join(pr.jacobian_to_bytes(p, z=rnd.randrange(2, pr.P)) for p in aff)
    sc = rand_scalars(rnd, m)
    assert oracle.compress_jac(engine.msm_from_projective(jac, sc)) == oracle.compress_jac(oracle.msm_from_projective(jac, sc))


@pytest.mark.parametrize("n", [1, 2, 5, 128, 512])
def test_fold(engine, oracle, pool, n):
    rnd = random.Random(100 + n)
    L = bytearray(pool[:96 * n])
    R = bytearray(pool[96 * 600:96 * (600 + n)])
    if n >= 5:
        L[0:96] = bytes(96)                      # L = infinity (T/U blinder slots)
        R[96:192] = bytes(96)                    # R = infinity
        L[192:288] = bytes(96)
        R[192:288] = bytes(96)                   # both infinity
    L, R = bytes(L), bytes(R)
    for gamma in (rnd.randrange(pr.R_ORDER), 0, 1, pr.R_ORDER - 1):
        g = pr.fr_to_bytes(gamma)
        assert engine.fold(L, R, g) == oracle.fold(L, R, g)
        if n > 5:
            break


def test_fold_cancels_to_infinity(engine, oracle, pool):
    # L = -gamma*R  =>  the fold output is the point at infinity (all-zero affine encoding)
    gamma = 0x1234567890abcdef1234567890abcdef
    R = pool[:96]
    gR = pr.mul(pr.affine_from_bytes(R), gamma)
    L = pr.affine_to_bytes(pr.neg(gR))
    assert engine.fold(L, R, pr.fr_to_bytes(gamma)) == bytes(96)
    # L = gamma*R  =>  doubling branch of the final mixed addition
    L2 = pr.affine_to_bytes(gR)
    assert engine.fold(L2, R, pr.fr_to_bytes(gamma)) == oracle.fold(L2, R, pr.fr_to_bytes(gamma))


@pytest.mark.parametrize("n", [1, 33, 256])
def test_scalar_mul_batch(engine, oracle, pool, n):
    rnd = random.Random(200 + n)
    pts = bytearray(pool[:96 * n])
    sc = bytearray(rand_scalars(rnd, n))
    if n > 3:
        pts[96:192] = bytes(96)
        sc[64:96] = bytes(32)
    assert engine.scalar_mul_batch(bytes(pts), bytes(sc)) == oracle.scalar_mul_batch(bytes(pts), bytes(sc))


def test_normalize_and_compress(engine, oracle, pool):
    rnd = random.Random(300)
    aff = [pr.affine_from_bytes(pool[96 * i:96 * i + 96]) for i in range(40)]
    aff[0] = pr.INF
    aff[17] = pr.INF
    jac = b"".join(pr.jacobian_to_bytes(p, z=rnd.randrange(1, pr.P)) for p in aff)
    assert engine.normalize_batch(jac) == b"".join(pr.affine_to_bytes(p) for p in aff)
    assert engine.compress_batch(jac) == b"".join(pr.compress(p) for p in aff)
    assert engine.normalize_batch(jac) == oracle.normalize_batch(jac)


def test_generator_kat_through_gpu(engine, oracle):
    # /root/reference/src/whisk.rs:363-368
    g = oracle.generator()
    jac = g + pr.fp_to_mont_bytes(1)
    want = open(os.path.join(HERE, "golden", "g1_generator_compressed.hex")).read().strip()
    assert engine.compress_batch(jac).hex() == want


@pytest.fixture(params=["large_path", "large_path_batched_affine", "chunked_small_path"])
def msm_path(request, engine):
    """One MSM of 2^13 .. 2^16 pairs through BOTH implementations: the sort-based large Pippenger (k_bigmsm.cu; forced from 2^13 here, the
    default crossover is 2^16) and the chunked small-MSM kernels (the default below 2^16)."""
    engine.set_big_msm_min(8192 if request.param.startswith("large_path") else 0)
    engine.set_big_ba_min(8192 if request.param == "large_path_batched_affine" else 1 << 30)
    yield request.param
    engine.set_big_msm_min(0)
    engine.set_big_ba_min(0)


@pytest.mark.parametrize("n", [8192, 8192 + 37, 20000])
def test_msm_large_pippenger(engine, oracle, pool, n, msm_path):
    """Bit-exact vs the oracle, with infinity bases, zero scalars and small scalars mixed in."""
    rnd = random.Random(n)
    pts = bytearray((pool * (n // 2100 + 1))[:96 * n])
    sc = bytearray(rand_scalars(rnd, n))
    for i in (0, 17, n - 1):
        pts[96 * i:96 * i + 96] = bytes(96)
    for i in (3, 18, n - 2):
        sc[32 * i:32 * i + 32] = bytes(32)
    for i in range(100, 140):
        sc[32 * i:32 * i + 32] = pr.fr_to_bytes(rnd.randrange(1 << 20))
    sc[32 * 50:32 * 51] = pr.fr_to_bytes(pr.R_ORDER - 1)
    pts, sc = bytes(pts), bytes(sc)
    assert oracle.compress_jac(engine.msm(pts, sc)) == oracle.compress_jac(oracle.msm(pts, sc, threads=8))


@pytest.mark.parametrize("kind", ["all_equal", "small32", "half_shared", "two_values"])
def test_msm_large_pippenger_skewed_scalars(engine, oracle, pool, kind, msm_path):
    """Skewed scalar distributions on the large-Pippenger path (BASELINE config 5's extra distributions; the all-equal case is the
    SamePerm MSM shape, /root/reference/src/same_permutation_argument.rs:75-76): whole windows collapse into one bucket, which goes
    through the heavy-bucket work list (k_big_heavy / k_big_heavy_fold, also for the top window).  Bit-exact vs the oracle."""
    n = 16384 + 123
    rnd = random.Random(hash(kind) & 0xFFFF)
    pts = (pool * (n // 2100 + 1))[:96 * n]
    one = pr.fr_to_bytes(rnd.randrange(1, pr.R_ORDER))
    if kind == "all_equal":
        sc = one * n
    elif kind == "small32":
        sc = b"".join(pr.fr_to_bytes(rnd.randrange(1 << 32)) for _ in range(n))
    elif kind == "half_shared":
        sc = b"".join(one if i % 2 else pr.fr_to_bytes(rnd.randrange(pr.R_ORDER)) for i in range(n))
    else:
        two = pr.fr_to_bytes(pr.R_ORDER - 2)
        sc = b"".join(one if rnd.random() < 0.7 else two for _ in range(n))
    assert oracle.compress_jac(engine.msm(pts, sc)) == oracle.compress_jac(oracle.msm(pts, sc, threads=8))


@pytest.mark.parametrize("n", [131072 + 11, 300000, 1 << 20])
def test_msm_large_pippenger_round_geometry(engine, oracle, n):
    """The bucket sums of the large path are rounds of batched affine additions (k_batchaff.cu) whose thread count and additions per inversion
    depend on the round's size: sizes on either side of every regime (K = minimum, K growing with the round, K = maximum), on DISTINCT random
    subgroup points (multiples of the generator made on the device), bit-exact vs the oracle's `util::msm` (/root/reference/src/util.rs:19-22)."""
    import numpy as np
    rng = np.random.Generator(np.random.Philox(key=n))
    t = rng.integers(0, 2 ** 64, size=(n, 4), dtype=np.uint64)
    t[:, 3] &= np.uint64((1 << 62) - 1)
    s = rng.integers(0, 2 ** 64, size=(n, 4), dtype=np.uint64)
    s[:, 3] &= np.uint64((1 << 62) - 1)
    pts = engine.scalar_mul_batch(oracle.generator() * n, t.tobytes())
    sc = s.tobytes()
    assert oracle.compress_jac(engine.msm(pts, sc)) == oracle.compress_jac(oracle.msm(pts, sc, threads=16))
