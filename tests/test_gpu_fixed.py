"""GPU: fixed-base MSM over the HBM-resident digit table (cdp_fixed_table / cdp_msm_fixed*) against the CPU oracle.
The table path must return the same group elements as `util::msm` (/root/reference/src/util.rs:19-22) over the same
points; comparisons are on 48-byte compressed encodings (bit-exact)."""
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


@pytest.fixture(scope="module")
def bases(oracle):
    rnd = random.Random(77)
    pts = bytearray(oracle.scalar_mul_batch(oracle.generator() * 40, rand_scalars(rnd, 40)))
    pts[96 * 7:96 * 8] = bytes(96)  # one base at infinity
    return bytes(pts)


@pytest.fixture(scope="module", params=[19, 16, 8, 5])
def table(request, engine, bases):
    t = engine.fixed_table_create(bases, request.param)
    yield t
    t.close()


def test_table_size(engine, bases):
    t = engine.fixed_table_create(bases[:96 * 3], 16)
    stride = 96 if os.environ.get("CDP_FIXED_STRIDE") == "96" else 128  # entries are 128 bytes apart by default: one DRAM line per gather
    assert t.nbytes == 3 * 16 * 32768 * stride
    t.close()


@pytest.mark.parametrize("n,off", [(1, 0), (1, 39), (2, 3), (7, 5), (33, 0), (40, 0)])
def test_fixed_matches_oracle(engine, oracle, bases, table, n, off):
    rnd = random.Random(100 * n + off)
    sc = rand_scalars(rnd, n)
    got = engine.msm_fixed(table, off, sc)
    want = oracle.msm(bases[96 * off:96 * (off + n)], sc)
    assert oracle.compress_jac(got) == oracle.compress_jac(want)


def test_fixed_edge_scalars(engine, oracle, bases, table):
    """0, 1, r-1, 2^255-ish patterns, all-ones windows (maximal carries of the signed recoding), repeated scalars."""
    vals = [0, 1, 2, pr.R_ORDER - 1, pr.R_ORDER - 2, 2**128, 2**128 - 1, 0x8000 << 16, 0x7FFF7FFF7FFF7FFF, (1 << 254) + 12345,
            int("f" * 63, 16) % pr.R_ORDER, 0x8000800080008000800080008000800080008000800080008000800080008000 % pr.R_ORDER,
            0x7FFF8000, 0xFFFF, 0x10000, 0xFFFFFFFF]
    for k in vals:
        got = engine.msm_fixed(table, 2, pr.fr_to_bytes(k))
        want = oracle.msm(bases[96 * 2:96 * 3], pr.fr_to_bytes(k))
        assert oracle.compress_jac(got) == oracle.compress_jac(want), hex(k)
    sc = b"".join(pr.fr_to_bytes(v) for v in vals)
    got = engine.msm_fixed(table, 8, sc)
    assert oracle.compress_jac(got) == oracle.compress_jac(oracle.msm_naive(bases[96 * 8:96 * (8 + len(vals))], sc))
    beta = pr.fr_to_bytes(random.Random(5).randrange(pr.R_ORDER))  # all-equal scalars (same_permutation_argument.rs:75-76)
    got = engine.msm_fixed(table, 0, beta * 40)
    assert oracle.compress_jac(got) == oracle.compress_jac(oracle.msm(bases, beta * 40))
    assert pr.jacobian_from_bytes(engine.msm_fixed(table, 0, b"")) is pr.INF
    assert pr.jacobian_from_bytes(engine.msm_fixed(table, 7, pr.fr_to_bytes(12345))) is pr.INF  # infinity base


def test_fixed_segments_select_remap_extra(engine, oracle, bases, table):
    """The segment forms the batched prover uses: L / R half selection by an index bit (round MSMs over the original bases),
    a base list with a gap (G_with_blinders), and the extra `+ ip * H` pair."""
    from curdleproofs_b200 import FixedSeg
    rnd = random.Random(31)
    n = 32
    sc = rand_scalars(rnd, n + 2)
    segs, want = [], []

    def pts_of(idx):
        return b"".join(bases[96 * i:96 * i + 96] for i in idx)

    def sc_of(idx):
        return b"".join(sc[32 * i:32 * i + 32] for i in idx)

    out = 0
    for h in (16, 8, 4, 2, 1):
        for val in (0, h):
            idx = [j for j in range(n) if (j & h) == val]
            extra = 1 + 36 if val else 0
            segs.append(FixedSeg(base_off=0, scalars_off=0, n=len(idx), sel_h=h, sel_val=val, remap_from=0xFFFFFFFF, remap_delta=0,
                                 extra_base=extra, extra_scalar=n + 1, out_idx=out))
            p, s = pts_of(idx), sc_of(idx)
            if extra:
                p += bases[96 * 36:96 * 37]
                s += sc[32 * (n + 1):32 * (n + 2)]
            want.append(oracle.msm(p, s))
            out += 1
    # gap: positions 0..9 -> bases 4..13, positions 10..11 -> bases 17..18 ; combined with a selection bit
    for h, val in ((0, 0), (2, 2), (2, 0)):
        pos = [j for j in range(12) if h == 0 or (j & h) == val]
        segs.append(FixedSeg(base_off=4, scalars_off=3, n=len(pos), sel_h=h, sel_val=val, remap_from=10, remap_delta=3,
                             extra_base=0, extra_scalar=0, out_idx=out))
        want.append(oracle.msm(pts_of([4 + j + (3 if j >= 10 else 0) for j in pos]), sc_of([3 + j for j in pos])))
        out += 1
    # plain device-resident points added to the sum (D = B + fixed part; A' = A + T_1 + U_1 with no fixed pairs at all)
    var = bases[96 * 20:96 * 25] + bytes(96)   # five points and one at infinity
    segs.append(FixedSeg(base_off=1, scalars_off=0, n=2, sel_h=0, sel_val=0, remap_from=0xFFFFFFFF, remap_delta=0, extra_base=0, extra_scalar=0,
                         out_idx=out, addv_off=1, addv_n=1))
    want.append(oracle.msm(bases[96:96 * 3] + var[96:96 * 2], sc[:64] + pr.fr_to_bytes(1)))
    out += 1
    segs.append(FixedSeg(base_off=0, scalars_off=0, n=0, sel_h=0, sel_val=0, remap_from=0xFFFFFFFF, remap_delta=0, extra_base=0, extra_scalar=0,
                         out_idx=out, addv_off=0, addv_n=6))
    want.append(oracle.msm(var, pr.fr_to_bytes(1) * 6))
    out += 1
    got = engine.msm_fixed_batch(table, sc, segs, var)
    for i, (g, w) in enumerate(zip(got, want)):
        assert oracle.compress_jac(g) == oracle.compress_jac(w), i


def test_fixed_tree_of_batched_affine_additions(engine, oracle, bases, table):
    """cdp_msm_fixed_batch_dev_tree: the table points of every segment summed as a tree of batched affine additions (k_fixed.cu), against the
    oracle's `util::msm` (/root/reference/src/util.rs:19-22): segments of unequal length (shorter than the padding), an extra pair, half
    selection, plain points, scalars with zero digits / all-equal / 0 / r - 1, an empty segment, and the lane-kernel fallback for short ones."""
    from curdleproofs_b200 import FixedSeg
    rnd = random.Random(8)
    nb = len(bases) // 96
    n = min(nb - 1, 40)
    vals = [rnd.randrange(pr.R_ORDER) for _ in range(n + 2)]
    vals[3] = 0
    vals[4] = pr.R_ORDER - 1
    vals[5] = 1 << 16
    vals[6] = (1 << 200) + 5          # long runs of zero digits
    vals[7] = vals[8] = vals[9]       # equal scalars on different bases
    sc = b"".join(pr.fr_to_bytes(v) for v in vals)
    var = bases[96 * 2:96 * 5]

    def want_of(idx, extra=None, addv=b""):
        p = b"".join(bases[96 * i:96 * i + 96] for i in idx)
        s = b"".join(sc[32 * i:32 * i + 32] for i in idx)
        if extra is not None:
            p += bases[96 * extra[0]:96 * extra[0] + 96]
            s += sc[32 * extra[1]:32 * extra[1] + 32]
        p += addv
        s += pr.fr_to_bytes(1) * (len(addv) // 96)
        return oracle.msm(p, s) if p else None

    segs, want = [], []
    def add(**kw):
        base = dict(base_off=0, scalars_off=0, n=0, sel_h=0, sel_val=0, remap_from=0xFFFFFFFF, remap_delta=0, extra_base=0, extra_scalar=0,
                    out_idx=len(segs))
        base.update(kw)
        segs.append(FixedSeg(**base))
    add(n=n); want.append(want_of(range(n)))
    add(n=n - 1, extra_base=1 + 2, extra_scalar=n + 1); want.append(want_of(range(n - 1), extra=(2, n + 1)))
    add(n=n // 2, sel_h=4, sel_val=4); want.append(want_of([j for j in range(n) if j & 4][:n // 2]))
    add(n=17, addv_off=0, addv_n=3); want.append(want_of(range(17), addv=var))
    add(n=0); want.append(None)
    add(n=1, scalars_off=3); want.append(None)   # scalar 0 on base 0: infinity
    for mp in (n, n + 7):
        got = engine.msm_fixed_batch(table, sc, segs, var, tree_max_pairs=mp)
        for i, (g, w) in enumerate(zip(got, want)):
            if w is None:
                assert pr.jacobian_from_bytes(g) is pr.INF, i
            else:
                assert oracle.compress_jac(g) == oracle.compress_jac(w), (mp, i)
    # fewer than 4 * 2^rounds items per segment: served by the lane kernel
    got = engine.msm_fixed_batch(table, sc, segs[3:4], var, tree_max_pairs=1)
    assert oracle.compress_jac(got[3]) == oracle.compress_jac(want[3])


def test_tuning_setters_reject_bad_arguments(engine):
    """cdp_set_big_msm_min / cdp_set_big_ba_min: sizes below 2048 pairs are refused (the sort-based path has no sensible geometry there), 0 restores the default."""
    from curdleproofs_b200 import CdpError
    for setter in (engine.set_big_msm_min, engine.set_big_ba_min):
        with pytest.raises(CdpError):
            setter(100)
        setter(4096)
        setter(0)
