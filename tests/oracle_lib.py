"""ctypes wrapper around oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY (never imported by curdleproofs_b200/)."""
import ctypes
import os
import subprocess
from ctypes import POINTER, c_char_p, c_double, c_int, c_size_t, c_uint32, c_uint64, c_void_p

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "liboracle.so")


def build():
    subprocess.run(["make", "-C", ORACLE_DIR, "-s"], check=True)


def _u8(n):
    return (ctypes.c_uint8 * max(1, n))()


def _in(b):
    return (ctypes.c_uint8 * max(1, len(b))).from_buffer_copy(bytes(b) if len(b) else b"\0")


class Oracle:
    def __init__(self):
        if not os.path.exists(LIB):
            build()
        L = self.L = ctypes.CDLL(LIB)
        L.oracle_msm.argtypes = [c_void_p, c_void_p, c_size_t, c_void_p, c_int]
        L.oracle_msm_naive.argtypes = [c_void_p, c_void_p, c_size_t, c_void_p]
        L.oracle_msm_from_projective.argtypes = [c_void_p, c_void_p, c_size_t, c_void_p]
        L.oracle_scalar_mul_batch.argtypes = [c_void_p, c_void_p, c_size_t, c_void_p]
        L.oracle_fold.argtypes = [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
        L.oracle_normalize_batch.argtypes = [c_void_p, c_size_t, c_void_p]
        L.oracle_compress.argtypes = [c_void_p, c_size_t, c_void_p]
        L.oracle_decompress.argtypes = [c_void_p, c_size_t, c_void_p, c_int]
        L.oracle_on_curve.argtypes = [c_void_p, c_size_t]
        L.oracle_merlin_kat.argtypes = [c_char_p, c_char_p, c_char_p, c_size_t, c_char_p, c_void_p, c_size_t]
        L.oracle_stdrng_u32s.argtypes = [c_uint64, c_void_p, c_size_t]
        L.oracle_whisk_shuffle_proof_seed0.argtypes = [c_size_t, c_void_p, c_void_p, POINTER(c_int), c_int]
        L.oracle_generate_crs_points.argtypes = [c_size_t, c_void_p]
        L.oracle_random_instance.argtypes = [c_size_t, c_void_p, c_uint64, c_int] + [c_void_p] * 8 + [c_int]
        L.oracle_prove.argtypes = [c_size_t] + [c_void_p] * 9 + [c_uint64, c_void_p, c_int]
        L.oracle_verify.argtypes = [c_size_t] + [c_void_p] * 7 + [c_uint64, c_int, c_void_p, c_void_p, POINTER(c_size_t)]
        L.oracle_proof_size.argtypes = [c_size_t]
        L.oracle_proof_size.restype = c_size_t
        L.oracle_time_prove.argtypes = [c_size_t] + [c_void_p] * 9 + [c_int, c_int, c_void_p]
        L.oracle_time_prove.restype = c_double
        L.oracle_time_verify.argtypes = [c_size_t] + [c_void_p] * 7 + [c_int, c_int, POINTER(c_int)]
        L.oracle_time_verify.restype = c_double
        L.oracle_time_msm.argtypes = [c_void_p, c_void_p, c_size_t, c_int, c_int, c_void_p]
        L.oracle_time_msm.restype = c_double

    # ---- boundary ops (include/cdp_msm.h layouts) ----
    def msm(self, pts, scalars, threads=1):
        n = len(pts) // 96
        out = _u8(144)
        self.L.oracle_msm(_in(pts), _in(scalars), n, out, threads)
        return bytes(out)

    def msm_naive(self, pts, scalars):
        n = len(pts) // 96
        out = _u8(144)
        self.L.oracle_msm_naive(_in(pts), _in(scalars), n, out)
        return bytes(out)

    def msm_from_projective(self, jac, scalars):
        n = len(jac) // 144
        out = _u8(144)
        self.L.oracle_msm_from_projective(_in(jac), _in(scalars), n, out)
        return bytes(out)

    def scalar_mul_batch(self, pts, scalars):
        n = len(pts) // 96
        out = _u8(96 * n)
        self.L.oracle_scalar_mul_batch(_in(pts), _in(scalars), n, out)
        return bytes(out)[:96 * n]

    def fold(self, L_, R_, gamma):
        n = len(L_) // 96
        out = _u8(96 * n)
        self.L.oracle_fold(_in(L_), _in(R_), _in(gamma), n, out)
        return bytes(out)[:96 * n]

    def normalize_batch(self, jac):
        n = len(jac) // 144
        out = _u8(96 * n)
        self.L.oracle_normalize_batch(_in(jac), n, out)
        return bytes(out)[:96 * n]

    def compress(self, affine):
        n = len(affine) // 96
        out = _u8(48 * n)
        self.L.oracle_compress(_in(affine), n, out)
        return bytes(out)[:48 * n]

    def compress_jac(self, jac):
        return self.compress(self.normalize_batch(jac))

    def decompress(self, comp, check_subgroup=True):
        n = len(comp) // 48
        out = _u8(96 * n)
        rc = self.L.oracle_decompress(_in(comp), n, out, int(check_subgroup))
        if rc != 0:
            raise ValueError("invalid encoding")
        return bytes(out)[:96 * n]

    def on_curve(self, affine):
        return bool(self.L.oracle_on_curve(_in(affine), len(affine) // 96))

    def generator(self):
        out = _u8(96)
        self.L.oracle_generator(out)
        return bytes(out)

    # ---- protocol ----
    def stdrng_seed_bytes(self, seed):
        """The 32-byte ChaCha12 key `StdRng::seed_from_u64(seed)` expands to (rand_core's PCG32 fill): the same stream through `from_seed`."""
        out = _u8(32)
        self.L.oracle_stdrng_seed_bytes.argtypes = [c_uint64, c_void_p]
        self.L.oracle_stdrng_seed_bytes(seed, out)
        return bytes(out)

    def crs_points(self, ell):
        out = _u8(96 * (ell + 7))
        self.L.oracle_generate_crs_points(ell, out)
        return bytes(out)

    def random_instance(self, ell, crs_pts, seed, fast_points=True, threads=1):
        R, S, T, U = (_u8(96 * ell) for _ in range(4))
        M, k, mb = _u8(144), _u8(32), _u8(128)
        perm = (c_uint32 * ell)()
        self.L.oracle_random_instance(ell, _in(crs_pts), seed, int(fast_points), R, S, T, U, M, perm, k, mb, threads)
        return dict(ell=ell, crs=bytes(crs_pts), R=bytes(R), S=bytes(S), T=bytes(T), U=bytes(U), M=bytes(M), perm=list(perm), k=bytes(k),
                    m_blinders=bytes(mb))

    def prove(self, inst, rng_seed, threads=1):
        ell = inst["ell"]
        size = self.L.oracle_proof_size(ell)
        out = _u8(size)
        perm = (c_uint32 * ell)(*inst["perm"])
        n = self.L.oracle_prove(ell, _in(inst["crs"]), _in(inst["R"]), _in(inst["S"]), _in(inst["T"]), _in(inst["U"]), _in(inst["M"]), perm,
                                _in(inst["k"]), _in(inst["m_blinders"]), rng_seed, out, threads)
        assert n == size
        return bytes(out)

    def verify(self, inst, proof, rng_seed=1, threads=1, export_acc=False):
        ell = inst["ell"]
        if export_acc:
            cap = 5 * ell + 16
            ab, asc, an = _u8(96 * cap), _u8(32 * cap), c_size_t(0)
            rc = self.L.oracle_verify(ell, _in(inst["crs"]), _in(inst["R"]), _in(inst["S"]), _in(inst["T"]), _in(inst["U"]), _in(inst["M"]),
                                      _in(proof), rng_seed, threads, ab, asc, ctypes.byref(an))
            return rc, bytes(ab)[:96 * an.value], bytes(asc)[:32 * an.value]
        return self.L.oracle_verify(ell, _in(inst["crs"]), _in(inst["R"]), _in(inst["S"]), _in(inst["T"]), _in(inst["U"]), _in(inst["M"]),
                                    _in(proof), rng_seed, threads, None, None, None)

    def whisk_shuffle_proof_seed0(self, ell=124, threads=1, want_instance=False):
        size = 48 + self.L.oracle_proof_size(ell)
        out = _u8(size)
        inst = _u8(384 * ell + 144 + 4 * ell + 32 + 128 + 16) if want_instance else None
        ok = c_int(-1)
        n = self.L.oracle_whisk_shuffle_proof_seed0(ell, out, inst, ctypes.byref(ok), threads)
        assert n == size
        d = None
        if want_instance:
            raw = bytes(inst)
            o = 384 * ell
            d = dict(ell=ell, crs=self.crs_points(ell), R=raw[:96 * ell], S=raw[96 * ell:192 * ell], T=raw[192 * ell:288 * ell],
                     U=raw[288 * ell:o], M=raw[o:o + 144],
                     perm=[int.from_bytes(raw[o + 144 + 4 * i:o + 148 + 4 * i], "little") for i in range(ell)],
                     k=raw[o + 144 + 4 * ell:o + 176 + 4 * ell], m_blinders=raw[o + 176 + 4 * ell:o + 304 + 4 * ell],
                     rng_words=int.from_bytes(raw[o + 304 + 4 * ell:o + 312 + 4 * ell], "little"),
                     whisk_entry_words=int.from_bytes(raw[o + 312 + 4 * ell:o + 320 + 4 * ell], "little"))
        return bytes(out), bool(ok.value), d

    def whisk_tracker_inputs_seed0(self):
        out = _u8(184)
        self.L.oracle_whisk_tracker_inputs_seed0(out)
        raw = bytes(out)
        return dict(k=raw[:32], tracker=raw[32:128], k_commitment=raw[128:176], rng_words=int.from_bytes(raw[176:184], "little"))

    def whisk_tracker_proof_seed0(self):
        out = _u8(128)
        self.L.oracle_whisk_tracker_proof_seed0(out)
        return bytes(out)
