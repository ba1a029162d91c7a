"""ctypes binding of include/cdp_msm.h.  Byte layouts are documented in that header."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_size_t, c_uint32, c_uint64, c_void_p

FP_BYTES, SCALAR_BYTES, AFFINE_BYTES, JACOBIAN_BYTES, COMPRESSED_BYTES = 48, 32, 96, 144, 48

_HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path() -> str:
    return os.path.join(_HERE, "libcdp_b200.so")


class CdpError(RuntimeError):
    pass


class FixedSeg(ctypes.Structure):
    """cdp_fixed_seg (include/cdp_msm.h)."""
    _fields_ = [("base_off", c_uint32), ("scalars_off", c_uint32), ("n", c_uint32), ("sel_h", c_uint32), ("sel_val", c_uint32),
                ("remap_from", c_uint32), ("remap_delta", c_uint32), ("extra_base", c_uint32), ("extra_scalar", c_uint32),
                ("out_idx", c_uint32), ("addv_off", c_uint32), ("addv_n", c_uint32), ("pos_off", c_uint32), ("pos_stride", c_uint32)]


class _MsmDesc(ctypes.Structure):
    _fields_ = [("affine_pts", c_void_p), ("scalars", c_void_p), ("n", c_size_t)]


_SIGS = {
    "cdp_ctx_create": (c_int, [POINTER(c_void_p), c_int, c_void_p]),
    "cdp_ctx_destroy": (None, [c_void_p]),
    "cdp_last_error": (c_char_p, [c_void_p]),
    "cdp_launch_count": (c_uint64, [c_void_p]),
    "cdp_sync": (c_int, [c_void_p]),
    "cdp_ctx_device": (c_int, [c_void_p]),
    "cdp_msm": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "cdp_msm_from_projective": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "cdp_msm_batch": (c_int, [c_void_p, POINTER(_MsmDesc), c_size_t, c_void_p]),
    "cdp_fold": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "cdp_scalar_mul_batch": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "cdp_normalize_batch": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "cdp_compress_batch": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "cdp_dev_alloc": (c_void_p, [c_void_p, c_size_t]),
    "cdp_dev_free": (None, [c_void_p, c_void_p]),
    "cdp_h2d": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t]),
    "cdp_d2h": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t]),
    "cdp_host_alloc": (c_void_p, [c_void_p, c_size_t]),
    "cdp_host_free": (None, [c_void_p, c_void_p]),
    "cdp_msm_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "cdp_sum_jacobian_dev": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "cdp_sum_groups_dev": (c_int, [c_void_p, c_void_p, c_size_t, c_size_t, c_size_t, c_void_p]),
    "cdp_sum_groups2_dev": (c_int, [c_void_p, c_void_p, c_size_t, c_size_t, c_size_t, c_void_p, c_size_t, c_size_t, c_size_t, c_size_t, c_void_p]),
    "cdp_msm_batch_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_size_t, c_size_t, c_void_p]),
    "cdp_fixed_table_create": (c_int, [c_void_p, c_void_p, c_size_t, c_int, POINTER(c_void_p)]),
    "cdp_fixed_table_destroy": (None, [c_void_p, c_void_p]),
    "cdp_fixed_table_bytes": (c_size_t, [c_void_p]),
    "cdp_fixed_table_bases": (c_size_t, [c_void_p]),
    "cdp_msm_fixed": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_void_p]),
    "cdp_msm_fixed_batch_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_size_t, c_void_p, c_void_p]),
    "cdp_msm_fixed_batch_dev_tree": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_size_t, c_void_p, c_void_p, c_size_t]),
    "cdp_msm_fixed_batch_dev_lanes": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_size_t, c_void_p, c_void_p, c_int]),
    "cdp_profile_enable": (c_int, [c_void_p, c_int]),
    "cdp_profile_reset": (c_int, [c_void_p]),
    "cdp_profile_read": (c_int, [c_void_p, POINTER(ctypes.c_double), POINTER(c_uint64), POINTER(c_uint64)]),
    "cdp_smul_jobs_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_size_t]),
    "cdp_gather_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t]),
    "cdp_decompress_batch": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "cdp_decompress_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "cdp_transcript_open_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_size_t, c_void_p, c_void_p]),
    "cdp_compress_affine_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "cdp_normalize_dev": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "cdp_bench_kernel": (c_int, [c_void_p, c_int, c_int, c_int, c_int, POINTER(c_float)]),
    "cdp_prove_random_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_size_t, c_void_p]),
    "cdp_prove_random_scalars": (c_size_t, [c_size_t]),
    "cdp_set_big_msm_min": (c_int, [c_void_p, c_size_t]),
    "cdp_set_big_ba_min": (c_int, [c_void_p, c_size_t]),
    "cdp_host_is_pinned": (c_int, [c_void_p]),
    "cdp_h2d_2d": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_size_t, c_size_t]),
    "cdp_comm_unique_id": (c_int, [c_void_p]),
    "cdp_comm_create": (c_int, [POINTER(c_void_p), c_void_p, c_void_p, c_int, c_int]),
    "cdp_comm_create_all": (c_int, [POINTER(c_void_p), POINTER(c_void_p), c_int]),
    "cdp_comm_destroy": (None, [c_void_p]),
    "cdp_comm_rank": (c_int, [c_void_p]),
    "cdp_comm_size": (c_int, [c_void_p]),
    "cdp_comm_last_error": (c_char_p, [c_void_p]),
    "cdp_shard_range": (None, [c_size_t, c_int, c_int, POINTER(c_size_t), POINTER(c_size_t)]),
    "cdp_msm_sharded_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "cdp_msm_sharded_group": (c_int, [POINTER(c_void_p), c_int, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_size_t), POINTER(c_void_p)]),
    "cdp_allreduce_jacobian_dev": (c_int, [c_void_p, c_void_p, c_void_p]),
}

_LIB = None


def load_library() -> ctypes.CDLL:
    """Load libcdp_b200.so (fails loudly when it has not been built -- there is no fallback)."""
    global _LIB
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            raise CdpError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(make -C curdleproofs_b200/csrc). There is no CPU fallback.")
        lib = ctypes.CDLL(path)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = lib
    return _LIB


def _buf(b) -> ctypes.Array:
    if isinstance(b, (bytes, bytearray, memoryview)):
        return (ctypes.c_uint8 * len(b)).from_buffer_copy(bytes(b))
    raise TypeError("expected a bytes-like object")


class Engine:
    """One context on one GPU (``cdp_ctx``)."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self._lib = load_library()
        h = c_void_p()
        rc = self._lib.cdp_ctx_create(ctypes.byref(h), device, c_void_p(stream) if stream else None)
        if rc != 0:
            raise CdpError(f"cdp_ctx_create failed (code {rc}): no usable sm_100 CUDA device {device}; there is no CPU fallback")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            for p in getattr(self, "_pinned", []):
                self._lib.cdp_host_free(self._h, p)
            self._pinned = []
            self._lib.cdp_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise CdpError(f"{what} failed (code {rc}): {self._lib.cdp_last_error(self._h).decode()}")

    @property
    def handle(self):
        return self._h

    @property
    def lib(self):
        return self._lib

    @property
    def launch_count(self) -> int:
        return int(self._lib.cdp_launch_count(self._h))

    def pinned_array(self, data: bytes):
        """A page-locked host buffer (cdp_host_alloc) holding `data`, as a ctypes array: inputs handed to the batched prover / verifier from
        such buffers go to the device by direct DMA, without a staging pass.  Freed with the Engine (keep the Engine alive while in use)."""
        n = max(1, len(data))
        p = self._lib.cdp_host_alloc(self._h, n)
        if not p:
            raise CdpError("cdp_host_alloc failed")
        arr = (ctypes.c_uint8 * n).from_address(p)
        ctypes.memmove(arr, data, len(data))
        self._pinned = getattr(self, "_pinned", [])
        self._pinned.append(p)
        return arr

    def set_big_msm_min(self, n_pairs: int):
        """Pairs from which one MSM takes the sort-based large Pippenger on this engine (0 = the built-in 2^16)."""
        self._check(self._lib.cdp_set_big_msm_min(self._h, n_pairs), "cdp_set_big_msm_min")

    def prove_random(self, keys: bytes, ell: int, skip_words=None) -> bytes:
        """The prover's `Fr::rand` draws for len(keys) / 32 proofs from their ChaCha12 keys (cdp_prove_random_dev): per proof
        cdp_prove_random_scalars(ell) raw 32-byte Montgomery representations in the layout of cdp_prove_dev.d_random."""
        import struct
        lib, h = self._lib, self._h
        batch = len(keys) // 32
        nrnd = int(lib.cdp_prove_random_scalars(ell))
        d_k = lib.cdp_dev_alloc(h, 40 * batch)
        d_o = lib.cdp_dev_alloc(h, 32 * nrnd * batch)
        try:
            blob = keys + struct.pack("<%dQ" % batch, *(skip_words or [0] * batch))
            self._check(lib.cdp_h2d(h, d_k, _buf(blob), len(blob)), "cdp_h2d")
            self._check(lib.cdp_prove_random_dev(h, d_k, (d_k + 32 * batch) if skip_words else None, batch, ell, d_o), "cdp_prove_random_dev")
            out = (ctypes.c_uint8 * (32 * nrnd * batch))()
            self._check(lib.cdp_d2h(h, out, d_o, 32 * nrnd * batch), "cdp_d2h")
            self.sync()
        finally:
            lib.cdp_dev_free(h, d_k)
            lib.cdp_dev_free(h, d_o)
        return bytes(out)

    def set_big_ba_min(self, n_pairs: int):
        """Pairs from which the large Pippenger sums its buckets by rounds of batched affine additions (0 = the built-in 2^19)."""
        self._check(self._lib.cdp_set_big_ba_min(self._h, n_pairs), "cdp_set_big_ba_min")

    def sync(self):
        self._check(self._lib.cdp_sync(self._h), "cdp_sync")

    # ---- drop-ins ------------------------------------------------------------------------------------
    def msm(self, points: bytes, scalars: bytes) -> bytes:
        """`util::msm` (src/util.rs:19-22): affine points (96 B each), canonical scalars (32 B each) -> Jacobian (144 B).
        Like the reference (`assert_eq!`, src/util.rs:20) a length mismatch is an error."""
        if len(points) % AFFINE_BYTES or len(scalars) % SCALAR_BYTES:
            raise ValueError("malformed input length")
        n = len(points) // AFFINE_BYTES
        if n != len(scalars) // SCALAR_BYTES:
            raise ValueError("number of points != number of scalars")
        out = (ctypes.c_uint8 * JACOBIAN_BYTES)()
        self._check(self._lib.cdp_msm(self._h, _buf(points), _buf(scalars), n, out), "cdp_msm")
        return bytes(out)

    def msm_from_projective(self, points: bytes, scalars: bytes) -> bytes:
        """`util::msm_from_projective` (src/util.rs:25-29)."""
        if len(points) % JACOBIAN_BYTES or len(scalars) % SCALAR_BYTES:
            raise ValueError("malformed input length")
        n = len(points) // JACOBIAN_BYTES
        if n != len(scalars) // SCALAR_BYTES:
            raise ValueError("number of points != number of scalars")
        out = (ctypes.c_uint8 * JACOBIAN_BYTES)()
        self._check(self._lib.cdp_msm_from_projective(self._h, _buf(points), _buf(scalars), n, out), "cdp_msm_from_projective")
        return bytes(out)

    def msm_batch(self, items) -> list[bytes]:
        """Independent MSMs [(points, scalars), ...] in one call."""
        keep = []
        descs = (_MsmDesc * len(items))()
        for i, (p, s) in enumerate(items):
            n = len(p) // AFFINE_BYTES
            if n != len(s) // SCALAR_BYTES or len(p) % AFFINE_BYTES or len(s) % SCALAR_BYTES:
                raise ValueError("number of points != number of scalars")
            pb, sb = _buf(p) if n else None, _buf(s) if n else None
            keep.append((pb, sb))
            descs[i].affine_pts = ctypes.cast(pb, c_void_p) if n else None
            descs[i].scalars = ctypes.cast(sb, c_void_p) if n else None
            descs[i].n = n
        out = (ctypes.c_uint8 * (JACOBIAN_BYTES * max(1, len(items))))()
        self._check(self._lib.cdp_msm_batch(self._h, descs, len(items), out), "cdp_msm_batch")
        raw = bytes(out)
        return [raw[i * JACOBIAN_BYTES:(i + 1) * JACOBIAN_BYTES] for i in range(len(items))]

    def fold(self, L: bytes, R: bytes, gamma: bytes) -> bytes:
        """(L[i] + gamma * R[i]).into_affine() -- src/inner_product_argument.rs:177-178, src/same_multiscalar_argument.rs:128-130."""
        if len(L) != len(R) or len(L) % AFFINE_BYTES or len(gamma) != SCALAR_BYTES:
            raise ValueError("malformed input length")
        n = len(L) // AFFINE_BYTES
        out = (ctypes.c_uint8 * max(1, len(L)))()
        self._check(self._lib.cdp_fold(self._h, _buf(L) if n else None, _buf(R) if n else None, _buf(gamma), n, out), "cdp_fold")
        return bytes(out)[:len(L)]

    def scalar_mul_batch(self, points: bytes, scalars: bytes) -> bytes:
        n = len(points) // AFFINE_BYTES
        if len(points) % AFFINE_BYTES or len(scalars) != n * SCALAR_BYTES:
            raise ValueError("number of points != number of scalars")
        out = (ctypes.c_uint8 * max(1, len(points)))()
        self._check(self._lib.cdp_scalar_mul_batch(self._h, _buf(points) if n else None, _buf(scalars) if n else None, n, out),
                    "cdp_scalar_mul_batch")
        return bytes(out)[:len(points)]

    def normalize_batch(self, jac: bytes) -> bytes:
        n = len(jac) // JACOBIAN_BYTES
        if len(jac) % JACOBIAN_BYTES:
            raise ValueError("malformed input length")
        out = (ctypes.c_uint8 * max(1, n * AFFINE_BYTES))()
        self._check(self._lib.cdp_normalize_batch(self._h, _buf(jac) if n else None, n, out), "cdp_normalize_batch")
        return bytes(out)[:n * AFFINE_BYTES]

    def compress_batch(self, jac: bytes) -> bytes:
        n = len(jac) // JACOBIAN_BYTES
        if len(jac) % JACOBIAN_BYTES:
            raise ValueError("malformed input length")
        out = (ctypes.c_uint8 * max(1, n * COMPRESSED_BYTES))()
        self._check(self._lib.cdp_compress_batch(self._h, _buf(jac) if n else None, n, out), "cdp_compress_batch")
        return bytes(out)[:n * COMPRESSED_BYTES]

    def decompress_batch(self, comp: bytes):
        """48-byte encodings -> (affine bytes, status list); status 0 ok / 1 malformed / 2 not on curve / 3 not in subgroup."""
        n = len(comp) // COMPRESSED_BYTES
        if len(comp) % COMPRESSED_BYTES:
            raise ValueError("malformed input length")
        out = (ctypes.c_uint8 * max(1, n * AFFINE_BYTES))()
        st = (ctypes.c_uint8 * max(1, n))()
        rc = self._lib.cdp_decompress_batch(self._h, _buf(comp) if n else None, n, out, st)
        if rc not in (0, 4):
            self._check(rc, "cdp_decompress_batch")
        return bytes(out)[:n * AFFINE_BYTES], list(st)[:n]

    def transcript_open(self, comp_vecs: bytes, comp_M: bytes, ell: int):
        """Device-side transcript opening (cdp_transcript_open_dev): comp_vecs = batch x 4 x ell encodings (R | S | T | U per proof),
        comp_M = batch encodings.  Returns (vec_a bytes: batch x ell x 32, states: batch x 208 bytes)."""
        B = len(comp_M) // COMPRESSED_BYTES
        if len(comp_vecs) != B * 4 * ell * COMPRESSED_BYTES:
            raise ValueError("malformed input length")
        lib, h = self._lib, self._h
        d_v, d_m = lib.cdp_dev_alloc(h, len(comp_vecs)), lib.cdp_dev_alloc(h, len(comp_M))
        d_a, d_s = lib.cdp_dev_alloc(h, B * ell * 32), lib.cdp_dev_alloc(h, B * 208)
        try:
            bv, bm = _buf(comp_vecs), _buf(comp_M)
            self._check(lib.cdp_h2d(h, d_v, bv, len(comp_vecs)), "cdp_h2d")
            self._check(lib.cdp_h2d(h, d_m, bm, len(comp_M)), "cdp_h2d")
            self._check(lib.cdp_transcript_open_dev(h, d_v, d_m, ell, B, d_a, d_s), "cdp_transcript_open_dev")
            oa, os_ = (ctypes.c_uint8 * (B * ell * 32))(), (ctypes.c_uint8 * (B * 208))()
            self._check(lib.cdp_d2h(h, oa, d_a, B * ell * 32), "cdp_d2h")
            self._check(lib.cdp_d2h(h, os_, d_s, B * 208), "cdp_d2h")
            self.sync()
        finally:
            for d in (d_v, d_m, d_a, d_s):
                lib.cdp_dev_free(h, d)
        return bytes(oa), bytes(os_)

    # ---- fixed-base MSM over a digit table of CRS points ------------------------------------------------
    def fixed_table_create(self, points: bytes, window_bits: int = 16) -> "FixedTable":
        return FixedTable(self, points, window_bits)

    def msm_fixed(self, table: "FixedTable", base_off: int, scalars: bytes) -> bytes:
        """`util::msm(&crs_points[base_off .. base_off + n], scalars)` through the digit table."""
        if len(scalars) % SCALAR_BYTES:
            raise ValueError("malformed input length")
        n = len(scalars) // SCALAR_BYTES
        out = (ctypes.c_uint8 * JACOBIAN_BYTES)()
        self._check(self._lib.cdp_msm_fixed(self._h, table.handle, base_off, _buf(scalars) if n else None, n, out), "cdp_msm_fixed")
        return bytes(out)

    def msm_fixed_batch(self, table: "FixedTable", scalars: bytes, segs: list, var_pts: bytes = b"", tree_max_pairs: int = 0) -> list[bytes]:
        """A batch of cdp_fixed_seg segments over one scalar array (host convenience over cdp_msm_fixed_batch_dev); with
        ``tree_max_pairs`` (>= the longest segment, extra pair included) through cdp_msm_fixed_batch_dev_tree."""
        lib, h = self._lib, self._h
        n_out = max(s.out_idx for s in segs) + 1
        arr = (FixedSeg * len(segs))(*segs)
        d_sc = lib.cdp_dev_alloc(h, max(1, len(scalars)))
        d_vp = lib.cdp_dev_alloc(h, max(1, len(var_pts)))
        d_sg = lib.cdp_dev_alloc(h, ctypes.sizeof(arr))
        d_out = lib.cdp_dev_alloc(h, n_out * JACOBIAN_BYTES)
        try:
            sb = _buf(scalars)
            self._check(lib.cdp_h2d(h, d_sc, sb, len(scalars)), "cdp_h2d")
            self._check(lib.cdp_h2d(h, d_sg, arr, ctypes.sizeof(arr)), "cdp_h2d")
            if var_pts:
                vb = _buf(var_pts)
                self._check(lib.cdp_h2d(h, d_vp, vb, len(var_pts)), "cdp_h2d")
            self.sync()
            if tree_max_pairs:
                self._check(lib.cdp_msm_fixed_batch_dev_tree(h, table.handle, d_sc, d_sg, len(segs), 0, d_vp, d_out, tree_max_pairs), "cdp_msm_fixed_batch_dev_tree")
            else:
                self._check(lib.cdp_msm_fixed_batch_dev(h, table.handle, d_sc, d_sg, len(segs), 0, d_vp, d_out), "cdp_msm_fixed_batch_dev")
            out = (ctypes.c_uint8 * (n_out * JACOBIAN_BYTES))()
            self._check(lib.cdp_d2h(h, out, d_out, n_out * JACOBIAN_BYTES), "cdp_d2h")
            self.sync()
        finally:
            for d in (d_sc, d_sg, d_out, d_vp):
                lib.cdp_dev_free(h, d)
        raw = bytes(out)
        return [raw[i * JACOBIAN_BYTES:(i + 1) * JACOBIAN_BYTES] for i in range(n_out)]

    PROFILE_KINDS = ("msm_buckets", "msm_combine", "smul", "normalize", "other", "msm_fixed", "prove_stage", "transcript")

    def profile_enable(self, on: bool = True):
        self._check(self._lib.cdp_profile_enable(self._h, int(on)), "cdp_profile_enable")

    def profile_reset(self):
        self._check(self._lib.cdp_profile_reset(self._h), "cdp_profile_reset")

    def profile_read(self) -> dict:
        """{kind: {"ms": device time, "launches": count, "units": work items}} accumulated since the last reset."""
        ms = (ctypes.c_double * len(self.PROFILE_KINDS))()
        ln = (c_uint64 * len(self.PROFILE_KINDS))()
        un = (c_uint64 * len(self.PROFILE_KINDS))()
        self._check(self._lib.cdp_profile_read(self._h, ms, ln, un), "cdp_profile_read")
        return {k: {"ms": ms[i], "launches": int(ln[i]), "units": int(un[i])} for i, k in enumerate(self.PROFILE_KINDS)}

    def bench_kernel(self, which: int, blocks: int, threads: int, iters: int) -> float:
        ms = c_float()
        self._check(self._lib.cdp_bench_kernel(self._h, which, blocks, threads, iters, ctypes.byref(ms)), "cdp_bench_kernel")
        return float(ms.value)


class FixedTable:
    """Device-resident digit table of fixed bases (``cdp_fixed_table``): d * 2^(c w) * B for every base, window and digit."""

    def __init__(self, engine: Engine, points: bytes, window_bits: int = 16):
        if len(points) % AFFINE_BYTES or not points:
            raise ValueError("malformed input length")
        self._engine = engine
        h = c_void_p()
        engine._check(engine.lib.cdp_fixed_table_create(engine.handle, _buf(points), len(points) // AFFINE_BYTES, window_bits, ctypes.byref(h)),
                      "cdp_fixed_table_create")
        self._h = h

    @property
    def handle(self):
        return self._h

    @property
    def nbytes(self) -> int:
        return int(self._engine.lib.cdp_fixed_table_bytes(self._h))

    def close(self):
        if getattr(self, "_h", None) and self._engine.handle:
            self._engine.lib.cdp_fixed_table_destroy(self._engine.handle, self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
