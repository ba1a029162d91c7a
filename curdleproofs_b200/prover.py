"""ctypes binding of include/cdp_prover.h: the batched prover, mirroring `CurdleproofsProof::new`
(/root/reference/src/curdleproofs.rs:59-184) for a batch of independent shuffles."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_size_t, c_uint32, c_uint64, c_void_p

from .engine import AFFINE_BYTES, JACOBIAN_BYTES, SCALAR_BYTES, CdpError, Engine, _HERE, load_library


class _ProveInputs(ctypes.Structure):
    _fields_ = [("vec_R", c_void_p), ("vec_S", c_void_p), ("vec_T", c_void_p), ("vec_U", c_void_p), ("M", c_void_p),
                ("permutation", c_void_p), ("k", c_void_p), ("vec_m_blinders", c_void_p), ("rng_seed", c_void_p),
                ("rng_skip_words", c_void_p), ("rng_key", c_void_p)]


_PLIB = None


def load_prover_library() -> ctypes.CDLL:
    global _PLIB
    if _PLIB is None:
        load_library()  # libcdp_b200.so first (fails loudly if missing)
        path = os.path.join(_HERE, "libcdp_prover.so")
        if not os.path.exists(path):
            raise CdpError(f"{path} is missing: run __graft_entry__.build()")
        lib = ctypes.CDLL(path)
        lib.cdp_proof_size.restype = c_size_t
        lib.cdp_proof_size.argtypes = [c_size_t]
        lib.cdp_prover_create.restype = c_int
        lib.cdp_prover_create.argtypes = [POINTER(c_void_p), c_void_p, c_size_t, c_void_p, c_size_t, c_int]
        lib.cdp_prover_create_lanes.restype = c_int
        lib.cdp_prover_create_lanes.argtypes = [POINTER(c_void_p), c_void_p, c_size_t, c_void_p, c_size_t, c_int, c_int]
        lib.cdp_prover_lane_count.restype = c_int
        lib.cdp_prover_lane_count.argtypes = [c_void_p]
        lib.cdp_prover_lane_ctx.restype = c_void_p
        lib.cdp_prover_lane_ctx.argtypes = [c_void_p, c_int]
        lib.cdp_prover_destroy.argtypes = [c_void_p]
        lib.cdp_prover_last_error.restype = c_char_p
        lib.cdp_prover_last_error.argtypes = [c_void_p]
        lib.cdp_prove_batch.restype = c_int
        lib.cdp_prove_batch.argtypes = [c_void_p, c_size_t, POINTER(_ProveInputs), c_void_p]
        lib.cdp_prover_last_timing.argtypes = [c_void_p, POINTER(c_double)]
        lib.cdp_whisk_shuffle_proof_size.restype = c_size_t
        lib.cdp_whisk_shuffle_proof_size.argtypes = [c_size_t]
        lib.cdp_whisk_generate_shuffle_proofs.restype = c_int
        lib.cdp_whisk_generate_shuffle_proofs.argtypes = [c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
        lib.cdp_whisk_verify_shuffle_proofs.restype = c_int
        lib.cdp_whisk_verify_shuffle_proofs.argtypes = [c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
        lib.cdp_prover_last_traffic.argtypes = [c_void_p, POINTER(c_uint64)]
        lib.cdp_prover_set_serial.argtypes = [c_void_p, c_int]
        lib.cdp_prover_table_bytes.restype = c_size_t
        lib.cdp_prover_table_bytes.argtypes = [c_void_p]
        _PLIB = lib
    return _PLIB


def _arr(b: bytes):
    return (ctypes.c_uint8 * max(1, len(b))).from_buffer_copy(b if len(b) else b"\0")


class BatchProver:
    """`CurdleproofsProof::new` for `batch` shuffles at a time on one GPU.

    crs_points: ell + 7 affine points, `CurdleproofsCrs::from_points` order (src/crs.rs:37-58)."""

    def __init__(self, engine: Engine, ell: int, crs_points: bytes, max_batch: int, host_threads: int = 0, lanes: int = 0):
        self._lib = load_prover_library()
        self.engine = engine
        self.ell = ell
        self.max_batch = max_batch
        self.proof_size = int(self._lib.cdp_proof_size(ell))
        if len(crs_points) != (ell + 7) * AFFINE_BYTES:
            raise ValueError("crs_points must hold ell + 7 affine points")
        h = c_void_p()
        rc = self._lib.cdp_prover_create_lanes(ctypes.byref(h), engine.handle, ell, _arr(crs_points), max_batch, host_threads, lanes)
        if rc != 0:
            raise CdpError(f"cdp_prover_create failed (code {rc})")
        self._h = h
        self.lanes = int(self._lib.cdp_prover_lane_count(h))
        self._lane_ctx = [c_void_p(self._lib.cdp_prover_lane_ctx(h, i)) for i in range(self.lanes)]

    def close(self):
        if getattr(self, "_h", None):
            self._lib.cdp_prover_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def prove_batch(self, instances, rng_seeds=None, rng_skip_words=None, rng_keys=None) -> list[bytes]:
        """instances: list of dicts with R, S, T, U (ell affine each), M (jacobian), perm (list of int), k, m_blinders.
        The reference's `rng` argument, per proof: rng_keys[i] (32 bytes, `StdRng::from_seed`) or rng_seeds[i] (`seed_from_u64`: test vectors
        only, 64 bits of entropy); with neither, every proof draws a fresh key from the operating system."""
        B = len(instances)
        ell = self.ell
        cat = lambda key: b"".join(i[key] for i in instances)  # noqa: E731
        bufs = dict(R=_arr(cat("R")), S=_arr(cat("S")), T=_arr(cat("T")), U=_arr(cat("U")), M=_arr(cat("M")), k=_arr(cat("k")),
                    mb=_arr(cat("m_blinders")))
        perm = (c_uint32 * (B * ell))(*[x for i in instances for x in i["perm"]])
        seeds = (c_uint64 * B)(*rng_seeds) if rng_seeds is not None else None
        skips = (c_uint64 * B)(*rng_skip_words) if rng_skip_words is not None else None
        keys = _arr(b"".join(rng_keys)) if rng_keys is not None else None
        return self.prove_raw(B, bufs["R"], bufs["S"], bufs["T"], bufs["U"], bufs["M"], perm, bufs["k"], bufs["mb"], seeds, skips, keys=keys)

    def prove_raw(self, B, R, S, T, U, M, perm, k, mb, seeds, skips=None, out=None, split=True, keys=None):
        # R is None: reuse the instance batch staged in HBM by the previous call ("inputs already resident" mode)
        vp = lambda x: ctypes.cast(x, c_void_p) if x is not None else None  # noqa: E731
        inp = _ProveInputs(vp(R), vp(S), vp(T), vp(U), vp(M), ctypes.cast(perm, c_void_p), ctypes.cast(k, c_void_p), ctypes.cast(mb, c_void_p),
                           vp(seeds), vp(skips), vp(keys))
        if out is None:
            out = (ctypes.c_uint8 * (B * self.proof_size))()
        rc = self._lib.cdp_prove_batch(self._h, B, ctypes.byref(inp), out)
        if rc != 0:
            raise CdpError(f"cdp_prove_batch failed (code {rc}): {self._lib.cdp_prover_last_error(self._h).decode()}")
        if not split:
            return out
        raw = bytes(out)
        return [raw[i * self.proof_size:(i + 1) * self.proof_size] for i in range(B)]

    def whisk_generate_shuffle_proofs(self, pre_trackers: list, rng_seeds, rng_skip_words=None):
        """`generate_whisk_shuffle_proof` (/root/reference/src/whisk.rs:144-179) for a batch.  pre_trackers[b] = ell * 96 bytes
        (r_G || k_r_G per tracker).  Returns (post_trackers, whisk_shuffle_proof_bytes) per shuffle."""
        B, ell = len(pre_trackers), self.ell
        wsz = int(self._lib.cdp_whisk_shuffle_proof_size(ell))
        pre = _arr(b"".join(pre_trackers))
        seeds = (c_uint64 * B)(*rng_seeds)
        skips = (c_uint64 * B)(*rng_skip_words) if rng_skip_words is not None else None
        post = (ctypes.c_uint8 * (B * ell * 96))()
        out = (ctypes.c_uint8 * (B * wsz))()
        rc = self._lib.cdp_whisk_generate_shuffle_proofs(self._h, B, pre, seeds, skips, post, out)
        if rc != 0:
            raise CdpError(f"cdp_whisk_generate_shuffle_proofs failed (code {rc}): {self._lib.cdp_prover_last_error(self._h).decode()}")
        post, out = bytes(post), bytes(out)
        return [(post[b * ell * 96:(b + 1) * ell * 96], out[b * wsz:(b + 1) * wsz]) for b in range(B)]

    # ---- accounting across the lanes' contexts (lane 0 is the caller's Engine) ----
    @property
    def launch_count(self) -> int:
        return sum(int(self.engine.lib.cdp_launch_count(c)) for c in self._lane_ctx)

    def set_serial(self, on: bool):
        """Diagnostics: run the lanes one after the other, so that per-kernel profile times are those of kernels running alone."""
        self._lib.cdp_prover_set_serial(self._h, int(on))

    @property
    def table_bytes(self) -> int:
        return int(self._lib.cdp_prover_table_bytes(self._h))

    def profile_enable(self, on: bool = True):
        for c in self._lane_ctx:
            self.engine.lib.cdp_profile_enable(c, int(on))

    def profile_reset(self):
        for c in self._lane_ctx:
            self.engine.lib.cdp_profile_reset(c)

    def profile_read(self) -> dict:
        """Per-kernel device time summed over lanes (lanes overlap on the GPU, so the sum can exceed the wall time)."""
        tot = {k: {"ms": 0.0, "launches": 0, "units": 0} for k in Engine.PROFILE_KINDS}
        for c in self._lane_ctx:
            ms = (c_double * len(Engine.PROFILE_KINDS))()
            ln = (c_uint64 * len(Engine.PROFILE_KINDS))()
            un = (c_uint64 * len(Engine.PROFILE_KINDS))()
            self.engine.lib.cdp_profile_read(c, ms, ln, un)
            for i, k in enumerate(Engine.PROFILE_KINDS):
                tot[k]["ms"] += ms[i]
                tot[k]["launches"] += int(ln[i])
                tot[k]["units"] += int(un[i])
        return tot

    def last_traffic(self) -> dict:
        t = (c_uint64 * 2)()
        self._lib.cdp_prover_last_traffic(self._h, t)
        return {"h2d_bytes": int(t[0]), "d2h_bytes": int(t[1])}

    def last_timing(self) -> dict:
        t = (c_double * 4)()
        self._lib.cdp_prover_last_timing(self._h, t)
        return {"total_ms": t[0], "host_ms": t[1], "gpu_wait_ms": t[2], "copy_issue_ms": t[3]}


def whisk_generate_tracker_proofs(engine: Engine, trackers: list, ks: list, rng_seeds, rng_skip_words=None) -> list:
    """`generate_whisk_tracker_proof` (/root/reference/src/whisk.rs:231-263) for a batch: 128-byte proofs A || B || s."""
    lib = load_prover_library()
    lib.cdp_whisk_generate_tracker_proofs.restype = c_int
    lib.cdp_whisk_generate_tracker_proofs.argtypes = [c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    B = len(trackers)
    out = (ctypes.c_uint8 * (128 * B))()
    seeds = (c_uint64 * B)(*rng_seeds)
    skips = (c_uint64 * B)(*rng_skip_words) if rng_skip_words is not None else None
    rc = lib.cdp_whisk_generate_tracker_proofs(engine.handle, B, _arr(b"".join(trackers)), _arr(b"".join(ks)), seeds, skips, out)
    if rc != 0:
        raise CdpError(f"cdp_whisk_generate_tracker_proofs failed (code {rc})")
    raw = bytes(out)
    return [raw[128 * i:128 * i + 128] for i in range(B)]


def whisk_verify_tracker_proofs(engine: Engine, trackers: list, k_commitments: list, proofs: list) -> list:
    """`is_valid_whisk_tracker_proof` (/root/reference/src/whisk.rs:183-229) for a batch: 1 valid, 0 invalid, 2 does not deserialise."""
    lib = load_prover_library()
    lib.cdp_whisk_verify_tracker_proofs.restype = c_int
    lib.cdp_whisk_verify_tracker_proofs.argtypes = [c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p]
    B = len(proofs)
    out = (ctypes.c_uint8 * B)()
    rc = lib.cdp_whisk_verify_tracker_proofs(engine.handle, B, _arr(b"".join(trackers)), _arr(b"".join(k_commitments)), _arr(b"".join(proofs)), out)
    if rc != 0:
        raise CdpError(f"cdp_whisk_verify_tracker_proofs failed (code {rc})")
    return list(out)


class _VerifyInputs(ctypes.Structure):
    _fields_ = [("vec_R", c_void_p), ("vec_S", c_void_p), ("vec_T", c_void_p), ("vec_U", c_void_p), ("M", c_void_p), ("proofs", c_void_p),
                ("rng_seed", c_void_p)]


class BatchVerifier:
    """`CurdleproofsProof::deserialize` + `verify` (/root/reference/src/curdleproofs.rs:197-323) for a batch of proofs on one GPU.
    Results: 1 = Ok(()), 0 = Err(VerificationError), 2 = proof does not deserialise."""

    def __init__(self, engine: Engine, ell: int, crs_points: bytes, max_batch: int, host_threads: int = 0, lanes: int = 0):
        self._lib = lib = load_prover_library()
        lib.cdp_verifier_create.restype = c_int
        lib.cdp_verifier_create.argtypes = [POINTER(c_void_p), c_void_p, c_size_t, c_void_p, c_size_t, c_int, c_int]
        lib.cdp_verifier_destroy.argtypes = [c_void_p]
        lib.cdp_verifier_last_error.restype = c_char_p
        lib.cdp_verifier_last_error.argtypes = [c_void_p]
        lib.cdp_verifier_last_timing.argtypes = [c_void_p, POINTER(c_double)]
        lib.cdp_verify_batch.restype = c_int
        lib.cdp_verify_batch.argtypes = [c_void_p, c_size_t, POINTER(_VerifyInputs), c_void_p]
        self.engine, self.ell, self.max_batch = engine, ell, max_batch
        self.proof_size = int(lib.cdp_proof_size(ell))
        if len(crs_points) != (ell + 7) * AFFINE_BYTES:
            raise ValueError("crs_points must hold ell + 7 affine points")
        h = c_void_p()
        rc = lib.cdp_verifier_create(ctypes.byref(h), engine.handle, ell, _arr(crs_points), max_batch, host_threads, lanes)
        if rc != 0:
            raise CdpError(f"cdp_verifier_create failed (code {rc})")
        self._h = h
        lib.cdp_verifier_lane_count.restype = c_int
        lib.cdp_verifier_lane_count.argtypes = [c_void_p]
        self.lanes = int(lib.cdp_verifier_lane_count(h))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.cdp_verifier_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def merge_stats(self) -> dict:
        """Lane sub-batches accepted by the merged accumulated check / re-checked proof by proof, since creation."""
        out = (c_uint64 * 2)()
        self._lib.cdp_verifier_merge_stats.argtypes = [c_void_p, POINTER(c_uint64)]
        self._lib.cdp_verifier_merge_stats.restype = None
        self._lib.cdp_verifier_merge_stats(self._h, out)
        return {"merged": int(out[0]), "fallback": int(out[1])}

    def last_timing(self) -> dict:
        t = (c_double * 3)()
        self._lib.cdp_verifier_last_timing(self._h, t)
        return {"total_ms": t[0], "host_ms": t[1], "gpu_wait_ms": t[2]}

    def verify_batch(self, instances, proofs, rng_seeds=None, comm=None) -> list[int]:
        B = len(instances)
        cat = lambda key: b"".join(i[key] for i in instances)  # noqa: E731
        R, S, T, U, M = (_arr(cat(k)) for k in ("R", "S", "T", "U", "M"))
        P = _arr(b"".join(proofs))
        seeds = (c_uint64 * B)(*rng_seeds) if rng_seeds is not None else None
        return list(self.verify_raw(B, R, S, T, U, M, P, seeds, comm=comm))

    def whisk_verify_shuffle_proofs(self, pre_trackers: list, post_trackers: list, proofs: list, rng_seeds=None) -> list:
        """`is_valid_whisk_shuffle_proof` (/root/reference/src/whisk.rs:106-130) for a batch: 1 valid, 0 invalid, 2 does not deserialise."""
        B = len(proofs)
        seeds = (c_uint64 * B)(*rng_seeds) if rng_seeds is not None else None
        out = (ctypes.c_uint8 * B)()
        rc = self._lib.cdp_whisk_verify_shuffle_proofs(self._h, B, _arr(b"".join(pre_trackers)), _arr(b"".join(post_trackers)),
                                                       _arr(b"".join(proofs)), seeds, out)
        if rc != 0:
            raise CdpError(f"cdp_whisk_verify_shuffle_proofs failed (code {rc}): {self._lib.cdp_verifier_last_error(self._h).decode()}")
        return list(out)

    def global_stats(self) -> dict:
        """Sharded calls accepted by the cross-rank accumulated sum / decided locally instead, since creation."""
        out = (c_uint64 * 2)()
        self._lib.cdp_verifier_global_stats.argtypes = [c_void_p, POINTER(c_uint64)]
        self._lib.cdp_verifier_global_stats.restype = None
        self._lib.cdp_verifier_global_stats(self._h, out)
        return {"accepted": int(out[0]), "local": int(out[1])}

    def verify_raw(self, B, R, S, T, U, M, proofs, seeds=None, out=None, comm=None):
        """comm (curdleproofs_b200.sharded.Comm on this verifier's Engine): the accumulated check sharded over the ranks
        (cdp_verify_batch_sharded): one all-gather of partial sums decides the proofs of all ranks; a collective call."""
        vp = lambda x: ctypes.cast(x, c_void_p) if x is not None else None  # noqa: E731
        inp = _VerifyInputs(vp(R), vp(S), vp(T), vp(U), vp(M), vp(proofs), vp(seeds))
        if out is None:
            out = (ctypes.c_uint8 * B)()
        if comm is not None:
            self._lib.cdp_verify_batch_sharded.restype = c_int
            self._lib.cdp_verify_batch_sharded.argtypes = [c_void_p, c_void_p, c_size_t, POINTER(_VerifyInputs), c_void_p]
            rc = self._lib.cdp_verify_batch_sharded(self._h, comm._h, B, ctypes.byref(inp), out)
            if rc != 0:
                raise CdpError(f"cdp_verify_batch_sharded failed (code {rc}): {self._lib.cdp_verifier_last_error(self._h).decode()}")
            return out
        rc = self._lib.cdp_verify_batch(self._h, B, ctypes.byref(inp), out)
        if rc != 0:
            raise CdpError(f"cdp_verify_batch failed (code {rc}): {self._lib.cdp_verifier_last_error(self._h).decode()}")
        return out
