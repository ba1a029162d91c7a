"""curdleproofs_b200 -- B200 (sm_100a) engine for the Curdleproofs MSM / fold hot path.

The product is the C-ABI library ``libcdp_b200.so`` (``include/cdp_msm.h``); this package is the thin
ctypes mirror of the reference's operator interface for that path:

    reference (Rust)                                   here
    ------------------------------------------------   -------------------------------
    util::msm(points, scalars)            util.rs:19    Engine.msm(points, scalars)
    util::msm(&crs.vec_G[a..b], scalars)  (CRS bases)   Engine.msm_fixed(table, a, scalars)
    util::msm_from_projective(...)        util.rs:25    Engine.msm_from_projective(...)
    fold loops  (L + gamma*R).into_affine()             Engine.fold(L, R, gamma)
    (s_i * P_i).into_affine()                           Engine.scalar_mul_batch(points, scalars)
    G1Projective::normalize_batch         util.rs:27    Engine.normalize_batch(points)
    serialize_compressed                                Engine.compress_batch(points)
    CurdleproofsCrs / CurdleproofsCrsHex  crs.rs:17-142   CurdleproofsCrs.from_points / to_json / from_json

There is no CPU fallback: constructing an ``Engine`` without the built extension or without a CUDA
device raises.
"""
from .engine import Engine, CdpError, FixedSeg, FixedTable, lib_path, load_library  # noqa: F401
from .crs import CrsError, CurdleproofsCrs  # noqa: F401
from .prover import (BatchProver, BatchVerifier, load_prover_library, whisk_generate_tracker_proofs,  # noqa: F401
                     whisk_verify_tracker_proofs)

__all__ = ["Engine", "CdpError", "FixedSeg", "FixedTable", "lib_path", "load_library", "BatchProver", "BatchVerifier", "load_prover_library",
           "whisk_generate_tracker_proofs", "whisk_verify_tracker_proofs", "CurdleproofsCrs", "CrsError"]
