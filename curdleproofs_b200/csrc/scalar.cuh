// Scalar recoding helpers shared by the kernels (GLV split, endomorphism constant).
#pragma once
#include "g1.cuh"
#include "glv_split.cuh"

namespace cdp {

// GLV split (k = k2 * lambda + k1, both halves below 2^128): glv_split.cuh

__device__ __forceinline__ void fp_mul_beta(fp &r, const fp &a) {
    fp beta;
#pragma unroll
    for (int i = 0; i < 12; i++) beta.v[i] = FP_BETA_MONT[i];
    fp_mul(r, a, beta);
}

}  // namespace cdp
