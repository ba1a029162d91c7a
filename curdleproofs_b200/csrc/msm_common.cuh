// Small device helpers shared by the MSM kernels.
#pragma once
#include "launch.h"
#include "scalar.cuh"

namespace cdp {

__device__ __forceinline__ void shfl_down_g1j(g1j &o, const g1j &p, int delta, int width) {
#pragma unroll
    for (int i = 0; i < 12; i++) {
        o.X.v[i] = __shfl_down_sync(0xffffffffu, p.X.v[i], delta, width);
        o.Y.v[i] = __shfl_down_sync(0xffffffffu, p.Y.v[i], delta, width);
        o.Z.v[i] = __shfl_down_sync(0xffffffffu, p.Z.v[i], delta, width);
    }
}

}  // namespace cdp
