#include "launch.h"
#include "scalar.cuh"

namespace cdp {

// ------------------------------------------------------------------------------------------------ k_smul_add
// One thread per element.  out_jac[i] = (add_pts ? add_pts[i] : O) + scalars[sidx ? sidx[i] : (bcast ? 0 : i)] * pts[i].
// Joint double-and-add over the two 128-bit GLV halves; the three possible addends per step are
//   (1,0) -> P = (x, y)      (0,1) -> phi(P) = (beta x, y)      (1,1) -> P + phi(P) = -phi^2(P) = (beta^2 x, -y)
// so each step is one doubling plus at most one mixed addition.  Within one fold all threads of a proof share the
// scalar, so the add/skip pattern is warp-uniform.
__global__ void __launch_bounds__(128) k_smul_add(const uint32_t *__restrict__ pts, const uint32_t *__restrict__ scalars,
                                                  const uint32_t *__restrict__ sidx, int bcast,
                                                  const uint32_t *__restrict__ add_pts, uint32_t *__restrict__ out_jac, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t si = sidx ? sidx[i] : (bcast ? 0u : i);
    uint32_t k[8];
    {
        const uint4 *sp = reinterpret_cast<const uint4 *>(scalars + 8 * (size_t)si);
        uint4 a = sp[0], b = sp[1];
        k[0] = a.x; k[1] = a.y; k[2] = a.z; k[3] = a.w; k[4] = b.x; k[5] = b.y; k[6] = b.z; k[7] = b.w;
    }
    glv_t g;
    glv_split(g, k);
    g1a P;
    g1a_load(P, pts + 24 * (size_t)i);
    fp bx, bbx, ny;
    fp_mul_beta(bx, P.x);
    fp_mul_beta(bbx, bx);
    fp_neg(ny, P.y);
    g1j acc;
    g1j_set_inf(acc);
#pragma unroll 1
    for (int bit = 127; bit >= 0; bit--) {
        g1j_dbl(acc, acc);
        uint32_t b1 = g.k1[3] >> 31, b2 = g.k2[3] >> 31;
        g.k1[3] = (g.k1[3] << 1) | (g.k1[2] >> 31); g.k1[2] = (g.k1[2] << 1) | (g.k1[1] >> 31);
        g.k1[1] = (g.k1[1] << 1) | (g.k1[0] >> 31); g.k1[0] <<= 1;
        g.k2[3] = (g.k2[3] << 1) | (g.k2[2] >> 31); g.k2[2] = (g.k2[2] << 1) | (g.k2[1] >> 31);
        g.k2[1] = (g.k2[1] << 1) | (g.k2[0] >> 31); g.k2[0] <<= 1;
        if (b1 | b2) {
            g1a q;
            fp_select(q.x, P.x, bx, b2 != 0);
            fp_select(q.x, q.x, bbx, (b1 & b2) != 0);
            fp_select(q.y, P.y, ny, (b1 & b2) != 0);
            g1j_add_mixed(acc, acc, q);
        }
    }
    if (add_pts) {
        g1a A;
        g1a_load(A, add_pts + 24 * (size_t)i);
        g1j_add_mixed(acc, acc, A);
    }
    g1j_store(out_jac + 36 * (size_t)i, acc);
}

cudaError_t launch_smul_add(cudaStream_t st, const uint32_t *pts, const uint32_t *scalars, const uint32_t *sidx, int bcast,
                            const uint32_t *add_pts, uint32_t *out_jac, uint32_t n) {
    k_smul_add<<<(n + 127) / 128, 128, 0, st>>>(pts, scalars, sidx, bcast, add_pts, out_jac, n);
    return cudaGetLastError();
}

}  // namespace cdp
