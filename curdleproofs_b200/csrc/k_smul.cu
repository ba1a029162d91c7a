// k_smul_jobs: batched scalar multiplication / fold.
//   out_jac[i] = (add ? pts[add_off + e] : O) + scalars[scalar_off + e * scalar_stride] * pts[src_off + e]
// for element e of job j (i = j * elems_per_job + e).  One job = one vector of one proof:
//   fold loops        (L + gamma * R).into_affine()   /root/reference/src/inner_product_argument.rs:174-179,
//                                                      src/same_multiscalar_argument.rs:126-131   (scalar_stride = 0)
//   CRS rescale       beta^-(i+1) * G_i                src/grand_product_argument.rs:92-102       (scalar_stride = 1)
//   shuffling         k * R_i, k * S_i                 src/util.rs:94-95                          (scalar_stride = 0)
#include "launch.h"
#include "scalar.cuh"

namespace cdp {

// One thread per element.  Joint double-and-add over the two 128-bit GLV halves; the three possible addends per step:
//   (1,0) -> P = (x, y)      (0,1) -> phi(P) = (beta x, y)      (1,1) -> P + phi(P) = -phi^2(P) = (beta^2 x, -y)
// so each step is one doubling plus at most one mixed addition.  Within one fold job all threads share the scalar, so
// the add/skip pattern is warp-uniform whenever a job spans whole warps.
template <int OCC>
__global__ void __launch_bounds__(128, OCC) k_smul_jobs(const uint32_t *__restrict__ pts, const uint32_t *__restrict__ scalars,
                                                   const smul_job_t *__restrict__ jobs, uint32_t elems_per_job, uint32_t total,
                                                   uint32_t *__restrict__ out_jac) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const uint32_t j = i / elems_per_job, e = i - j * elems_per_job;
    const smul_job_t job = jobs[j];
    uint32_t k[8];
    {
        const uint4 *sp = reinterpret_cast<const uint4 *>(scalars + 8 * ((size_t)job.scalar_off + (size_t)e * job.scalar_stride));
        uint4 a = sp[0], b = sp[1];
        k[0] = a.x; k[1] = a.y; k[2] = a.z; k[3] = a.w; k[4] = b.x; k[5] = b.y; k[6] = b.z; k[7] = b.w;
    }
    glv_t g;
    glv_split(g, k);
    g1a P;
    g1a_load(P, pts + 24 * ((size_t)job.src_off + e));
    fp bx, bbx, ny;
    fp_mul_beta(bx, P.x);
    fp_mul_beta(bbx, bx);
    fp_neg(ny, P.y);
    g1j acc;
    g1j_set_inf(acc);
    // iteration 128 is the optional final "+ L" (one inlined addition serves both uses)
#pragma unroll 1
    for (int bit = 127; bit >= -1; bit--) {
        g1a q;
        bool do_add;
        if (bit >= 0) {
            g1j_dbl(acc, acc);
            uint32_t b1 = g.k1[3] >> 31, b2 = g.k2[3] >> 31;
            g.k1[3] = (g.k1[3] << 1) | (g.k1[2] >> 31); g.k1[2] = (g.k1[2] << 1) | (g.k1[1] >> 31);
            g.k1[1] = (g.k1[1] << 1) | (g.k1[0] >> 31); g.k1[0] <<= 1;
            g.k2[3] = (g.k2[3] << 1) | (g.k2[2] >> 31); g.k2[2] = (g.k2[2] << 1) | (g.k2[1] >> 31);
            g.k2[1] = (g.k2[1] << 1) | (g.k2[0] >> 31); g.k2[0] <<= 1;
            do_add = (b1 | b2) != 0;
            fp_select(q.x, P.x, bx, b2 != 0);
            fp_select(q.x, q.x, bbx, (b1 & b2) != 0);
            fp_select(q.y, P.y, ny, (b1 & b2) != 0);
        } else {
            do_add = job.add_off != 0xFFFFFFFFu;
            if (do_add) g1a_load(q, pts + 24 * ((size_t)job.add_off + e));
        }
        if (do_add) g1j_add_mixed(acc, acc, q);
    }
    g1j_store(out_jac + 36 * (size_t)i, acc);
}

cudaError_t launch_smul_jobs(cudaStream_t st, const uint32_t *pts, const uint32_t *scalars, const smul_job_t *jobs, uint32_t n_jobs,
                             uint32_t elems_per_job, uint32_t *out_jac) {
    uint32_t total = n_jobs * elems_per_job;
    if (total == 0) return cudaSuccess;
    const int occ = tuned_occupancy("CDP_OCC_SMUL", 3);
    if (occ == 5) k_smul_jobs<5><<<(total + 127) / 128, 128, 0, st>>>(pts, scalars, jobs, elems_per_job, total, out_jac);
    else if (occ == 4) k_smul_jobs<4><<<(total + 127) / 128, 128, 0, st>>>(pts, scalars, jobs, elems_per_job, total, out_jac);
    else k_smul_jobs<3><<<(total + 127) / 128, 128, 0, st>>>(pts, scalars, jobs, elems_per_job, total, out_jac);
    return cudaGetLastError();
}

}  // namespace cdp
