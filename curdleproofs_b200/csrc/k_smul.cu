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

// One thread per element.  The two 128-bit GLV halves (k1, k2) are recoded into the joint sparse form (Solinas): digits in {0, +-1},
// on average only every second column non-zero (plain binary: three out of four).  The addends of a column:
//   (+-1, 0) -> +-P = (x, +-y)      (0, +-1) -> +-phi(P) = (beta x, +-y)      +-(1, 1) -> +-(P + phi(P)) = -+phi^2(P) = (beta^2 x, -+y)
// are affine and free; +-(1, -1) -> +-(P - phi(P)) is computed once per element and made affine (one inversion by division steps,
// fp_inv_safegcd.cuh: integer-ALU work beside a multiply-pipe-bound loop), so that its ~16 additions are mixed ones as well (7M + 4S
// instead of 11M + 5S).  So: 128 doublings and ~64 mixed additions per element instead of 128 doublings and ~96.  Within one fold
// job all threads share the scalar, so the column pattern is warp-uniform whenever a job spans whole warps.
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
    // ---- joint sparse form, least significant column first; kept as four bit masks (non-zero / negative, per half), 130 columns
    uint32_t nz0[5] = {0, 0, 0, 0, 0}, sg0[5] = {0, 0, 0, 0, 0}, nz1[5] = {0, 0, 0, 0, 0}, sg1[5] = {0, 0, 0, 0, 0};
    int top = -1;
    {
        uint32_t a0 = g.k1[0], a1 = g.k1[1], a2 = g.k1[2], a3 = g.k1[3], b0 = g.k2[0], b1 = g.k2[1], b2 = g.k2[2], b3 = g.k2[3];
        uint32_t d0 = 0, d1 = 0;
#pragma unroll 1
        for (int col = 0; col < 130; col++) {
            if ((a0 | a1 | a2 | a3 | b0 | b1 | b2 | b3 | d0 | d1) == 0) break;
            const uint32_t l0 = (a0 + d0) & 7, l1 = (b0 + d1) & 7;   // only the three low bits of k + d matter
            int u0 = 0, u1 = 0;
            if (l0 & 1) {
                u0 = 2 - (int)(l0 & 3);                                // l0 mods 4: 1 -> 1, 3 -> -1
                if ((l0 == 3 || l0 == 5) && (l1 & 3) == 2) u0 = -u0;
            }
            if (l1 & 1) {
                u1 = 2 - (int)(l1 & 3);
                if ((l1 == 3 || l1 == 5) && (l0 & 3) == 2) u1 = -u1;
            }
            if (2 * (int)d0 == 1 + u0) d0 = 1 - d0;
            if (2 * (int)d1 == 1 + u1) d1 = 1 - d1;
            const uint32_t bit = 1u << (col & 31);
            if (u0) { nz0[col >> 5] |= bit; if (u0 < 0) sg0[col >> 5] |= bit; }
            if (u1) { nz1[col >> 5] |= bit; if (u1 < 0) sg1[col >> 5] |= bit; }
            if (u0 | u1) top = col;
            a0 = (a0 >> 1) | (a1 << 31); a1 = (a1 >> 1) | (a2 << 31); a2 = (a2 >> 1) | (a3 << 31); a3 >>= 1;
            b0 = (b0 >> 1) | (b1 << 31); b1 = (b1 >> 1) | (b2 << 31); b2 = (b2 >> 1) | (b3 << 31); b3 >>= 1;
        }
    }
    g1a P;
    g1a_load(P, pts + 24 * ((size_t)job.src_off + e));
    fp bx, bbx, ny;
    fp_mul_beta(bx, P.x);
    fp_mul_beta(bbx, bx);
    fp_neg(ny, P.y);
    // D = P - phi(P) = P + (beta x, -y), affine; infinity stays infinity (x = y = 0)
    g1a D;
    {
        g1j Pj, Dj;
        g1j_from_affine(Pj, P);
        g1a nphi;
        nphi.x = bx; nphi.y = ny;
        if (g1a_is_inf(P)) g1a_set_inf(nphi);
        g1j_add_mixed(Dj, Pj, nphi);
        fp zi, zi2;
        fp_inv(zi, Dj.Z);  // 0 -> 0: infinity comes out as (0, 0)
        fp_sqr(zi2, zi);
        fp_mul(D.x, Dj.X, zi2);
        fp_mul(zi2, zi2, zi);
        fp_mul(D.y, Dj.Y, zi2);
    }
    g1j acc;
    g1j_set_inf(acc);
    // column -1 is the optional final "+ L" (one inlined mixed addition serves both uses)
#pragma unroll 1
    for (int col = top; col >= -1; col--) {
        g1a q;
        bool do_madd = false;
        if (col >= 0) {
            g1j_dbl(acc, acc);
            const uint32_t w = (uint32_t)col >> 5, bit = 1u << (col & 31);
            const bool z0 = (nz0[w] & bit) != 0, z1 = (nz1[w] & bit) != 0, n0 = (sg0[w] & bit) != 0, n1 = (sg1[w] & bit) != 0;
            if (z0 && z1 && n0 != n1) {
                do_madd = true;
                q = D;
                if (n0) fp_neg(q.y, q.y);                              // (-1, +1) = -(P - phi P)
            } else if (z0 || z1) {
                do_madd = true;
                const bool both = z0 && z1;                            // same signs: +-(P + phi P) = (beta^2 x, -+y)
                const bool neg = z0 ? n0 : n1;
                fp_select(q.x, P.x, bx, !z0);
                fp_select(q.x, q.x, bbx, both);
                fp_select(q.y, P.y, ny, neg != both);
            }
        } else {
            do_madd = job.add_off != 0xFFFFFFFFu;
            if (do_madd) g1a_load(q, pts + 24 * ((size_t)job.add_off + e));
        }
        if (do_madd) g1j_add_mixed(acc, acc, q);
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
