// BLS12-381 base field Fp for sm_100a: 12 x 32-bit limbs held in registers, Montgomery form (R = 2^384),
// values always fully reduced to [0, p).
//
// This is the device counterpart of what the reference obtains from ark-ff's `Fp384` (Montgomery, 6 x u64 --
// the same bytes as 12 x u32 little-endian) underneath every group operation reached from
// `util::msm` (/root/reference/src/util.rs:19-22) and the fold loops
// (/root/reference/src/inner_product_argument.rs:174-179, src/same_multiscalar_argument.rs:126-131).
//
// Multiplication is a ROW-wise Montgomery product with two interleaved accumulator arrays (even / odd limbs of the multiplicand): every
// row is four independent carry chains of six IMAD.WIDE.U32(.X), carries in the condition-code register, p folded in as immediates --
// 288 wide multiply-adds per product and enough instruction-level parallelism for one warp per scheduler to keep the pipe 63 % busy.
// Squaring is row-wise too (fp_sqr_rows.inc, generated and model-checked by tools/gen_fp_sqr.py): 222 wide multiply-adds.  Both are
// called (register ABI), not inlined, so that the hot set of every kernel fits the instruction cache.
// No tensor cores: 381-bit modular integer arithmetic is not a dense contraction.
#pragma once
#include <stdint.h>
#include "constants.cuh"
#include "fp_inv_euclid.cuh"
#include "fp_inv_safegcd.cuh"

namespace cdp {

struct fp {
    uint32_t v[12];
};

// p as immediates (ptxas folds these into the instruction stream / constant bank)
#define CDP_P(i) (fp_p_limb(i))

__device__ __forceinline__ void fp_set_zero(fp &r) {
#pragma unroll
    for (int i = 0; i < 12; i++) r.v[i] = 0;
}
__device__ __forceinline__ void fp_set_one(fp &r) {
#pragma unroll
    for (int i = 0; i < 12; i++) r.v[i] = fp_r_mod_p_limb(i);
}
__device__ __forceinline__ bool fp_is_zero(const fp &a) {
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) acc |= a.v[i];
    return acc == 0;
}
__device__ __forceinline__ bool fp_eq(const fp &a, const fp &b) {
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) acc |= a.v[i] ^ b.v[i];
    return acc == 0;
}
__device__ __forceinline__ void fp_select(fp &r, const fp &a, const fp &b, bool take_b) {
#pragma unroll
    for (int i = 0; i < 12; i++) r.v[i] = take_b ? b.v[i] : a.v[i];
}

// t = a - p, returns borrow (1 when a < p)
__device__ __forceinline__ uint32_t fp_sub_p(fp &t, const fp &a) {
    uint32_t borrow;
    asm("sub.cc.u32 %0, %13, %25;\n\t"
        "subc.cc.u32 %1, %14, %26;\n\t"
        "subc.cc.u32 %2, %15, %27;\n\t"
        "subc.cc.u32 %3, %16, %28;\n\t"
        "subc.cc.u32 %4, %17, %29;\n\t"
        "subc.cc.u32 %5, %18, %30;\n\t"
        "subc.cc.u32 %6, %19, %31;\n\t"
        "subc.cc.u32 %7, %20, %32;\n\t"
        "subc.cc.u32 %8, %21, %33;\n\t"
        "subc.cc.u32 %9, %22, %34;\n\t"
        "subc.cc.u32 %10, %23, %35;\n\t"
        "subc.cc.u32 %11, %24, %36;\n\t"
        "subc.u32 %12, 0, 0;"
        : "=r"(t.v[0]), "=r"(t.v[1]), "=r"(t.v[2]), "=r"(t.v[3]), "=r"(t.v[4]), "=r"(t.v[5]), "=r"(t.v[6]), "=r"(t.v[7]),
          "=r"(t.v[8]), "=r"(t.v[9]), "=r"(t.v[10]), "=r"(t.v[11]), "=r"(borrow)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]), "r"(a.v[8]),
          "r"(a.v[9]), "r"(a.v[10]), "r"(a.v[11]), "r"(CDP_P(0)), "r"(CDP_P(1)), "r"(CDP_P(2)), "r"(CDP_P(3)), "r"(CDP_P(4)),
          "r"(CDP_P(5)), "r"(CDP_P(6)), "r"(CDP_P(7)), "r"(CDP_P(8)), "r"(CDP_P(9)), "r"(CDP_P(10)), "r"(CDP_P(11)));
    return borrow;  // 0xffffffff when a < p, else 0
}

__device__ __forceinline__ void fp_add(fp &r, const fp &a, const fp &b) {
    fp s, t;
    asm("add.cc.u32 %0, %12, %24;\n\t"
        "addc.cc.u32 %1, %13, %25;\n\t"
        "addc.cc.u32 %2, %14, %26;\n\t"
        "addc.cc.u32 %3, %15, %27;\n\t"
        "addc.cc.u32 %4, %16, %28;\n\t"
        "addc.cc.u32 %5, %17, %29;\n\t"
        "addc.cc.u32 %6, %18, %30;\n\t"
        "addc.cc.u32 %7, %19, %31;\n\t"
        "addc.cc.u32 %8, %20, %32;\n\t"
        "addc.cc.u32 %9, %21, %33;\n\t"
        "addc.cc.u32 %10, %22, %34;\n\t"
        "addc.u32 %11, %23, %35;"
        : "=r"(s.v[0]), "=r"(s.v[1]), "=r"(s.v[2]), "=r"(s.v[3]), "=r"(s.v[4]), "=r"(s.v[5]), "=r"(s.v[6]), "=r"(s.v[7]),
          "=r"(s.v[8]), "=r"(s.v[9]), "=r"(s.v[10]), "=r"(s.v[11])
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]), "r"(a.v[8]),
          "r"(a.v[9]), "r"(a.v[10]), "r"(a.v[11]), "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]),
          "r"(b.v[6]), "r"(b.v[7]), "r"(b.v[8]), "r"(b.v[9]), "r"(b.v[10]), "r"(b.v[11]));
    uint32_t borrow = fp_sub_p(t, s);
    fp_select(r, t, s, borrow != 0);
}

__device__ __forceinline__ void fp_sub(fp &r, const fp &a, const fp &b) {
    fp d;
    uint32_t borrow;
    asm("sub.cc.u32 %0, %13, %25;\n\t"
        "subc.cc.u32 %1, %14, %26;\n\t"
        "subc.cc.u32 %2, %15, %27;\n\t"
        "subc.cc.u32 %3, %16, %28;\n\t"
        "subc.cc.u32 %4, %17, %29;\n\t"
        "subc.cc.u32 %5, %18, %30;\n\t"
        "subc.cc.u32 %6, %19, %31;\n\t"
        "subc.cc.u32 %7, %20, %32;\n\t"
        "subc.cc.u32 %8, %21, %33;\n\t"
        "subc.cc.u32 %9, %22, %34;\n\t"
        "subc.cc.u32 %10, %23, %35;\n\t"
        "subc.cc.u32 %11, %24, %36;\n\t"
        "subc.u32 %12, 0, 0;"
        : "=r"(d.v[0]), "=r"(d.v[1]), "=r"(d.v[2]), "=r"(d.v[3]), "=r"(d.v[4]), "=r"(d.v[5]), "=r"(d.v[6]), "=r"(d.v[7]),
          "=r"(d.v[8]), "=r"(d.v[9]), "=r"(d.v[10]), "=r"(d.v[11]), "=r"(borrow)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]), "r"(a.v[8]),
          "r"(a.v[9]), "r"(a.v[10]), "r"(a.v[11]), "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]),
          "r"(b.v[6]), "r"(b.v[7]), "r"(b.v[8]), "r"(b.v[9]), "r"(b.v[10]), "r"(b.v[11]));
    // add back (p & borrow)
    asm("add.cc.u32 %0, %0, %12;\n\t"
        "addc.cc.u32 %1, %1, %13;\n\t"
        "addc.cc.u32 %2, %2, %14;\n\t"
        "addc.cc.u32 %3, %3, %15;\n\t"
        "addc.cc.u32 %4, %4, %16;\n\t"
        "addc.cc.u32 %5, %5, %17;\n\t"
        "addc.cc.u32 %6, %6, %18;\n\t"
        "addc.cc.u32 %7, %7, %19;\n\t"
        "addc.cc.u32 %8, %8, %20;\n\t"
        "addc.cc.u32 %9, %9, %21;\n\t"
        "addc.cc.u32 %10, %10, %22;\n\t"
        "addc.u32 %11, %11, %23;"
        : "+r"(d.v[0]), "+r"(d.v[1]), "+r"(d.v[2]), "+r"(d.v[3]), "+r"(d.v[4]), "+r"(d.v[5]), "+r"(d.v[6]), "+r"(d.v[7]),
          "+r"(d.v[8]), "+r"(d.v[9]), "+r"(d.v[10]), "+r"(d.v[11])
        : "r"(CDP_P(0) & borrow), "r"(CDP_P(1) & borrow), "r"(CDP_P(2) & borrow), "r"(CDP_P(3) & borrow), "r"(CDP_P(4) & borrow),
          "r"(CDP_P(5) & borrow), "r"(CDP_P(6) & borrow), "r"(CDP_P(7) & borrow), "r"(CDP_P(8) & borrow), "r"(CDP_P(9) & borrow),
          "r"(CDP_P(10) & borrow), "r"(CDP_P(11) & borrow));
    r = d;
}
__device__ __forceinline__ void fp_dbl(fp &r, const fp &a) { fp_add(r, a, a); }
__device__ __forceinline__ void fp_neg(fp &r, const fp &a) {
    fp z;
    fp_set_zero(z);
    fp_sub(r, z, a);  // 0 - 0 = 0 stays canonical
}

// acc(96 bit) += x * y
__device__ __forceinline__ void mac96(uint32_t &a0, uint32_t &a1, uint32_t &a2, uint32_t x, uint32_t y) {
    asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
        "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
        "addc.u32 %2, %2, 0;"
        : "+r"(a0), "+r"(a1), "+r"(a2)
        : "r"(x), "r"(y));
}
// acc(96 bit) += 2 * s(96 bit)   (squaring: the column's cross terms are summed once, then doubled here)
__device__ __forceinline__ void acc96_add2(uint32_t &a0, uint32_t &a1, uint32_t &a2, uint32_t s0, uint32_t s1, uint32_t s2) {
    asm("add.cc.u32 %0, %0, %3;\n\t"
        "addc.cc.u32 %1, %1, %4;\n\t"
        "addc.u32 %2, %2, %5;\n\t"
        "add.cc.u32 %0, %0, %3;\n\t"
        "addc.cc.u32 %1, %1, %4;\n\t"
        "addc.u32 %2, %2, %5;"
        : "+r"(a0), "+r"(a1), "+r"(a2)
        : "r"(s0), "r"(s1), "r"(s2));
}

// Montgomery reduction tail shared by mul and sqr is inlined in both (the m[] words stay in registers).
__device__ __forceinline__ void fp_final_sub(fp &r) {
    fp t;
    uint32_t borrow = fp_sub_p(t, r);
    fp_select(r, t, r, borrow != 0);
}

// ---- Montgomery product, row-wise with two interleaved 64-bit-aligned accumulator arrays ("even / odd" form).
// The column-wise product further down funnels every partial product through ONE 96-bit accumulator: 300 multiply-adds in a single
// carry chain, so a warp can issue one only every ~6 cycles and the kernels built on it (168 registers, 3 warps per scheduler) keep the
// integer pipe ~40% busy (profiles/r01_ncu_msm_buckets_v3.txt: issue active 39%, stall_wait 3.5).  Here the running value is
//   T = sum_k X[k] 2^(32k) + sum_k Y[k] 2^(32(k+1))
// and the partial products a_j * b_i of one row fall into X (even j) or Y (odd j) at 64-bit aligned positions, so a row is four
// independent carry chains of six multiply-adds each (a_even*b_i, a_odd*b_i, p_even*m, p_odd*m), and consecutive rows overlap word by
// word.  Dividing by 2^32 after each row swaps the roles of the arrays: the old Y is the new X, and the old X (whose word 0 is now zero),
// moved down two words in place, is the new Y.
#define CDP_MAD6(acc, m0, m1, m2, m3, m4, m5, s, top)                                                                                  \
    asm("mad.lo.cc.u32 %0, %13, %19, %0;\n\t"                                                                                           \
        "madc.hi.cc.u32 %1, %13, %19, %1;\n\t"                                                                                          \
        "madc.lo.cc.u32 %2, %14, %19, %2;\n\t"                                                                                          \
        "madc.hi.cc.u32 %3, %14, %19, %3;\n\t"                                                                                          \
        "madc.lo.cc.u32 %4, %15, %19, %4;\n\t"                                                                                          \
        "madc.hi.cc.u32 %5, %15, %19, %5;\n\t"                                                                                          \
        "madc.lo.cc.u32 %6, %16, %19, %6;\n\t"                                                                                          \
        "madc.hi.cc.u32 %7, %16, %19, %7;\n\t"                                                                                          \
        "madc.lo.cc.u32 %8, %17, %19, %8;\n\t"                                                                                          \
        "madc.hi.cc.u32 %9, %17, %19, %9;\n\t"                                                                                          \
        "madc.lo.cc.u32 %10, %18, %19, %10;\n\t"                                                                                        \
        "madc.hi.cc.u32 %11, %18, %19, %11;\n\t"                                                                                        \
        "addc.u32 %12, %12, 0;"                                                                                                         \
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "+r"(acc[8]), \
          "+r"(acc[9]), "+r"(acc[10]), "+r"(acc[11]), "+r"(top)                                                                         \
        : "r"(m0), "r"(m1), "r"(m2), "r"(m3), "r"(m4), "r"(m5), "r"(s))
// Y[0] += X[1]; then X <- (X >> 64) + a_odd * bi (in place, ascending: every word is read before it is overwritten), carry-in from the add
#define CDP_MAD6_RSHIFT(X, y0, m0, m1, m2, m3, m4, m5, s)                                                                              \
    asm("add.cc.u32 %12, %12, %1;\n\t"                                                                                                 \
        "madc.lo.cc.u32 %0, %13, %19, %2;\n\t"                                                                                          \
        "madc.hi.cc.u32 %1, %13, %19, %3;\n\t"                                                                                          \
        "madc.lo.cc.u32 %2, %14, %19, %4;\n\t"                                                                                          \
        "madc.hi.cc.u32 %3, %14, %19, %5;\n\t"                                                                                          \
        "madc.lo.cc.u32 %4, %15, %19, %6;\n\t"                                                                                          \
        "madc.hi.cc.u32 %5, %15, %19, %7;\n\t"                                                                                          \
        "madc.lo.cc.u32 %6, %16, %19, %8;\n\t"                                                                                          \
        "madc.hi.cc.u32 %7, %16, %19, %9;\n\t"                                                                                          \
        "madc.lo.cc.u32 %8, %17, %19, %10;\n\t"                                                                                         \
        "madc.hi.cc.u32 %9, %17, %19, %11;\n\t"                                                                                         \
        "madc.lo.cc.u32 %10, %18, %19, 0;\n\t"                                                                                          \
        "madc.hi.u32 %11, %18, %19, 0;"                                                                                                 \
        : "+r"(X[0]), "+r"(X[1]), "+r"(X[2]), "+r"(X[3]), "+r"(X[4]), "+r"(X[5]), "+r"(X[6]), "+r"(X[7]), "+r"(X[8]), "+r"(X[9]),       \
          "+r"(X[10]), "+r"(X[11]), "+r"(y0)                                                                                            \
        : "r"(m0), "r"(m1), "r"(m2), "r"(m3), "r"(m4), "r"(m5), "r"(s))
// one Montgomery step on (X aligned, Y offset): m = X[0] * (-p^-1); Y += p_odd * m; X += p_even * m (X[0] becomes 0), carry into Y[11]
#define CDP_REDC_ROW(X, Y)                                                                                                   \
    {                                                                                                                        \
        const uint32_t mm = X[0] * FP_INV32;                                                                                 \
        uint32_t drop = 0;                                                                                                   \
        CDP_MAD6(Y, CDP_P(1), CDP_P(3), CDP_P(5), CDP_P(7), CDP_P(9), CDP_P(11), mm, drop);                                  \
        CDP_MAD6(X, CDP_P(0), CDP_P(2), CDP_P(4), CDP_P(6), CDP_P(8), CDP_P(10), mm, Y[11]);                                 \
    }
// row i >= 1: X / Y are the aligned / offset arrays left by the previous row
#define CDP_MUL_ROW(X, Y, bi)                                                                                                \
    {                                                                                                                        \
        CDP_MAD6_RSHIFT(X, Y[0], a.v[1], a.v[3], a.v[5], a.v[7], a.v[9], a.v[11], bi);                                       \
        CDP_MAD6(Y, a.v[0], a.v[2], a.v[4], a.v[6], a.v[8], a.v[10], bi, X[11]);                                             \
        CDP_REDC_ROW(Y, X)                                                                                                   \
    }
__device__ __forceinline__ void fp_mul_eo(fp &r, const fp &a, const fp &b) {
    uint32_t E[12], O[12];
#pragma unroll
    for (int j = 0; j < 12; j += 2) {
        uint64_t pe = (uint64_t)a.v[j] * b.v[0], po = (uint64_t)a.v[j + 1] * b.v[0];
        E[j] = (uint32_t)pe; E[j + 1] = (uint32_t)(pe >> 32);
        O[j] = (uint32_t)po; O[j + 1] = (uint32_t)(po >> 32);
    }
    CDP_REDC_ROW(E, O)
    CDP_MUL_ROW(E, O, b.v[1]) CDP_MUL_ROW(O, E, b.v[2]) CDP_MUL_ROW(E, O, b.v[3]) CDP_MUL_ROW(O, E, b.v[4])
    CDP_MUL_ROW(E, O, b.v[5]) CDP_MUL_ROW(O, E, b.v[6]) CDP_MUL_ROW(E, O, b.v[7]) CDP_MUL_ROW(O, E, b.v[8])
    CDP_MUL_ROW(E, O, b.v[9]) CDP_MUL_ROW(O, E, b.v[10]) CDP_MUL_ROW(E, O, b.v[11])
    // after row 11 the aligned array is O (its word 0 is zero) and the offset array is E: result = E + (O >> 32)
    fp out;
    asm("add.cc.u32 %0, %12, %24;\n\t"
        "addc.cc.u32 %1, %13, %25;\n\t"
        "addc.cc.u32 %2, %14, %26;\n\t"
        "addc.cc.u32 %3, %15, %27;\n\t"
        "addc.cc.u32 %4, %16, %28;\n\t"
        "addc.cc.u32 %5, %17, %29;\n\t"
        "addc.cc.u32 %6, %18, %30;\n\t"
        "addc.cc.u32 %7, %19, %31;\n\t"
        "addc.cc.u32 %8, %20, %32;\n\t"
        "addc.cc.u32 %9, %21, %33;\n\t"
        "addc.cc.u32 %10, %22, %34;\n\t"
        "addc.u32 %11, %23, 0;"
        : "=r"(out.v[0]), "=r"(out.v[1]), "=r"(out.v[2]), "=r"(out.v[3]), "=r"(out.v[4]), "=r"(out.v[5]), "=r"(out.v[6]), "=r"(out.v[7]),
          "=r"(out.v[8]), "=r"(out.v[9]), "=r"(out.v[10]), "=r"(out.v[11])
        : "r"(E[0]), "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]), "r"(E[8]), "r"(E[9]), "r"(E[10]), "r"(E[11]),
          "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]), "r"(O[7]), "r"(O[8]), "r"(O[9]), "r"(O[10]), "r"(O[11]));
    fp_final_sub(out);
    r = out;
}

// ---- Montgomery squaring, row-wise (generated: tools/gen_fp_sqr.py, which also checks the exact instruction sequence on a register /
// carry-flag model).  The 24-word square is built from the 66 off-diagonal products in even / odd arrays (plain IMAD.WIDE carry
// chains), doubled with funnel shifts, the 12 diagonal squares added in one chain; then twelve even / odd reduction rows on the low
// half and the high half added at the end: 222 wide multiply-adds and ~130 other instructions.  The column-wise squaring below
// (fp_sqr_inl) pays one IADD3.X per product for its third accumulator word -- 234 + ~380 instructions, a quarter of all instructions
// of a mixed addition (profiles/r01_ncu_fixed_msm_v2.txt).
// one reduction row: aligned array X (X[0] == 0 after the previous row), offset array Y; m = new lowest word * (-p^-1)
#define CDP_SQR_RED_ROW(X, Y)                                                                                                \
    {                                                                                                                        \
        const uint32_t mr = (Y[0] + X[1]) * FP_INV32;                                                                        \
        CDP_MAD6_RSHIFT(X, Y[0], CDP_P(1), CDP_P(3), CDP_P(5), CDP_P(7), CDP_P(9), CDP_P(11), mr);                           \
        CDP_MAD6(Y, CDP_P(0), CDP_P(2), CDP_P(4), CDP_P(6), CDP_P(8), CDP_P(10), mr, X[11]);                                 \
    }
__device__ __forceinline__ void fp_sqr_rw(fp &r, const fp &a) {
#include "fp_sqr_rows.inc"
    fp out;
    asm("add.cc.u32 %0, %12, %24;\n\t"
        "addc.cc.u32 %1, %13, %25;\n\t"
        "addc.cc.u32 %2, %14, %26;\n\t"
        "addc.cc.u32 %3, %15, %27;\n\t"
        "addc.cc.u32 %4, %16, %28;\n\t"
        "addc.cc.u32 %5, %17, %29;\n\t"
        "addc.cc.u32 %6, %18, %30;\n\t"
        "addc.cc.u32 %7, %19, %31;\n\t"
        "addc.cc.u32 %8, %20, %32;\n\t"
        "addc.cc.u32 %9, %21, %33;\n\t"
        "addc.cc.u32 %10, %22, %34;\n\t"
        "addc.u32 %11, %23, 0;"
        : "=r"(out.v[0]), "=r"(out.v[1]), "=r"(out.v[2]), "=r"(out.v[3]), "=r"(out.v[4]), "=r"(out.v[5]), "=r"(out.v[6]), "=r"(out.v[7]),
          "=r"(out.v[8]), "=r"(out.v[9]), "=r"(out.v[10]), "=r"(out.v[11])
        : "r"(E[0]), "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]), "r"(E[8]), "r"(E[9]), "r"(E[10]), "r"(E[11]),
          "r"(Y[1]), "r"(Y[2]), "r"(Y[3]), "r"(Y[4]), "r"(Y[5]), "r"(Y[6]), "r"(Y[7]), "r"(Y[8]), "r"(Y[9]), "r"(Y[10]), "r"(Y[11]));
    asm("add.cc.u32 %0, %0, %12;\n\t"
        "addc.cc.u32 %1, %1, %13;\n\t"
        "addc.cc.u32 %2, %2, %14;\n\t"
        "addc.cc.u32 %3, %3, %15;\n\t"
        "addc.cc.u32 %4, %4, %16;\n\t"
        "addc.cc.u32 %5, %5, %17;\n\t"
        "addc.cc.u32 %6, %6, %18;\n\t"
        "addc.cc.u32 %7, %7, %19;\n\t"
        "addc.cc.u32 %8, %8, %20;\n\t"
        "addc.cc.u32 %9, %9, %21;\n\t"
        "addc.cc.u32 %10, %10, %22;\n\t"
        "addc.u32 %11, %11, %23;"
        : "+r"(out.v[0]), "+r"(out.v[1]), "+r"(out.v[2]), "+r"(out.v[3]), "+r"(out.v[4]), "+r"(out.v[5]), "+r"(out.v[6]), "+r"(out.v[7]),
          "+r"(out.v[8]), "+r"(out.v[9]), "+r"(out.v[10]), "+r"(out.v[11])
        : "r"(E[12]), "r"(E[13]), "r"(E[14]), "r"(E[15]), "r"(E[16]), "r"(E[17]), "r"(E[18]), "r"(E[19]), "r"(E[20]), "r"(E[21]), "r"(E[22]),
          "r"(E[23]));
    fp_final_sub(out);
    r = out;
}

__device__ __forceinline__ void fp_mul_inl(fp &r, const fp &a, const fp &b) {
    uint32_t m[12];
    fp out;
    uint32_t a0 = 0, a1 = 0, a2 = 0;
#pragma unroll
    for (int c = 0; c < 12; c++) {
#pragma unroll
        for (int j = 0; j <= c; j++) mac96(a0, a1, a2, a.v[j], b.v[c - j]);
#pragma unroll
        for (int j = 0; j < c; j++) mac96(a0, a1, a2, m[j], CDP_P(c - j));
        m[c] = a0 * FP_INV32;
        mac96(a0, a1, a2, m[c], CDP_P(0));
        a0 = a1; a1 = a2; a2 = 0;
    }
#pragma unroll
    for (int c = 12; c < 24; c++) {
#pragma unroll
        for (int j = c - 11; j < 12; j++) mac96(a0, a1, a2, a.v[j], b.v[c - j]);
#pragma unroll
        for (int j = c - 11; j < 12; j++) mac96(a0, a1, a2, m[j], CDP_P(c - j));
        out.v[c - 12] = a0;
        a0 = a1; a1 = a2; a2 = 0;
    }
    fp_final_sub(out);
    r = out;
}

__device__ __forceinline__ void fp_sqr_inl(fp &r, const fp &a) {
    uint32_t m[12];
    fp out;
    uint32_t a0 = 0, a1 = 0, a2 = 0;
#pragma unroll
    for (int c = 0; c < 12; c++) {
        uint32_t s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
        for (int j = 0; 2 * j < c; j++) mac96(s0, s1, s2, a.v[j], a.v[c - j]);
        if (c > 0) acc96_add2(a0, a1, a2, s0, s1, s2);
        if ((c & 1) == 0) mac96(a0, a1, a2, a.v[c / 2], a.v[c / 2]);
#pragma unroll
        for (int j = 0; j < c; j++) mac96(a0, a1, a2, m[j], CDP_P(c - j));
        m[c] = a0 * FP_INV32;
        mac96(a0, a1, a2, m[c], CDP_P(0));
        a0 = a1; a1 = a2; a2 = 0;
    }
#pragma unroll
    for (int c = 12; c < 24; c++) {
        uint32_t s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
        for (int j = c - 11; 2 * j < c; j++) mac96(s0, s1, s2, a.v[j], a.v[c - j]);
        if (c < 22) acc96_add2(a0, a1, a2, s0, s1, s2);
        if ((c & 1) == 0) mac96(a0, a1, a2, a.v[c / 2], a.v[c / 2]);
#pragma unroll
        for (int j = c - 11; j < 12; j++) mac96(a0, a1, a2, m[j], CDP_P(c - j));
        out.v[c - 12] = a0;
        a0 = a1; a1 = a2; a2 = 0;
    }
    fp_final_sub(out);
    r = out;
}

// Out-of-line entry points.  Fully inlined, a kernel built on these products is 250-330 KB of straight-line SASS and
// stalls on instruction fetch (ncu: smsp__average_warps_issue_stalled_no_instruction 2.5-5 per issue, profiles/r01_*_v0).
// Called, the whole hot working set (mul + sqr + the point-formula glue) is ~20 KB and stays in the 32 KB L1.5 I-cache.
// Arguments and result travel in registers (by-value structs), nothing goes through local memory.
#ifndef CDP_INLINE_FP_MUL
static __device__ __noinline__ fp fp_mul_fn(const fp a, const fp b) {
    fp r;
#ifdef CDP_FP_MUL_COLUMNWISE
    fp_mul_inl(r, a, b);
#else
    fp_mul_eo(r, a, b);
#endif
    return r;
}
static __device__ __noinline__ fp fp_sqr_fn(const fp a) {
    fp r;
#if defined(CDP_FP_SQR_VIA_MUL)
    fp_mul_eo(r, a, a);
#elif defined(CDP_FP_SQR_COLUMNWISE)
    fp_sqr_inl(r, a);
#else
    fp_sqr_rw(r, a);
#endif
    return r;
}
__device__ __forceinline__ void fp_mul(fp &r, const fp &a, const fp &b) { r = fp_mul_fn(a, b); }
__device__ __forceinline__ void fp_sqr(fp &r, const fp &a) { r = fp_sqr_fn(a); }
#else
__device__ __forceinline__ void fp_mul(fp &r, const fp &a, const fp &b) { fp_mul_inl(r, a, b); }
__device__ __forceinline__ void fp_sqr(fp &r, const fp &a) { fp_sqr_inl(r, a); }
#endif

// Montgomery <-> canonical
__device__ __forceinline__ void fp_from_mont(fp &r, const fp &a) {
    fp one;
    fp_set_zero(one);
    one.v[0] = 1;
    fp_mul(r, a, one);
}
__device__ __forceinline__ void fp_to_mont(fp &r, const fp &a) {
    fp r2;
#pragma unroll
    for (int i = 0; i < 12; i++) r2.v[i] = FP_R2_MOD_P[i];
    fp_mul(r, a, r2);
}

// a^e for a fixed 384-bit exponent held in constant memory: uniform control flow across the warp (4-bit fixed window).
static __device__ __noinline__ void fp_pow_fixed(fp &r, const fp &a, const uint32_t *e) {
    fp tbl[16];  // tbl[k] = a^k ; lives in local memory -- indexed dynamically, 15 muls to build
    fp_set_one(tbl[0]);
    tbl[1] = a;
#pragma unroll 1
    for (int k = 2; k < 16; k++) fp_mul(tbl[k], tbl[k - 1], a);
    fp acc;
    fp_set_one(acc);
    bool started = false;
#pragma unroll 1
    for (int w = 95; w >= 0; w--) {  // 96 nibbles of the 384-bit exponent
        uint32_t nib = (e[w >> 3] >> ((w & 7) * 4)) & 0xF;
        if (started) {
            fp_sqr(acc, acc); fp_sqr(acc, acc); fp_sqr(acc, acc); fp_sqr(acc, acc);
        }
        if (nib) {
            fp_mul(acc, acc, tbl[nib]);
            started = true;
        }
    }
    r = acc;
}
// a^(p-2) (Fermat inverse; 0 -> 0)
__device__ __forceinline__ void fp_inv_fermat(fp &r, const fp &a) { fp_pow_fixed(r, a, FP_P_MINUS_2); }
// a^-1 (0 -> 0), Montgomery form in and out: the integer inverse of aR is a^-1 R^-1 (binary Euclid, fp_inv_euclid.cuh: ~10x less latency than
// the Fermat ladder above and off the multiply pipe), one product with R^3 = 2^1152 mod p gives a^-1 R
static __device__ __noinline__ void fp_inv_euclid_fn(fp *r, const fp *a) {
    fp t, r3;
#ifdef CDP_FP_INV_BINARY_EUCLID
    euclid::inverse_int(t.v, a->v, [](bool done) { return done; });
#else
    safegcd::inverse_int(t.v, a->v, [](bool done) { return done; });  // ~10x fewer instructions than the binary Euclid (fp_inv_safegcd.cuh)
#endif
    const uint32_t R3[12] = {0xd94ca1e0u, 0xed48ac6bu, 0x03a7adf8u, 0x315f831eu, 0x615e29ddu, 0x9a53352au,
                             0x921e1761u, 0x34c04e5eu, 0x65724728u, 0x2512d435u, 0x91755d4du, 0x0aa63460u};
#pragma unroll
    for (int i = 0; i < 12; i++) r3.v[i] = R3[i];
    fp_mul(*r, t, r3);
}
__device__ __forceinline__ void fp_inv(fp &r, const fp &a) {
#ifdef CDP_FP_INV_FERMAT
    fp_inv_fermat(r, a);
#else
    fp t = a;
    fp_inv_euclid_fn(&r, &t);
#endif
}
// square root for p = 3 mod 4: candidate a^((p+1)/4); returns false when a is not a square
__device__ __forceinline__ bool fp_sqrt(fp &r, const fp &a) {
    fp s, s2;
    fp_pow_fixed(s, a, FP_P_PLUS_1_DIV_4);
    fp_sqr(s2, s);
    r = s;
    return fp_eq(s2, a);
}

// true when the canonical value of y is > (p-1)/2, i.e. y > -y : the "sign" bit of the ZCash encoding
__device__ __forceinline__ bool fp_canon_is_lexicographically_largest(const fp &y_canon) {
    // compare with (p-1)/2 from the top limb down
    bool gt = false, decided = false;
#pragma unroll
    for (int i = 11; i >= 0; i--) {
        uint32_t h = FP_P_MINUS_1_DIV_2[i];
        if (!decided && y_canon.v[i] != h) {
            gt = y_canon.v[i] > h;
            decided = true;
        }
    }
    return gt;
}

__device__ __forceinline__ void fp_load(fp &r, const uint32_t *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 a = q[0], b = q[1], c = q[2];
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    r.v[8] = c.x; r.v[9] = c.y; r.v[10] = c.z; r.v[11] = c.w;
}
__device__ __forceinline__ void fp_store(uint32_t *p, const fp &r) {
    uint4 *q = reinterpret_cast<uint4 *>(p);
    q[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
    q[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
    q[2] = make_uint4(r.v[8], r.v[9], r.v[10], r.v[11]);
}

}  // namespace cdp
