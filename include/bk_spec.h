/*
 * bk_spec.h — NORMATIVE scalar arithmetic of the B200 PGBART sampler.
 *
 * This header is the written-down "spec" that SURVEY.md App. A.9 asks for:
 * "Oracle and kernels share one header".  It holds ONLY order-free scalar
 * definitions (counter-based RNG, uniform/normal conversion, fixed-point
 * quantisation, transcendental functions built from IEEE +,*,fma, and the
 * per-leaf closed forms).  It holds NO algorithm structure: tree growth,
 * member selection, partitioning, reductions, resampling and bookkeeping are
 * written twice, independently — once as CUDA (pymc_bart_b200/csrc/) and once
 * as plain C (oracle/) — and compared bit for bit by tests/.
 *
 * Why every result is bit-reproducible on gcc/x86 and nvcc/sm_100a:
 *   - integers: Philox4x32-10, fixed-point sums (order independent);
 *   - fp: only IEEE-754 correctly rounded primitives (+ - * / sqrt fma) in a
 *     FIXED expression order.  On the device the BK_Dxxx and BK_Fxxx macros map to the
 *     never-contracted __dadd_rn/__fmaf_rn/... intrinsics; on the host they
 *     are plain operators, and every host translation unit including this
 *     header MUST be compiled with -ffp-contract=off.
 *   - no libm transcendental is used anywhere on the path: bk_exp, bk_log,
 *     bk_cos2pi below are polynomial kernels with literal coefficients.
 *
 * Reference anchors (pymc-devs/pymc-bart @4daa2e2; the sampler arithmetic lives
 * in the un-vendored `bartrs` dependency — requirements.txt:6 — so the
 * definitions here restate SURVEY.md Appendix A, not in-tree code):
 *   depth prior .......... pymc_bart/bart.py:107-109
 *   leaf init Y.mean() ... pymc_bart/bart.py:148
 */
#ifndef BK_SPEC_H
#define BK_SPEC_H

#include <stdint.h>

#if defined(__CUDACC__)
#define BK_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#define BK_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define BK_DADD(a, b) __dadd_rn((a), (b))
#define BK_DSUB(a, b) __dsub_rn((a), (b))
#define BK_DMUL(a, b) __dmul_rn((a), (b))
#define BK_DDIV(a, b) __ddiv_rn((a), (b))
#define BK_DFMA(a, b, c) __fma_rn((a), (b), (c))
#define BK_DSQRT(a) __dsqrt_rn((a))
#define BK_DRINT(a) rint((a))
#define BK_FADD(a, b) __fadd_rn((a), (b))
#define BK_FSUB(a, b) __fsub_rn((a), (b))
#define BK_FMUL(a, b) __fmul_rn((a), (b))
#define BK_FDIV(a, b) __fdiv_rn((a), (b))
#define BK_FFMA(a, b, c) __fmaf_rn((a), (b), (c))
#define BK_FSQRT(a) __fsqrt_rn((a))
#else
#define BK_DADD(a, b) ((double)(a) + (double)(b))
#define BK_DSUB(a, b) ((double)(a) - (double)(b))
#define BK_DMUL(a, b) ((double)(a) * (double)(b))
#define BK_DDIV(a, b) ((double)(a) / (double)(b))
#define BK_DFMA(a, b, c) fma((double)(a), (double)(b), (double)(c))
#define BK_DSQRT(a) sqrt((double)(a))
#define BK_DRINT(a) rint((double)(a))
#define BK_FADD(a, b) ((float)(a) + (float)(b))
#define BK_FSUB(a, b) ((float)(a) - (float)(b))
#define BK_FMUL(a, b) ((float)(a) * (float)(b))
#define BK_FDIV(a, b) ((float)(a) / (float)(b))
#define BK_FFMA(a, b, c) fmaf((float)(a), (float)(b), (float)(c))
#define BK_FSQRT(a) sqrtf((float)(a))
#endif

/* ------------------------------------------------------------------ limits */
#define BK_MAX_NODES 255      /* node ids 0..254 live in one byte             */
#define BK_LIMBO 255          /* leaf id of a row dropped by a NaN covariate  */
#define BK_QBITS 29           /* |q| <= 2^29-1 so 16 squared terms fit in u64 */
#define BK_QMAX ((int32_t)((1 << BK_QBITS) - 1))
#define BK_MAX_DEPTH_TABLE 256
#define BK_SPLIT_TRIES 4      /* candidate members per split-value draw (the 4 words of one Philox block); a candidate
                                 whose covariate is missing (NaN) is skipped, 4 misses leave the node a leaf         */

/* RNG purposes */
#define BK_U_LEAF 0u      /* grow-or-stay-leaf test of the popped node */
#define BK_U_VAR 1u       /* split variable                            */
#define BK_U_VAL 2u       /* split value = k-th member                 */
#define BK_Z_LEFT 3u      /* leaf value noise, left child              */
#define BK_Z_RIGHT 4u     /* leaf value noise, right child             */
#define BK_U_RESAMPLE 5u  /* in-loop systematic resampling (particle=0) */
#define BK_U_FINAL 6u     /* final systematic resampling               */
#define BK_U_PICK 7u      /* final position pick                       */

/* SubsetSplit (docs/api_reference.rst:16 `SubsetSplitRule`): categorical covariates hold integer
 * category codes 0..BK_SUBSET_MAX_CATS-1 (the host maps arbitrary category values to codes); the split "value" of a node
 * is the SET of categories that go left, carried as the float whose integer value is the set's bit mask (< 2^24: exact). */
#define BK_SUBSET_MAX_CATS 24
#define BK_MAX_SUBSET_COLS 8  /* columns with the subset rule per model (per-row presence masks are kept for each) */

/* likelihood families */
#define BK_LIK_NORMAL 0
#define BK_LIK_BERNOULLI_LOGIT 1
/* shared-tree multi-output families (every leaf carries one value per output; the reference's tested multi-output
 * models, tests/test_bart.py:107-123 and :140-164) */
#define BK_LIK_NORMAL_HETERO 2   /* y ~ Normal(f[0], |f[1]|),    BART(shape=(2, n)) */
#define BK_LIK_CATEGORICAL 3     /* y ~ Categorical(softmax(f)), BART(shape=(k, n)), y in {0..k-1} */
#define BK_MAX_OUTPUTS 7         /* leaf values per leaf (one 64-byte device node holds 7 floats) */

/* --------------------------------------------------------- Philox4x32-10 */
typedef struct { uint32_t v[4]; } bk_u32x4;

BK_HD uint32_t bk_mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

BK_HD bk_u32x4 bk_philox(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1,
                         uint32_t c2, uint32_t c3) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  const uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = bk_mulhi32(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = bk_mulhi32(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n1 = lo1;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    uint32_t n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  bk_u32x4 out;
  out.v[0] = c0; out.v[1] = c1; out.v[2] = c2; out.v[3] = c3;
  return out;
}

/* Addressable stream: key (seed, chain), counter
 * (draw, group<<16 | tree, round<<16 | particle, purpose). */
BK_HD bk_u32x4 bk_rng(uint32_t seed, uint32_t chain, uint32_t draw, uint32_t group,
                      uint32_t tree, uint32_t round, uint32_t particle,
                      uint32_t purpose) {
  return bk_philox(seed, chain, draw, (group << 16) | (tree & 0xFFFFu),
                   (round << 16) | (particle & 0xFFFFu), purpose);
}

/* x * 2^-32 in [0,1), exact in double */
BK_HD double bk_u01(uint32_t x) { return BK_DMUL((double)x, 2.3283064365386963e-10); }
/* floor(x * n / 2^32): uniform index in [0, n) without any rounding */
BK_HD uint32_t bk_index(uint32_t x, uint32_t n) { return bk_mulhi32(x, n); }

/* ------------------------------------------------------------ SubsetSplit */
/* category code of a covariate value: the integers 0..23; anything else (NaN included) has no category (-1) */
BK_HD int bk_subset_code(float x) {
  if (!(x >= 0.0f && x < (float)BK_SUBSET_MAX_CATS)) return -1;
  const int c = (int)x;
  return ((float)c == x) ? c : -1;
}
/* does a row with covariate x go left at a subset split whose node carries `split`? */
BK_HD int bk_subset_left(float x, float split) {
  const int c = bk_subset_code(x);
  return c >= 0 && ((((uint32_t)split) >> c) & 1u);
}
/* The set drawn for a node whose members show the categories `present` (bit c = some member with a value has category
 * c): a uniformly drawn NON-EMPTY subset of the present categories without the largest one — every split of the
 * present categories into two non-empty groups is drawn with the same probability and the right child is never empty
 * (the historical rule: `unique(values)[:-1]`, redrawn until non-empty).  pick = 1 + floor(r * n_sub / 2^32) indexes
 * the n_sub = 2^(u-1) - 1 non-empty subsets of the u - 1 candidates; bit i of pick decides the i-th smallest
 * candidate.  Returns 0 (no split: the node stays a leaf) when fewer than two categories are present. */
BK_HD uint32_t bk_subset_draw(uint32_t present, uint32_t r) {
  present &= (1u << BK_SUBSET_MAX_CATS) - 1u;
  int u = 0, top = -1;
  for (int c = 0; c < BK_SUBSET_MAX_CATS; ++c) if ((present >> c) & 1u) { u += 1; top = c; }
  if (u < 2) return 0u;
  const uint32_t cand = present & ~(1u << top);
  const uint32_t n_sub = (1u << (u - 1)) - 1u;
  uint32_t pick = bk_index(r, n_sub) + 1u;
  uint32_t out = 0u;
  for (int c = 0; c < BK_SUBSET_MAX_CATS; ++c)
    if ((cand >> c) & 1u) { if (pick & 1u) out |= 1u << c; pick >>= 1; }
  return out;
}

/* ------------------------------------------------------- bit-level helpers */
BK_HD uint64_t bk_d2bits(double d) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(d);
#else
  union { double d; uint64_t u; } c; c.d = d; return c.u;
#endif
}
BK_HD double bk_bits2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  union { double d; uint64_t u; } c; c.u = u; return c.d;
#endif
}

BK_HD float bk_bits2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}

/* ------------------------------------------------------------ exp, log, cos */
#define BK_LN2_HI 6.93147180369123816490e-01
#define BK_LN2_LO 1.90821492927058770002e-10
#define BK_INV_LN2 1.4426950408889634
#define BK_HALF_LOG_2PI 0.9189385332046727
#define BK_PI_4 0.7853981633974483

/* exp(x); x < -708 flushes to 0, x > 709 saturates (never needed: callers pass x <= 0) */
BK_HD double bk_exp(double x) {
  if (!(x >= -708.0)) return 0.0;
  if (x > 709.0) x = 709.0;
  double kf = BK_DRINT(BK_DMUL(x, BK_INV_LN2));
  double r = BK_DFMA(-kf, BK_LN2_HI, x);
  r = BK_DFMA(-kf, BK_LN2_LO, r);
  double p = 1.6059043836821613e-10;            /* 1/13! */
  p = BK_DFMA(p, r, 2.08767569878681e-09);      /* 1/12! */
  p = BK_DFMA(p, r, 2.505210838544172e-08);     /* 1/11! */
  p = BK_DFMA(p, r, 2.755731922398589e-07);     /* 1/10! */
  p = BK_DFMA(p, r, 2.7557319223985893e-06);    /* 1/9!  */
  p = BK_DFMA(p, r, 2.48015873015873e-05);      /* 1/8!  */
  p = BK_DFMA(p, r, 0.0001984126984126984);     /* 1/7!  */
  p = BK_DFMA(p, r, 0.001388888888888889);      /* 1/6!  */
  p = BK_DFMA(p, r, 0.008333333333333333);      /* 1/5!  */
  p = BK_DFMA(p, r, 0.041666666666666664);      /* 1/4!  */
  p = BK_DFMA(p, r, 0.16666666666666666);       /* 1/3!  */
  p = BK_DFMA(p, r, 0.5);
  p = BK_DFMA(p, r, 1.0);
  p = BK_DFMA(p, r, 1.0);
  int64_t k = (int64_t)kf;                      /* |k| <= 1023 */
  return BK_DMUL(p, bk_bits2d((uint64_t)(k + 1023) << 52));
}

/* natural log of a positive normal double (x <= 0 or subnormal -> -745.0 floor) */
BK_HD double bk_log(double x) {
  if (!(x >= 2.2250738585072014e-308)) return -745.0;
  uint64_t b = bk_d2bits(x);
  int64_t e = (int64_t)((b >> 52) & 0x7FFu) - 1023;
  double m = bk_bits2d((b & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull);
  if (m > 1.4142135623730951) { m = BK_DMUL(m, 0.5); e += 1; }
  double f = BK_DSUB(m, 1.0);
  double s = BK_DDIV(f, BK_DADD(2.0, f));
  double z = BK_DMUL(s, s);
  double p = 0.04;                               /* 1/25 */
  p = BK_DFMA(p, z, 0.043478260869565216);       /* 1/23 */
  p = BK_DFMA(p, z, 0.047619047619047616);       /* 1/21 */
  p = BK_DFMA(p, z, 0.05263157894736842);        /* 1/19 */
  p = BK_DFMA(p, z, 0.058823529411764705);       /* 1/17 */
  p = BK_DFMA(p, z, 0.06666666666666667);        /* 1/15 */
  p = BK_DFMA(p, z, 0.07692307692307693);        /* 1/13 */
  p = BK_DFMA(p, z, 0.09090909090909091);        /* 1/11 */
  p = BK_DFMA(p, z, 0.1111111111111111);         /* 1/9  */
  p = BK_DFMA(p, z, 0.14285714285714285);        /* 1/7  */
  p = BK_DFMA(p, z, 0.2);                        /* 1/5  */
  p = BK_DFMA(p, z, 0.3333333333333333);         /* 1/3  */
  p = BK_DFMA(p, z, 1.0);
  double lm = BK_DMUL(BK_DMUL(2.0, s), p);       /* log(m) = 2 atanh(s) */
  double ef = (double)e;
  return BK_DFMA(ef, BK_LN2_HI, BK_DFMA(ef, BK_LN2_LO, lm));
}

BK_HD double bk_sin_q(double x) { /* x in [0, pi/4] */
  double z = BK_DMUL(x, x);
  double p = 2.8114572543455206e-15;             /* 1/17! */
  p = BK_DFMA(p, z, -7.647163731819816e-13);
  p = BK_DFMA(p, z, 1.6059043836821613e-10);
  p = BK_DFMA(p, z, -2.505210838544172e-08);
  p = BK_DFMA(p, z, 2.7557319223985893e-06);
  p = BK_DFMA(p, z, -0.0001984126984126984);
  p = BK_DFMA(p, z, 0.008333333333333333);
  p = BK_DFMA(p, z, -0.16666666666666666);
  p = BK_DFMA(p, z, 1.0);
  return BK_DMUL(x, p);
}
BK_HD double bk_cos_q(double x) { /* x in [0, pi/4] */
  double z = BK_DMUL(x, x);
  double p = -1.5619206968586225e-16;            /* -1/18! */
  p = BK_DFMA(p, z, 4.779477332387385e-14);
  p = BK_DFMA(p, z, -1.1470745597729725e-11);
  p = BK_DFMA(p, z, 2.08767569878681e-09);
  p = BK_DFMA(p, z, -2.755731922398589e-07);
  p = BK_DFMA(p, z, 2.48015873015873e-05);
  p = BK_DFMA(p, z, -0.001388888888888889);
  p = BK_DFMA(p, z, 0.041666666666666664);
  p = BK_DFMA(p, z, -0.5);
  p = BK_DFMA(p, z, 1.0);
  return p;
}
/* cos(2*pi*x/2^32): exact octant reduction on the integer, then a Taylor kernel */
BK_HD double bk_cos2pi(uint32_t x) {
  uint32_t oct = x >> 29;
  double t = BK_DMUL((double)(x & 0x1FFFFFFFu), 1.862645149230957e-09); /* 2^-29 */
  double phi = BK_DMUL(t, BK_PI_4);
  double psi = BK_DMUL(BK_DSUB(1.0, t), BK_PI_4);
  switch (oct) {
    case 0: return bk_cos_q(phi);
    case 1: return bk_sin_q(psi);
    case 2: return -bk_sin_q(phi);
    case 3: return -bk_cos_q(psi);
    case 4: return -bk_cos_q(phi);
    case 5: return -bk_sin_q(psi);
    case 6: return bk_sin_q(phi);
    default: return bk_cos_q(psi);
  }
}
/* one standard normal from one Philox block (Box-Muller, lanes 0 and 1) */
BK_HD double bk_normal(bk_u32x4 w) {
  double u1 = BK_DMUL(BK_DADD((double)w.v[0], 1.0), 2.3283064365386963e-10); /* (0,1] */
  double rad = BK_DSQRT(BK_DMUL(-2.0, bk_log(u1)));
  return BK_DMUL(rad, bk_cos2pi(w.v[1]));
}

/* ------------------------------------------------------ fixed-point sums */
/* q = clamp(rint(v * 2^qshift)); qscale = 2^qshift as float (exact) */
BK_HD int32_t bk_quant(float v, float qscale) {
  float t = BK_FMUL(v, qscale);
  const float QF = 536870912.0f; /* 2^29 */
  t = t > QF ? QF : t;
  t = t < -QF ? -QF : t;
#if defined(__CUDA_ARCH__)
  int32_t q = __float2int_rn(t);
#else
  int32_t q = (int32_t)lrintf(t);
#endif
  q = q > BK_QMAX ? BK_QMAX : q;
  q = q < -BK_QMAX ? -BK_QMAX : q;
  return q;
}

/* 128-bit unsigned sum of squares kept as (hi, lo) 64-bit words */
typedef struct { uint64_t hi, lo; } bk_u128;
BK_HD bk_u128 bk_u128_make(uint64_t hi, uint64_t lo) { bk_u128 r; r.hi = hi; r.lo = lo; return r; }
BK_HD bk_u128 bk_u128_add(bk_u128 a, bk_u128 b) {
  bk_u128 r; r.lo = a.lo + b.lo; r.hi = a.hi + b.hi + (r.lo < a.lo ? 1u : 0u); return r;
}
BK_HD bk_u128 bk_u128_sub(bk_u128 a, bk_u128 b) {
  bk_u128 r; r.lo = a.lo - b.lo; r.hi = a.hi - b.hi - (a.lo < b.lo ? 1u : 0u); return r;
}
/* kernels accumulate the low and high 32-bit halves of each partial separately:
 * value = acc_hi * 2^32 + acc_lo */
BK_HD bk_u128 bk_u128_from_split(uint64_t acc_hi, uint64_t acc_lo) {
  bk_u128 a = bk_u128_make(acc_hi >> 32, acc_hi << 32);
  return bk_u128_add(a, bk_u128_make(0, acc_lo));
}
BK_HD double bk_u128_to_double(bk_u128 a) {
  return BK_DFMA((double)a.hi, 18446744073709551616.0, (double)a.lo);
}

/* per-leaf sufficient statistics in fixed point */
typedef struct {
  int32_t n;      /* members                         */
  int64_t sst;    /* sum q(sum_trees)                */
  int64_t sr;     /* sum q(r), r = y - sum_trees_noi */
} bk_stats;

BK_HD bk_stats bk_stats_sub(bk_stats a, bk_stats b) {
  bk_stats r; r.n = a.n - b.n; r.sst = a.sst - b.sst; r.sr = a.sr - b.sr; return r;
}

/* leaf value: mean(sum_trees over members)/m + z*leaf_sd, 0 for an empty leaf
 * (SURVEY.md App. A.5).  inv_qm = 2^-qshift / m (one rounded double, computed once by the caller).
 * The only division is a correctly rounded float one: fp64 division costs ~1000 cycles on a B200 SM. */
BK_HD float bk_leaf_value(int32_t n, int64_t sst, double inv_qm, double z, float leaf_sd) {
  if (n <= 0) return 0.0f;
  float mean = BK_FDIV((float)BK_DMUL((double)sst, inv_qm), (float)n);
  double v = BK_DFMA(z, (double)leaf_sd, (double)mean);
  return (float)v;
}

/* Gaussian: for a leaf with value mu, sum over members of (r - mu)^2 = R2_leaf - g with the leaf "gain"
 *   g = mu * (2 * sum r - n * mu).
 * Summed over the leaves of a tree the R2_leaf add up to the sum of squares of ALL rows, which does not depend on the
 * tree: ssq(tree) = R2_total - sum over leaves of g.  A particle therefore carries only its gain sum (updated on a
 * split as ((G - g_parent) + g_left) + g_right) and no per-leaf sum of squares exists anywhere on the path. */
BK_HD double bk_leaf_gain(bk_stats s, float mu, double inv_qscale) {
  double r1 = BK_DMUL((double)s.sr, inv_qscale);
  double m = (double)mu;
  double t = BK_DFMA(-m, (double)s.n, BK_DMUL(2.0, r1));
  return BK_DMUL(m, t);
}
/* sum of squares of all rows' residuals from the exact 128-bit integer sum of q(r)^2 */
BK_HD double bk_total_r2(bk_u128 sr2, double inv_qscale) {
  return BK_DMUL(BK_DMUL(bk_u128_to_double(sr2), inv_qscale), inv_qscale);
}
BK_HD double bk_ssq_from_gain(double r2_total, double gain) { return BK_DSUB(r2_total, gain); }
/* Gaussian log-likelihood of all N rows given ssq = sum (r - mu_leaf)^2:
 * lw = -ssq * inv2s2 + c with the two per-step constants below */
BK_HD double bk_normal_inv2s2(float sigma) {
  double s = (double)sigma;
  return BK_DDIV(0.5, BK_DMUL(s, s));
}
BK_HD double bk_normal_const(float sigma, double n_rows) {
  return BK_DMUL(-n_rows, BK_DADD(bk_log((double)sigma), BK_HALF_LOG_2PI));
}
BK_HD double bk_normal_loglik_pre(double ssq, double inv2s2, double c) { return BK_DFMA(-ssq, inv2s2, c); }
BK_HD double bk_normal_loglik(double ssq, float sigma, double n_rows) {
  return bk_normal_loglik_pre(ssq, bk_normal_inv2s2(sigma), bk_normal_const(sigma, n_rows));
}

/* Bernoulli-logit per-row term y*f - softplus(f), in float with a fixed
 * polynomial kernel; quantised by the caller with qscale_ll */
BK_HD float bk_exp_neg_f(float a) { /* exp(-a), a >= 0 */
  if (a > 87.0f) return 0.0f;
  float x = -a;
  float kf = rintf(BK_FMUL(x, 1.4426950408889634f));
  float r = BK_FFMA(-kf, 0.693145751953125f, x);        /* ln2 hi (12 bits) */
  r = BK_FFMA(-kf, 1.4286068203094173e-06f, r);         /* ln2 lo           */
  float p = 0.0001984126984126984f;
  p = BK_FFMA(p, r, 0.001388888888888889f);
  p = BK_FFMA(p, r, 0.008333333333333333f);
  p = BK_FFMA(p, r, 0.041666666666666664f);
  p = BK_FFMA(p, r, 0.16666666666666666f);
  p = BK_FFMA(p, r, 0.5f);
  p = BK_FFMA(p, r, 1.0f);
  p = BK_FFMA(p, r, 1.0f);
  int32_t k = (int32_t)kf;                              /* -126 <= k <= 0 */
  return BK_FMUL(p, bk_bits2f((uint32_t)(k + 127) << 23));
}
BK_HD float bk_log1p_01(float t) { /* log(1+t), t in [0,1] */
  /* 1+t in [1,2]: m = (1+t) or (1+t)/2 to stay within [sqrt(.5), sqrt(2)] */
  float m = BK_FADD(1.0f, t);
  float e = 0.0f;
  if (m > 1.4142135f) { m = BK_FMUL(m, 0.5f); e = 0.6931471805599453f; }
  float f = BK_FSUB(m, 1.0f);
  float s = BK_FDIV(f, BK_FADD(2.0f, f));
  float z = BK_FMUL(s, s);
  float p = 0.1111111111111111f;
  p = BK_FFMA(p, z, 0.14285714285714285f);
  p = BK_FFMA(p, z, 0.2f);
  p = BK_FFMA(p, z, 0.3333333333333333f);
  p = BK_FFMA(p, z, 1.0f);
  return BK_FFMA(BK_FMUL(2.0f, s), p, e);
}
BK_HD float bk_bernoulli_logit_term(float y, float f) {
  float a = f < 0.0f ? -f : f;
  float sp = BK_FADD(f > 0.0f ? f : 0.0f, bk_log1p_01(bk_exp_neg_f(a)));
  return BK_FSUB(BK_FMUL(y, f), sp);
}

/* Bernoulli-logit likelihood in fixed point (sum_i y_i*f_i - softplus(f_i)).
 * The row's linear predictor is f = noi + leaf value, one float add; its log-likelihood term is
 * quantised to a multiple of 2^-BK_LL_QSHIFT and summed as int64, so a node's / particle's
 * log-likelihood is an exact integer whatever the reduction order.  |term| <= 512 after clamping. */
#define BK_LL_QSHIFT 20
BK_HD int32_t bk_bern_q(float y, float noi, float value) {
  return bk_quant(bk_bernoulli_logit_term(y, BK_FADD(noi, value)), 1048576.0f);
}
/* log-likelihood of a particle from the integer sum of its leaves' terms (exact: |llq| < 2^53) */
BK_HD double bk_bern_loglik(double llq) { return BK_DMUL(llq, 9.5367431640625e-07); }

/* natural log of a positive normal float (float kernel, fixed order; |error| ~ 1e-7 relative) */
BK_HD uint32_t bk_f2bits(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  union { float f; uint32_t u; } c; c.f = f; return c.u;
#endif
}
BK_HD float bk_logf(float x) {
  const uint32_t b = bk_f2bits(x);
  int32_t e = (int32_t)((b >> 23) & 0xFFu) - 127;
  float m = bk_bits2f((b & 0x007FFFFFu) | 0x3F800000u);
  if (m > 1.4142135f) { m = BK_FMUL(m, 0.5f); e += 1; }
  const float f = BK_FSUB(m, 1.0f);
  const float s = BK_FDIV(f, BK_FADD(2.0f, f));
  const float z = BK_FMUL(s, s);
  float p = 0.1111111111111111f;
  p = BK_FFMA(p, z, 0.14285714285714285f);
  p = BK_FFMA(p, z, 0.2f);
  p = BK_FFMA(p, z, 0.3333333333333333f);
  p = BK_FFMA(p, z, 1.0f);
  const float lm = BK_FMUL(BK_FMUL(2.0f, s), p);
  const float ef = (float)e;
  return BK_FFMA(ef, 0.693145751953125f, BK_FFMA(ef, 1.4286068203094173e-06f, lm));
}

/* Per-row log-likelihood term of the shared-tree multi-output families at the linear predictors f[0..K-1]
 * (the weight is the full-model data log-likelihood).  Float kernels in a fixed order; the caller
 * quantises with bk_lik_q and sums integers, exactly like the Bernoulli path. */
BK_HD float bk_lik_term(int lik, int K, float y, const float* f) {
  if (lik == BK_LIK_NORMAL_HETERO) {       /* -1/2 ((y - f0)/|f1|)^2 - log|f1| - 1/2 log(2 pi) */
    float a = f[1] < 0.0f ? -f[1] : f[1];
    a = a < 1e-20f ? 1e-20f : a;
    const float z = BK_FDIV(BK_FSUB(y, f[0]), a);
    return BK_FSUB(BK_FSUB(BK_FMUL(-0.5f, BK_FMUL(z, z)), bk_logf(a)), 0.9189385332046727f);
  }
  /* BK_LIK_CATEGORICAL: f[y] - logsumexp(f) */
  float mx = f[0];
  for (int j = 1; j < K; ++j) mx = f[j] > mx ? f[j] : mx;
  float sum = 0.0f;
  for (int j = 0; j < K; ++j) sum = BK_FADD(sum, bk_exp_neg_f(BK_FSUB(mx, f[j])));
  int yi = (int)y;
  yi = yi < 0 ? 0 : (yi >= K ? K - 1 : yi);
  float fy = f[0];
  for (int j = 1; j < K; ++j) fy = j == yi ? f[j] : fy;
  return BK_FSUB(BK_FSUB(fy, mx), bk_logf(sum));
}
BK_HD int32_t bk_lik_q(int lik, int K, float y, const float* f) { return bk_quant(bk_lik_term(lik, K, y, f), 1048576.0f); }

/* Particle weights and systematic resampling in FIXED POINT (w = exp(lw - max) + 1e-12,
 * normalise, inverse-CDF walk over the points (u + i)/L).
 *   W_i   = rint(exp(-(max - lw_i)) * 2^40) + 1       exp in float (bk_exp_neg_f); the +1 (2^-40 ~ 9e-13) is the floor
 *   S_j   = W_0 + ... + W_j                           exact 64-bit integers: any scan order gives the same bits
 *   point i of L:  A_i = i * 2^32 + u32               (u32 = the raw 32-bit uniform)
 *   ancestor(i) = first j with  A_i * S_last <= S_j * L * 2^32   (128-bit integers), capped at L - 1
 * which is the walk `while (point > cum[idx]) idx++` on cum_j = S_j / S_last, point = (u + i)/L without a single
 * rounding.  No fp64 division or exponential is left on the per-round critical path. */
#define BK_W_SHIFT 40
BK_HD uint64_t bk_weight_fix(double lw, double lw_max) {
  float t = (float)BK_DSUB(lw_max, lw);             /* >= 0 */
  float e = bk_exp_neg_f(t < 0.0f ? 0.0f : t);      /* (0, 1] or 0 after underflow */
  /* e * 2^40 is exact in float; round-to-nearest-even conversion on both sides */
#if defined(__CUDA_ARCH__)
  return (uint64_t)__float2ull_rn(BK_FMUL(e, 1099511627776.0f)) + 1u;
#else
  return (uint64_t)llrintf(BK_FMUL(e, 1099511627776.0f)) + 1u;
#endif
}
BK_HD uint64_t bk_mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
  return __umul64hi(a, b);
#else
  return (uint64_t)(((unsigned __int128)a * (unsigned __int128)b) >> 64);
#endif
}
/* left-hand side of the ancestor test for point i: A_i * S_last as (hi, lo) */
BK_HD bk_u128 bk_resample_point(uint32_t i, uint32_t u32, uint64_t s_last) {
  uint64_t a = ((uint64_t)i << 32) | (uint64_t)u32;
  return bk_u128_make(bk_mulhi64(a, s_last), a * s_last);
}
/* point <= S_j * L * 2^32 ?   (S_j * L < 2^63 for L <= 128) */
BK_HD int bk_resample_le(bk_u128 point, uint64_t s_j, uint32_t L) {
  uint64_t sl = s_j * (uint64_t)L;
  uint64_t hi = sl >> 32, lo = sl << 32;
  return point.hi < hi || (point.hi == hi && point.lo <= lo);
}

#endif /* BK_SPEC_H */
