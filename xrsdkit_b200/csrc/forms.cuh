// forms.cuh -- per-q device functions of the non-crystalline path (FP64).
//
// Each function states the reference code it reproduces (paths relative to the
// reference root).  Operation order follows the reference where the formula is
// ill-conditioned (hard_sphere_sf), so that differences stay at the level of the
// sin/cos implementations.
#pragma once
#include <math.h>

namespace xrsd {

__device__ __forceinline__ double sqr(double x) { return x * x; }

// sin(x) - x cos(x) for |x| < SPHERE_SERIES_X by its Maclaurin series (relative accuracy ~1e-16,
// where the closed form cancels).  t = x^3 (1/3 - x^2/30 + x^4/840 - ...), coefficient k:
// (-1)^(k+1) 2k/(2k+1)!; six terms suffice below x = 0.25 (next term < 1e-18 relative).
constexpr double SPHERE_SERIES_X = 0.25;
__device__ __forceinline__ double sphere_t_series(double x) {
  // Estrin evaluation (dependency depth 3): p(y) = sum_k c_k y^k, y = x^2
  const double y = x * x, y2 = y * y, y4 = y2 * y2;
  const double a0 = fma(-3.333333333333333e-02, y, 3.333333333333333e-01);      //  2/3!  - 4/5! y
  const double a1 = fma(-2.2045855379188714e-05, y, 1.1904761904761906e-03);    //  6/7!  - 8/9! y
  const double a2 = fma(-1.9270852604185937e-09, y, 2.505210838544172e-07);     // 10/11! - 12/13! y
  return fma(a2, y4, fma(a1, y2, a0)) * y * x;
}

// sphere amplitude 3 (sin x - x cos x) / x^3 : scattering/form_factors.py:20-28.
// q == 0 is the caller's business (only q[0]==0 is special-cased by the reference).
__device__ __forceinline__ double sphere_ff(double q, double r) {
  const double x = q * r;
  double s, c;
  sincos(x, &s, &c);
  const double x3 = __dmul_rn(__dmul_rn(x, x), x);
  return __dmul_rn(__dmul_rn(3.0, __dsub_rn(s, __dmul_rn(x, c))), 1.0 / x3);   // 3.*(sin-x cos)*x**-3
}

// normalised 4-Gaussian atomic form factor: scattering/form_factors.py:8-18
__device__ __forceinline__ double atomic_ff(double q, double Z, const double *a, const double *b) {
  const double g = q * 1.0 / (2.0 * M_PI);
  const double s = g * 1.0 / 2.0;
  const double s2 = s * s;
  double acc = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) acc += a[i] * exp(-1.0 * b[i] * s2);
  return (Z - 41.78214 * s2 * acc) / Z;
}

// Guinier-Porod, I(0)=1: scattering/__init__.py:69-108.  `q_splice` and `porod` are the
// per-system constants gp_constants() returns.
struct GPConst { double q_splice, porod, rg2, D; };

__device__ __forceinline__ GPConst gp_constants(double rg, double D) {
  GPConst k;
  k.q_splice = 1.0 / rg * sqrt(3.0 / 2.0 * D);
  k.porod = 1.0 * exp(-1.0 / 2.0 * D) * pow(3.0 / 2.0 * D, 1.0 / 2.0 * D) * 1.0 / pow(rg, D);
  k.rg2 = rg * rg;
  k.D = D;
  return k;
}

__device__ __forceinline__ double guinier_porod(double q, const GPConst &k) {
  if (q <= k.q_splice) return exp(-1.0 / 3.0 * (q * q) * k.rg2);
  if (q > k.q_splice) {
    // porod_factor * 1/(q**D).  D = 4 is the default (and fixed by default): q^4 by two squarings is
    // within 1 ulp of pow(); other exponents go through exp(D log q) (error ~|D log q| ulp), half the
    // cost of the correctly-rounded-ish pow()
    if (k.D == 4.0) { const double q2 = q * q; return k.porod * 1.0 / (q2 * q2); }
    return k.porod * 1.0 / exp(k.D * log(q));
  }
  return 0.0;  // NaN q: neither mask selects it, the reference leaves the zero fill
}

// Percus-Yevick hard spheres, S(q)/S(0): scattering/structure_factors.py:3-47
struct HSConst { double d, p, l1, l2, nc0; };

__device__ __forceinline__ HSConst hs_constants(double r_hard, double v_fraction) {
  HSConst k;
  const double p = v_fraction;
  k.p = p;
  k.d = 2.0 * r_hard;
  const double omp2 = (1.0 - p) * (1.0 - p);
  const double omp4 = omp2 * omp2;                    // (1-p)**4
  k.l1 = sqr(1.0 + 2.0 * p) / omp4;
  k.l2 = -1.0 * sqr(1.0 + p / 2.0) / omp4;
  k.nc0 = -2.0 * p * (4.0 * k.l1 + 18.0 * p * k.l2 + p * k.l1);
  return k;
}

// Returns nc(q d); S(q)/S(0) = (1 - nc0)/(1 - nc).  The three cancelling numerators are formed
// unfused in the reference's order (an FMA contraction there would move the result
// by more than the tolerance at small qd); their quotients by qd^3, qd^4, qd^6 are taken through
// one reciprocal of qd, which perturbs each O(1) bracket by ~1 ulp and is not amplified.
// (s, c) = sin, cos of q d, supplied by the caller (who may carry them from span to span by rotation)
__device__ __forceinline__ double hard_sphere_nc(double q, const HSConst &k, bool is_q0_zero, double s, double c) {
  if (is_q0_zero) return k.nc0;
  const double qd = q * k.d;
  const double qd2 = __dmul_rn(qd, qd), qd3 = __dmul_rn(qd2, qd), qd4 = __dmul_rn(qd2, qd2);
  const double p = k.p;
  const double r1 = 1.0 / qd, r2 = r1 * r1, r3 = r2 * r1, r4 = r2 * r2, r6 = r3 * r3;
  const double x1 = __dsub_rn(s, __dmul_rn(qd, c)) * r3;
  double t2 = __dsub_rn(__dmul_rn(qd2, c), __dmul_rn(__dmul_rn(2.0, qd), s));
  t2 = __dadd_rn(__dsub_rn(t2, __dmul_rn(2.0, c)), 2.0);
  const double x2 = t2 * r4;
  double t3 = __dsub_rn(__dmul_rn(qd4, c), __dmul_rn(__dmul_rn(4.0, qd3), s));
  t3 = __dsub_rn(t3, __dmul_rn(__dmul_rn(12.0, qd2), c));
  t3 = __dadd_rn(t3, __dmul_rn(__dmul_rn(24.0, qd), s));
  t3 = __dsub_rn(__dadd_rn(t3, __dmul_rn(24.0, c)), 24.0);
  const double x3 = t3 * r6;
  double br = __dmul_rn(k.l1, x1);
  br = __dsub_rn(br, __dmul_rn(__dmul_rn(__dmul_rn(6.0, p), k.l2), x2));
  br = __dsub_rn(br, __dmul_rn(__dmul_rn(p, k.l1) / 2.0, x3));
  return __dmul_rn(__dmul_rn(-24.0, p), br);
}

}  // namespace xrsd
