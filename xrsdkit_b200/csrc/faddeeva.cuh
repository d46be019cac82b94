// faddeeva.cuh -- Re w(x + i y), the real part of the Faddeeva function, FP64, for the
// Voigt profile of tools/peak_math.py:38-47 (the reference calls scipy.special.wofz;
// scipy is an unpinned third-party dependency whose source is not in the reference).
//
// Own implementation of two published expansions, chosen because y = gamma/(sigma sqrt2)
// is CONSTANT for a whole pattern, so everything that depends on y alone is a
// per-pattern table (VoigtConst) staged once in shared memory:
//
//   |z|^2 >= 64 : asymptotic series  w(z) ~ i/(sqrt(pi) z) * sum_k (2k-1)!!/(2 z^2)^k, k <= 5, 7 or 12 by |z|^2;
//                 its real part through a real three-term recurrence (rew_far).
//   otherwise   : the exponential series of Zaghloul & Ali (ACM TOMS 38, 15 (2011),
//                 Algorithm 916, eq. 13/14) with spacing a = 0.5183..., real part:
//                   Re w = e^{-x^2} { [erfcx(y) - c y S1] cos(2xy) + c x sin(xy) sinc(xy)
//                                     + (c y / 2) sum_n K_n (E^n + E^-n) },
//                 K_n = e^{-a^2 n^2}/(a^2 n^2 + y^2), S1 = sum K_n, E = e^{2 a x}, c = 2a/pi.
//                 All terms of the last sum are positive, so the result keeps RELATIVE
//                 accuracy in the far Gaussian wings where Re w ~ y/(sqrt(pi) x^2) << 1.
//
// Measured against scipy 1.18.1 wofz on x in [0, 1e7], y in [1e-12, 1e6]: max relative
// difference 5e-13 (at the |z| = 8 hand-over), typically 2e-14 (tests/test_gpu_parity.py::test_faddeeva_against_scipy).
#pragma once
#include <math.h>

namespace xrsd {

constexpr int VOIGT_NPOS = 32;   // terms of sum_n K_n E^n   (|x| < 8: e^{-(an-x)^2} < 1e-30 beyond)
constexpr int VOIGT_NNEG = 13;   // terms of sum_n K_n E^-n  (e^{-a^2 n^2} < 1e-19 beyond)
constexpr double VOIGT_A = 0.518321480430085929872;
constexpr double VOIGT_C = 2.0 * VOIGT_A / M_PI;

// per-pattern constants (depend on y only)
struct VoigtConst {
  double y;
  double head;                // erfcx(y) - c y S1
  double cy_half;             // c y / 2
  double K[VOIGT_NPOS];
};

// fills `vc` cooperatively: call with all threads of the CTA, then __syncthreads().
__device__ __forceinline__ void voigt_constants(VoigtConst *vc, double y, int tid) {
  static_assert(VOIGT_NPOS == 32, "one warp computes the K_n and their sum");
  if (tid < VOIGT_NPOS) {                               // warp 0, all lanes
    const double an = VOIGT_A * (tid + 1);
    const double kn = exp(-an * an) / (an * an + y * y);
    vc->K[tid] = kn;
    double s1 = kn;                                      // S1 = sum_n K_n by a shuffle tree
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    if (tid == 0) {
      vc->y = y;
      vc->head = erfcx(y) - VOIGT_C * y * s1;
      vc->cy_half = 0.5 * VOIGT_C * y;
    }
  }
}

// asymptotic branch, |z|^2 >= 64:  w(z) ~ i/(sqrt(pi) z) sum_k c_k / z^2k,  c_k = (2k-1)!!/2^k, truncated where the
// first neglected term is < 1e-13 relative (k <= 5, 7 or 12 by |z|^2).  Only the real part is needed, and
//     Re[i / z^(2k+1)] = Im(z^(2k+1)) / |z|^(4k+2) =: u_k
// obeys a REAL three-term recurrence (z^2 and conj(z)^2 are the roots of its characteristic polynomial):
//     u_k+1 = (2 Re z^2 / |z|^4) u_k - u_k-1 / |z|^4,     u_0 = y / |z|^2,  u_-1 = -y,
// whose two solutions have equal modulus (errors grow linearly in k, no cancellation: every u_k carries the factor y
// that makes the far wing of a narrow-Lorentzian Voigt profile small).  The sum is taken by Clenshaw's backward
// recurrence over that three-term relation: two FMAs per term with the coefficient as an immediate (forming the u_k and
// accumulating c_k u_k took three FP64 instructions per term, a complex Horner / Estrin evaluation four plus
// two-constant linear factors), and 2 Re z^2 / |z|^4 = 2 t - 4 y^2 t^2 needs x no more: 18 FP64 instructions for k <= 5
// (24 before; this branch is ~40 % of the instructions of a crystalline pattern, profiles/r02_ncu_cfg3.txt).
// Max relative difference to scipy's wofz 1.3e-13 (at |z| = 8, the truncation), the same as the forward form.
__device__ __forceinline__ double rew_far_raw(double /* x: enters through z2 only */, double y, double z2) {
  // Returns sqrt(pi) Re w (the caller scales, or -- in a ratio of profile sums -- does not have to).
  // 1/|z|^2 for 64 <= |z|^2 < inf: the hardware seed (MUFU.RCP64H, relative error <= 1e-6) and ONE third-order step
  // e <- e + e^2, t <- t + t e: error e^3 ~ 1e-18 plus rounding -- measured 1.11e-16 over 4e6 arguments in
  // [64, 6.4e8], the same as with the second Newton step __drcp_rn adds for correct rounding (and without its test
  // and slow path for subnormal / huge arguments: 4 instead of 11 instructions)
  double t;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(t) : "d"(z2));
  {
    double e = fma(-z2, t, 1.0);
    e = fma(e, e, e);
    t = fma(t, e, t);
  }
  const double b = t * t;                               // 1/|z|^4
  // al = 2 Re z^2 / |z|^4 with Re z^2 = |z|^2 - 2 y^2:  2 t - 4 y^2 b  (x does not enter again)
  const double al = fma(-4.0 * (y * y), b, t + t);
  // S = sum_k c_k u_k, c_k = (2k-1)!! / 2^k, by Clenshaw's backward recurrence for u_k+1 = al u_k - b u_k-1:
  //     B_k = c_k + al B_k+1 - b B_k+2,      S = B_0 u_0 + B_1 (u_1 - al u_0) = y (B_0 t + B_1 b)
  // -- two FMAs per term where forming u_k and accumulating c_k u_k took a product and two FMAs; the B_k are dominated
  // by their c_k (al, b <= 2/64), so nothing cancels.  The truncation (k <= 5 / 7 / 12 by |z|^2) sets where the chain starts.
  double B1, B2;                                        // B_k+1, B_k+2
#define XRSD_FAR_STEP(ck) { const double Bn = fma(al, B1, fma(-b, B2, ck)); B2 = B1; B1 = Bn; }
  if (z2 < 1024.0) {
    if (z2 < 256.0) {                                   // down to |z| = 8: k = 12 .. 8
      B2 = 77205601.37329102;
      B1 = fma(al, B2, 6713530.554199219);
      XRSD_FAR_STEP(639383.8623046875) XRSD_FAR_STEP(67303.564453125) XRSD_FAR_STEP(7918.06640625)
      XRSD_FAR_STEP(1055.7421875) XRSD_FAR_STEP(162.421875)
    } else {                                            // k = 7, 6
      B2 = 1055.7421875;
      B1 = fma(al, B2, 162.421875);
    }
    XRSD_FAR_STEP(29.53125) XRSD_FAR_STEP(6.5625)
  } else {                                              // k = 5, 4
    B2 = 29.53125;
    B1 = fma(al, B2, 6.5625);
  }
  XRSD_FAR_STEP(1.875) XRSD_FAR_STEP(0.75) XRSD_FAR_STEP(0.5) XRSD_FAR_STEP(1.0)
#undef XRSD_FAR_STEP
  return y * fma(B2, b, B1 * t);                        // B1 = B_0, B2 = B_1
}
constexpr double XRSD_INV_SQRT_PI = 0.56418958354775628695, XRSD_SQRT_PI = 1.7724538509055160273;
__device__ __forceinline__ double rew_far(double x, double y, double z2) { return rew_far_raw(x, y, z2) * XRSD_INV_SQRT_PI; }

// exponential-series branch, |z|^2 < 64 (so |x| < 8).  x must be >= 0 (Re w is even in x).
static __device__ __noinline__ double rew_near(double x, const VoigtConst *vc) {
  const double y = vc->y;
  const double ex2 = exp(-x * x);
  const double E = exp(2.0 * VOIGT_A * x);
  const double Ei = 1.0 / E;
  double sp = vc->K[VOIGT_NPOS - 1];
#pragma unroll
  for (int k = VOIGT_NPOS - 2; k >= 0; --k) sp = fma(sp, E, vc->K[k]);
  sp *= E;
  double sm = vc->K[VOIGT_NNEG - 1];
#pragma unroll
  for (int k = VOIGT_NNEG - 2; k >= 0; --k) sm = fma(sm, Ei, vc->K[k]);
  sm *= Ei;
  const double xy = x * y;
  double sxy, cxy;
  sincos(xy, &sxy, &cxy);
  const double cos2 = fma(-2.0 * sxy, sxy, 1.0);                 // cos(2xy)
  const double sinc = (fabs(xy) > 1.0e-4) ? sxy / xy : fma(-xy * xy, 1.0 / 6.0, 1.0);
  return ex2 * (vc->head * cos2 + VOIGT_C * x * sxy * sinc + vc->cy_half * (sp + sm));
}

// Per-pattern table of the near branch: y is fixed, so Re w(x + i y), 0 <= x < 8, is ONE smooth function of
// x for the whole pattern.  It is tabulated as a degree-13 Chebyshev expansion on each of 16 intervals of
// width 1/2 (14 x 16 evaluations of rew_near per pattern instead of one per (q, shell) pair near a peak
// core); max relative deviation from scipy's wofz 6e-14 for y in [1e-4, 8) (narrow intervals keep the
// dynamic range of e^{-x^2} inside one interval small, so the error stays relative).
constexpr int VOIGT_TAB_N = 14;          // nodes per interval (degree 13: as accurate as degree 16 here -- 4e-14 for y >= 1e-4 --
                                         // and 14 x 16 = 224 node evaluations are two passes of 128 threads, 17 x 16 were three)
constexpr int VOIGT_TAB_INTERVALS = 16;
constexpr double VOIGT_TAB_MIN_Y = 1.0e-4;     // below: e^{-x^2} dominates out to x ~ 4.5 and the table loses relative accuracy
constexpr int VOIGT_TAB_DOUBLES = VOIGT_TAB_N * VOIGT_TAB_INTERVALS;      // nodes of one table build
// A far-field coefficient row is stored with a pitch of 18 doubles (16-byte aligned): pairs (e_j, o_j), ONE LDS.128
// per pair -- 8 loads per evaluation (the coefficient loads were 5 % of a crystalline pattern's instructions).
constexpr int CHEB_ROW = 18;
constexpr int VOIGT_TAB_STORE = VOIGT_TAB_N * VOIGT_TAB_INTERVALS;        // table rows: (e_0, o_0, ..., e_6, o_6), 14 doubles (16-byte multiples)

// Cooperative build (all `nthreads` threads): mat / nodes = XRSD_VT_M / XRSD_VT_NODES of cheb_matrices.cuh (node values ->
// even / odd power-series coefficients in u = 2 t^2 - 1, rows e_0, o_0, e_1, o_1, ...); tmp holds VOIGT_TAB_DOUBLES
// doubles; tab receives, per interval, the row (e_0, o_0, ..., e_6, o_6) that eo14() evaluates.
__device__ __forceinline__ void voigt_table(double *tab, double *tmp, const double *mat, const double *nodes, const VoigtConst *vc,
                                            int tid, int nthreads) {
  for (int it = tid; it < VOIGT_TAB_DOUBLES; it += nthreads) {
    const int k = it / VOIGT_TAB_N, m = it - k * VOIGT_TAB_N;
    tmp[it] = rew_near(0.5 * ((double)k + 0.5 + 0.5 * nodes[m]), vc);
  }
  __syncthreads();
  for (int it = tid; it < VOIGT_TAB_DOUBLES; it += nthreads) {
    const int k = it / VOIGT_TAB_N, j = it - k * VOIGT_TAB_N;
    double a = 0.0;
#pragma unroll
    for (int m = 0; m < VOIGT_TAB_N; ++m) a = fma(tmp[k * VOIGT_TAB_N + m], mat[j * VOIGT_TAB_N + m], a);
    tab[k * VOIGT_TAB_N + j] = a * XRSD_SQRT_PI;      // the table holds sqrt(pi) Re w, like rew_far_raw (see rew_raw)
  }
  __syncthreads();
}

// The interpolants are evaluated as E(u) + t O(u), u = 2 t^2 - 1, |t| <= 1: two independent Horner chains over the
// even / odd power-series coefficients the matrices of cheb_matrices.cuh produce straight from the node values (one FMA
// per coefficient; the Clenshaw form of the same interpolant, two chains in u as well, took an FMA and a subtraction
// per coefficient -- 33 against 15 FP64 instructions for 15 coefficients).  A row holds (e_0, o_0, e_1, o_1, ...),
// 16-byte aligned, so every LDS.128 feeds both chains.  These sums sit in loops bound by the FP64 pipe.
template <int N>       // N = 15: e_0 .. e_7, o_0 .. o_6 (c[15] is never read)
__device__ __forceinline__ double far_poly(const double *__restrict__ c, double t) {
  static_assert(N == 15 && N + 1 <= CHEB_ROW, "15 coefficients in a CHEB_ROW row");
  const double2 *c2 = reinterpret_cast<const double2 *>(c);
  const double u = fma(2.0 * t, t, -1.0);
  double E = c2[7].x;                                    // e_7
  double2 p = c2[6];
  double O = p.y;                                        // o_6
  E = fma(E, u, p.x);
#pragma unroll
  for (int j = 5; j >= 0; --j) {
    p = c2[j];
    E = fma(E, u, p.x);
    O = fma(O, u, p.y);
  }
  return fma(t, O, E);
}

// the rows of the near-branch Voigt table: (e_0, o_0, ..., e_6, o_6)
__device__ __forceinline__ double eo14(const double *__restrict__ c, double t) {
  static_assert(VOIGT_TAB_N == 14, "degree 13");
  const double2 *c2 = reinterpret_cast<const double2 *>(c);
  const double u = fma(2.0 * t, t, -1.0);
  double2 p = c2[6];
  double E = p.x, O = p.y;
#pragma unroll
  for (int j = 5; j >= 0; --j) {
    p = c2[j];
    E = fma(E, u, p.x);
    O = fma(O, u, p.y);
  }
  return fma(t, O, E);
}

__device__ __forceinline__ double rew_table(double x, const double *tab) {
  const double x2 = x + x;                               // 0 <= x < 8: interval k = floor(2x), t = 2(2x - k) - 1
  const int k = (int)x2;
  const double t = 2.0 * (x2 - (double)k) - 1.0;
  return eo14(tab + k * VOIGT_TAB_N, t);
}

// Re w(x + i y) with y = vc->y > 0.  (y == 0 is the Gaussian exp(-x^2): rew_any below; the crystalline sums route that
// case to the closed form before they get here.)
__device__ __forceinline__ double rew(double x, const VoigtConst *vc) {
  x = fabs(x);
  const double y = vc->y;
  const double z2 = fma(x, x, y * y);
  if (z2 >= 64.0) return rew_far(x, y, z2);
  return rew_near(x, vc);
}

// sqrt(pi) Re w(x + i y), y > 0; `tab` (optional) is the pattern's near-branch table.  The crystalline sums use this
// un-normalised form: I(q) is a RATIO of two profile sums (the pattern over its I(0) normalisation,
// scattering/__init__.py:305,331), so the factors 1/sqrt(pi) and 1/(sigma sqrt(2 pi)) of the Voigt profile cancel
// and are not applied per evaluation (two multiplications of ~30).
__device__ __forceinline__ double rew_raw(double x, const VoigtConst *vc, const double *tab = nullptr) {
  x = fabs(x);
  const double y = vc->y;
  const double z2 = fma(x, x, y * y);
  if (z2 >= 64.0) return rew_far_raw(x, y, z2);
  return tab ? rew_table(x, tab) : rew_near(x, vc) * XRSD_SQRT_PI;
}

// any y >= 0
__device__ __forceinline__ double rew_any(double x, const VoigtConst *vc) {
  if (vc->y == 0.0 && x * x >= 64.0) return exp(-x * x);
  return rew(x, vc);
}

}  // namespace xrsd
