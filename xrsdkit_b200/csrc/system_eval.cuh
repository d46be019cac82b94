// system_eval.cuh -- I(q) of one system (noise model + populations) for a span of q held
// by one CTA, accumulated into a shared-memory pattern buffer.
//
// Replaces System.compute_intensity (system/__init__.py:112-130) and everything below it:
//   system/noise.py:50-63, scattering/__init__.py:12-66 (dispatcher), :69-108 (Guinier-Porod),
//   :111-161 (normal size distribution), :286-333 (per-reflection factors, profiles, sum),
//   scattering/structure_factors.py:3-47, scattering/form_factors.py:8-28,
//   tools/peak_math.py:24-47.
//
// One CTA owns q[q_begin, q_begin+q_count) of system b.  Populations are processed in
// layout order (the reference's dict order), each in two steps: a cooperative prologue
// that stages what depends on the parameters only (size quadrature, shell weights,
// Voigt constants, the I(0) normalisation) in shared memory, then a loop in which each
// thread evaluates QPT q-points against the staged table.
//
// Hot loops:
//  * r_normal: radii are equally spaced, so sin/cos(q r_i) are advanced by a rotation
//    (4 DP ops) instead of a sincos per (q, r); for x = q r < 0.5, where the closed form
//    cancels, the Maclaurin series is used instead (forms.cuh).  Weights carry 1/r^6 so
//    the loop has no division: 8 DP instructions per (q, r) pair.
//  * crystalline: sum over |G|-shells of W_s * profile(q - q_s); Voigt via faddeeva.cuh.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <type_traits>
#include "../../include/xrsd_b200.h"
#include "faddeeva.cuh"
#include "cheb_matrices.cuh"
#include "forms.cuh"
#include "lattice.cuh"

namespace xrsd {

constexpr int EVAL_THREADS = 128;
constexpr int SHELL_CHUNK = 128;
#ifndef XRSD_RNORMAL_UNROLL
#define XRSD_RNORMAL_UNROLL 16    // radii per trip of the size-distribution loop (cfg2, same box: 1 -> 1.69 ms, 2 -> 1.53, 4 -> 1.49, 8 -> 1.48, 16 -> 1.46, 25 -> 1.47, 32 -> 1.50)
#endif
constexpr int RNORMAL_UNROLL = XRSD_RNORMAL_UNROLL;
#ifndef XRSD_F32_MIN_X
#define XRSD_F32_MIN_X 2.0      // single-precision path: smallest q r that continues in FP32
#endif

// shared-memory working set of one CTA (besides the pattern buffer)
constexpr int CHEB_N = 15;                       // Chebyshev nodes of the far-field interpolant (degree 14; cheb_matrices.cuh is
                                                 // generated for this N).  Same box, cfg3 kernel and worst golden parity: 17 nodes
                                                 // 2.10 ms / 8.9e-14, 15 nodes 1.98 ms / 2.8e-13, 13 nodes 1.89 ms / 1.8e-11 (gate: 1e-10)
#ifndef XRSD_HX_MIN_CTAS
#define XRSD_HX_MIN_CTAS 8                        // resident CTAs per SM the crystalline kernels are compiled for (64 registers; 7 x 72 and 6 x 80 measured 4 % and 9 % slower)
#endif
#ifndef XRSD_CHEB_ROUND
#define XRSD_CHEB_ROUND 32
#endif
constexpr int CHEB_ROUND = XRSD_CHEB_ROUND;                   // runs of EVAL_THREADS q-points handled between two barriers
                                                 // (measured: 8 -> 16 +3.5 %, 16 -> 32 +1.7 % since the loops got leaner, 64 loses occupancy: -6 %)

struct ChebShared {                               // only in kernels that can meet a crystalline population
  double cheb_cos[CHEB_N * CHEB_N];              // XRSD_FAR_M: node sums -> even / odd power-series coefficients (cheb_matrices.cuh)
  double cheb_node[CHEB_N];                      // the nodes cos(pi (m + 1/2) / N)
  double vt_cos[VOIGT_TAB_N * VOIGT_TAB_N];      // XRSD_VT_M: the same for the nodes of the Voigt near-branch table
  double vt_node[VOIGT_TAB_N];
  double cheb_part[2 * CHEB_ROUND * CHEB_N];     // node sums, two shell-parity halves per node
  alignas(16) double cheb_coef[CHEB_ROUND * CHEB_ROW];   // rows (e_0, o_0, e_1, ...) of far_poly (faddeeva.cuh)
  // per run of a round, double-buffered by round parity (the next round's entries are prepared while this round's
  // points are evaluated): the contiguous range [near_lo, near_hi) of shells summed directly, centre, half-width, 1/half-width
  int near_lo[2][CHEB_ROUND], near_hi[2][CHEB_ROUND];
  double run_c[2][CHEB_ROUND], run_H[2][CHEB_ROUND], run_invH[2][CHEB_ROUND];
  alignas(16) double voigt_tab[VOIGT_TAB_STORE];   // near-branch Faddeeva table of the current population (faddeeva.cuh)
};
static_assert(VOIGT_TAB_DOUBLES <= 2 * CHEB_ROUND * CHEB_N, "cheb_part doubles as the table's node buffer");
struct NoCheb {};

struct EvalCore {
  VoigtConst vc;
  Cell cell;
  double red[EVAL_THREADS / 32];
};
template <bool HAS_XTAL>
struct EvalSharedT : EvalCore, std::conditional<HAS_XTAL, ChebShared, NoCheb>::type {};
using EvalShared = EvalCore;                     // what the helpers below need

__device__ __forceinline__ double block_sum(double v, EvalShared *S) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) S->red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < EVAL_THREADS / 32; ++w) t += S->red[w];
  return t;
}

// peak profiles: tools/peak_math.py:24-47.  The per-pattern constants are folded into
// reciprocals once (make_profile) so that the per-(q, shell) evaluation has no division except
// the 1/|z|^2 of the Faddeeva far field; this reorders the reference's x/sigma/sqrt(2) by <= 2 ulp.
struct Profile {
  int kind;
  double a;          // voigt: 1/(sigma sqrt 2)      gaussian: 1/h            lorentzian: h^2
  double b;          // voigt: 1/(sigma sqrt(2 pi))  gaussian: sqrt(ln2/pi)/h lorentzian: h/pi
};

// returns y = gamma/(sigma sqrt 2) for voigt (0 otherwise)
__device__ __forceinline__ double make_profile(Profile &P, int kind, double p0, double p1) {
  P.kind = kind;
  if (kind == XRSD_PROFILE_VOIGT) {
    const double sigma = p0 / 1.1774100225154747;                  // hwhm_g / sqrt(2 ln 2)
    P.a = 1.0 / sigma / M_SQRT2;
    P.b = 1.0 / sigma / 2.5066282746310002;                        // 1/sigma/sqrt(2 pi)
    return p1 / sigma / M_SQRT2;
  }
  if (kind == XRSD_PROFILE_GAUSSIAN) {
    P.a = 1.0 / p0;
    P.b = 0.46971863934982566 / p0;                                // sqrt(ln2/pi)/h
  } else {
    P.a = p0 * p0;
    P.b = p0 / M_PI;
  }
  return 0.0;
}

// VOIGT is a compile-time property of the loops that call this (the crystalline branch is instantiated once for Voigt
// profiles and once for the two closed forms): the kind test per (q, shell) evaluation was 5 % of the kernel's
// instructions (profiles/sass_by_line.py on the cfg3 capture).
template <bool VOIGT>
__device__ __forceinline__ double profile_eval(const Profile &P, const VoigtConst *vc, double x, const double *vtab = nullptr) {
  if constexpr (VOIGT) return rew_raw(x * P.a, vc, vtab);      // un-normalised: the constant factors cancel in I(q) (faddeeva.cuh)
  if (P.kind == XRSD_PROFILE_GAUSSIAN) {
    const double u = x * P.a;
    return P.b * exp(-(u * u) * 0.6931471805599453);               // exp(-(x/h)**2 * ln 2)
  }
  return P.b / fma(x, x, P.a);
}
// the same for an argument known to lie in the far branch (|z|^2 >= 64): the shells a far-field node sum runs over are
// at least 12 core widths from every node, so the branch test of rew_raw is not needed there
template <bool VOIGT>
__device__ __forceinline__ double profile_far(const Profile &P, const VoigtConst *vc, double x) {
  if constexpr (VOIGT) {
    const double a = fabs(x * P.a), y = vc->y;
    return rew_far_raw(a, y, fma(a, a, y * y));
  } else {
    return profile_eval<false>(P, vc, x);
  }
}
__device__ __forceinline__ double profile_eval(const Profile &P, const VoigtConst *vc, double x, const double *vtab = nullptr) {
  return P.kind == XRSD_PROFILE_VOIGT ? rew_any(x * P.a, vc) * P.b : profile_eval<false>(P, vc, x, vtab);
}

// form factor of species `a` of a crystalline population at q
__device__ __forceinline__ double species_ff(const xrsd_pop_t &pop, int a, const double *prm, double q) {
  if (pop.species_form[a] == XRSD_FORM_SPHERE) return sphere_ff(q, prm[pop.p_r]);   // q == 0 -> NaN, as the reference
  return atomic_ff(q, pop.atom_Z[a], pop.atom_a[a], pop.atom_b[a]);
}

__device__ __forceinline__ int pair_index(int n_sp, int a, int b) { return a * n_sp - a * (a - 1) / 2 + (b - a); }

// ------------------------------------------------------------------------------------------
// Accumulate I(q) of system `b` for q[q_begin .. q_begin+q_count) into I_sm[0 .. q_count).
// All EVAL_THREADS threads of the CTA must call this.  `scratch` holds scratch_doubles doubles.
// ------------------------------------------------------------------------------------------
template <int QPT, bool RADIAL_MULTI, bool HAS_XTAL = true, bool F32 = false>
__device__ void accumulate_system(const xrsd_layout_t &L, const double *__restrict__ prm, const xrsd_tables_t &T, int b,
                                  const double *__restrict__ q, int n_q, int q_begin, int q_count,
                                  double *__restrict__ I_sm, double *__restrict__ scratch, int scratch_doubles,
                                  EvalSharedT<HAS_XTAL> *S, const double *__restrict__ prm_topo = nullptr,
                                  double *__restrict__ cum_out = nullptr, long long cum_stride = 0, int only_pop = -1) {
  // only_pop >= 0: evaluate that population (or the noise model, index 0) alone -- what a finite-difference variant
  // needs when its parameter belongs to one population: the other terms cancel in the difference.
  // cum_out (optional, global memory): after population ip has been added, the running total of this
  // CTA's span is written to cum_out[ip * cum_stride + i].  The Jacobian uses the differences as the
  // per-population contributions, which give the derivatives of the linear parameters (I0) for free.
  // prm_topo (optional) is the parameter row that decides the DISCRETE part of the evaluation --
  // the number of radii of each size quadrature -- while `prm` supplies the values.  The variants of
  // a finite-difference Jacobian pass their unperturbed row here so that a step in r or sigma cannot
  // flip ceil() and put a jump into the difference quotient.
  if (prm_topo == nullptr) prm_topo = prm;
  const int tid = threadIdx.x;
  // Nothing is written until a population has per-q values: a leading flat noise level is carried as a constant
  // (`pending`) and folded into the first per-q write, which is then a plain store -- no zero fill, no separate
  // noise pass, and in direct mode (I_sm in global memory) no read for the first population.
  double pending = 0.0;
  bool fresh = true;                   // CTA-uniform: I_sm holds nothing yet
  const int span = EVAL_THREADS * QPT;
  // strictly ascending q over the span enables the far-field interpolation of crystalline sums
  bool q_sorted = false;
  if constexpr (HAS_XTAL) {
    if (L.q_step > 0.0) {                       // a uniform ascending grid (checked on the host): nothing to test
      q_sorted = true;
    } else {
      bool mine = true;
      for (int i = tid; i + 1 < q_count; i += EVAL_THREADS) mine = mine && (q[q_begin + i] < q[q_begin + i + 1]);
      q_sorted = __syncthreads_and(mine) != 0;
    }
    for (int i = tid; i < CHEB_N * CHEB_N; i += EVAL_THREADS) S->cheb_cos[i] = XRSD_FAR_M[i];
    for (int i = tid; i < VOIGT_TAB_N * VOIGT_TAB_N; i += EVAL_THREADS) S->vt_cos[i] = XRSD_VT_M[i];
    if (tid < CHEB_N) S->cheb_node[tid] = XRSD_FAR_NODES[tid];
    if (tid >= 32 && tid < 32 + VOIGT_TAB_N) S->vt_node[tid - 32] = XRSD_VT_NODES[tid - 32];
  }

  for (int ip = 0; ip < L.n_pops; ++ip) {
    if (only_pop >= 0 && ip != only_pop) continue;
    const xrsd_pop_t &pop = L.pops[ip];
    __syncthreads();
    bool wrote = false;                // this population stored per-q values for the whole span
    switch (pop.kind) {
      // ---------------------------------------------------------------- noise (system/noise.py:50-63)
      case XRSD_POP_NOISE_FLAT: {
        const double I0 = prm[pop.p_I0];
        if (fresh) {
          pending += I0 * 1.0;
        } else {
          for (int i = tid; i < q_count; i += EVAL_THREADS) I_sm[i] += I0 * 1.0;
        }
      } break;
      case XRSD_POP_NOISE_LOW_Q: {
        const double I0 = prm[pop.p_I0], f = prm[pop.p_flat_fraction];
        const double beam = I0 * (1.0 - f);
        const GPConst gp = gp_constants(prm[pop.p_rg], prm[pop.p_D]);
        for (int i = tid; i < q_count; i += EVAL_THREADS) {
          double v = fresh ? pending : I_sm[i];
          v += beam * guinier_porod(q[q_begin + i], gp);
          v += I0 * f * 1.0;
          I_sm[i] = v;
        }
        wrote = true;
      } break;
      // ---------------------------------------------------------------- diffuse / disordered
      case XRSD_POP_NONCRYSTALLINE: {
        wrote = true;
        const double I0 = prm[pop.p_I0];
        HSConst hs;
        if (pop.disordered) hs = hs_constants(prm[pop.p_r_hard], prm[pop.p_v_fraction]);
        GPConst gp;
        if (pop.form == XRSD_FORM_GUINIER_POROD) gp = gp_constants(prm[pop.p_rg], prm[pop.p_D]);
        // size quadrature prologue (scattering/__init__.py:148-160, tools/__init__.py:168-173)
        bool quad = false;
        int n_r = 0;
        double r0 = 0.0, dstep = 0.0, norm = 1.0, r_first = 0.0;
        if (pop.form == XRSD_FORM_SPHERE || pop.form == XRSD_FORM_SPHERE_R_NORMAL) r0 = prm[pop.p_r];
        if (pop.form == XRSD_FORM_SPHERE_R_NORMAL) {
          const double sigma = prm[pop.p_sigma];
          if (!(sigma < 1.0e-9)) {
            quad = true;
            // unfused arithmetic: a fused multiply-add here can flip ceil() and change the
            // number of radii relative to numpy (tools/__init__.py:168-173, np.arange)
            const double sigma_r = __dmul_rn(sigma, r0);
            const double dr = __dmul_rn(sigma_r, pop.sampling_step);
            const double r_min = fmax(__dsub_rn(r0, __dmul_rn(pop.sampling_width, sigma_r)), dr);
            const double r_max = __dadd_rn(r0, __dmul_rn(pop.sampling_width, sigma_r));
            // numpy.arange(r_min, r_max, dr): len = ceil((stop-start)/step), r_i = start + i*((start+step)-start)
            const double len = ceil(__dsub_rn(r_max, r_min) / dr);
            double len_t = len;
            if (prm_topo != prm) {
              const double r0t = prm_topo[pop.p_r];
              const double srt = __dmul_rn(prm_topo[pop.p_sigma], r0t);
              const double drt = __dmul_rn(srt, pop.sampling_step);
              const double lo_t = fmax(__dsub_rn(r0t, __dmul_rn(pop.sampling_width, srt)), drt);
              const double hi_t = __dadd_rn(r0t, __dmul_rn(pop.sampling_width, srt));
              len_t = ceil(__dsub_rn(hi_t, lo_t) / drt);
            }
            // the scratch holds 3 doubles per radius: (r_i, rho_i) as doubles, then -- single-precision path -- as float2
            n_r = (len_t > 0.0) ? (int)fmin(len_t, (double)(scratch_doubles / 3)) : 0;
            dstep = __dsub_rn(__dadd_rn(r_min, dr), r_min);
            r_first = r_min;
            const double two_s2 = 2.0 * (sigma_r * sigma_r);
            double part = 0.0;
            for (int i = tid; i < n_r; i += EVAL_THREADS) {
              const double ri = __dadd_rn(r_min, __dmul_rn((double)i, dstep));   // unfused, as numpy.arange
              const double rho = exp(-1.0 * ((r0 - ri) * (r0 - ri)) / two_s2);
              scratch[2 * i] = ri;
              scratch[2 * i + 1] = rho;
              if constexpr (F32) reinterpret_cast<float2 *>(scratch + 2 * n_r)[i] = make_float2((float)ri, (float)rho);
              const double r3 = ri * ri * ri;
              part += rho * (r3 * r3);
            }
            norm = block_sum(part, S);     // sum_i rho_i r_i^6  (common factors of I0_i cancel)
          }
        }
        __syncthreads();
        // On a uniform q grid consecutive spans of one thread differ by span*q_step, so the starting
        // phases (q r_first, q dstep) of the next span follow from the previous ones by a fixed rotation
        double s0[QPT], c0[QPT], sd0[QPT], cd0[QPT], sh[QPT], ch[QPT];
        double rot_s = 0.0, rot_c = 1.0, rotd_s = 0.0, rotd_c = 1.0, roth_s = 0.0, roth_c = 1.0;
        const bool carry = quad && (L.q_step > 0.0);
        const bool carry_hs = pop.disordered && (L.q_step > 0.0);     // same for sin/cos(q d) of the hard-sphere S(q)
        if (carry) {
          const double dq = (double)span * L.q_step;
          sincos(dq * r_first, &rot_s, &rot_c);
          sincos(dq * dstep, &rotd_s, &rotd_c);
        }
        if (carry_hs) sincos((double)span * L.q_step * hs.d, &roth_s, &roth_c);
        for (int off = 0; off < q_count; off += span) {
          double qv[QPT], ff2[QPT], den[QPT];
          bool ok[QPT], q0[QPT];
#pragma unroll
          for (int j = 0; j < QPT; ++j) {
            const int i = off + tid + j * EVAL_THREADS;
            ok[j] = i < q_count;
            qv[j] = q[q_begin + (ok[j] ? i : q_count - 1)];
            q0[j] = L.q0_is_zero && (q_begin + i == 0);
            ff2[j] = 0.0;
            den[j] = 1.0;
          }
          if (pop.form == XRSD_FORM_ATOMIC) {
#pragma unroll
            for (int j = 0; j < QPT; ++j) ff2[j] = sqr(atomic_ff(qv[j], pop.atom_Z[0], pop.atom_a[0], pop.atom_b[0]));
          } else if (pop.form == XRSD_FORM_GUINIER_POROD) {
#pragma unroll
            for (int j = 0; j < QPT; ++j) ff2[j] = guinier_porod(qv[j], gp);
          } else if (!quad) {
#pragma unroll
            for (int j = 0; j < QPT; ++j) ff2[j] = q0[j] ? 1.0 : sqr(sphere_ff(qv[j], r0));
          } else {
            // rotation recurrence over the radius grid
            double s[QPT], c[QPT], sd[QPT], cd[QPT], acc[QPT];
            double qmin_t = qv[0];
#pragma unroll
            for (int j = 0; j < QPT; ++j) {
              if (carry && off > 0) {
                const double sn = fma(s0[j], rot_c, c0[j] * rot_s);
                c0[j] = fma(c0[j], rot_c, -(s0[j] * rot_s));
                s0[j] = sn;
                const double sdn = fma(sd0[j], rotd_c, cd0[j] * rotd_s);
                cd0[j] = fma(cd0[j], rotd_c, -(sd0[j] * rotd_s));
                sd0[j] = sdn;
              } else {
                sincos(qv[j] * r_first, &s0[j], &c0[j]);
                sincos(qv[j] * dstep, &sd0[j], &cd0[j]);
              }
              s[j] = s0[j]; c[j] = c0[j]; sd[j] = sd0[j]; cd[j] = cd0[j];
              acc[j] = 0.0;
              qmin_t = fmin(qmin_t, qv[j]);
            }
            for (int o = 16; o > 0; o >>= 1) qmin_t = fmin(qmin_t, __shfl_xor_sync(0xffffffffu, qmin_t, o));
            int i = 0;
            // phase A: some lane of the warp still has x < SPHERE_SERIES_X -> series where needed
            for (; i < n_r; ++i) {
              const double ri = scratch[2 * i], rho = scratch[2 * i + 1];
              if (!(qmin_t * ri < SPHERE_SERIES_X)) break;
#pragma unroll
              for (int j = 0; j < QPT; ++j) {
                const double x = qv[j] * ri;
                double t = fma(-x, c[j], s[j]);
                if (x < SPHERE_SERIES_X) t = sphere_t_series(x);
                acc[j] = fma(rho, t * t, acc[j]);
                const double sn = fma(s[j], cd[j], c[j] * sd[j]);
                c[j] = fma(c[j], cd[j], -(s[j] * sd[j]));
                s[j] = sn;
              }
            }
            if constexpr (F32) {
              // The optional single-precision path (BASELINE north_star: <= 1e-5 relative).  Radii with q r >= XRSD_F32_MIN_X for
              // the whole warp -- where t = sin x - x cos x is O(x) and an absolute error of a few 1e-6 in sin / cos
              // does no harm -- continue the SAME recurrence in FP32: 8 FP32 instructions per (q, r) pair on the
              // full-rate pipe instead of 8 FP64.  Everything else stays FP64: the radii below that (t cancels like
              // x^3/3 there), the starting phases and their span-to-span carry (an FP32 product q r is off by 4e-6 rad
              // at x = 35), the normalisation, the Percus-Yevick factor, the accumulation of the populations.
              for (; i < n_r; ++i) {       // phase B64: FP64 recurrence up to q r = 2
                const double2 rr = *reinterpret_cast<const double2 *>(scratch + 2 * i);
                if (!(qmin_t * rr.x < XRSD_F32_MIN_X)) break;
#pragma unroll
                for (int j = 0; j < QPT; ++j) {
                  const double x = qv[j] * rr.x;
                  const double t = fma(-x, c[j], s[j]);
                  acc[j] = fma(rr.y, t * t, acc[j]);
                  const double sn = fma(s[j], cd[j], c[j] * sd[j]);
                  c[j] = fma(c[j], cd[j], -(s[j] * sd[j]));
                  s[j] = sn;
                }
              }
              float sf[QPT], cf[QPT], sdf[QPT], cdf[QPT], accf[QPT], qf[QPT];
#pragma unroll
              for (int j = 0; j < QPT; ++j) {
                sf[j] = (float)s[j]; cf[j] = (float)c[j]; sdf[j] = (float)sd[j]; cdf[j] = (float)cd[j];
                qf[j] = (float)qv[j];
                accf[j] = 0.0f;
              }
              const float2 *rec32 = reinterpret_cast<const float2 *>(scratch + 2 * n_r);
#pragma unroll RNORMAL_UNROLL
              for (; i < n_r; ++i) {       // phase B32
                const float2 rw = rec32[i];
                const float rf = rw.x, wf = rw.y;
#pragma unroll
                for (int j = 0; j < QPT; ++j) {
                  const float x = qf[j] * rf;
                  const float t = fmaf(-x, cf[j], sf[j]);
                  accf[j] = fmaf(wf, t * t, accf[j]);
                  const float sn = fmaf(sf[j], cdf[j], cf[j] * sdf[j]);
                  cf[j] = fmaf(cf[j], cdf[j], -(sf[j] * sdf[j]));
                  sf[j] = sn;
                }
              }
#pragma unroll
              for (int j = 0; j < QPT; ++j) acc[j] += (double)accf[j];
            }
            // phase B: pure recurrence, 8 DP instructions per (q, r)
#pragma unroll RNORMAL_UNROLL
            for (; i < n_r; ++i) {
              const double2 rr = *reinterpret_cast<const double2 *>(scratch + 2 * i);
#pragma unroll
              for (int j = 0; j < QPT; ++j) {
                const double x = qv[j] * rr.x;
                const double t = fma(-x, c[j], s[j]);
                acc[j] = fma(rr.y, t * t, acc[j]);
                const double sn = fma(s[j], cd[j], c[j] * sd[j]);
                c[j] = fma(c[j], cd[j], -(s[j] * sd[j]));
                s[j] = sn;
              }
            }
#pragma unroll
            for (int j = 0; j < QPT; ++j) {
              // ff2 = 9 acc / (q^6 norm), kept as numerator / denominator: one division below
              const double q2 = qv[j] * qv[j];
              ff2[j] = q0[j] ? 1.0 : 9.0 * acc[j];
              den[j] = q0[j] ? 1.0 : (q2 * q2 * q2) * norm;
            }
          }
          if (pop.disordered) {
#pragma unroll
            for (int j = 0; j < QPT; ++j) {
              if (carry_hs && off > 0) {
                const double sn = fma(sh[j], roth_c, ch[j] * roth_s);
                ch[j] = fma(ch[j], roth_c, -(sh[j] * roth_s));
                sh[j] = sn;
              } else {
                sincos(qv[j] * hs.d, &sh[j], &ch[j]);
              }
            }
          }
#pragma unroll
          for (int j = 0; j < QPT; ++j) {
            if (!ok[j]) continue;
            const int i = off + tid + j * EVAL_THREADS;
            double num = I0 * ff2[j], dn = den[j];
            if (pop.disordered) {                                  // I0 * sf * ff_sqr, sf = (1-nc0)/(1-nc)
              num *= (1.0 - hs.nc0);
              dn *= (1.0 - hard_sphere_nc(qv[j], hs, q0[j], sh[j], ch[j]));
            }
            I_sm[i] = (fresh ? pending : I_sm[i]) + num / dn;
          }
        }
      } break;
      // ---------------------------------------------------------------- crystalline
      case XRSD_POP_CRYSTALLINE: if constexpr (HAS_XTAL) {
        const int t_own = b * L.n_xtal + pop.table_slot;
        const int t = T.table_of != nullptr ? T.table_of[t_own] : T.own_base + t_own;      // storage row (xrsd_tables_t)
        const int n_shell = T.n_shell[t];
        if (n_shell <= 0) break;                                   // scattering/__init__.py:277-278
        wrote = true;
        const int n_sp = pop.n_species;
        const int n_pairs = n_sp * (n_sp + 1) / 2;
        const int32_t *sh_hkl = T.shell_hkl + (size_t)t * T.cap_shell * 4;
        const double *sh_geo = T.shell_geo + (size_t)t * T.cap_shell * T.n_pairs_max;
        const double wl = L.wavelength;
        const bool radial = pop.sf_radial != 0;
        const bool radial_multi = RADIAL_MULTI && radial && n_sp > 1;
        Profile P;
        if (tid == 0) make_cell(pop, prm, &S->cell);
        // A Voigt profile of zero Lorentzian width IS the Gaussian of the same hwhm_g (Re w(x) = exp(-x^2) on the real
        // axis; tools/peak_math.py:24-47 agree to rounding), so it takes the closed form: the Faddeeva code then never
        // sees y == 0 and needs no test for it.
        const bool voigt = pop.profile == XRSD_PROFILE_VOIGT && prm[pop.p_hwhm_l] != 0.0;
        if (voigt) {
          const double y = make_profile(P, XRSD_PROFILE_VOIGT, prm[pop.p_hwhm_g], prm[pop.p_hwhm_l]);
          voigt_constants(&S->vc, y, tid);
        } else if (pop.profile == XRSD_PROFILE_VOIGT) {
          make_profile(P, XRSD_PROFILE_GAUSSIAN, prm[pop.p_hwhm_g], 0.0);
        } else {
          make_profile(P, pop.profile, prm[pop.p_hwhm], 0.0);
        }
        __syncthreads();
        auto body = [&](auto voigt_c) {
        constexpr bool VG = decltype(voigt_c)::value;
        // pass A: normalisation I(0) = sum_s W0_s * profile(0 - q_s)   (scattering/__init__.py:305,331)
        double part = 0.0;
        // does any peak core (|q - q_s| < 8 sigma sqrt2, the Faddeeva near branch) reach into this span?
        const bool voigt_near_possible = VG && S->vc.y < 8.0 && S->vc.y >= VOIGT_TAB_MIN_Y;
        const double core_reach = voigt_near_possible ? 8.0 / P.a : 0.0;
        const double span_lo = q_sorted ? q[q_begin] : -INFINITY, span_hi = q_sorted ? q[q_begin + q_count - 1] : INFINITY;
        bool core_here = false;
        for (int s = tid; s < n_shell; s += EVAL_THREADS) {
          const double qs = 2.0 * M_PI * g_norm(S->cell.b, (double)sh_hkl[4 * s], (double)sh_hkl[4 * s + 1], (double)sh_hkl[4 * s + 2]);
          double ltz = 1.0;
          if (pop.lorentz) {
            const double th = asin(wl * qs / (4.0 * M_PI));
            ltz = 1.0 / (sin(th) * sin(2.0 * th));
          }
          double w0 = 0.0;
          int ipq = 0;
          for (int a = 0; a < n_sp; ++a)
            for (int bb = a; bb < n_sp; ++bb, ++ipq) w0 += ((a == bb) ? 1.0 : 2.0) * sh_geo[(size_t)s * T.n_pairs_max + ipq];
          part += ltz * w0 * profile_eval<VG>(P, &S->vc, 0.0 - qs);
          core_here = core_here || (qs > span_lo - core_reach && qs < span_hi + core_reach);
        }
        const double norm = block_sum(part, S);
        const double *vtab = nullptr;
        if (voigt_near_possible && __syncthreads_or(core_here)) {
          voigt_table(S->voigt_tab, S->cheb_part, S->vt_cos, S->vt_node, &S->vc, tid, EVAL_THREADS);
          vtab = S->voigt_tab;
        }
        // pass B: chunks of shells staged as (q_s, W_s[...]) in scratch
        const int wstride = radial_multi ? n_pairs : 1;
        const int rec = 1 + wstride;
        int chunk = scratch_doubles / rec;
        if (chunk > SHELL_CHUNK) chunk = SHELL_CHUNK;
        const bool stage_once = n_shell <= chunk;
        // (q_s, W_s[...]) of shells [s0, s0+ns) into scratch
        auto stage_shells = [&](int s0, int ns) {
            for (int sl = tid; sl < ns; sl += EVAL_THREADS) {
            const int s = s0 + sl;
            const double qs = 2.0 * M_PI * g_norm(S->cell.b, (double)sh_hkl[4 * s], (double)sh_hkl[4 * s + 1], (double)sh_hkl[4 * s + 2]);
            double ltz = 1.0;
            if (pop.lorentz) {
              const double th = asin(wl * qs / (4.0 * M_PI));
              ltz = 1.0 / (sin(th) * sin(2.0 * th));
            }
            scratch[sl * rec] = qs;
            if (radial_multi) {
              int ipq = 0;
              for (int a = 0; a < n_sp; ++a)
                for (int bb = a; bb < n_sp; ++bb, ++ipq)
                  scratch[sl * rec + 1 + ipq] = ltz * ((a == bb) ? 1.0 : 2.0) * sh_geo[(size_t)s * T.n_pairs_max + ipq];
            } else if (radial) {
              scratch[sl * rec + 1] = ltz * sh_geo[(size_t)s * T.n_pairs_max];
            } else {
              double ffs[XRSD_MAX_SPECIES];
              for (int a = 0; a < n_sp; ++a) ffs[a] = species_ff(pop, a, prm, qs);
              double w = 0.0;
              int ipq = 0;
              for (int a = 0; a < n_sp; ++a)
                for (int bb = a; bb < n_sp; ++bb, ++ipq)
                  w += ((a == bb) ? 1.0 : 2.0) * ffs[a] * ffs[bb] * sh_geo[(size_t)s * T.n_pairs_max + ipq];
              scratch[sl * rec + 1] = ltz * w;
            }
          }
        };
        const double I0 = prm[pop.p_I0];
        // per-pattern reciprocals: the per-q finish has no division (<= 2 ulp from the reference's order)
        const double inv_norm = 1.0 / norm, wl_4pi = wl / (4.0 * M_PI);
        // per-q finish shared by both paths: radial form factor, polarisation, normalisation, I0
        auto finish_q = [&](int i, double qi, double a_sum) {
          if (radial && !radial_multi) {
            const bool q0 = L.q0_is_zero && (q_begin + i == 0);
            const double f = (q0 && pop.species_form[0] == XRSD_FORM_SPHERE) ? 1.0 : species_ff(pop, 0, prm, qi);
            a_sum *= f * f;
          }
          double pz = 1.0;
          if (pop.polz) {
            // 0.5 (1 + cos^2(2 theta)), theta = asin(lambda q / 4 pi)  (scattering/__init__.py:234-238):
            // cos(2 theta) = 1 - 2 sin^2(theta), so no transcendental is needed; |sin| > 1 gives the
            // reference's NaN
            const double st = wl_4pi * qi;
            const double c2 = (fabs(st) <= 1.0) ? fma(-2.0 * st, st, 1.0) : NAN;
            pz = 0.5 * (1.0 + c2 * c2);
          }
          I_sm[i] = (fresh ? pending : I_sm[i]) + I0 * (pz * a_sum * inv_norm);   // I0 * (pz*I/I0_norm)
        };
        // ---- far-field interpolation path ------------------------------------------------------
        // For a run of EVAL_THREADS consecutive q-points (centre c, half-width H) every shell with
        // |q_s - c| >= max(4 H, H + 12 core widths) is "far": the sum of all far shells is smooth over
        // the run (nearest singular structure >= 3 H from its edge: Bernstein parameter rho >= 7.9, rho^-15 = 3.6e-14), so it is evaluated at CHEB_N Chebyshev
        // nodes only and interpolated (degree 14: worst golden parity 2.8e-13, tests/test_gpu_parity.py),
        // while the few near shells are summed directly.  This cuts the profile evaluations per q from
        // n_shell to about n_near + CHEB_N n_far / EVAL_THREADS.
        const bool interp = stage_once && !radial_multi && q_sorted && P.kind != XRSD_PROFILE_GAUSSIAN &&
                            n_shell >= 8 && q_count >= 32;
        if (interp) {
          __syncthreads();
          stage_shells(0, n_shell);
          __syncthreads();
          const double core = VG ? fmax(1.0, S->vc.y) / P.a : sqrt(P.a);
          // rounds of CHEB_ROUND runs: (0) near range of every run, (1) node sums of all runs, (2) coefficients,
          // (3) per-q evaluation; three barriers per round instead of four per run
          auto prepare_round = [&](int r0, int buf) {       // (0): one thread per run; the table is sorted by |G|, so the
            const int n_run = min(CHEB_ROUND, (q_count - r0 + EVAL_THREADS - 1) / EVAL_THREADS);      // near shells are one range
            if (tid < n_run) {
              const int run = tid;
              const int i0 = r0 + run * EVAL_THREADS, i1 = min(q_count, i0 + EVAL_THREADS);
              const double qa = q[q_begin + i0], qb = q[q_begin + i1 - 1];
              const double c = 0.5 * (qa + qb), H = 0.5 * (qb - qa);
              const double dn = fmax(4.0 * H, H + 12.0 * core);     // far: >= 12 core widths from every node, |z|^2 >= 144
              int lo = n_shell, hi = 0;
              for (int sl = 0; sl < n_shell; ++sl)
                if (fabs(scratch[2 * sl] - c) < dn) { lo = min(lo, sl); hi = max(hi, sl + 1); }
              if (hi == 0) lo = hi = n_shell;               // no near shell: everything is far
              S->near_lo[buf][run] = lo;
              S->near_hi[buf][run] = hi;
              S->run_c[buf][run] = c;
              S->run_H[buf][run] = H;
              S->run_invH[buf][run] = (H > 0.0) ? 1.0 / H : 0.0;
            }
          };
          prepare_round(0, 0);
          __syncthreads();
          int buf = 0;
          for (int r0 = 0; r0 < q_count; r0 += CHEB_ROUND * EVAL_THREADS, buf ^= 1) {
            const int n_run = min(CHEB_ROUND, (q_count - r0 + EVAL_THREADS - 1) / EVAL_THREADS);
            // (1) item = (run, node m, shell parity g): the far shells below and above the near range, no test per shell
            for (int item = tid; item < n_run * CHEB_N * 2; item += EVAL_THREADS) {
              const int run = item / (2 * CHEB_N), rem = item - run * 2 * CHEB_N;
              const int g = rem / CHEB_N, m = rem - g * CHEB_N;
              const double xm = fma(S->run_H[buf][run], S->cheb_node[m], S->run_c[buf][run]);
              const int lo = S->near_lo[buf][run], hi = S->near_hi[buf][run];
              double part = 0.0;
              for (int sl = g; sl < lo; sl += 2) {
                const double2 sw = *reinterpret_cast<const double2 *>(scratch + 2 * sl);
                part = fma(sw.y, profile_far<VG>(P, &S->vc, xm - sw.x), part);
              }
              for (int sl = hi + ((hi ^ g) & 1); sl < n_shell; sl += 2) {
                const double2 sw = *reinterpret_cast<const double2 *>(scratch + 2 * sl);
                part = fma(sw.y, profile_far<VG>(P, &S->vc, xm - sw.x), part);
              }
              S->cheb_part[item] = part;
            }
            __syncthreads();
            // (2) coefficient (run, k) of the even / odd power series in u = 2 t^2 - 1: row k of XRSD_FAR_M times the node sums
            for (int item = tid; item < n_run * CHEB_N; item += EVAL_THREADS) {
              const int run = item / CHEB_N, k = item - run * CHEB_N;
              const double *pa = S->cheb_part + run * 2 * CHEB_N;
              double ak = 0.0;
#pragma unroll
              for (int m = 0; m < CHEB_N; ++m) ak = fma(pa[m] + pa[CHEB_N + m], S->cheb_cos[k * CHEB_N + m], ak);
              S->cheb_coef[run * CHEB_ROW + k] = ak;
            }
            if (r0 + CHEB_ROUND * EVAL_THREADS < q_count) prepare_round(r0 + CHEB_ROUND * EVAL_THREADS, buf ^ 1);
            __syncthreads();
            // (3) one q-point per thread and run
            for (int run = 0; run < n_run; ++run) {
              const int i0 = r0 + run * EVAL_THREADS, i1 = min(q_count, i0 + EVAL_THREADS);
              const int i = i0 + tid;
              if (i >= i1) continue;
              const double qi = q[q_begin + i];
              // the interpolant at t in [-1, 1] (two Horner chains in u: faddeeva.cuh)
              const double t = (qi - S->run_c[buf][run]) * S->run_invH[buf][run];
              double a_sum = far_poly<CHEB_N>(S->cheb_coef + run * CHEB_ROW, t);
              const int s_lo = S->near_lo[buf][run], s_hi = S->near_hi[buf][run];
              for (int sl = s_lo; sl < s_hi; ++sl) {
                const double2 sw = *reinterpret_cast<const double2 *>(scratch + 2 * sl);
                a_sum = fma(sw.y, profile_eval<VG>(P, &S->vc, qi - sw.x, vtab), a_sum);
              }
              finish_q(i, qi, a_sum);
            }
            __syncthreads();
          }
          return;
        }
        for (int off = 0; off < q_count; off += span) {
          double qv[QPT], acc[QPT];
          double accp[QPT][RADIAL_MULTI ? XRSD_MAX_PAIRS : 1];
          bool ok[QPT];
#pragma unroll
          for (int j = 0; j < QPT; ++j) {
            const int i = off + tid + j * EVAL_THREADS;
            ok[j] = i < q_count;
            qv[j] = q[q_begin + (ok[j] ? i : q_count - 1)];
            acc[j] = 0.0;
          }
#pragma unroll
          for (int j = 0; j < QPT; ++j)
#pragma unroll
            for (int k = 0; k < (RADIAL_MULTI ? XRSD_MAX_PAIRS : 1); ++k) accp[j][k] = 0.0;
          for (int s0 = 0; s0 < n_shell; s0 += chunk) {
            const int ns = min(chunk, n_shell - s0);
            // (q_s, W_s) of this chunk; when all shells fit in one chunk they are staged once per
            // population instead of once per q-span
            if (!(stage_once && off > 0)) {
            __syncthreads();
            stage_shells(s0, ns);
            __syncthreads();
            }
            if (!radial_multi) {
              for (int sl = 0; sl < ns; ++sl) {
                const double2 sw = *reinterpret_cast<const double2 *>(scratch + 2 * sl);
#pragma unroll
                for (int j = 0; j < QPT; ++j) acc[j] = fma(sw.y, profile_eval<VG>(P, &S->vc, qv[j] - sw.x, vtab), acc[j]);
              }
            } else {
              for (int sl = 0; sl < ns; ++sl) {
                const double qs = scratch[sl * rec];
#pragma unroll
                for (int j = 0; j < QPT; ++j) {
                  const double v = profile_eval<VG>(P, &S->vc, qv[j] - qs, vtab);
#pragma unroll
                  for (int k = 0; k < (RADIAL_MULTI ? XRSD_MAX_PAIRS : 1); ++k)
                    if (k < n_pairs) accp[j][k] = fma(scratch[sl * rec + 1 + k], v, accp[j][k]);
                }
              }
            }
          }
#pragma unroll
          for (int j = 0; j < QPT; ++j) {
            if (!ok[j]) continue;
            const int i = off + tid + j * EVAL_THREADS;
            double a_sum = acc[j];
            if (radial) {
              double ffq[XRSD_MAX_SPECIES];
              const bool q0 = L.q0_is_zero && (q_begin + i == 0);
              for (int a = 0; a < n_sp; ++a) ffq[a] = (q0 && pop.species_form[a] == XRSD_FORM_SPHERE) ? 1.0 : species_ff(pop, a, prm, qv[j]);
              if (radial_multi) {
                a_sum = 0.0;
#pragma unroll
                for (int a = 0; a < XRSD_MAX_SPECIES; ++a)
#pragma unroll
                  for (int bb = a; bb < XRSD_MAX_SPECIES; ++bb)
                    if (RADIAL_MULTI && bb < n_sp) a_sum += ffq[a] * ffq[bb] * accp[j][RADIAL_MULTI ? pair_index(n_sp, a, bb) : 0];
              } else {
                a_sum = acc[j] * (ffq[0] * ffq[0]);
              }
            }
            double pz = 1.0;
            if (pop.polz) {
              const double st = wl * qv[j] / (4.0 * M_PI);          // see finish_q
              const double c2 = (fabs(st) <= 1.0) ? fma(-2.0 * st, st, 1.0) : NAN;
              pz = 0.5 * (1.0 + c2 * c2);
            }
            I_sm[i] = (fresh ? pending : I_sm[i]) + I0 * (pz * a_sum / norm);     // I0 * (pz*I/I0_norm)
          }
        }
        };      // body
        if (voigt) body(std::true_type{}); else body(std::false_type{});
      } break;
      default: break;
    }
    if (wrote) { fresh = false; pending = 0.0; }
    if (cum_out != nullptr) {
      double *dst = cum_out + (size_t)ip * cum_stride;
      for (int i = tid; i < q_count; i += EVAL_THREADS) dst[i] = fresh ? pending : I_sm[i];   // each thread owns i = tid (mod EVAL_THREADS)
    }
  }
  if (fresh)                           // no population wrote (noise only, or every table empty)
    for (int i = tid; i < q_count; i += EVAL_THREADS) I_sm[i] = pending;
  __syncthreads();
}

}  // namespace xrsd
