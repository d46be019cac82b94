// fit_kernels.cu -- batched fit objective, finite-difference normal equations and the Levenberg-Marquardt loop
// kernels: one CTA per (system, q-span) for the objective and the Jacobian variants, one CTA per system for the trial
// point, one thread per system for the LM algebra.
//
//   residual_kernel          <- System.evaluate_residual, system/__init__.py:134-187, and compute_chi2,
//                               tools/__init__.py:104-125 (scalar weighted chi^2, including the reference's
//                               `wts *= dI**2` convention); I(q) is produced in shared memory and consumed in place
//   jacobian_variants_kernel <- what one lmfit objective sweep would need (system/__init__.py:189-195 evaluates one
//   jacobian_reduce_kernel      parameter set per call; here every perturbed pattern of every system in one launch);
//                               the variant patterns go through an HBM workspace (resident for the whole fit)
//   objective_prepare_kernel, lm_propose_kernel, lm_trial_update_kernel, lm_accept_kernel
//                            <- replace the lmfit Nelder-Mead driver of system/__init__.py:262-266 (no oracle in the
//                               reference; parity is defined on outcomes, and measured against batched Nelder-Mead
//                               on the same systems in bench.py)
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include "launch_util.cuh"
#include "system_eval.cuh"

namespace xrsd {

struct FreeSet {
  int n_free;
  int idx[XRSD_MAX_FREE];       // parameter slot of free parameter j
  int n_var;                    // evaluated variants: 1 (base) + number of non-linear free parameters
  int var_of[XRSD_MAX_FREE];    // variant index (>= 1) of a non-linear parameter, 0 for a linear one
  int pop_of[XRSD_MAX_FREE];    // linear parameter (an I0): index of its population, else -1
  int free_of_var[XRSD_MAX_FREE + 1];   // inverse of var_of
  int vpop_of[XRSD_MAX_FREE];   // non-linear parameter used by ONE population (incl. what is tied to it): that population's
                                // index -- its variant evaluates the population alone -- else -1 (whole system)
  int sf_only[XRSD_MAX_FREE];   // 1: the parameter is r_hard or v_fraction of a disordered population and drives nothing
                                // else -- it acts through the hard-sphere factor S(q) alone, so with a resident base its
                                // variant is I_pop * S'(q) / S(q) and needs no size-distribution loop
  // the reduction forms every difference quotient as  scale_j * ((slab[ra] - slab[rb]) + slab[rc]):  slab rows of
  // the workspace, -1 = a row of zeros; scale_j = 1, or h / I0 for a linear parameter (lin_j set)
  int ra[XRSD_MAX_FREE], rb[XRSD_MAX_FREE], rc[XRSD_MAX_FREE], lin_j[XRSD_MAX_FREE];
  int n_tie;                    // constraint terms, applied in order (apply_ties below).  tie_add: 0 "dst = scale * src + offset",
                                // 1 "dst += scale * src + offset" (the further terms of a linear combination), 2 "dst = scale *
                                // (top of the expression stack) + offset", >= XRSD_TIE_OP0: a word of an expression program
                                // (non-linear constraint_expr: push parameter tie_src / push constant tie_scale / operator)
  int tie_dst[XRSD_MAX_TIES], tie_src[XRSD_MAX_TIES], tie_add[XRSD_MAX_TIES];
  int tie_to[XRSD_MAX_TIES];    // the slot a term that READS parameter tie_src feeds (-1: the term reads no parameter)
  double tie_scale[XRSD_MAX_TIES], tie_off[XRSD_MAX_TIES];
};

// does population (or noise model) `p` read parameter slot `slot`?
static bool pop_uses_slot(const xrsd_pop_t &p, int slot) {
  if (p.kind == XRSD_POP_NONE) return false;
  const int32_t scalars[] = {p.p_I0, p.p_r, p.p_sigma, p.p_rg, p.p_D, p.p_r_hard, p.p_v_fraction, p.p_flat_fraction,
                             p.p_hwhm, p.p_hwhm_g, p.p_hwhm_l};
  for (int32_t s : scalars)
    if (s == slot) return true;
  for (int i = 0; i < 6; ++i)
    if (p.p_lat[i] == slot) return true;
  for (int a = 0; a < XRSD_MAX_SPECIES; ++a)
    for (int j = 0; j < 3; ++j)
      if (p.p_uvw[a][j] == slot) return true;
  return false;
}

// An I0 multiplies its whole population (scattering/__init__.py:48,64,66; system/noise.py:56-62), so
// dI/dI0 = I_pop / I0 needs no extra evaluation: the base variant records the running total after
// each population and the reducer forms f(I + h I_pop / I0) from it.
static FreeSet make_free_set(const xrsd_layout_t *L, const int32_t *free_idx, int n_free, const int32_t *ties = nullptr,
                             const double *tie_coef = nullptr, int n_tie = 0) {
  FreeSet F;
  memset(&F, 0, sizeof(F));
  F.n_free = n_free;
  F.n_var = 1;
  F.n_tie = (ties && n_tie > 0) ? std::min(n_tie, (int)XRSD_MAX_TIES) : 0;
  for (int t = 0; t < F.n_tie; ++t) {
    const int d = ties[2 * t];
    F.tie_src[t] = ties[2 * t + 1];
    F.tie_scale[t] = tie_coef ? tie_coef[2 * t] : 1.0;
    F.tie_off[t] = tie_coef ? tie_coef[2 * t + 1] : 0.0;
    F.tie_to[t] = -1;
    if (d <= -XRSD_TIE_WORD) {                           // a word of an expression program: -(XRSD_TIE_WORD + op)
      F.tie_add[t] = XRSD_TIE_OP0 + (-d - XRSD_TIE_WORD);
      F.tie_dst[t] = -1;
    } else if (d < 0) {                                  // dst given as -(slot + 1): accumulate
      F.tie_add[t] = 1;
      F.tie_dst[t] = -d - 1;
      F.tie_to[t] = F.tie_dst[t];
    } else {                                             // src >= 0: linear term; src == -1: the value on the stack
      F.tie_add[t] = F.tie_src[t] < 0 ? 2 : 0;
      F.tie_dst[t] = d;
      if (F.tie_src[t] >= 0) F.tie_to[t] = d;
    }
  }
  // the parameters an expression program pushes feed the slot its closing "store" term sets
  for (int t = F.n_tie - 1, to = -1; t >= 0; --t) {
    if (F.tie_add[t] == 2) to = F.tie_dst[t];
    else if (F.tie_add[t] < XRSD_TIE_OP0) to = -1;
    else if (F.tie_add[t] == XRSD_TIE_OP0 + XRSD_TIE_PUSHP) F.tie_to[t] = to;
  }
  for (int j = 0; j < XRSD_MAX_FREE; ++j) { F.pop_of[j] = -1; F.vpop_of[j] = -1; }
  for (int j = 0; j < n_free; ++j) {
    // free_idx[j] = -(slot + 1) asks for a finite-difference variant even if the parameter is an I0 (the caller does
    // that when a system starts at I0 == 0, where I_pop / I0 carries no derivative)
    const bool forced = free_idx[j] < 0;
    const int slot = forced ? -free_idx[j] - 1 : free_idx[j];
    F.idx[j] = slot;
    int pop = -1;
    if (L)
      for (int k = 0; k < L->n_pops; ++k)
        if (L->pops[k].kind != XRSD_POP_NONE && L->pops[k].p_I0 == slot) pop = k;
    // a slot shared by two populations (not produced by the lowering) would not be linear in one term
    int uses = 0;
    if (L)
      for (int k = 0; k < L->n_pops; ++k) uses += (L->pops[k].p_I0 == slot);
    bool tied = false;      // a parameter that drives another one is not linear in a single term
    for (int t = 0; t < F.n_tie; ++t) tied = tied || (F.tie_to[t] >= 0 && F.tie_src[t] == slot);
    if (pop >= 0 && uses == 1 && !tied && !forced) {
      F.pop_of[j] = pop;
      F.var_of[j] = 0;
    } else {
      F.var_of[j] = F.n_var;
      F.free_of_var[F.n_var] = j;
      ++F.n_var;
      // populations that see this parameter or anything tied to it
      if (L) {
        int users = 0, which = -1;
        for (int k = 0; k < L->n_pops; ++k) {
          bool u = pop_uses_slot(L->pops[k], slot);
          for (int t = 0; t < F.n_tie; ++t) u = u || (F.tie_to[t] >= 0 && F.tie_src[t] == slot && pop_uses_slot(L->pops[k], F.tie_to[t]));
          if (u) { ++users; which = k; }
        }
        if (users == 1) F.vpop_of[j] = which;
        if (users == 1 && !tied && which > 0) {
          const xrsd_pop_t &pp = L->pops[which];
          if (pp.kind == XRSD_POP_NONCRYSTALLINE && pp.disordered && (slot == pp.p_r_hard || slot == pp.p_v_fraction) &&
              pp.p_r_hard != pp.p_v_fraction && slot != pp.p_I0 && slot != pp.p_r && slot != pp.p_sigma && slot != pp.p_rg && slot != pp.p_D)
            F.sf_only[j] = 1;
        }
      }
    }
  }
  for (int j = 0; j < XRSD_MAX_FREE; ++j) {
    F.ra[j] = F.rb[j] = F.rc[j] = -1;
    F.lin_j[j] = 0;
    if (j >= n_free) continue;
    const int nv = F.n_var;
    if (F.var_of[j] > 0) {                  // variant slab minus the base's term: whole pattern, or one population's
      F.ra[j] = F.var_of[j];
      const int kv = F.vpop_of[j];
      if (kv < 0) { F.rb[j] = 0; }
      else { F.rb[j] = nv + kv; F.rc[j] = kv > 0 ? nv + kv - 1 : -1; }
    } else {                                // linear: (h / I0) * the population's own term, from the running totals
      const int kp = F.pop_of[j];
      F.lin_j[j] = 1;
      F.ra[j] = nv + kp;
      F.rb[j] = kp > 0 ? nv + kp - 1 : -1;
    }
  }
  return F;
}

// The constraint terms of a system, in order, on one parameter row (System.fit hands `constraint_expr` strings to
// lmfit, system/__init__.py:197-209, which evaluates them with asteval after every change of the free parameters).
// Linear expressions are one fused multiply-add per term; anything else was compiled by the host
// (xrsdkit_b200/lowering.py: compile_constraint) into a postfix program over a small value stack.
__device__ __forceinline__ void apply_ties(const FreeSet &F, double *p) {
  double st[XRSD_TIE_STACK];
  int sp = 0;
  for (int t = 0; t < F.n_tie; ++t) {
    const int op = F.tie_add[t];
    if (op == 0) { p[F.tie_dst[t]] = fma(F.tie_scale[t], p[F.tie_src[t]], F.tie_off[t]); continue; }
    if (op == 1) { p[F.tie_dst[t]] = fma(F.tie_scale[t], p[F.tie_src[t]], F.tie_off[t] + p[F.tie_dst[t]]); continue; }
    if (op == 2) { p[F.tie_dst[t]] = fma(F.tie_scale[t], sp > 0 ? st[sp - 1] : 0.0, F.tie_off[t]); sp = 0; continue; }
    const int w = op - XRSD_TIE_OP0;
    if (w == XRSD_TIE_PUSHC) { if (sp < XRSD_TIE_STACK) st[sp++] = F.tie_scale[t]; continue; }
    if (w == XRSD_TIE_PUSHP) { if (sp < XRSD_TIE_STACK) st[sp++] = p[F.tie_src[t]]; continue; }
    if (w >= XRSD_TIE_ADD && w <= XRSD_TIE_MAX) {         // binary: a (op) b with b on top
      if (sp < 2) continue;
      const double b = st[--sp], a = st[sp - 1];
      double r;
      switch (w) {
        case XRSD_TIE_ADD: r = a + b; break;
        case XRSD_TIE_SUB: r = a - b; break;
        case XRSD_TIE_MUL: r = a * b; break;
        case XRSD_TIE_DIV: r = a / b; break;
        case XRSD_TIE_POW: r = pow(a, b); break;
        case XRSD_TIE_MIN: r = fmin(a, b); break;
        default: r = fmax(a, b); break;
      }
      st[sp - 1] = r;
      continue;
    }
    if (sp < 1) continue;
    const double a = st[sp - 1];
    double r = a;
    switch (w) {
      case XRSD_TIE_NEG: r = -a; break;
      case XRSD_TIE_SQRT: r = sqrt(a); break;
      case XRSD_TIE_EXP: r = exp(a); break;
      case XRSD_TIE_LOG: r = log(a); break;
      case XRSD_TIE_LOG10: r = log10(a); break;
      case XRSD_TIE_SIN: r = sin(a); break;
      case XRSD_TIE_COS: r = cos(a); break;
      case XRSD_TIE_TAN: r = tan(a); break;
      case XRSD_TIE_ABS: r = fabs(a); break;
      case XRSD_TIE_ASIN: r = asin(a); break;
      case XRSD_TIE_ACOS: r = acos(a); break;
      case XRSD_TIE_ATAN: r = atan(a); break;
      default: break;
    }
    st[sp - 1] = r;
  }
}

// weight and mask of one q-point (system/__init__.py:164-172)
__device__ __forceinline__ bool fit_weight(const xrsd_objective_t &O, double qv, double Iobs, const double *dI, size_t k,
                                           double &w) {
  if (O.prepared) {            // xrsd_objective_prepare did this once per fit: weight stream, -1 on masked points
    w = dI[k];
    return !(w < 0.0);
  }
  const bool fit = (Iobs > 0.0) && (qv >= O.q_lo) && (qv <= O.q_hi);
  w = 1.0;
  if (O.error_weighted) {
    double d;
    if (O.has_dI) d = dI[k];
    else d = fit ? sqrt(Iobs) : NAN;
    w = w * (d * d);
  }
  return fit;
}

// f(I_obs) of the objective: the logarithm (or the value) of the observed intensity, or the prepared stream
__device__ __forceinline__ double fit_target(const xrsd_objective_t &O, double Iobs) {
  return (O.logI_weighted && !O.prepared) ? log(Iobs) : Iobs;
}

// Grid = batch x q-spans (same span policy as the intensity kernel).  Each CTA evaluates its span into
// shared memory, reduces its part of sum(w d^2) and sum(w), adds both to a per-system accumulator with
// FP64 atomics; the span that takes the last ticket of a system writes chi2 = num/den and re-zeroes the
// accumulator.  I(q) never goes to HBM.
template <bool RM, bool HX>
__global__ void __launch_bounds__(EVAL_THREADS, 6)
residual_kernel(const __grid_constant__ xrsd_layout_t L, const xrsd_objective_t O, const double *__restrict__ q, int n_q,
                const double *__restrict__ params, int ldp, int batch, const xrsd_tables_t T,
                const double *__restrict__ Iobs, const double *__restrict__ dI, long long ld_obs,
                double *__restrict__ chi2, int scratch_doubles, const uint8_t *__restrict__ active,
                double *__restrict__ acc_g, unsigned *__restrict__ tickets, int q_per_cta, int n_tiles) {
  extern __shared__ __align__(16) double smem_d[];
  __shared__ EvalSharedT<HX> S;
  __shared__ int is_last;
  const long long bid = blockIdx.x;
  const int b = (int)(bid / n_tiles);
  const int tile = (int)(bid - (long long)b * n_tiles);
  if (b >= batch) return;
  if (active != nullptr && !active[b]) return;
  const int q_begin = tile * q_per_cta;
  const int q_count = min(q_per_cta, n_q - q_begin);
  double *scratch = smem_d;                   // at the base: a compile-time address inside the profile loops
  double *I_sm = smem_d + scratch_doubles;
  accumulate_system<2, RM, HX>(L, params + (size_t)b * ldp, T, b, q, n_q, q_begin, q_count, I_sm, scratch, scratch_doubles, &S);
  const double *obs = Iobs + (size_t)b * ld_obs + q_begin;
  const double *err = dI ? dI + (size_t)b * ld_obs + q_begin : nullptr;
  double num = 0.0, den = 0.0;
  for (int i = threadIdx.x; i < q_count; i += EVAL_THREADS) {
    const double Io = obs[i], Ic = I_sm[i];
    double w;
    bool fit = fit_weight(O, q[q_begin + i], Io, err, i, w);
    double d;
    if (O.logI_weighted) {
      fit = fit && (Ic > 0.0);
      d = fit ? (log(Ic) - fit_target(O, Io)) : 0.0;
    } else {
      d = Ic - Io;
    }
    if (fit) { num += (d * d) * w; den += w; }
  }
  num = block_sum(num, &S);
  den = block_sum(den, &S);
  // the spans of a system meet in partial sums kept per span and added up IN SPAN ORDER by whichever span arrives
  // last: the result does not depend on the order of arrival (bit-identical from run to run)
  if (threadIdx.x == 0) {
    double *part = acc_g + 2 * ((size_t)b * n_tiles + tile);
    part[0] = num;
    part[1] = den;
    __threadfence();
    is_last = (atomicAdd(&tickets[b], 1u) == (unsigned)(n_tiles - 1));
    if (is_last) {
      __threadfence();
      double n_ = 0.0, d_ = 0.0;
      for (int t = 0; t < n_tiles; ++t) {
        n_ += __ldcg(&acc_g[2 * ((size_t)b * n_tiles + t)]);
        d_ += __ldcg(&acc_g[2 * ((size_t)b * n_tiles + t) + 1]);
      }
      chi2[b] = (d_ == 0.0 && n_ == 0.0) ? 0.0 : n_ / d_;      // no point in the fit mask: the reference's compute_chi2 sums nothing -> 0.0
      tickets[b] = 0;
    }
  }
}

// Normal equations by forward differences: two launches.
//
// (1) jacobian_variants_kernel, grid = batch x evaluated variants x q-spans: every CTA evaluates one
//     parameter variant of one system over one q-span exactly like the intensity kernel does and parks
//     the raw I(q) in an HBM scratch [batch][n_var + n_pops][n_q] (the base variant also writes its
//     running total after each population).  At cfg4 sizes that is ~0.6 MB per system, written and
//     read once: ~0.2 us of HBM time against ~6 us of FP64 work.
// (2) jacobian_reduce_kernel, grid = batch x q-chunks: forms the weighted residual r and the difference
//     quotients D_j per q-point (log mode: D_j = log1p((I_j - I_0)/I_0), no cancellation between two
//     logarithms), accumulates JTJ / JTr / chi2 in registers, folds them per CTA and adds them to a
//     per-system accumulator with FP64 atomics; the chunk that takes the last ticket normalises and
//     writes the outputs.  (A first version reduced inside the variant kernel by the last-finishing
//     CTA of each system: that serial tail cost 25 % of the whole Jacobian.)
__host__ __device__ __forceinline__ int np_ceil2(int n) { return (n + 1) & ~1; }

template <bool RM, bool HX>
__global__ void __launch_bounds__(EVAL_THREADS, HX ? XRSD_HX_MIN_CTAS : 6)   // crystalline layouts, 10^4 cfg4 systems, same box: 8 CTAs x 64 registers 15.5 ms, 10 x 48 (spills) 16.1 (round 1: 11.8 / 11.5 per 4069 systems, 12 x 40 14.4, 6 x 80 7 % behind 8)
jacobian_variants_kernel(const __grid_constant__ xrsd_layout_t L, const double *__restrict__ q, int n_q,
                         const double *__restrict__ params, int ldp, int batch, const xrsd_tables_t T, const FreeSet F,
                         double rel_step, double *__restrict__ Fvar, int q_per_cta, int n_tiles, int pattern_doubles,
                         int scratch_doubles, const uint8_t *__restrict__ active, int skip_base) {
  // pattern_doubles == 0: direct mode, the variant's slab in HBM is the accumulation buffer (see intensity_kernel.cu)
  extern __shared__ __align__(16) double smem_d[];
  __shared__ EvalSharedT<HX> S;
  const int tid = threadIdx.x;
  const int nv = F.n_var, n_slab = nv + L.n_pops;
  const long long bid = blockIdx.x;
  const int per_sys = nv * n_tiles;
  const int b = (int)(bid / per_sys);
  const int rem = (int)(bid - (long long)b * per_sys);
  const int v = rem / n_tiles, tile = rem - v * n_tiles;
  if (b >= batch) return;
  if (active != nullptr && !active[b]) return;
  if (skip_base && v == 0) return;            // base pattern and running totals are already in the workspace
  const int np = L.n_params;
  // shared memory: shell scratch at the base (a compile-time address), the parameter row, then -- staged mode only --
  // the pattern span.  Direct mode is the crystalline instantiation (HX), see intensity_kernel.cu.
  if ((pattern_doubles == 0) != HX) return;
  double *scratch = smem_d;
  double *prm_s = smem_d + scratch_doubles;
  double *I_sm = prm_s + ((np_ceil2(L.n_params)));
  const double *prm = params + (size_t)b * ldp;
  for (int i = tid; i < np; i += EVAL_THREADS) prm_s[i] = prm[i];
  __syncthreads();
  if (v > 0 && tid == 0) {
    const int j = F.free_of_var[v];
    const double p0 = prm[F.idx[j]];
    prm_s[F.idx[j]] = p0 + ((p0 != 0.0) ? fabs(p0) * rel_step : rel_step);
    apply_ties(F, prm_s);                                                         // tied parameters move along
  }
  __syncthreads();
  const int q_begin = tile * q_per_cta;
  const int q_count = min(q_per_cta, n_q - q_begin);
  // slabs of this system: [0, nv) variants (slab 0 = base), [nv, nv+n_pops) the base variant's running
  // totals after each population
  double *slab0 = Fvar + (size_t)b * n_slab * n_q;
  double *dst = slab0 + (size_t)v * n_q + q_begin;
  if constexpr (HX) I_sm = dst;
  // a variant whose parameter belongs to one population evaluates that population alone (the reduction takes the
  // difference against the base's own term for it, from the running totals)
  const int only_pop = v > 0 ? F.vpop_of[F.free_of_var[v]] : -1;
  if constexpr (!HX) {
    // A hard-sphere parameter (r_hard, v_fraction) acts through S(q) alone: I_pop = I0 S(q) <ff^2>(q).  With a resident
    // base (the running totals of the unperturbed system are in the workspace) its variant is the base's own term
    // times S'(q) / S(q) -- two Percus-Yevick evaluations per q-point instead of the size-distribution loop.
    if (v > 0 && skip_base && F.sf_only[F.free_of_var[v]]) {
      const xrsd_pop_t &pop = L.pops[only_pop];
      const HSConst h0 = hs_constants(prm[pop.p_r_hard], prm[pop.p_v_fraction]);
      const HSConst h1 = hs_constants(prm_s[pop.p_r_hard], prm_s[pop.p_v_fraction]);
      const double *tot_hi = slab0 + (size_t)(nv + only_pop) * n_q + q_begin;
      const double *tot_lo = slab0 + (size_t)(nv + only_pop - 1) * n_q + q_begin;
      const bool same_d = h0.d == h1.d;
      const double f0 = 1.0 - h0.nc0, f1 = 1.0 - h1.nc0;
      for (int i = tid; i < q_count; i += EVAL_THREADS) {
        const double qi = q[q_begin + i];
        const bool q0 = L.q0_is_zero && (q_begin + i == 0);
        double s0, c0, s1, c1;
        sincos(qi * h0.d, &s0, &c0);
        if (same_d) { s1 = s0; c1 = c0; } else sincos(qi * h1.d, &s1, &c1);
        const double d0 = 1.0 - hard_sphere_nc(qi, h0, q0, s0, c0), d1 = 1.0 - hard_sphere_nc(qi, h1, q0, s1, c1);
        dst[i] = (tot_hi[i] - tot_lo[i]) * ((f1 * d0) / (f0 * d1));      // I_pop S'/S,  S = (1 - nc0) / (1 - nc)
      }
      return;
    }
  }
  accumulate_system<2, RM, HX>(L, prm_s, T, b, q, n_q, q_begin, q_count, I_sm, scratch, scratch_doubles, &S, prm,
                               v == 0 ? slab0 + (size_t)nv * n_q + q_begin : nullptr, n_q, only_pop);
  if constexpr (!HX)
    for (int i = tid; i < q_count; i += EVAL_THREADS) dst[i] = I_sm[i];
}

constexpr int JR_THREADS = 128;
constexpr int JR_STAGES = 4;      // q-points in flight per thread (ring of cp.async groups)

// 8-byte asynchronous global -> shared copy (LDGSTS) and its group primitives
__device__ __forceinline__ void cp_async8(double *dst_smem, const double *src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// RB rows of JTJ are accumulated per pass over the chunk (RB * NF accumulators live in registers)
// n_free <= 8: capped at 128 registers (4 CTAs/SM; the compiler otherwise takes 232 for 2 -- measured +8 % on cfg4
// with 140 bytes of spills); the 16-parameter instantiation spills 1.5 KB under that cap and keeps 2 CTAs/SM.
template <int NF, int RB>
#ifndef XRSD_JR_MIN_CTAS
#define XRSD_JR_MIN_CTAS 4
#endif
__global__ void __launch_bounds__(JR_THREADS, NF <= 8 ? XRSD_JR_MIN_CTAS : 2)
jacobian_reduce_kernel(int n_pops, const xrsd_objective_t O, const double *__restrict__ q, int n_q,
                       const double *__restrict__ params, int ldp, int batch, const double *__restrict__ Iobs,
                       const double *__restrict__ dI, long long ld_obs, const FreeSet F, double rel_step,
                       const double *__restrict__ Fvar, double *__restrict__ acc_g, unsigned *__restrict__ tickets,
                       double *__restrict__ JTJ, double *__restrict__ JTr, double *__restrict__ chi2, int chunk_q,
                       int n_chunks, const uint8_t *__restrict__ active) {
  constexpr int NACC = NF * NF + NF + 2;          // full square (upper part used), JTr, chi2 numerator, sum w
  // ring[stage][row][tid]: every thread streams its own q-points, JR_STAGES deep, through private slots --
  // rows 0..n_slab-1 are the variant / running-total slabs, then I_obs, then dI.  The copies are asynchronous
  // (LDGSTS), so a thread has JR_STAGES * n_rows * 8 B in flight instead of the few loads the compiler can hoist
  // around 150 live accumulators (the first version was latency-bound at 0.45 TB/s).
  extern __shared__ __align__(16) double ring[];
  __shared__ double fold[JR_THREADS / 32][RB * NF + NF + 2];
  __shared__ int4 dsc[NF];          // ring offsets (doubles) of the three slab rows of every difference quotient
  __shared__ int is_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x / n_chunks, chunk = blockIdx.x - b * n_chunks;
  if (b >= batch) return;
  if (active != nullptr && !active[b]) return;       // outputs of a converged system keep their last values
  const int nf = F.n_free, nv = F.n_var, n_slab = nv + n_pops;
  const double *prm = params + (size_t)b * ldp;
  const double *obs = Iobs + (size_t)b * ld_obs;
  const double *err = dI ? dI + (size_t)b * ld_obs : nullptr;
  const double *Fb = Fvar + (size_t)b * n_slab * n_q;
  double step[NF], scale[NF];
#pragma unroll
  for (int j = 0; j < NF; ++j) {
    const double p0 = (j < nf) ? prm[F.idx[j]] : 1.0;
    step[j] = (p0 != 0.0) ? fabs(p0) * rel_step : rel_step;
    // h / I0 for a linear parameter (0 at I0 == 0: no derivative there, the caller forces a variant instead), else 1
    scale[j] = (j >= nf) ? 0.0 : (F.lin_j[j] ? ((p0 != 0.0) ? step[j] / p0 : 0.0) : 1.0);
  }
  const int q_lo = chunk * chunk_q, q_hi = min(n_q, q_lo + chunk_q);
  double *accb = acc_g + ((size_t)b * n_chunks + chunk) * NACC;      // this chunk's partial sums
  const int n_rows = n_slab + 1 + (err ? 1 : 0);
  const int z_row = n_rows;                          // one more row per stage that stays zero: the "absent term" of a quotient
  for (int i = tid; i < JR_STAGES * JR_THREADS; i += JR_THREADS)
    ring[((size_t)(i / JR_THREADS) * (n_rows + 1) + z_row) * JR_THREADS + (i % JR_THREADS)] = 0.0;
  // the slab rows of each quotient as ring offsets, once per CTA: inside the loop the FreeSet lookups (constant bank
  // load, test for the absent row, multiply -- per parameter and q-point) were 13 % of the kernel's instructions
  if (tid < NF)
    dsc[tid] = make_int4((F.ra[tid] < 0 ? z_row : F.ra[tid]) * JR_THREADS, (F.rb[tid] < 0 ? z_row : F.rb[tid]) * JR_THREADS,
                         (F.rc[tid] < 0 ? z_row : F.rc[tid]) * JR_THREADS, 0);
  __syncthreads();
  const int n_it = (q_hi - q_lo + JR_THREADS - 1) / JR_THREADS;
  auto issue = [&](int it) {             // stage the q-point of iteration `it` (if any); always closes a group
    const int i = q_lo + it * JR_THREADS + tid;
    if (it < n_it && i < q_hi) {
      double *slot = ring + (size_t)(it % JR_STAGES) * (n_rows + 1) * JR_THREADS + tid;
      for (int r = 0; r < n_slab; ++r) cp_async8(slot + r * JR_THREADS, Fb + (size_t)r * n_q + i);
      cp_async8(slot + n_slab * JR_THREADS, obs + i);
      if (err) cp_async8(slot + (n_slab + 1) * JR_THREADS, err + i);
    }
    cp_async_commit();
  };
  for (int row0 = 0; row0 < nf; row0 += RB) {
    double aJ[RB][NF], ar[RB], c2 = 0.0, sw = 0.0;
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      ar[r] = 0.0;
#pragma unroll
      for (int k = 0; k < NF; ++k) aJ[r][k] = 0.0;
    }
#pragma unroll
    for (int st = 0; st < JR_STAGES; ++st) issue(st);
    for (int it = 0; it < n_it; ++it) {
      cp_async_wait<JR_STAGES - 1>();                 // the group of iteration `it` has landed
      const int i = q_lo + it * JR_THREADS + tid;
      const double *Fs = ring + (size_t)(it % JR_STAGES) * (n_rows + 1) * JR_THREADS + tid;   // row r at Fs[r * JR_THREADS]
      bool fit = i < q_hi;
      double w = 0.0, Io = 0.0, I0v = 0.0;
      if (fit) {
        Io = Fs[n_slab * JR_THREADS];
        fit = fit_weight(O, q[i], Io, err ? Fs + (n_slab + 1) * JR_THREADS : nullptr, 0, w);
        I0v = Fs[0];
        if (O.logI_weighted) fit = fit && (I0v > 0.0);
      }
      if (!fit) { issue(it + JR_STAGES); continue; }
      const double f0 = O.logI_weighted ? log(I0v) : I0v;
      const double fo = fit_target(O, Io);
      const double res = f0 - fo;
      const double inv = O.logI_weighted ? 1.0 / I0v : 1.0;
      // difference quotients (times the step): scale_j * ((slab[ra] - slab[rb]) + slab[rc]), see FreeSet; in log mode
      // log1p of the relative change -- no cancellation between two logarithms, and for the small changes a
      // finite-difference step makes (|x| < 0.01) three terms of its series instead of a logarithm per parameter
      double D[NF];
#pragma unroll
      for (int j = 0; j < NF; ++j) {
        const int4 o = dsc[j];
        double d = scale[j] * ((Fs[o.x] - Fs[o.y]) + Fs[o.z]) * inv;
        if (O.logI_weighted) {
          if (fabs(d) < 1.0e-2) d = d * fma(d, fma(d, 1.0 / 3.0, -0.5), 1.0);      // |error| < x^4 / 4
          else d = (d > -1.0) ? log1p(d) : 0.0;          // a variant intensity <= 0 carries no information (NaN: below)
        }
        D[j] = (d == d) ? d : 0.0;
      }
      issue(it + JR_STAGES);                          // the slot is free again: refill it
      if (row0 == 0) { c2 = fma(w, res * res, c2); sw += w; }
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const int j = row0 + r;
        if (j < nf) {
          double dj = 0.0;
#pragma unroll
          for (int k = 0; k < NF; ++k) if (k == j) dj = D[k];
          const double wdj = w * dj;
          ar[r] = fma(wdj, res, ar[r]);
#pragma unroll
          for (int k = (RB == NF ? r : 0); k < NF; ++k) aJ[r][k] = fma(wdj, D[k], aJ[r][k]);   // single pass: upper triangle only
        }
      }
    }
    cp_async_wait<0>();
    // fold over the CTA: warp shuffles, then one atomic per entry
#pragma unroll
    for (int r = 0; r < RB; ++r) {
#pragma unroll
      for (int k = 0; k < NF; ++k) {
        double v = aJ[r][k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) fold[warp][r * NF + k] = v;
      }
      double v = ar[r];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) fold[warp][RB * NF + r] = v;
    }
    for (int o = 16; o > 0; o >>= 1) { c2 += __shfl_xor_sync(0xffffffffu, c2, o); sw += __shfl_xor_sync(0xffffffffu, sw, o); }
    if (lane == 0) { fold[warp][RB * NF + NF] = c2; fold[warp][RB * NF + NF + 1] = sw; }
    __syncthreads();
    for (int e = tid; e < RB * NF + RB + 2; e += JR_THREADS) {
      double v = 0.0;
      int src = e, dst = -1;
      if (e < RB * NF) {
        const int r = e / NF, k = e - r * NF, j = row0 + r;
        if (j < nf && k >= j && k < nf) dst = j * NF + k;
      } else if (e < RB * NF + RB) {
        const int j = row0 + (e - RB * NF);
        if (j < nf) dst = NF * NF + j;
      } else if (row0 == 0) {
        src = RB * NF + NF + (e - RB * NF - RB);
        dst = NF * NF + NF + (e - RB * NF - RB);
      }
      if (dst >= 0) {
        for (int wv = 0; wv < JR_THREADS / 32; ++wv) v += fold[wv][src];
        accb[dst] = v;
      }
    }
    __syncthreads();
  }
  // ---- the chunk that arrives last adds the partial sums of all chunks IN CHUNK ORDER (the result does not depend on
  //      the order of arrival: bit-identical from run to run), normalises and writes the outputs
  __threadfence();
  __syncthreads();
  if (tid == 0) is_last = (atomicAdd(&tickets[b], 1u) == (unsigned)(n_chunks - 1));
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const double *acc0 = acc_g + (size_t)b * n_chunks * NACC;
  auto total = [&](int e) {
    double v = 0.0;
    for (int c = 0; c < n_chunks; ++c) v += __ldcg(&acc0[(size_t)c * NACC + e]);
    return v;
  };
  double swt = total(NF * NF + NF + 1);
  const double c2t = total(NF * NF + NF);
  if (swt == 0.0 && c2t == 0.0) swt = 1.0;      // empty fit mask: chi2 = 0, no gradient (see residual_kernel)
  for (int e = tid; e < nf * nf; e += JR_THREADS) {
    const int j = e / nf, k = e - j * nf;
    const int jj = min(j, k), kk = max(j, k);
    double sj = 1.0, sk = 1.0;
#pragma unroll
    for (int m = 0; m < NF; ++m) { if (m == jj) sj = step[m]; if (m == kk) sk = step[m]; }
    JTJ[((size_t)b * nf + j) * nf + k] = total(jj * NF + kk) / (sj * sk) / swt;
  }
  if (tid < nf) {
    double sj = 1.0;
#pragma unroll
    for (int m = 0; m < NF; ++m) if (m == tid) sj = step[m];
    JTr[(size_t)b * nf + tid] = total(NF * NF + tid) / sj / swt;
  }
  if (tid == 0) {
    chi2[b] = c2t / swt;
    tickets[b] = 0;
  }
}

// ---------------------------------------------------------------------------- reduction on the FP64 tensor cores
// JTJ = sum_q (w D)(D)^T is a rank-k update of an NF x NF matrix with k = n_q: GEMM-shaped work.  Accumulating it in
// registers costs 46 accumulators per thread (n_free <= 8), which under any occupancy worth having left the compiler
// re-deriving loop-invariant addresses in every iteration (a quarter of the instructions of the register version,
// profiles/sass_by_line.py on the cfg4 capture, FP64 pipe 22 %).  Here every thread still forms the residual and the
// difference quotients of ITS q-point, but hands them to the warp through a small shared-memory tile, and the warp
// accumulates with mma.sync.m8n8k4.f64 (DMMA): A = (w D)^T as 8 parameters x 4 q-points, B = D as 4 q-points x 8
// parameters, C = the 8 x 8 block of JTJ, two accumulator registers per thread; JTr is the first column of a second
// product with B = (res, 0, ..., 0).  Tile layout V[row][(q % 4) * 8 + q / 4] with a row pitch of 34 doubles: row j < NF
// holds D_j, row NF the residual, row NF + 1 the weight; a thread's eight operands of one row are contiguous (4 x
// LDS.128, conflict-free), and so are the 32 stores of one row.  The sums of the four warps and of the q-chunks of a
// system are added in a fixed order: bit-identical from run to run, as the register version.
#ifndef XRSD_JM_STAGES
#define XRSD_JM_STAGES 2
#endif
constexpr int JM_STAGES = XRSD_JM_STAGES;
constexpr int JM_PITCH = 34;

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NF>
__global__ void __launch_bounds__(JR_THREADS, NF <= 8 ? 5 : 3)
jacobian_reduce_mma_kernel(int n_pops, const xrsd_objective_t O, const double *__restrict__ q, int n_q,
                           const double *__restrict__ params, int ldp, int batch, const double *__restrict__ Iobs,
                           const double *__restrict__ dI, long long ld_obs, const FreeSet F, double rel_step,
                           const double *__restrict__ Fvar, double *__restrict__ acc_g, unsigned *__restrict__ tickets,
                           double *__restrict__ JTJ, double *__restrict__ JTr, double *__restrict__ chi2, int chunk_q,
                           int n_chunks, const uint8_t *__restrict__ active) {
  constexpr int NACC = NF * NF + NF + 2;          // same partial-sum layout as jacobian_reduce_kernel
  constexpr int NB = NF / 8;                      // 8 x 8 blocks per side
  constexpr int NWARP = JR_THREADS / 32;
  extern __shared__ __align__(16) double ring[];  // [stage][row][thread], then the warps' tiles
  __shared__ double fold[NWARP][NACC];
  __shared__ int4 dsc[NF];
  __shared__ int is_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x / n_chunks, chunk = blockIdx.x - b * n_chunks;
  if (b >= batch) return;
  if (active != nullptr && !active[b]) return;
  const int nf = F.n_free, nv = F.n_var, n_slab = nv + n_pops;
  const double *prm = params + (size_t)b * ldp;
  const double *obs = Iobs + (size_t)b * ld_obs;
  const double *err = dI ? dI + (size_t)b * ld_obs : nullptr;
  const double *Fb = Fvar + (size_t)b * n_slab * n_q;
  const int q_lo = chunk * chunk_q, q_hi = min(n_q, q_lo + chunk_q);
  const int n_rows = n_slab + 1 + (err ? 1 : 0);
  const int z_row = n_rows;
  double *tile = ring + (size_t)JM_STAGES * (n_rows + 1) * JR_THREADS + (size_t)warp * (NF + 2) * JM_PITCH;
  for (int i = tid; i < JM_STAGES * JR_THREADS; i += JR_THREADS)
    ring[((size_t)(i / JR_THREADS) * (n_rows + 1) + z_row) * JR_THREADS + (i % JR_THREADS)] = 0.0;
  if (tid < NF)
    dsc[tid] = make_int4((F.ra[tid] < 0 ? z_row : F.ra[tid]) * JR_THREADS * 8, (F.rb[tid] < 0 ? z_row : F.rb[tid]) * JR_THREADS * 8,
                         (F.rc[tid] < 0 ? z_row : F.rc[tid]) * JR_THREADS * 8, 0);      // BYTE offsets into the thread's stage
  __syncthreads();
  double scale[NF];                  // h / I0 for a linear parameter (0 at I0 == 0), 1 for a variant slab, 0 beyond n_free
#pragma unroll
  for (int j = 0; j < NF; ++j) {
    const double p0 = (j < nf) ? prm[F.idx[j]] : 1.0;
    const double hstep = (p0 != 0.0) ? fabs(p0) * rel_step : rel_step;
    scale[j] = (j >= nf) ? 0.0 : (F.lin_j[j] ? ((p0 != 0.0) ? hstep / p0 : 0.0) : 1.0);
  }
  const int n_it = (q_hi - q_lo + JR_THREADS - 1) / JR_THREADS;
  auto issue = [&](int it) {
    const int i = q_lo + it * JR_THREADS + tid;
    if (it < n_it && i < q_hi) {
      double *slot = ring + (size_t)(it % JM_STAGES) * (n_rows + 1) * JR_THREADS + tid;
      for (int r = 0; r < n_slab; ++r) cp_async8(slot + r * JR_THREADS, Fb + (size_t)r * n_q + i);
      cp_async8(slot + n_slab * JR_THREADS, obs + i);
      if (err) cp_async8(slot + (n_slab + 1) * JR_THREADS, err + i);
    }
    cp_async_commit();
  };
  // this thread's place in the tile (store) and in the fragments (load)
  const int st_col = (lane & 3) * 8 + (lane >> 2);          // q-point `lane` of the warp's 32
  const int fm = lane >> 2, fk = lane & 3;                  // fragment row (parameter within a block) / q-point within a group of 4
  double C[NB][NB][2], R[NB][2], c2 = 0.0, sw = 0.0;
#pragma unroll
  for (int x = 0; x < NB; ++x) {
    R[x][0] = R[x][1] = 0.0;
#pragma unroll
    for (int y = 0; y < NB; ++y) C[x][y][0] = C[x][y][1] = 0.0;
  }
#pragma unroll
  for (int st = 0; st < JM_STAGES; ++st) issue(st);
  for (int it = 0; it < n_it; ++it) {
    cp_async_wait<JM_STAGES - 1>();
    const int i = q_lo + it * JR_THREADS + tid;
    const double *Fs = ring + (size_t)(it % JM_STAGES) * (n_rows + 1) * JR_THREADS + tid;
    bool fit = i < q_hi;
    double w = 0.0, Io = 0.0, I0v = 0.0;
    if (fit) {
      Io = Fs[n_slab * JR_THREADS];
      fit = fit_weight(O, q[i], Io, err ? Fs + (n_slab + 1) * JR_THREADS : nullptr, 0, w);
      I0v = Fs[0];
      if (O.logI_weighted) fit = fit && (I0v > 0.0);
    }
    double res = 0.0;
    if (fit) {
      const double f0 = O.logI_weighted ? log(I0v) : I0v;
      res = f0 - fit_target(O, Io);
      const double inv = O.logI_weighted ? 1.0 / I0v : 1.0;
      const char *Fb8 = reinterpret_cast<const char *>(Fs);
      auto at = [&](int byte_off) { return *reinterpret_cast<const double *>(Fb8 + byte_off); };
#pragma unroll
      for (int j = 0; j < NF; ++j) {
        const int4 o = dsc[j];
        double d = scale[j] * ((at(o.x) - at(o.y)) + at(o.z)) * inv;
        if (O.logI_weighted) {
          // log1p of the relative change: three terms of its series for the small changes a finite-difference step
          // makes (|error| < x^4 / 4), the library function otherwise -- selected without a branch around the series
          const double ser = d * fma(d, fma(d, 1.0 / 3.0, -0.5), 1.0);
          if (!(fabs(d) < 1.0e-2)) d = (d > -1.0) ? log1p(d) : 0.0;      // rare; a variant intensity <= 0 carries no information
          else d = ser;
        }
        tile[j * JM_PITCH + st_col] = (d == d) ? d : 0.0;
      }
      c2 = fma(w, res * res, c2);
      sw += w;
    } else {
      w = 0.0;
#pragma unroll
      for (int j = 0; j < NF; ++j) tile[j * JM_PITCH + st_col] = 0.0;
    }
    tile[NF * JM_PITCH + st_col] = res;
    tile[(NF + 1) * JM_PITCH + st_col] = w;
    issue(it + JM_STAGES);
    __syncwarp();
    // fragments: eight groups of four q-points
    double wq[8], rs[8];
    {
      const double2 *pw = reinterpret_cast<const double2 *>(tile + (NF + 1) * JM_PITCH + fk * 8);
      const double2 *pr = reinterpret_cast<const double2 *>(tile + NF * JM_PITCH + fk * 8);
#pragma unroll
      for (int g2 = 0; g2 < 4; ++g2) {
        const double2 a = pw[g2], r = pr[g2];
        wq[2 * g2] = a.x; wq[2 * g2 + 1] = a.y;
        rs[2 * g2] = (fm == 0) ? r.x : 0.0; rs[2 * g2 + 1] = (fm == 0) ? r.y : 0.0;
      }
    }
    double xv[NB][8];
#pragma unroll
    for (int x = 0; x < NB; ++x) {
      const double2 *px = reinterpret_cast<const double2 *>(tile + (x * 8 + fm) * JM_PITCH + fk * 8);
#pragma unroll
      for (int g2 = 0; g2 < 4; ++g2) {
        const double2 a = px[g2];
        xv[x][2 * g2] = a.x; xv[x][2 * g2 + 1] = a.y;
      }
    }
#pragma unroll
    for (int g = 0; g < 8; ++g) {
#pragma unroll
      for (int x = 0; x < NB; ++x) {
        const double a = wq[g] * xv[x][g];
#pragma unroll
        for (int y = x; y < NB; ++y) dmma884(C[x][y][0], C[x][y][1], a, xv[y][g]);
        dmma884(R[x][0], R[x][1], a, rs[g]);
      }
    }
    __syncwarp();                          // the tile is free for the next iteration's stores
  }
  cp_async_wait<0>();
  // ---- warp results -> fold[warp][NACC layout]; C[x][y] fragment: row x*8 + fm, columns y*8 + 2 fk + {0, 1}
  for (int e = lane; e < NACC; e += 32) fold[warp][e] = 0.0;
  __syncwarp();
#pragma unroll
  for (int x = 0; x < NB; ++x) {
#pragma unroll
    for (int y = x; y < NB; ++y) {
      fold[warp][(x * 8 + fm) * NF + y * 8 + 2 * fk] = C[x][y][0];
      fold[warp][(x * 8 + fm) * NF + y * 8 + 2 * fk + 1] = C[x][y][1];
    }
    if (fk == 0) fold[warp][NF * NF + x * 8 + fm] = R[x][0];      // column 0 of the second product
  }
  for (int o = 16; o > 0; o >>= 1) { c2 += __shfl_xor_sync(0xffffffffu, c2, o); sw += __shfl_xor_sync(0xffffffffu, sw, o); }
  if (lane == 0) { fold[warp][NF * NF + NF] = c2; fold[warp][NF * NF + NF + 1] = sw; }
  __syncthreads();
  double *accb = acc_g + ((size_t)b * n_chunks + chunk) * NACC;
  for (int e = tid; e < NACC; e += JR_THREADS) {
    double v = 0.0;
#pragma unroll
    for (int wv = 0; wv < NWARP; ++wv) v += fold[wv][e];
    accb[e] = v;
  }
  // ---- last chunk of the system: chunks added in order, normalised, written (as jacobian_reduce_kernel)
  __threadfence();
  __syncthreads();
  if (tid == 0) is_last = (atomicAdd(&tickets[b], 1u) == (unsigned)(n_chunks - 1));
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const double *acc0 = acc_g + (size_t)b * n_chunks * NACC;
  auto total = [&](int e) {
    double v = 0.0;
    for (int c = 0; c < n_chunks; ++c) v += __ldcg(&acc0[(size_t)c * NACC + e]);
    return v;
  };
  auto step_of = [&](int j) {
    const double p0 = prm[F.idx[j]];
    return (p0 != 0.0) ? fabs(p0) * rel_step : rel_step;
  };
  double swt = total(NF * NF + NF + 1);
  const double c2t = total(NF * NF + NF);
  if (swt == 0.0 && c2t == 0.0) swt = 1.0;
  for (int e = tid; e < nf * nf; e += JR_THREADS) {
    const int j = e / nf, k = e - j * nf;
    const int jj = min(j, k), kk = max(j, k);
    JTJ[((size_t)b * nf + j) * nf + k] = total(jj * NF + kk) / (step_of(jj) * step_of(kk)) / swt;
  }
  if (tid < nf) JTr[(size_t)b * nf + tid] = total(NF * NF + tid) / step_of(tid) / swt;
  if (tid == 0) {
    chi2[b] = c2t / swt;
    tickets[b] = 0;
  }
}

// ---------------------------------------------------------------------------- LM algebra
__global__ void lm_propose_kernel(const double *__restrict__ params, int ldp, int batch, const FreeSet F,
                                  const double *__restrict__ JTJ, const double *__restrict__ JTr,
                                  const double *__restrict__ lambda, const double *__restrict__ lo,
                                  const double *__restrict__ hi, const uint8_t *__restrict__ active,
                                  double *__restrict__ trial, double max_rel_step, int *__restrict__ iter_counter) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b == 0 && iter_counter != nullptr) iter_counter[0] += 1;      // the iteration this launch opens (read by lm_trial_update)
  if (b >= batch) return;
  const int nf = F.n_free;
  const double *p = params + (size_t)b * ldp;
  double *t = trial + (size_t)b * ldp;
  for (int k = 0; k < ldp; ++k) t[k] = p[k];
  if (active && !active[b]) return;
  double A[XRSD_MAX_FREE][XRSD_MAX_FREE], g[XRSD_MAX_FREE];
  bool frozen[XRSD_MAX_FREE];
  const double lam = lambda[b];
  for (int j = 0; j < nf; ++j) {
    g[j] = -JTr[(size_t)b * nf + j];
    // a parameter sitting on a bound whose gradient pushes outward is held (projected LM)
    const double pj = p[F.idx[j]];
    const double l = lo[(size_t)b * nf + j], h = hi[(size_t)b * nf + j];
    frozen[j] = (pj <= l && g[j] < 0.0) || (pj >= h && g[j] > 0.0);
    for (int k = 0; k < nf; ++k) A[j][k] = JTJ[((size_t)b * nf + j) * nf + k];
  }
  for (int j = 0; j < nf; ++j) {
    const double d = A[j][j];
    A[j][j] = d + lam * (d > 0.0 ? d : 1.0);
    if (frozen[j] || !(d > 0.0) || !(g[j] == g[j])) {
      for (int k = 0; k < nf; ++k) { A[j][k] = 0.0; A[k][j] = 0.0; }
      A[j][j] = 1.0;
      g[j] = 0.0;
    }
  }
  // Cholesky A = L L^T (in place, lower), then two triangular solves
  bool ok = true;
  for (int j = 0; j < nf && ok; ++j) {
    double d = A[j][j];
    for (int k = 0; k < j; ++k) d -= A[j][k] * A[j][k];
    if (!(d > 0.0)) { ok = false; break; }
    d = sqrt(d);
    A[j][j] = d;
    for (int i = j + 1; i < nf; ++i) {
      double s = A[i][j];
      for (int k = 0; k < j; ++k) s -= A[i][k] * A[j][k];
      A[i][j] = s / d;
    }
  }
  if (!ok) return;   // trial == params: the update kernel will raise lambda
  for (int i = 0; i < nf; ++i) {
    double s = g[i];
    for (int k = 0; k < i; ++k) s -= A[i][k] * g[k];
    g[i] = s / A[i][i];
  }
  for (int i = nf - 1; i >= 0; --i) {
    double s = g[i];
    for (int k = i + 1; k < nf; ++k) s -= A[k][i] * g[k];
    g[i] = s / A[i][i];
  }
  for (int j = 0; j < nf; ++j) {
    double dj = g[j];
    if (max_rel_step > 0.0) {
      // relative trust region: one step changes a parameter by at most max_rel_step of its magnitude (a damped
      // Gauss-Newton step far outside the linear range -- a peak width grown 100x -- costs a whole-pattern
      // evaluation in the slow regime of the profile sum and is rejected anyway)
      // (a parameter sitting at exactly 0 has no scale of its own: its first step is free)
      const double cap = max_rel_step * fabs(p[F.idx[j]]);
      if (cap > 0.0) dj = fmin(fmax(dj, -cap), cap);
    }
    double v = p[F.idx[j]] + dj;
    const double l = lo[(size_t)b * nf + j], h = hi[(size_t)b * nf + j];
    // a positive parameter is not put ON a zero lower bound (an I0 at exactly 0 has no derivative left): it shrinks
    // by 1000x per step instead and reaches the bound geometrically
    if (l == 0.0 && v <= 0.0 && p[F.idx[j]] > 0.0) v = 1.0e-3 * p[F.idx[j]];
    v = fmin(fmax(v, l), h);
    if (v == v) t[F.idx[j]] = v;
  }
  apply_ties(F, t);
}

__global__ void lm_update_kernel(double *__restrict__ params, int ldp, int batch, const double *__restrict__ trial,
                                 double *__restrict__ chi2, const double *__restrict__ chi2_trial,
                                 double *__restrict__ lambda, uint8_t *__restrict__ active, int *__restrict__ n_accept,
                                 double up, double down, double ftol, double lambda_max) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  if (active && !active[b]) return;
  const double c0 = chi2[b], c1 = chi2_trial[b];
  if (c1 < c0) {   // false for NaN trials
    for (int k = 0; k < ldp; ++k) params[(size_t)b * ldp + k] = trial[(size_t)b * ldp + k];
    chi2[b] = c1;
    lambda[b] = fmax(lambda[b] / down, 1.0e-12);
    if (n_accept) n_accept[b] += 1;
    if ((c0 - c1) <= ftol * c0 && active) active[b] = 0;
  } else {
    lambda[b] = lambda[b] * up;
    if (lambda[b] > lambda_max && active) active[b] = 0;
  }
}

// chi2[b] of patterns that are already in memory (I[b][ld_I]): the objective of residual_kernel without the
// evaluation.  The resident LM loop uses it for the trial point, whose pattern the variants kernel has just left
// in the second workspace.  One CTA per system; 2 (3 with dI) streams of 8 B per q-point and two logs.
constexpr int CHI2_THREADS = 256;
constexpr int COPY_THREADS = 256, COPY_CHUNK = 2048;
__global__ void __launch_bounds__(CHI2_THREADS)
pattern_chi2_kernel(const xrsd_objective_t O, const double *__restrict__ q, int n_q, const double *__restrict__ I, long long ld_I,
                    int batch, const double *__restrict__ Iobs, const double *__restrict__ dI, long long ld_obs,
                    double *__restrict__ chi2, const uint8_t *__restrict__ active) {
  __shared__ double red[2][CHI2_THREADS / 32];
  const int b = blockIdx.x;
  if (b >= batch) return;
  if (active != nullptr && !active[b]) return;
  const double *pat = I + (size_t)b * ld_I;
  const double *obs = Iobs + (size_t)b * ld_obs;
  const double *err = dI ? dI + (size_t)b * ld_obs : nullptr;
  double num = 0.0, den = 0.0;
#pragma unroll 4
  for (int i = threadIdx.x; i < n_q; i += CHI2_THREADS) {
    const double Io = obs[i], Ic = pat[i];
    double w;
    bool fit = fit_weight(O, q[i], Io, err, i, w);
    double d;
    if (O.logI_weighted) {
      fit = fit && (Ic > 0.0);
      d = fit ? (log(Ic) - fit_target(O, Io)) : 0.0;
    } else {
      d = Ic - Io;
    }
    if (fit) { num += (d * d) * w; den += w; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    num += __shfl_xor_sync(0xffffffffu, num, o);
    den += __shfl_xor_sync(0xffffffffu, den, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = num; red[1][threadIdx.x >> 5] = den; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double n_ = 0.0, d_ = 0.0;
#pragma unroll
    for (int w = 0; w < CHI2_THREADS / 32; ++w) { n_ += red[0][w]; d_ += red[1][w]; }
    chi2[b] = (d_ == 0.0 && n_ == 0.0) ? 0.0 : n_ / d_;
  }
}

// Once per fit: f(I_obs) and the weight of every point (system/__init__.py:164-172), so the kernels of the loop read two
// streams instead of evaluating a logarithm, a square root and the q-range test per point and iteration.
//   fo[r][i] = log(I_obs) (log mode, 0 where I_obs <= 0) or I_obs;   w[r][i] = the point's weight, -1 when masked out
__global__ void __launch_bounds__(256)
objective_prepare_kernel(const xrsd_objective_t O, const double *__restrict__ q, int n_q, const double *__restrict__ Iobs,
                         const double *__restrict__ dI, long long ld_obs, int rows, double *__restrict__ fo,
                         double *__restrict__ wt, long long ld_out) {
  const int r = blockIdx.x;
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < n_q; i += gridDim.y * blockDim.x) {
    const double Io = Iobs[(size_t)r * ld_obs + i];
    double w;
    const bool fit = fit_weight(O, q[i], Io, dI ? dI + (size_t)r * ld_obs : nullptr, i, w);
    fo[(size_t)r * ld_out + i] = O.logI_weighted ? (Io > 0.0 ? log(Io) : 0.0) : Io;
    wt[(size_t)r * ld_out + i] = fit ? w : -1.0;
  }
}

// The trial point of one LM iteration: chi2 of its stored pattern (as pattern_chi2_kernel; or the value already in
// chi2_trial when no pattern is given), then accept / reject for that system -- parameters, chi2, damping, the
// flags the next launches are masked with, and the reason a system left the loop:
//   reason 0 = still running, 1 = converged (relative change of chi2 <= ftol, or of every parameter <= xtol),
//          2 = damping beyond lambda_max (no acceptable step left), 3 = objective not finite at the accepted point
// A trial point whose reflection table could not be built (status != 0, e.g. the cell outgrew the candidate box of
// this iteration) counts as chi2 = inf; flags[0] records that it happened.
__global__ void __launch_bounds__(CHI2_THREADS)
lm_trial_update_kernel(const xrsd_objective_t O, const double *__restrict__ q, int n_q, const double *__restrict__ I,
                       long long ld_I, int batch, const double *__restrict__ Iobs, const double *__restrict__ dI,
                       long long ld_obs, double *__restrict__ params, int ldp, const double *__restrict__ trial,
                       double *__restrict__ chi2, double *__restrict__ chi2_trial, double *__restrict__ lambda,
                       uint8_t *__restrict__ active, uint8_t *__restrict__ accepted, uint8_t *__restrict__ needjac,
                       uint8_t *__restrict__ reason, int *__restrict__ n_accept, const xrsd_tables_t Tt, int n_xtal,
                       int *__restrict__ flags, int tag, double up, double down, double ftol, double xtol, double lambda_max) {
  __shared__ double red[2][CHI2_THREADS / 32];
  __shared__ int take;
  const int b = blockIdx.x;
  if (b >= batch) return;
  if (!active[b]) {
    if (threadIdx.x == 0) { accepted[b] = 0; needjac[b] = 0; }
    return;
  }
  if (I != nullptr) {
    const double *pat = I + (size_t)b * ld_I;
    const double *obs = Iobs + (size_t)b * ld_obs;
    const double *err = dI ? dI + (size_t)b * ld_obs : nullptr;
    double num = 0.0, den = 0.0;
#pragma unroll 4
    for (int i = threadIdx.x; i < n_q; i += CHI2_THREADS) {
      const double Io = obs[i], Ic = pat[i];
      double w;
      bool fit = fit_weight(O, O.prepared ? 0.0 : q[i], Io, err, i, w);
      double d;
      if (O.logI_weighted) {
        fit = fit && (Ic > 0.0);
        d = fit ? (log(Ic) - fit_target(O, Io)) : 0.0;
      } else {
        d = Ic - Io;
      }
      if (fit) { num += (d * d) * w; den += w; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      num += __shfl_xor_sync(0xffffffffu, num, o);
      den += __shfl_xor_sync(0xffffffffu, den, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = num; red[1][threadIdx.x >> 5] = den; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double c1;
    if (I != nullptr) {
      double n_ = 0.0, d_ = 0.0;
#pragma unroll
      for (int w = 0; w < CHI2_THREADS / 32; ++w) { n_ += red[0][w]; d_ += red[1][w]; }
      c1 = (d_ > 0.0) ? n_ / d_ : INFINITY;       // a trial point with nothing left in the fit mask is not an improvement
    } else {
      c1 = chi2_trial[b];
    }
    if (n_xtal > 0 && Tt.status != nullptr) {
      int st = 0;
      for (int s = 0; s < n_xtal; ++s) {
        const int t = b * n_xtal + s;
        st |= Tt.status[Tt.table_of != nullptr ? Tt.table_of[t] : Tt.own_base + t];
      }
      if (st != 0) {
        c1 = INFINITY;
        if (flags != nullptr) atomicOr(&flags[0], st);
      }
    }
    chi2_trial[b] = c1;
    const double c0 = chi2[b];
    // largest relative parameter change of the step
    double rel = 0.0;
    for (int k = 0; k < ldp; ++k) {
      const double a = params[(size_t)b * ldp + k], t = trial[(size_t)b * ldp + k];
      if (a != t) rel = fmax(rel, fabs(t - a) / fmax(fabs(a), 1.0e-300));
    }
    int acc = 0, why = 0;
    if (c1 < c0) {   // false for NaN trials
      acc = 1;
      chi2[b] = c1;
      lambda[b] = fmax(lambda[b] / down, 1.0e-12);
      n_accept[b] += 1;
      if ((c0 - c1) <= ftol * c0 || rel <= xtol) why = 1;
    } else {
      lambda[b] = lambda[b] * up;
      if (c1 == c1 && c1 != INFINITY && (c1 - c0) <= ftol * c0 && rel > 0.0) why = 1;   // as good as the current point
      else if (rel > 0.0 && rel <= xtol) why = 1;
      else if (lambda[b] > lambda_max) why = 2;
      if (!(c0 == c0) || c0 == INFINITY) why = 3;
      if (c0 == 0.0) why = 1;                        // nothing left to improve (exact data, or an empty fit mask)
    }
    take = acc;
    accepted[b] = (uint8_t)acc;
    if (why) { active[b] = 0; reason[b] = (uint8_t)why; }
    else if (flags != nullptr) atomicMax(&flags[1], tag > 0 ? tag : flags[2]);      // still running after this iteration
    needjac[b] = (uint8_t)(acc && !why);
  }
  __syncthreads();
  if (take)
    for (int k = threadIdx.x; k < ldp; k += CHI2_THREADS) params[(size_t)b * ldp + k] = trial[(size_t)b * ldp + k];
}

// Accepted systems take over what the trial evaluation left behind: its pattern and running totals become the base
// rows of the resident Jacobian workspace (slab s of the 1 + n_pops trial slabs goes to slab 0 or n_var + s - 1), and
// its reflection tables become the system's tables -- for a shared table that is one row index, for an own table the
// row's contents.  One launch, grid = batch x (slab chunks + n_xtal).
__global__ void __launch_bounds__(COPY_THREADS)
lm_accept_kernel(const uint8_t *__restrict__ accepted, int batch, double *__restrict__ dst, long long dst_stride,
                 const double *__restrict__ src, long long src_stride, int n_q, int n_var, int n_pops, int chunks_per_slab,
                 const xrsd_tables_t T, const xrsd_tables_t Tt, int n_xtal) {
  const int per_sys = (dst != nullptr ? (1 + n_pops) * chunks_per_slab : 0) + n_xtal;
  const int b = blockIdx.x / per_sys;
  int c = blockIdx.x - b * per_sys;
  if (b >= batch || !accepted[b]) return;
  if (dst != nullptr) {
    if (c < (1 + n_pops) * chunks_per_slab) {
      const int s = c / chunks_per_slab, part = c - s * chunks_per_slab;
      double *__restrict__ d = dst + (size_t)b * dst_stride + (size_t)(s == 0 ? 0 : n_var + s - 1) * n_q + (size_t)part * COPY_CHUNK;
      const double *__restrict__ sp = src + (size_t)b * src_stride + (size_t)s * n_q + (size_t)part * COPY_CHUNK;
      const int n = min(COPY_CHUNK, n_q - part * COPY_CHUNK);
      double v[COPY_CHUNK / COPY_THREADS];
#pragma unroll
      for (int k = 0; k < COPY_CHUNK / COPY_THREADS; ++k) {
        const int i = threadIdx.x + k * COPY_THREADS;
        if (i < n) v[k] = sp[i];
      }
#pragma unroll
      for (int k = 0; k < COPY_CHUNK / COPY_THREADS; ++k) {
        const int i = threadIdx.x + k * COPY_THREADS;
        if (i < n) d[i] = v[k];
      }
      return;
    }
    c -= (1 + n_pops) * chunks_per_slab;
  }
  // table slot c
  const int t = b * n_xtal + c;
  const int src_row = Tt.table_of != nullptr ? Tt.table_of[t] : Tt.own_base + t;
  if (T.table_of != nullptr && Tt.n_shared > 0 && src_row >= Tt.shared_base) {
    if (threadIdx.x == 0) T.table_of[t] = src_row;
    return;
  }
  const int dst_row = T.own_base + t;
  const int ns = min(Tt.n_shell[src_row], T.cap_shell);
  const int32_t *sh = Tt.shell_hkl + (size_t)src_row * Tt.cap_shell * 4;
  int32_t *dh = T.shell_hkl + (size_t)dst_row * T.cap_shell * 4;
  for (int i = threadIdx.x; i < ns * 4; i += COPY_THREADS) dh[i] = sh[i];
  const double *sg = Tt.shell_geo + (size_t)src_row * Tt.cap_shell * Tt.n_pairs_max;
  double *dg = T.shell_geo + (size_t)dst_row * T.cap_shell * T.n_pairs_max;
  for (int i = threadIdx.x; i < ns * T.n_pairs_max; i += COPY_THREADS) dg[i] = sg[i];
  if (threadIdx.x == 0) {
    T.n_shell[dst_row] = Tt.n_shell[src_row];
    T.n_refl[dst_row] = Tt.n_refl[src_row];
    T.n_all[dst_row] = Tt.n_all[src_row];
    T.status[dst_row] = Tt.status[src_row];
    if (T.table_of != nullptr) T.table_of[t] = dst_row;
  }
}

// Rows of `src` flagged in `mask` replace the same rows of `dst` (row r starts at r*stride; `row_len` doubles are
// copied).  The LM loop uses it for accepted steps: the trial pattern and its running totals become the base rows
// of the resident Jacobian workspace.  Unflagged rows cost one byte read; flagged ones stream at HBM/L2 speed
// (each CTA copies COPY_CHUNK doubles with 8 independent loads per thread in flight).
__global__ void __launch_bounds__(COPY_THREADS)
copy_rows_masked_kernel(double *__restrict__ dst, long long dst_stride, const double *__restrict__ src, long long src_stride,
                        int row_len, int chunks, const uint8_t *__restrict__ mask) {
  const int r = blockIdx.x / chunks, c = blockIdx.x - r * chunks;
  if (!mask[r]) return;
  double *__restrict__ d = dst + (size_t)r * dst_stride + (size_t)c * COPY_CHUNK;
  const double *__restrict__ s = src + (size_t)r * src_stride + (size_t)c * COPY_CHUNK;
  const int n = min(COPY_CHUNK, row_len - c * COPY_CHUNK);
  double v[COPY_CHUNK / COPY_THREADS];
#pragma unroll
  for (int k = 0; k < COPY_CHUNK / COPY_THREADS; ++k) {
    const int i = threadIdx.x + k * COPY_THREADS;
    if (i < n) v[k] = s[i];
  }
#pragma unroll
  for (int k = 0; k < COPY_CHUNK / COPY_THREADS; ++k) {
    const int i = threadIdx.x + k * COPY_THREADS;
    if (i < n) d[i] = v[k];
  }
}

}  // namespace xrsd

using xrsd::FreeSet;

extern "C" cudaError_t xrsd_launch_residual(const xrsd_layout_t *layout, const xrsd_objective_t *obj, const double *d_q,
                                            int n_q, const double *d_params, int ldp, int batch,
                                            const xrsd_tables_t *tables, const double *d_Iobs, const double *d_dI,
                                            long long ld_obs, double *d_chi2, int q_per_cta, int scratch_doubles,
                                            int radial_multi, const uint8_t *d_active, double *d_acc, unsigned *d_tickets,
                                            cudaStream_t stream) {
  using namespace xrsd;
  const size_t smem = ((size_t)((q_per_cta + 1) & ~1) + (size_t)scratch_doubles) * sizeof(double);
  auto kern = radial_multi ? residual_kernel<true, true>
                           : (layout->n_xtal ? residual_kernel<false, true> : residual_kernel<false, false>);
  cudaError_t e = allow_max_dynamic_smem(kern);
  if (e != cudaSuccess) return e;
  if (batch <= 0) return cudaSuccess;
  const int n_tiles = (n_q + q_per_cta - 1) / q_per_cta;
  const long long n_blocks = (long long)batch * n_tiles;
  if (n_blocks > 2147483647LL) return cudaErrorInvalidConfiguration;
  kern<<<(unsigned)n_blocks, EVAL_THREADS, smem, stream>>>(*layout, *obj, d_q, n_q, d_params, ldp, batch, *tables, d_Iobs, d_dI,
                                                          ld_obs, d_chi2, scratch_doubles, d_active, d_acc, d_tickets, q_per_cta,
                                                          n_tiles);
  return cudaGetLastError();
}

extern "C" size_t xrsd_jacobian_smem(int n_params, int n_free, int q_per_cta, int scratch_doubles, int *pattern_doubles) {
  const int pat = q_per_cta < 0 ? 0 : (q_per_cta + 1) & ~1;        // q_per_cta < 0: direct mode, span = -q_per_cta
  if (pattern_doubles) *pattern_doubles = pat;
  return ((size_t)pat + (size_t)scratch_doubles + (size_t)((n_params + 1) & ~1)) * sizeof(double);
}

extern "C" int xrsd_jacobian_count_variants(const xrsd_layout_t *layout, const int32_t *free_idx, int n_free, const int32_t *ties,
                                            int n_tie) {
  return xrsd::make_free_set(layout, free_idx, n_free, ties, nullptr, n_tie).n_var;
}

extern "C" size_t xrsd_jacobian_acc_doubles(int n_free) {
  const int NF = n_free <= 8 ? 8 : XRSD_MAX_FREE;
  return (size_t)(NF * NF + NF + 2);
}

extern "C" cudaError_t xrsd_launch_jacobian(const xrsd_layout_t *layout, const xrsd_objective_t *obj, const double *d_q,
                                            int n_q, const double *d_params, int ldp, int batch,
                                            const xrsd_tables_t *tables, const double *d_Iobs, const double *d_dI,
                                            long long ld_obs, const int32_t *free_idx, int n_free, const int32_t *ties,
                                            const double *tie_coef, int n_tie, double rel_step, int base_ready,
                                            double *d_JTJ, double *d_JTr, double *d_chi2, double *d_Fvar, double *d_acc,
                                            unsigned *d_tickets, int q_per_cta, int scratch_doubles, int radial_multi,
                                            const uint8_t *d_active, cudaStream_t stream) {
  using namespace xrsd;
  const FreeSet F = make_free_set(layout, free_idx, n_free, ties, tie_coef, n_tie);
  int pat = 0;
  const size_t smem = xrsd_jacobian_smem(layout->n_params, n_free, q_per_cta, scratch_doubles, &pat);
  auto kern = radial_multi ? jacobian_variants_kernel<true, true>
                           : (layout->n_xtal ? jacobian_variants_kernel<false, true> : jacobian_variants_kernel<false, false>);
  cudaError_t e = allow_max_dynamic_smem(kern);
  if (e != cudaSuccess) return e;
  if (batch <= 0) return cudaSuccess;
  if (q_per_cta < 0) q_per_cta = -q_per_cta;
  const int n_tiles = (n_q + q_per_cta - 1) / q_per_cta;
  const long long n_blocks = (long long)batch * F.n_var * n_tiles;
  if (n_blocks > 2147483647LL) return cudaErrorInvalidConfiguration;
  kern<<<(unsigned)n_blocks, EVAL_THREADS, smem, stream>>>(*layout, d_q, n_q, d_params, ldp, batch, *tables, F, rel_step, d_Fvar,
                                                          q_per_cta, n_tiles, pat, scratch_doubles, d_active,
                                                          base_ready == XRSD_JAC_BASE_READY);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (base_ready == XRSD_JAC_PATTERNS_ONLY) return cudaSuccess;     // the caller only wants the slabs
  // q-chunk per reduction CTA: every CTA pays a prologue and a 74-value fold, so chunks are as long as the grid
  // allows -- whole patterns once the batch alone fills the GPU (2 CTAs/SM by registers), 1024 points for small batches
  int chunks_per_sys = (int)((4LL * 148 * 2 + batch - 1) / batch);
  chunks_per_sys = std::max(1, std::min(chunks_per_sys, (n_q + 1023) / 1024));
  const int chunk_q = (((n_q + chunks_per_sys - 1) / chunks_per_sys) + JR_THREADS - 1) / JR_THREADS * JR_THREADS;
  const int n_chunks = (n_q + chunk_q - 1) / chunk_q;
  if ((long long)batch * n_chunks > 2147483647LL) return cudaErrorInvalidConfiguration;
  static const bool use_mma = getenv("XRSD_JR_NO_MMA") == nullptr;      // A/B switch: the register-accumulator version
  if (use_mma) {
    const int NFt = n_free <= 8 ? 8 : XRSD_MAX_FREE;
    const size_t smem = ((size_t)JM_STAGES * (F.n_var + layout->n_pops + 3) * JR_THREADS + (size_t)(JR_THREADS / 32) * (NFt + 2) * JM_PITCH) *
                        sizeof(double);
    auto launch = [&](auto kern) -> cudaError_t {
      cudaError_t er = allow_max_dynamic_smem(kern);
      if (er != cudaSuccess) return er;
      kern<<<batch * n_chunks, JR_THREADS, smem, stream>>>(layout->n_pops, *obj, d_q, n_q, d_params, ldp, batch, d_Iobs, d_dI, ld_obs, F,
                                                          rel_step, d_Fvar, d_acc, d_tickets, d_JTJ, d_JTr, d_chi2, chunk_q, n_chunks,
                                                          d_active);
      return cudaGetLastError();
    };
    return n_free <= 8 ? launch(jacobian_reduce_mma_kernel<8>) : launch(jacobian_reduce_mma_kernel<XRSD_MAX_FREE>);
  }
  const size_t ring_bytes = (size_t)JR_STAGES * (F.n_var + layout->n_pops + 3) * JR_THREADS * sizeof(double);
  // opt in on every launch (microseconds; the attribute is per device and static shared memory counts against the
  // 48 KB default too)
  e = n_free <= 8 ? allow_max_dynamic_smem(jacobian_reduce_kernel<8, 8>) : allow_max_dynamic_smem(jacobian_reduce_kernel<XRSD_MAX_FREE, 4>);
  if (e != cudaSuccess) return e;
  if (n_free <= 8)
    jacobian_reduce_kernel<8, 8><<<batch * n_chunks, JR_THREADS, ring_bytes, stream>>>(
        layout->n_pops, *obj, d_q, n_q, d_params, ldp, batch, d_Iobs, d_dI, ld_obs, F, rel_step, d_Fvar, d_acc, d_tickets, d_JTJ,
        d_JTr, d_chi2, chunk_q, n_chunks, d_active);
  else
    jacobian_reduce_kernel<XRSD_MAX_FREE, 4><<<batch * n_chunks, JR_THREADS, ring_bytes, stream>>>(
        layout->n_pops, *obj, d_q, n_q, d_params, ldp, batch, d_Iobs, d_dI, ld_obs, F, rel_step, d_Fvar, d_acc, d_tickets, d_JTJ,
        d_JTr, d_chi2, chunk_q, n_chunks, d_active);
  return cudaGetLastError();
}

extern "C" cudaError_t xrsd_launch_lm_propose(const double *d_params, int ldp, int batch, const int32_t *free_idx, int n_free,
                                              const int32_t *ties, const double *tie_coef, int n_tie,
                                              const double *d_JTJ, const double *d_JTr, const double *d_lambda,
                                              const double *d_lo, const double *d_hi, const uint8_t *d_active,
                                              double *d_trial, double max_rel_step, int *d_iter, cudaStream_t stream) {
  const FreeSet F = xrsd::make_free_set(nullptr, free_idx, n_free, ties, tie_coef, n_tie);
  if (batch <= 0) return cudaSuccess;
  xrsd::lm_propose_kernel<<<(batch + 127) / 128, 128, 0, stream>>>(d_params, ldp, batch, F, d_JTJ, d_JTr, d_lambda, d_lo, d_hi,
                                                                   d_active, d_trial, max_rel_step, d_iter);
  return cudaGetLastError();
}

extern "C" cudaError_t xrsd_launch_lm_update(double *d_params, int ldp, int batch, const double *d_trial, double *d_chi2,
                                             const double *d_chi2_trial, double *d_lambda, uint8_t *d_active,
                                             int *d_n_accept, double up, double down, double ftol, double lambda_max,
                                             cudaStream_t stream) {
  if (batch <= 0) return cudaSuccess;
  xrsd::lm_update_kernel<<<(batch + 127) / 128, 128, 0, stream>>>(d_params, ldp, batch, d_trial, d_chi2, d_chi2_trial,
                                                                  d_lambda, d_active, d_n_accept, up, down, ftol, lambda_max);
  return cudaGetLastError();
}

extern "C" cudaError_t xrsd_launch_copy_rows_masked(double *d_dst, long long dst_stride, const double *d_src,
                                                    long long src_stride, long long row_len, int n_rows,
                                                    const uint8_t *d_mask, cudaStream_t stream) {
  if (n_rows <= 0 || row_len <= 0) return cudaSuccess;
  const long long chunks = (row_len + xrsd::COPY_CHUNK - 1) / xrsd::COPY_CHUNK;
  if (row_len > 2147483647LL || chunks * n_rows > 2147483647LL) return cudaErrorInvalidConfiguration;
  xrsd::copy_rows_masked_kernel<<<(unsigned)(chunks * n_rows), xrsd::COPY_THREADS, 0, stream>>>(
      d_dst, dst_stride, d_src, src_stride, (int)row_len, (int)chunks, d_mask);
  return cudaGetLastError();
}

extern "C" cudaError_t xrsd_launch_objective_prepare(const xrsd_objective_t *obj, const double *d_q, int n_q, const double *d_Iobs,
                                                     const double *d_dI, long long ld_obs, int rows, double *d_fo, double *d_w,
                                                     long long ld_out, cudaStream_t stream) {
  if (rows <= 0 || n_q <= 0) return cudaSuccess;
  const dim3 grid((unsigned)rows, (unsigned)std::min(64, (n_q + 255) / 256));
  xrsd::objective_prepare_kernel<<<grid, 256, 0, stream>>>(*obj, d_q, n_q, d_Iobs, d_dI, ld_obs, rows, d_fo, d_w, ld_out);
  return cudaGetLastError();
}

extern "C" cudaError_t xrsd_launch_lm_trial_update(const xrsd_objective_t *obj, const double *d_q, int n_q, const double *d_I,
                                                   long long ld_I, int batch, const double *d_Iobs, const double *d_dI,
                                                   long long ld_obs, double *d_params, int ldp, const double *d_trial,
                                                   double *d_chi2, double *d_chi2_trial, double *d_lambda, uint8_t *d_active,
                                                   uint8_t *d_accepted, uint8_t *d_needjac, uint8_t *d_reason, int *d_n_accept,
                                                   const xrsd_tables_t *trial_tables, int n_xtal, int *d_flags, int tag, double up,
                                                   double down, double ftol, double xtol, double lambda_max, cudaStream_t stream) {
  if (batch <= 0) return cudaSuccess;
  static const xrsd_tables_t none = {};
  xrsd::lm_trial_update_kernel<<<batch, xrsd::CHI2_THREADS, 0, stream>>>(
      *obj, d_q, n_q, d_I, ld_I, batch, d_Iobs, d_dI, ld_obs, d_params, ldp, d_trial, d_chi2, d_chi2_trial, d_lambda, d_active,
      d_accepted, d_needjac, d_reason, d_n_accept, trial_tables ? *trial_tables : none, trial_tables ? n_xtal : 0, d_flags, tag, up,
      down, ftol, xtol, lambda_max);
  return cudaGetLastError();
}

extern "C" cudaError_t xrsd_launch_lm_accept(const uint8_t *d_accepted, int batch, double *d_dst, long long dst_stride,
                                             const double *d_src, long long src_stride, int n_q, int n_var, int n_pops,
                                             const xrsd_tables_t *tables, const xrsd_tables_t *trial_tables, int n_xtal,
                                             cudaStream_t stream) {
  static const xrsd_tables_t none = {};
  if (!tables || !trial_tables) n_xtal = 0;
  const int chunks = (n_q + xrsd::COPY_CHUNK - 1) / xrsd::COPY_CHUNK;
  const long long per_sys = (d_dst ? (long long)(1 + n_pops) * chunks : 0) + n_xtal;
  if (batch <= 0 || per_sys <= 0) return cudaSuccess;
  if (per_sys * batch > 2147483647LL) return cudaErrorInvalidConfiguration;
  xrsd::lm_accept_kernel<<<(unsigned)(per_sys * batch), xrsd::COPY_THREADS, 0, stream>>>(
      d_accepted, batch, d_dst, dst_stride, d_src, src_stride, n_q, n_var, n_pops, chunks, tables ? *tables : none,
      trial_tables ? *trial_tables : none, n_xtal);
  return cudaGetLastError();
}

extern "C" cudaError_t xrsd_launch_pattern_chi2(const xrsd_objective_t *obj, const double *d_q, int n_q, const double *d_I,
                                                long long ld_I, int batch, const double *d_Iobs, const double *d_dI,
                                                long long ld_obs, double *d_chi2, const uint8_t *d_active, cudaStream_t stream) {
  if (batch <= 0) return cudaSuccess;
  xrsd::pattern_chi2_kernel<<<batch, xrsd::CHI2_THREADS, 0, stream>>>(*obj, d_q, n_q, d_I, ld_I, batch, d_Iobs, d_dI, ld_obs,
                                                                      d_chi2, d_active);
  return cudaGetLastError();
}
