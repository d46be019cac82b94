// fit_kernels.cu -- batched fit objective, finite-difference normal equations and the
// Levenberg-Marquardt step, one CTA (objective / Jacobian) or one thread (LM algebra) per
// system.  I(q) is produced by accumulate_system() in shared memory and consumed in the same
// kernel; it never goes to HBM.
//
//   residual_kernel  <- System.evaluate_residual, system/__init__.py:134-187, and
//                       compute_chi2, tools/__init__.py:104-125 (scalar weighted chi^2,
//                       including the reference's `wts *= dI**2` convention)
//   jacobian_kernel  <- what one lmfit objective sweep would need (system/__init__.py:189-195
//                       evaluates one parameter set per call; here n_free+1 sets per system)
//   lm_*             <- replaces the lmfit Nelder-Mead driver of system/__init__.py:262-266
//                       (no oracle in the reference; parity is defined on outcomes)
#include <string.h>
#include <algorithm>
#include "system_eval.cuh"

namespace xrsd {

struct FreeSet {
  int n_free;
  int idx[XRSD_MAX_FREE];       // parameter slot of free parameter j
  int n_var;                    // evaluated variants: 1 (base) + number of non-linear free parameters
  int var_of[XRSD_MAX_FREE];    // variant index (>= 1) of a non-linear parameter, 0 for a linear one
  int pop_of[XRSD_MAX_FREE];    // linear parameter (an I0): index of its population, else -1
  int free_of_var[XRSD_MAX_FREE + 1];   // inverse of var_of
};

// An I0 multiplies its whole population (scattering/__init__.py:48,64,66; system/noise.py:56-62), so
// dI/dI0 = I_pop / I0 needs no extra evaluation: the base variant records the running total after
// each population and the reducer forms f(I + h I_pop / I0) from it.
static FreeSet make_free_set(const xrsd_layout_t *L, const int32_t *free_idx, int n_free) {
  FreeSet F;
  memset(&F, 0, sizeof(F));
  F.n_free = n_free;
  F.n_var = 1;
  for (int j = 0; j < XRSD_MAX_FREE; ++j) { F.pop_of[j] = -1; }
  for (int j = 0; j < n_free; ++j) {
    F.idx[j] = free_idx[j];
    int pop = -1;
    if (L)
      for (int k = 0; k < L->n_pops; ++k)
        if (L->pops[k].kind != XRSD_POP_NONE && L->pops[k].p_I0 == free_idx[j]) pop = k;
    // a slot shared by two populations (not produced by the lowering) would not be linear in one term
    int uses = 0;
    if (L)
      for (int k = 0; k < L->n_pops; ++k) uses += (L->pops[k].p_I0 == free_idx[j]);
    if (pop >= 0 && uses == 1) {
      F.pop_of[j] = pop;
      F.var_of[j] = 0;
    } else {
      F.var_of[j] = F.n_var;
      F.free_of_var[F.n_var] = j;
      ++F.n_var;
    }
  }
  return F;
}

// weight and mask of one q-point (system/__init__.py:164-172)
__device__ __forceinline__ bool fit_weight(const xrsd_objective_t &O, double qv, double Iobs, const double *dI, size_t k,
                                           double &w) {
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

template <bool RM, bool HX>
__global__ void __launch_bounds__(EVAL_THREADS, HX ? 4 : 6)
residual_kernel(const __grid_constant__ xrsd_layout_t L, const xrsd_objective_t O, const double *__restrict__ q, int n_q,
                const double *__restrict__ params, int ldp, int batch, const xrsd_tables_t T,
                const double *__restrict__ Iobs, const double *__restrict__ dI, long long ld_obs,
                double *__restrict__ chi2, int scratch_doubles) {
  extern __shared__ __align__(16) double smem_d[];
  __shared__ EvalSharedT<HX> S;
  const int b = blockIdx.x;
  if (b >= batch) return;
  double *I_sm = smem_d;
  double *scratch = smem_d + ((n_q + 1) & ~1);
  accumulate_system<2, RM, HX>(L, params + (size_t)b * ldp, T, b, q, n_q, 0, n_q, I_sm, scratch, scratch_doubles, &S);
  const double *obs = Iobs + (size_t)b * ld_obs;
  const double *err = dI ? dI + (size_t)b * ld_obs : nullptr;
  double num = 0.0, den = 0.0;
  for (int i = threadIdx.x; i < n_q; i += EVAL_THREADS) {
    const double Io = obs[i], Ic = I_sm[i];
    double w;
    bool fit = fit_weight(O, q[i], Io, err, i, w);
    double d;
    if (O.logI_weighted) {
      fit = fit && (Ic > 0.0);
      d = fit ? (log(Ic) - log(Io)) : 0.0;
    } else {
      d = Ic - Io;
    }
    if (fit) { num += (d * d) * w; den += w; }
  }
  num = block_sum(num, &S);
  den = block_sum(den, &S);
  if (threadIdx.x == 0) chi2[b] = num / den;
}

// Normal equations by forward differences, ONE launch.  Grid = batch x (n_free+1) variants x q-spans:
// every CTA evaluates one parameter variant of one system over one q-span exactly like the intensity
// kernel does and parks f(I) (log I or I) in an HBM scratch [batch][n_free+1][n_q] (1.2 MB per system
// at cfg4 sizes, written and read once: ~0.3 us of HBM time against ~10 us of FP64 work).  The CTA
// that finishes last for a system (atomic ticket) reduces its variants to JTJ / JTr / chi2, staging
// 256 q-points of all variants at a time in shared memory.
template <bool RM, bool HX>
__global__ void __launch_bounds__(EVAL_THREADS, 6)
jacobian_kernel(const __grid_constant__ xrsd_layout_t L, const xrsd_objective_t O, const double *__restrict__ q, int n_q,
                const double *__restrict__ params, int ldp, int batch, const xrsd_tables_t T,
                const double *__restrict__ Iobs, const double *__restrict__ dI, long long ld_obs,
                const FreeSet F, double rel_step, double *__restrict__ JTJ, double *__restrict__ JTr,
                double *__restrict__ chi2, double *__restrict__ Fvar, unsigned *__restrict__ tickets,
                int q_per_cta, int n_tiles, int pattern_doubles, int scratch_doubles) {
  extern __shared__ __align__(16) double smem_d[];
  __shared__ EvalSharedT<HX> S;
  __shared__ double acc_sm[XRSD_MAX_FREE * (XRSD_MAX_FREE + 1) / 2 + XRSD_MAX_FREE + 2];
  __shared__ double steps[XRSD_MAX_FREE];
  __shared__ int is_last;
  const int tid = threadIdx.x;
  const int nf = F.n_free, nv = F.n_var, n_slab = nv + L.n_pops;
  const long long bid = blockIdx.x;
  const int per_sys = nv * n_tiles;
  const int b = (int)(bid / per_sys);
  const int rem = (int)(bid - (long long)b * per_sys);
  const int v = rem / n_tiles, tile = rem - v * n_tiles;
  if (b >= batch) return;
  const int np = L.n_params;
  double *I_sm = smem_d;
  double *scratch = smem_d + pattern_doubles;
  double *prm_s = scratch + scratch_doubles;
  const double *prm = params + (size_t)b * ldp;
  for (int i = tid; i < np; i += EVAL_THREADS) prm_s[i] = prm[i];
  if (tid < nf) {
    const double p0 = prm[F.idx[tid]];
    steps[tid] = (p0 != 0.0) ? fabs(p0) * rel_step : rel_step;
  }
  __syncthreads();
  if (v > 0 && tid == 0) {
    const int j = F.free_of_var[v];
    prm_s[F.idx[j]] = prm[F.idx[j]] + steps[j];
  }
  __syncthreads();
  const int q_begin = tile * q_per_cta;
  const int q_count = min(q_per_cta, n_q - q_begin);
  // slabs of this system in the scratch: [0, nv) variants (slab 0 = base, raw I), [nv, nv+n_pops) the
  // base variant's running totals after each population
  double *slab0 = Fvar + (size_t)b * n_slab * n_q;
  accumulate_system<2, RM, HX>(L, prm_s, T, b, q, n_q, q_begin, q_count, I_sm, scratch, scratch_doubles, &S, prm,
                               v == 0 ? slab0 + (size_t)nv * n_q + q_begin : nullptr, n_q);
  double *dst = slab0 + (size_t)v * n_q + q_begin;
  for (int i = tid; i < q_count; i += EVAL_THREADS) {
    const double x = I_sm[i];
    dst[i] = (v > 0 && O.logI_weighted) ? ((x > 0.0) ? log(x) : NAN) : x;
  }
  // ---- ticket: the last CTA of this system reduces
  __threadfence();
  __syncthreads();
  if (tid == 0) is_last = (atomicAdd(&tickets[b], 1u) == (unsigned)(per_sys - 1));
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const int n_tri = nf * (nf + 1) / 2;
  const int n_acc = n_tri + nf + 2;
  for (int i = tid; i < n_acc; i += EVAL_THREADS) acc_sm[i] = 0.0;
  const double *obs = Iobs + (size_t)b * ld_obs;
  const double *err = dI ? dI + (size_t)b * ld_obs : nullptr;
  const double *Fb = slab0;
  constexpr int RT = 256;                       // q-points per reduction stage
  double *Fv = smem_d;                          // [nf + 1][RT]: f of the base and of one variant per free parameter
  double *wts = Fv + (size_t)(nf + 1) * RT;
  double *fobs = wts + RT;
  const int warp = tid >> 5, lane = tid & 31;
  for (int q0 = 0; q0 < n_q; q0 += RT) {
    const int qc = min(RT, n_q - q0);
    __syncthreads();
    for (int i = tid; i < (nf + 1) * qc; i += EVAL_THREADS) {
      const int jj = i / qc, k = i - jj * qc;        // jj = 0: base, jj = j + 1: free parameter j
      const double I0v = __ldcg(Fb + q0 + k);        // written by other CTAs: read through L2
      double val;
      if (jj == 0) {
        val = O.logI_weighted ? ((I0v > 0.0) ? log(I0v) : NAN) : I0v;
      } else if (F.var_of[jj - 1] > 0) {
        val = __ldcg(Fb + (size_t)F.var_of[jj - 1] * n_q + q0 + k);
      } else {
        const int kp = F.pop_of[jj - 1];
        double Ik = __ldcg(Fb + (size_t)(nv + kp) * n_q + q0 + k);
        if (kp > 0) Ik -= __ldcg(Fb + (size_t)(nv + kp - 1) * n_q + q0 + k);
        const double pj = prm[F.idx[jj - 1]];
        const double Iv = (pj != 0.0) ? fma(steps[jj - 1] / pj, Ik, I0v) : I0v;
        val = O.logI_weighted ? ((Iv > 0.0) ? log(Iv) : NAN) : Iv;
      }
      Fv[jj * RT + k] = val;
    }
    __syncthreads();
    for (int i = tid; i < qc; i += EVAL_THREADS) {
      const double Io = obs[q0 + i];
      double w;
      bool fit = fit_weight(O, q[q0 + i], Io, err, q0 + i, w);
      const double f0 = Fv[i];
      if (O.logI_weighted) fit = fit && (f0 == f0);                  // I_comp > 0
      wts[i] = fit ? w : 0.0;
      fobs[i] = fit ? (O.logI_weighted ? log(Io) : Io) : 0.0;
    }
    __syncthreads();
    // entry e of acc_sm is owned by warp (e mod n_warps); lanes stride over the stage
    for (int e = warp; e < n_acc; e += EVAL_THREADS / 32) {
      int j = -1, k = -1, kind;   // kind 0: JTJ(j,k)  1: JTr(j)  2: chi2 numerator  3: sum of weights
      if (e < n_tri) {
        kind = 0;
        int r = e;
        j = 0;
        while (r >= nf - j) { r -= nf - j; ++j; }
        k = j + r;
      } else if (e < n_tri + nf) { kind = 1; j = e - n_tri; }
      else if (e == n_tri + nf) kind = 2;
      else kind = 3;
      double sum = 0.0;
      for (int i = lane; i < qc; i += 32) {
        const double w = wts[i];
        if (w == 0.0) continue;
        const double f0 = Fv[i];
        double term;
        if (kind == 0) {
          const double dj = Fv[(j + 1) * RT + i] - f0, dk = Fv[(k + 1) * RT + i] - f0;
          term = (dj == dj && dk == dk) ? dj * dk : 0.0;
        } else if (kind == 1) {
          const double dj = Fv[(j + 1) * RT + i] - f0;
          term = (dj == dj) ? dj * (f0 - fobs[i]) : 0.0;
        } else if (kind == 2) {
          term = (f0 - fobs[i]) * (f0 - fobs[i]);
        } else {
          term = 1.0;
        }
        sum = fma(w, term, sum);
      }
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if (lane == 0) acc_sm[e] += sum;
    }
  }
  __syncthreads();
  // normalise by the sum of weights and the steps; write out (full symmetric JTJ)
  const double sw = acc_sm[n_tri + nf + 1];
  for (int e = tid; e < n_tri; e += EVAL_THREADS) {
    int r = e, j = 0;
    while (r >= nf - j) { r -= nf - j; ++j; }
    const int k = j + r;
    const double val = acc_sm[e] / (steps[j] * steps[k]) / sw;
    JTJ[((size_t)b * nf + j) * nf + k] = val;
    JTJ[((size_t)b * nf + k) * nf + j] = val;
  }
  if (tid < nf) JTr[(size_t)b * nf + tid] = acc_sm[n_tri + tid] / steps[tid] / sw;
  if (tid == 0) {
    chi2[b] = acc_sm[n_tri + nf] / sw;
    tickets[b] = 0;                              // ready for the next launch
  }
}

// ---------------------------------------------------------------------------- LM algebra
__global__ void lm_propose_kernel(const double *__restrict__ params, int ldp, int batch, const FreeSet F,
                                  const double *__restrict__ JTJ, const double *__restrict__ JTr,
                                  const double *__restrict__ lambda, const double *__restrict__ lo,
                                  const double *__restrict__ hi, const uint8_t *__restrict__ active,
                                  double *__restrict__ trial) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
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
    double v = p[F.idx[j]] + g[j];
    const double l = lo[(size_t)b * nf + j], h = hi[(size_t)b * nf + j];
    v = fmin(fmax(v, l), h);
    if (v == v) t[F.idx[j]] = v;
  }
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

}  // namespace xrsd

using xrsd::FreeSet;

extern "C" cudaError_t xrsd_launch_residual(const xrsd_layout_t *layout, const xrsd_objective_t *obj, const double *d_q,
                                            int n_q, const double *d_params, int ldp, int batch,
                                            const xrsd_tables_t *tables, const double *d_Iobs, const double *d_dI,
                                            long long ld_obs, double *d_chi2, int scratch_doubles, int radial_multi,
                                            cudaStream_t stream) {
  using namespace xrsd;
  const size_t smem = ((size_t)((n_q + 1) & ~1) + (size_t)scratch_doubles) * sizeof(double);
  auto kern = radial_multi ? residual_kernel<true, true>
                           : (layout->n_xtal ? residual_kernel<false, true> : residual_kernel<false, false>);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (batch <= 0) return cudaSuccess;
  kern<<<batch, EVAL_THREADS, smem, stream>>>(*layout, *obj, d_q, n_q, d_params, ldp, batch, *tables, d_Iobs, d_dI, ld_obs,
                                             d_chi2, scratch_doubles);
  return cudaGetLastError();
}

extern "C" size_t xrsd_jacobian_smem(int n_params, int n_free, int q_per_cta, int scratch_doubles, int *pattern_doubles) {
  // pattern buffer: the q-span while evaluating, [(n_free+1)+2][256] while reducing
  int pat = std::max((q_per_cta + 1) & ~1, (n_free + 3) * 256);
  if (pattern_doubles) *pattern_doubles = pat;
  return ((size_t)pat + (size_t)scratch_doubles + (size_t)((n_params + 1) & ~1)) * sizeof(double);
}

extern "C" cudaError_t xrsd_launch_jacobian(const xrsd_layout_t *layout, const xrsd_objective_t *obj, const double *d_q,
                                            int n_q, const double *d_params, int ldp, int batch,
                                            const xrsd_tables_t *tables, const double *d_Iobs, const double *d_dI,
                                            long long ld_obs, const int32_t *free_idx, int n_free, double rel_step,
                                            double *d_JTJ, double *d_JTr, double *d_chi2, double *d_Fvar,
                                            unsigned *d_tickets, int q_per_cta, int scratch_doubles, int radial_multi,
                                            cudaStream_t stream) {
  using namespace xrsd;
  const FreeSet F = make_free_set(layout, free_idx, n_free);
  int pat = 0;
  const size_t smem = xrsd_jacobian_smem(layout->n_params, n_free, q_per_cta, scratch_doubles, &pat);
  auto kern = radial_multi ? jacobian_kernel<true, true>
                           : (layout->n_xtal ? jacobian_kernel<false, true> : jacobian_kernel<false, false>);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (batch <= 0) return cudaSuccess;
  const int n_tiles = (n_q + q_per_cta - 1) / q_per_cta;
  const long long n_blocks = (long long)batch * F.n_var * n_tiles;
  if (n_blocks > 2147483647LL) return cudaErrorInvalidConfiguration;
  kern<<<(unsigned)n_blocks, EVAL_THREADS, smem, stream>>>(*layout, *obj, d_q, n_q, d_params, ldp, batch, *tables, d_Iobs, d_dI,
                                                          ld_obs, F, rel_step, d_JTJ, d_JTr, d_chi2, d_Fvar, d_tickets,
                                                          q_per_cta, n_tiles, pat, scratch_doubles);
  return cudaGetLastError();
}

extern "C" cudaError_t xrsd_launch_lm_propose(const double *d_params, int ldp, int batch, const int32_t *free_idx, int n_free,
                                              const double *d_JTJ, const double *d_JTr, const double *d_lambda,
                                              const double *d_lo, const double *d_hi, const uint8_t *d_active,
                                              double *d_trial, cudaStream_t stream) {
  const FreeSet F = xrsd::make_free_set(nullptr, free_idx, n_free);
  if (batch <= 0) return cudaSuccess;
  xrsd::lm_propose_kernel<<<(batch + 127) / 128, 128, 0, stream>>>(d_params, ldp, batch, F, d_JTJ, d_JTr, d_lambda, d_lo, d_hi,
                                                                   d_active, d_trial);
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
