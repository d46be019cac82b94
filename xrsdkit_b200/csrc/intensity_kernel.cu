// intensity_kernel.cu -- batched I(q): grid = batch x q-tiles, one CTA per (system, q-span).
// Staged mode: the pattern span is accumulated in shared memory by accumulate_system() and written
// once, coalesced (8 B per (system, q) of HBM traffic; q and the parameter row are re-read from L2).
// Direct mode (layouts with crystalline populations): the span is accumulated in place in the output
// row (an L2-resident read-modify-write per population, ~1 % of the arithmetic), so the span length is
// not tied to shared memory and one CTA can take a whole pattern -- the per-population prologue
// (cell, shells, Voigt tables) is then paid once per pattern instead of once per 2048 q-points.
// Replaces System.compute_intensity, system/__init__.py:112-130.
#include "launch_util.cuh"
#include "system_eval.cuh"

#ifndef XRSD_NX_MIN_CTAS
#define XRSD_NX_MIN_CTAS 5      // layouts without crystalline populations: 96 registers
#endif
#ifndef XRSD_NX_QPT
#define XRSD_NX_QPT 2           // q-points per thread and pass of the size-distribution loop
#endif

namespace xrsd {

template <int QPT, bool RM, bool HX, bool F32 = false>
__global__ void __launch_bounds__(EVAL_THREADS, HX ? XRSD_HX_MIN_CTAS : XRSD_NX_MIN_CTAS)
intensity_kernel(const __grid_constant__ xrsd_layout_t L, const double *__restrict__ q, int n_q,
                 const double *__restrict__ params, int ldp, int batch, const xrsd_tables_t T,
                 double *__restrict__ I_out, long long ldI, int q_per_cta, int n_tiles, int scratch_doubles,
                 double *__restrict__ cum_out, int direct) {
  extern __shared__ __align__(16) double smem_d[];
  __shared__ EvalSharedT<HX> S;
  const long long bid = blockIdx.x;
  const int b = (int)(bid / n_tiles);
  const int tile = (int)(bid - (long long)b * n_tiles);
  if (b >= batch) return;
  const int q_begin = tile * q_per_cta;
  const int q_count = min(q_per_cta, n_q - q_begin);
  double *dst = I_out + (size_t)b * ldI + q_begin;
  // direct mode == a layout with crystalline populations (HX): a compile-time fact, so the accumulation buffer has a
  // known address space and the shell scratch sits at the base of the dynamic shared memory -- with `direct` as a
  // run-time flag the compiler re-derived the scratch address inside the profile loops (9 % of the instructions of
  // the cfg3 launch, profiles/sass_by_line.py on r02_cfg3)
  if (direct != (HX ? 1 : 0)) return;
  double *scratch = smem_d;
  double *I_sm = HX ? dst : smem_d + scratch_doubles;
  // cum_out (optional): [batch][n_pops][n_q] running totals after the noise model and after each
  // population -- their differences are the per-population patterns (plot decomposition, I0 rescaling)
  accumulate_system<QPT, RM, HX, F32>(L, params + (size_t)b * ldp, T, b, q, n_q, q_begin, q_count, I_sm, scratch, scratch_doubles, &S,
                                 nullptr, cum_out ? cum_out + (size_t)b * L.n_pops * n_q + q_begin : nullptr, n_q);
  if constexpr (!HX)
    for (int i = threadIdx.x; i < q_count; i += EVAL_THREADS) dst[i] = I_sm[i];
}

// DFMA-bound probe: 8 independent chains per thread, 4096 DFMA each
__global__ void __launch_bounds__(256) fp64_probe_kernel(double *out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1e-3, a2 = a0 + 2e-3, a3 = a0 + 3e-3;
  double a4 = a0 + 4e-3, a5 = a0 + 5e-3, a6 = a0 + 6e-3, a7 = a0 + 7e-3;
  const double m = 0.999999, c = 1e-7;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (s == 123.456) out[0] = s;
}

// peak profiles on their own (tools/peak_math.py:24-47): out[i] = profile(x[i])
__global__ void __launch_bounds__(EVAL_THREADS) profile_kernel(int kind, double p0, double p1, const double *__restrict__ x,
                                                               int n, double *__restrict__ out) {
  __shared__ VoigtConst vc;
  Profile P;
  const double y = make_profile(P, kind, p0, p1);
  if (kind == XRSD_PROFILE_VOIGT) voigt_constants(&vc, y, threadIdx.x);
  __syncthreads();
  for (int i = blockIdx.x * EVAL_THREADS + threadIdx.x; i < n; i += gridDim.x * EVAL_THREADS) out[i] = profile_eval(P, &vc, x[i]);
}

}  // namespace xrsd

extern "C" cudaError_t xrsd_launch_profile(int kind, double p0, double p1, const double *d_x, int n, double *d_out,
                                           cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  const int blocks = (n + xrsd::EVAL_THREADS - 1) / xrsd::EVAL_THREADS;
  xrsd::profile_kernel<<<blocks < 1184 ? blocks : 1184, xrsd::EVAL_THREADS, 0, stream>>>(kind, p0, p1, d_x, n, d_out);
  return cudaGetLastError();
}

extern "C" cudaError_t xrsd_launch_intensity(const xrsd_layout_t *layout, const double *d_q, int n_q, const double *d_params,
                                             int ldp, int batch, const xrsd_tables_t *tables, double *d_I, long long ldI,
                                             int q_per_cta, int scratch_doubles, int radial_multi, double *d_cum,
                                             int direct, int single_precision_loops, cudaStream_t stream) {
  using namespace xrsd;
  const int n_tiles = (n_q + q_per_cta - 1) / q_per_cta;
  const size_t smem = ((direct ? 0 : (size_t)((q_per_cta + 1) & ~1)) + (size_t)scratch_doubles) * sizeof(double);
  auto kern = radial_multi ? intensity_kernel<2, true, true>
                           : (layout->n_xtal ? intensity_kernel<2, false, true> : intensity_kernel<XRSD_NX_QPT, false, false>);
  if (single_precision_loops) {
    if (layout->n_xtal) return cudaErrorInvalidValue;      // the caller refuses such layouts before it gets here
    kern = intensity_kernel<XRSD_NX_QPT, false, false, true>;
  }
  cudaError_t e = xrsd::allow_max_dynamic_smem(kern);
  if (e != cudaSuccess) return e;
  const long long n_blocks = (long long)batch * n_tiles;
  if (n_blocks <= 0) return cudaSuccess;
  if (n_blocks > 2147483647LL) return cudaErrorInvalidConfiguration;
  kern<<<(unsigned)n_blocks, EVAL_THREADS, smem, stream>>>(*layout, d_q, n_q, d_params, ldp, batch, *tables, d_I, ldI,
                                                          q_per_cta, n_tiles, scratch_doubles, d_cum, direct);
  return cudaGetLastError();
}

extern "C" cudaError_t xrsd_launch_fp64_probe(double *d_out, int blocks, int iters, cudaStream_t stream) {
  xrsd::fp64_probe_kernel<<<blocks, 256, 0, stream>>>(d_out, iters);
  return cudaGetLastError();
}
