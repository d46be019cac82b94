// abi.cu -- the extern "C" boundary declared in include/xrsd_b200.h: argument checks,
// launch-shape policy, error strings, and the host-buffer convenience path.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "../../include/xrsd_b200.h"

extern "C" {
int64_t xrsd_table_smem_layout(int32_t max_cells, int32_t cap_list);
int64_t xrsd_table_scratch_bytes(int64_t cap_big);
cudaError_t xrsd_launch_tables(const xrsd_layout_t *, const double *, int, int, const xrsd_tables_t *, int, int, const uint8_t *,
                               cudaStream_t);
cudaError_t xrsd_launch_intensity(const xrsd_layout_t *, const double *, int, const double *, int, int, const xrsd_tables_t *,
                                  double *, long long, int, int, int, double *, int, int, cudaStream_t);
cudaError_t xrsd_launch_fp64_probe(double *, int, int, cudaStream_t);
cudaError_t xrsd_launch_profile(int, double, double, const double *, int, double *, cudaStream_t);
cudaError_t xrsd_launch_features(const double *, int, const double *, int64_t, int, double *, int64_t, cudaStream_t);
cudaError_t xrsd_launch_residual(const xrsd_layout_t *, const xrsd_objective_t *, const double *, int, const double *, int, int,
                                 const xrsd_tables_t *, const double *, const double *, long long, double *, int, int, int,
                                 const uint8_t *, double *, unsigned *, cudaStream_t);
size_t xrsd_jacobian_smem(int, int, int, int, int *);
cudaError_t xrsd_launch_jacobian(const xrsd_layout_t *, const xrsd_objective_t *, const double *, int, const double *, int, int,
                                 const xrsd_tables_t *, const double *, const double *, long long, const int32_t *, int,
                                 const int32_t *, const double *, int, double, int,
                                 double *, double *, double *, double *, double *, unsigned *, int, int, int, const uint8_t *,
                                 cudaStream_t);
size_t xrsd_jacobian_acc_doubles(int);
cudaError_t xrsd_launch_pattern_chi2(const xrsd_objective_t *, const double *, int, const double *, long long, int, const double *,
                                     const double *, long long, double *, const uint8_t *, cudaStream_t);
cudaError_t xrsd_launch_copy_rows_masked(double *, long long, const double *, long long, long long, int, const uint8_t *,
                                         cudaStream_t);
int xrsd_jacobian_count_variants(const xrsd_layout_t *, const int32_t *, int, const int32_t *, int);
cudaError_t xrsd_launch_lm_propose(const double *, int, int, const int32_t *, int, const int32_t *, const double *, int, const double *, const double *,
                                   const double *, const double *, const double *, const uint8_t *, double *, double, int *, cudaStream_t);
cudaError_t xrsd_launch_lm_update(double *, int, int, const double *, double *, const double *, double *, uint8_t *, int *,
                                  double, double, double, double, cudaStream_t);
cudaError_t xrsd_launch_objective_prepare(const xrsd_objective_t *, const double *, int, const double *, const double *, long long,
                                          int, double *, double *, long long, cudaStream_t);
cudaError_t xrsd_launch_lm_trial_update(const xrsd_objective_t *, const double *, int, const double *, long long, int,
                                        const double *, const double *, long long, double *, int, const double *, double *,
                                        double *, double *, uint8_t *, uint8_t *, uint8_t *, uint8_t *, int *,
                                        const xrsd_tables_t *, int, int *, int, double, double, double, double, double, cudaStream_t);
cudaError_t xrsd_launch_lm_accept(const uint8_t *, int, double *, long long, const double *, long long, int, int, int,
                                  const xrsd_tables_t *, const xrsd_tables_t *, int, cudaStream_t);
}

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int cuda_fail(cudaError_t e, const char *what) {
  return fail(XRSD_ECUDA, "%s: %s", what, cudaGetErrorString(e));
}

constexpr int kMaxSmem = 227 * 1024;
constexpr int kStaticSmem = 2048;        // static shared memory of the table kernel (TableShared: 1.4 KB)
constexpr int kStaticSmemEval = 16 * 1024;   // upper bound for the evaluation kernels (EvalShared with the Chebyshev / Voigt tables: 11.9 KB)

int check_layout(const xrsd_layout_t *L) {
  if (!L) return fail(XRSD_EINVAL, "layout is NULL");
  if (L->n_pops < 1 || L->n_pops > XRSD_MAX_POPS) return fail(XRSD_EINVAL, "layout.n_pops=%d out of range", L->n_pops);
  if (L->n_params < 1 || L->n_params > XRSD_MAX_PARAMS) return fail(XRSD_EINVAL, "layout.n_params=%d out of range", L->n_params);
  int n_x = 0;
  for (int i = 0; i < L->n_pops; ++i) {
    const xrsd_pop_t &p = L->pops[i];
    if (p.kind == XRSD_POP_CRYSTALLINE) {
      if (p.n_species < 1 || p.n_species > XRSD_MAX_SPECIES)
        return fail(XRSD_EINVAL, "population %d: n_species=%d (max %d)", i, p.n_species, XRSD_MAX_SPECIES);
      if (p.n_sites < 1 || p.n_sites > 8) return fail(XRSD_EINVAL, "population %d: n_sites=%d", i, p.n_sites);
      if (p.n_ops < 0 || p.n_ops > XRSD_MAX_OPS) return fail(XRSD_EINVAL, "population %d: n_ops=%d", i, p.n_ops);
      if (!(L->wavelength != 0.0))   // scattering/__init__.py:224-225 raises ValueError
        return fail(XRSD_EINVAL, "cannot compute diffraction with source wavelength of %g", L->wavelength);
      if (p.table_slot != n_x) return fail(XRSD_EINVAL, "population %d: table_slot=%d, expected %d", i, p.table_slot, n_x);
      ++n_x;
    }
  }
  if (n_x != L->n_xtal) return fail(XRSD_EINVAL, "layout.n_xtal=%d but %d crystalline populations", L->n_xtal, n_x);
  return XRSD_OK;
}

// constraint terms (include/xrsd_b200.h, XRSD_MAX_TIES): slots in range, program words known, stack never over- or underrun
int check_ties(const int32_t *ties, int n_tie, int n_params) {
  int depth = 0;
  for (int t = 0; t < n_tie; ++t) {
    const int d = ties[2 * t], a = ties[2 * t + 1];
    if (d <= -XRSD_TIE_WORD) {
      const int op = -d - XRSD_TIE_WORD;
      if (op > XRSD_TIE_ATAN) return fail(XRSD_EINVAL, "ties[%d]: unknown expression word %d", 2 * t, op);
      if (op == XRSD_TIE_PUSHP && (a < 0 || a >= n_params)) return fail(XRSD_EINVAL, "ties[%d]=%d", 2 * t + 1, a);
      if (op <= XRSD_TIE_PUSHP) ++depth;
      else if (op <= XRSD_TIE_MAX) { if (depth < 2) return fail(XRSD_EINVAL, "ties[%d]: expression stack underrun", 2 * t); --depth; }
      else if (depth < 1) return fail(XRSD_EINVAL, "ties[%d]: expression stack underrun", 2 * t);
      if (depth > XRSD_TIE_STACK) return fail(XRSD_ECAPACITY, "constraint expression deeper than %d values (XRSD_TIE_STACK)", XRSD_TIE_STACK);
      continue;
    }
    const int dst = d < 0 ? -d - 1 : d;
    if (dst >= n_params) return fail(XRSD_EINVAL, "ties[%d]=%d", 2 * t, d);
    if (a == -1 && d >= 0) {                          // store from the stack
      if (depth != 1) return fail(XRSD_EINVAL, "ties[%d]: a store needs exactly one value on the expression stack", 2 * t);
      depth = 0;
    } else if (a < 0 || a >= n_params) {
      return fail(XRSD_EINVAL, "ties[%d]=%d", 2 * t + 1, a);
    }
  }
  if (depth != 0) return fail(XRSD_EINVAL, "constraint expression without a closing store term");
  return XRSD_OK;
}

// scratch doubles accumulate_system() needs for this layout
int scratch_doubles_for(const xrsd_layout_t *L, int *radial_multi) {
  int need = 2 * 128 + 2;
  *radial_multi = 0;
  for (int i = 0; i < L->n_pops; ++i) {
    const xrsd_pop_t &p = L->pops[i];
    if (p.kind == XRSD_POP_NONCRYSTALLINE && p.form == XRSD_FORM_SPHERE_R_NORMAL) {
      const double n = ceil(2.0 * p.sampling_width / p.sampling_step) + 2.0;
      if (!(n >= 1.0) || n > 12000.0) return -1;
      need = std::max(need, 3 * (int)n);        // (r_i, rho_i) as doubles, then as float2 for the single-precision path
    }
    if (p.kind == XRSD_POP_CRYSTALLINE && p.sf_radial && p.n_species > 1) {
      *radial_multi = 1;
      const int np = p.n_species * (p.n_species + 1) / 2;
      need = std::max(need, (1 + np) * 128);
    }
  }
  return (need + 1) & ~1;
}

int check_tables(const xrsd_layout_t *L, const xrsd_tables_t *T) {
  if (L->n_xtal == 0) return XRSD_OK;
  if (!T) return fail(XRSD_EINVAL, "layout has crystalline populations but tables is NULL");
  if (!T->n_shell || !T->shell_hkl || !T->shell_geo || !T->status || !T->n_refl || !T->n_all)
    return fail(XRSD_EINVAL, "tables: missing device arrays");
  if (T->cap_shell < 1) return fail(XRSD_EINVAL, "tables.cap_shell=%d", T->cap_shell);
  for (int i = 0; i < L->n_pops; ++i) {
    const xrsd_pop_t &p = L->pops[i];
    if (p.kind == XRSD_POP_CRYSTALLINE && p.n_species * (p.n_species + 1) / 2 > T->n_pairs_max)
      return fail(XRSD_EINVAL, "tables.n_pairs_max=%d too small for population %d", T->n_pairs_max, i);
  }
  return XRSD_OK;
}

const xrsd_tables_t kNoTables = {};

// Upper bound of the candidate hkl box over a HOST parameter matrix: n_i <= ceil(G_max * edge_i) + 1
// (scattering/__init__.py:266-268; |a_i| is the cell edge for every lattice class, b_i.a_i = 1).
int host_max_cells(const xrsd_layout_t *L, const double *params, int ld, int batch) {
  long long best = 1;
  for (int ip = 0; ip < L->n_pops; ++ip) {
    const xrsd_pop_t &p = L->pops[ip];
    if (p.kind != XRSD_POP_CRYSTALLINE) continue;
    double emax[3] = {0.0, 0.0, 0.0};
    for (int b = 0; b < batch; ++b) {
      const double *row = params + (size_t)b * ld;
      const double a = p.p_lat[0] >= 0 ? fabs(row[p.p_lat[0]]) : 0.0;
      const double bb = p.p_lat[1] >= 0 ? fabs(row[p.p_lat[1]]) : a;
      const double c = p.p_lat[2] >= 0 ? fabs(row[p.p_lat[2]]) : a;
      double e[3] = {a, bb, c};
      if (p.lattice == XRSD_LAT_HCP) { e[1] = a; e[2] = a * sqrt(8.0 / 3.0); }
      if (p.lattice == XRSD_LAT_HEXAGONAL || p.lattice == XRSD_LAT_TETRAGONAL) e[1] = a;
      if (p.lattice == XRSD_LAT_CUBIC || p.lattice == XRSD_LAT_RHOMBOHEDRAL) { e[1] = a; e[2] = a; }
      for (int i = 0; i < 3; ++i)
        if (e[i] > emax[i]) emax[i] = e[i];
    }
    const double g_max = p.q_max / (2.0 * M_PI);
    long long cells = 1;
    for (int i = 0; i < 3; ++i) {
      double n = ceil(g_max * emax[i] * (1.0 + 1e-12)) + 1.0;
      if (!(n >= 1.0)) n = 1.0;
      if (n > 10000.0) n = 10000.0;
      cells *= (2 * (long long)n - 1);
    }
    if (cells > best) best = cells;
  }
  return (int)std::min<long long>(best, 1 << 30);
}

}  // namespace

extern "C" {

const char *xrsd_last_error(void) { return g_err; }
int xrsd_abi_version(void) { return XRSD_ABI_VERSION; }

int xrsd_struct_sizes(int32_t *sp, int32_t *sl, int32_t *st, int32_t *so) {
  if (sp) *sp = (int32_t)sizeof(xrsd_pop_t);
  if (sl) *sl = (int32_t)sizeof(xrsd_layout_t);
  if (st) *st = (int32_t)sizeof(xrsd_tables_t);
  if (so) *so = (int32_t)sizeof(xrsd_objective_t);
  return XRSD_OK;
}

int32_t xrsd_sizeof_lm_state(void) { return (int32_t)sizeof(xrsd_lm_state_t); }

int64_t xrsd_table_smem_bytes(int32_t n1, int32_t n2, int32_t n3) {
  const int64_t cells = (int64_t)(2 * n1 - 1) * (2 * n2 - 1) * (2 * n3 - 1);
  return xrsd_table_smem_layout((int32_t)std::min<int64_t>(cells, 1 << 30), 1024);
}

int xrsd_reflection_tables(const xrsd_layout_t *layout, const double *d_params, int32_t ld_params, int32_t batch,
                           const xrsd_tables_t *tables, int32_t max_cells, const uint8_t *d_active, void *stream) {
  int rc = check_layout(layout);
  if (rc) return rc;
  if (layout->n_xtal == 0 || batch <= 0) return XRSD_OK;
  rc = check_tables(layout, tables);
  if (rc) return rc;
  if (!d_params || ld_params < layout->n_params) return fail(XRSD_EINVAL, "params: NULL or ld_params < n_params");
  if (max_cells < 1) return fail(XRSD_EINVAL, "max_cells=%d", max_cells);
  // reflection-list capacity (power of two): the caller's hint, else 1024 when every crystalline
  // population is reduced by >= 3 operations (the reduced list is then a small fraction of the
  // shell) and 4096 otherwise; shrunk until it fits beside the candidate box
  int cap_list = tables->cap_list_hint;
  if (cap_list <= 0) {
    cap_list = 1024;
    for (int i = 0; i < layout->n_pops; ++i) {
      const xrsd_pop_t &p = layout->pops[i];
      if (p.kind == XRSD_POP_CRYSTALLINE && !(p.use_symmetry && p.n_ops >= 3)) cap_list = 4096;
    }
  }
  int pow2 = 64;
  while (pow2 < cap_list && pow2 < 4096) pow2 <<= 1;
  cap_list = pow2;
  while (cap_list > 64 && xrsd_table_smem_layout(max_cells, cap_list) + kStaticSmem > kMaxSmem) cap_list >>= 1;
  if (xrsd_table_smem_layout(max_cells, cap_list) + kStaticSmem > kMaxSmem)
    return fail(XRSD_ECAPACITY, "candidate box of %d cells does not fit in shared memory", max_cells);
  cudaError_t e = xrsd_launch_tables(layout, d_params, ld_params, batch, tables, max_cells, cap_list, d_active, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "reflection_table_kernel");
  return XRSD_OK;
}

static int pick_q_per_cta(int n_q, int batch, int scratch, bool has_xtal) {
  // One CTA holds a span of one pattern in shared memory.  Spans are multiples of 512 q-points;
  // aim for >= 24 CTAs per SM in the grid (several waves, so the last partial wave costs little)
  // and keep the span small enough for the resident CTAs: 4096 q-points (32 KB, five or six CTAs per SM), or 2048
  // for the residual kernel of layouts with crystalline populations (~12 KB of static shared memory on top).
  // (The intensity and Jacobian kernels of crystalline layouts accumulate in place: pick_direct_span.)
  const long long want_ctas = 148LL * 24;
  long long tiles = (want_ctas + batch - 1) / batch;
  if (tiles < 1) tiles = 1;
  long long per = ((long long)n_q + tiles - 1) / tiles;
  per = ((per + 511) / 512) * 512;
  const long long cap = std::max(512LL, std::min(has_xtal ? 2048LL : 4096LL, ((long long)(36 * 1024 / 8 - scratch) / 512) * 512));
  per = std::min(per, cap);
  per = std::min<long long>(per, ((long long)n_q + 511) / 512 * 512);
  return (int)std::max(512LL, per);
}

static int pick_direct_span(int n_q, long long units) {
  // Direct mode (crystalline layouts): the span is accumulated in the output row, so shared memory does not
  // bound it.  Whole patterns per CTA when the grid still holds >= 4 waves of 8 CTAs per SM (the prologue of
  // every population is then paid once per pattern); otherwise split, down to 512-point spans.  Measured on
  // cfg3: 2-4 waves are within 1 % of each other, 16 waves cost 10-30 % (B = 10^4 ... 2000).
  const long long want = 4LL * 148 * 8;
  long long tiles = (want + units - 1) / std::max(1LL, units);
  tiles = std::max(1LL, std::min(tiles, ((long long)n_q + 511) / 512));
  long long per = ((long long)n_q + tiles - 1) / tiles;
  per = ((per + 255) / 256) * 256;
  return (int)per;
}

static int intensity_impl(const xrsd_layout_t *layout, const double *d_q, int32_t n_q, const double *d_params,
                          int32_t ld_params, int32_t batch, const xrsd_tables_t *tables, double *d_I, int64_t ld_I,
                          double *d_cum, void *stream, int single_precision_loops = 0) {
  int rc = check_layout(layout);
  if (rc) return rc;
  if (single_precision_loops && layout->n_xtal > 0)
    return fail(XRSD_EINVAL, "the single-precision path covers layouts without crystalline populations");
  if (batch <= 0 || n_q <= 0) return XRSD_OK;
  rc = check_tables(layout, tables);
  if (rc) return rc;
  if (!d_q || !d_params || !d_I) return fail(XRSD_EINVAL, "NULL device pointer");
  if (ld_params < layout->n_params) return fail(XRSD_EINVAL, "ld_params < n_params");
  if (ld_I < n_q) return fail(XRSD_EINVAL, "ld_I < n_q");
  int rm;
  const int scratch = scratch_doubles_for(layout, &rm);
  if (scratch < 0) return fail(XRSD_ECAPACITY, "size quadrature too fine for shared memory (sampling_width/sampling_step)");
  const int direct = layout->n_xtal > 0;
  const int per = direct ? pick_direct_span(n_q, batch) : pick_q_per_cta(n_q, batch, scratch, false);
  cudaError_t e = xrsd_launch_intensity(layout, d_q, n_q, d_params, ld_params, batch, tables ? tables : &kNoTables, d_I,
                                        (long long)ld_I, per, scratch, rm, d_cum, direct, single_precision_loops,
                                        (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "intensity_kernel");
  return XRSD_OK;
}

int xrsd_intensity_batch_f32(const xrsd_layout_t *layout, const double *d_q, int32_t n_q, const double *d_params,
                             int32_t ld_params, int32_t batch, double *d_I, int64_t ld_I, void *stream) {
  return intensity_impl(layout, d_q, n_q, d_params, ld_params, batch, nullptr, d_I, ld_I, nullptr, stream, 1);
}

int xrsd_intensity_batch(const xrsd_layout_t *layout, const double *d_q, int32_t n_q, const double *d_params,
                         int32_t ld_params, int32_t batch, const xrsd_tables_t *tables, double *d_I, int64_t ld_I,
                         void *stream) {
  return intensity_impl(layout, d_q, n_q, d_params, ld_params, batch, tables, d_I, ld_I, nullptr, stream);
}

int xrsd_intensity_components_batch(const xrsd_layout_t *layout, const double *d_q, int32_t n_q, const double *d_params,
                                    int32_t ld_params, int32_t batch, const xrsd_tables_t *tables, double *d_I,
                                    int64_t ld_I, double *d_cum, void *stream) {
  if (!d_cum) return fail(XRSD_EINVAL, "d_cum is NULL");
  return intensity_impl(layout, d_q, n_q, d_params, ld_params, batch, tables, d_I, ld_I, d_cum, stream);
}

// partial sums (sum w d^2, sum w) per (system, span); spans are at least 512 q-points long (pick_q_per_cta)
static int64_t residual_partial_bytes(int32_t batch, int32_t n_q) {
  return (((int64_t)batch * (((int64_t)n_q + 511) / 512) * 16) + 255) & ~255LL;
}

int64_t xrsd_residual_workspace_bytes(int32_t batch, int32_t n_q) {
  // the partial sums, then one ticket per system (zero before the first call; the kernel leaves them zero)
  return residual_partial_bytes(batch, n_q) + (((int64_t)batch * 4 + 255) & ~255LL);
}

int xrsd_residual_batch(const xrsd_layout_t *layout, const xrsd_objective_t *obj, const double *d_q, int32_t n_q,
                        const double *d_params, int32_t ld_params, int32_t batch, const xrsd_tables_t *tables,
                        const double *d_Iobs, const double *d_dI, int64_t ld_obs, double *d_chi2, void *d_workspace,
                        const uint8_t *d_active, void *stream) {
  int rc = check_layout(layout);
  if (rc) return rc;
  if (batch <= 0 || n_q <= 0) return XRSD_OK;
  rc = check_tables(layout, tables);
  if (rc) return rc;
  if (!obj || !d_q || !d_params || !d_Iobs || !d_chi2 || !d_workspace) return fail(XRSD_EINVAL, "NULL pointer");
  if ((obj->has_dI || obj->prepared) && !d_dI) return fail(XRSD_EINVAL, "objective.has_dI / prepared set but d_dI is NULL");
  if (ld_obs != 0 && ld_obs < n_q) return fail(XRSD_EINVAL, "ld_obs < n_q");
  int rm;
  const int scratch = scratch_doubles_for(layout, &rm);
  if (scratch < 0) return fail(XRSD_ECAPACITY, "size quadrature too fine for shared memory");
  const int per = pick_q_per_cta(n_q, batch, scratch, layout->n_xtal > 0);
  double *d_acc = (double *)d_workspace;
  unsigned *d_tickets = (unsigned *)((char *)d_workspace + residual_partial_bytes(batch, n_q));
  cudaError_t e = xrsd_launch_residual(layout, obj, d_q, n_q, d_params, ld_params, batch, tables ? tables : &kNoTables,
                                       d_Iobs, (obj->has_dI || obj->prepared) ? d_dI : nullptr, (long long)ld_obs, d_chi2, per, scratch, rm, d_active,
                                       d_acc, d_tickets, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "residual_kernel");
  return XRSD_OK;
}

// partial normal equations per (system, q-chunk): the reduction's grid holds at most batch + 4*148*2 chunks
static int64_t jacobian_partial_bytes(int32_t batch, int32_t n_free) {
  return ((((int64_t)batch + 4 * 148 * 2 + 1) * (int64_t)xrsd_jacobian_acc_doubles(n_free) * 8) + 255) & ~255LL;
}

int64_t xrsd_jacobian_workspace_bytes(int32_t n_q, int32_t batch, int32_t n_free) {
  // [batch][(n_free+1) variants + XRSD_MAX_POPS running totals][n_q] doubles (upper bound: linear
  // parameters need no variant slab), then per system the normal-equation accumulators and one ticket
  // (both zero-initialised by the caller once; the reduce kernel re-zeroes them)
  const int64_t fbytes = (int64_t)batch * (n_free + 1 + XRSD_MAX_POPS) * n_q * 8;
  return fbytes + jacobian_partial_bytes(batch, n_free) + (((int64_t)batch * 4 + 255) & ~255LL);
}

int xrsd_jacobian_batch(const xrsd_layout_t *layout, const xrsd_objective_t *obj, const double *d_q, int32_t n_q,
                        const double *d_params, int32_t ld_params, int32_t batch, const xrsd_tables_t *tables,
                        const double *d_Iobs, const double *d_dI, int64_t ld_obs, const int32_t *free_idx, int32_t n_free,
                        const int32_t *ties, const double *tie_coef, int32_t n_tie, double rel_step, int32_t mode, double *d_JTJ, double *d_JTr, double *d_chi2,
                        void *d_workspace,
                        const uint8_t *d_active, void *stream) {
  int rc = check_layout(layout);
  if (rc) return rc;
  if (batch <= 0 || n_q <= 0) return XRSD_OK;
  rc = check_tables(layout, tables);
  if (rc) return rc;
  if (mode != XRSD_JAC_FULL && mode != XRSD_JAC_BASE_READY && mode != XRSD_JAC_PATTERNS_ONLY)
    return fail(XRSD_EINVAL, "mode=%d", mode);
  const bool reduce = mode != XRSD_JAC_PATTERNS_ONLY;
  if (!obj || !d_q || !d_params || (!free_idx && n_free > 0) || !d_workspace || (reduce && (!d_Iobs || !d_JTJ || !d_JTr || !d_chi2)))
    return fail(XRSD_EINVAL, "NULL pointer");
  // n_free == 0 is "the pattern and its running totals" and only makes sense without the reduction
  if (n_free < (reduce ? 1 : 0) || n_free > XRSD_MAX_FREE) return fail(XRSD_EINVAL, "n_free=%d (max %d)", n_free, XRSD_MAX_FREE);
  for (int i = 0; i < n_free; ++i) {
    const int slot = free_idx[i] < 0 ? -free_idx[i] - 1 : free_idx[i];      // -(slot+1): forced finite-difference variant
    if (slot >= layout->n_params) return fail(XRSD_EINVAL, "free_idx[%d]=%d", i, free_idx[i]);
  }
  if (reduce && (obj->has_dI || obj->prepared) && !d_dI)
    return fail(XRSD_EINVAL, "objective.has_dI / prepared set but d_dI is NULL");
  if (reduce && ld_obs != 0 && ld_obs < n_q) return fail(XRSD_EINVAL, "ld_obs < n_q");
  if (!(rel_step > 0.0)) return fail(XRSD_EINVAL, "rel_step must be > 0");
  if (n_tie < 0 || n_tie > XRSD_MAX_TIES || (n_tie > 0 && !ties)) return fail(XRSD_EINVAL, "n_tie=%d (max %d)", n_tie, XRSD_MAX_TIES);
  rc = check_ties(ties, n_tie, layout->n_params);
  if (rc) return rc;
  int rm;
  const int scratch = scratch_doubles_for(layout, &rm);
  if (scratch < 0) return fail(XRSD_ECAPACITY, "size quadrature too fine for shared memory");
  // crystalline layouts accumulate straight into the variant slabs (negative span = direct mode)
  const int per = layout->n_xtal > 0 ? -pick_direct_span(n_q, (long long)batch * (n_free + 1))
                                     : pick_q_per_cta(n_q, batch * (n_free + 1), scratch, false);
  if (xrsd_jacobian_smem(layout->n_params, n_free, per, scratch, nullptr) + kStaticSmemEval > (size_t)kMaxSmem)
    return fail(XRSD_ECAPACITY, "jacobian working set does not fit in shared memory");
  const int64_t fbytes = (int64_t)batch * (n_free + 1 + XRSD_MAX_POPS) * n_q * 8;
  const int64_t abytes = jacobian_partial_bytes(batch, n_free);
  double *d_Fvar = (double *)d_workspace;
  double *d_acc = (double *)((char *)d_workspace + fbytes);
  unsigned *d_tickets = (unsigned *)((char *)d_workspace + fbytes + abytes);
  cudaError_t e = xrsd_launch_jacobian(layout, obj, d_q, n_q, d_params, ld_params, batch, tables ? tables : &kNoTables, d_Iobs,
                                       (obj->has_dI || obj->prepared) ? d_dI : nullptr, (long long)ld_obs, free_idx, n_free, ties, tie_coef, n_tie, rel_step, mode, d_JTJ,
                                       d_JTr, d_chi2, d_Fvar, d_acc, d_tickets, per, scratch, rm, d_active, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "jacobian_kernel");
  return XRSD_OK;
}

int32_t xrsd_jacobian_n_variants(const xrsd_layout_t *layout, const int32_t *free_idx, int32_t n_free, const int32_t *ties,
                                 int32_t n_tie) {
  if (check_layout(layout) || !free_idx || n_free < 1 || n_free > XRSD_MAX_FREE) return -1;
  return xrsd_jacobian_count_variants(layout, free_idx, n_free, ties, n_tie);
}

int xrsd_lm_propose(const double *d_params, int32_t ld_params, int32_t batch, const int32_t *free_idx, int32_t n_free,
                    const int32_t *ties, const double *tie_coef, int32_t n_tie, const double *d_JTJ, const double *d_JTr, const double *d_lambda, const double *d_lo, const double *d_hi,
                    const uint8_t *d_active, double *d_trial, double max_rel_step, int32_t *d_iter, void *stream) {
  if (n_free < 1 || n_free > XRSD_MAX_FREE) return fail(XRSD_EINVAL, "n_free=%d (max %d)", n_free, XRSD_MAX_FREE);
  if (!(max_rel_step >= 0.0)) return fail(XRSD_EINVAL, "max_rel_step must be >= 0 (0 = no limit)");
  if (!d_params || !free_idx || !d_JTJ || !d_JTr || !d_lambda || !d_lo || !d_hi || !d_trial) return fail(XRSD_EINVAL, "NULL pointer");
  for (int i = 0; i < n_free; ++i)
    if ((free_idx[i] < 0 ? -free_idx[i] - 1 : free_idx[i]) >= ld_params) return fail(XRSD_EINVAL, "free_idx[%d]=%d", i, free_idx[i]);
  if (n_tie < 0 || n_tie > XRSD_MAX_TIES || (n_tie > 0 && !ties)) return fail(XRSD_EINVAL, "n_tie=%d (max %d)", n_tie, XRSD_MAX_TIES);
  {
    const int rc_t = check_ties(ties, n_tie, ld_params);
    if (rc_t) return rc_t;
  }
  cudaError_t e = xrsd_launch_lm_propose(d_params, ld_params, batch, free_idx, n_free, ties, tie_coef, n_tie, d_JTJ, d_JTr, d_lambda, d_lo, d_hi,
                                         d_active, d_trial, max_rel_step, d_iter, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "lm_propose_kernel");
  return XRSD_OK;
}

int xrsd_lm_update(double *d_params, int32_t ld_params, int32_t batch, const double *d_trial, double *d_chi2,
                   const double *d_chi2_trial, double *d_lambda, uint8_t *d_active, int32_t *d_n_accept, double up,
                   double down, double ftol, double lambda_max, void *stream) {
  if (!d_params || !d_trial || !d_chi2 || !d_chi2_trial || !d_lambda) return fail(XRSD_EINVAL, "NULL pointer");
  cudaError_t e = xrsd_launch_lm_update(d_params, ld_params, batch, d_trial, d_chi2, d_chi2_trial, d_lambda, d_active,
                                        d_n_accept, up, down, ftol, lambda_max, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "lm_update_kernel");
  return XRSD_OK;
}

int xrsd_pattern_chi2_batch(const xrsd_objective_t *obj, const double *d_q, int32_t n_q, const double *d_I, int64_t ld_I,
                            int32_t batch, const double *d_Iobs, const double *d_dI, int64_t ld_obs, double *d_chi2,
                            const uint8_t *d_active, void *stream) {
  if (batch <= 0 || n_q <= 0) return XRSD_OK;
  if (!obj || !d_q || !d_I || !d_Iobs || !d_chi2) return fail(XRSD_EINVAL, "NULL pointer");
  if ((obj->has_dI || obj->prepared) && !d_dI) return fail(XRSD_EINVAL, "objective.has_dI / prepared set but d_dI is NULL");
  if (ld_I < n_q) return fail(XRSD_EINVAL, "ld_I < n_q");
  if (ld_obs != 0 && ld_obs < n_q) return fail(XRSD_EINVAL, "ld_obs < n_q");
  cudaError_t e = xrsd_launch_pattern_chi2(obj, d_q, n_q, d_I, (long long)ld_I, batch, d_Iobs, (obj->has_dI || obj->prepared) ? d_dI : nullptr,
                                           (long long)ld_obs, d_chi2, d_active, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "pattern_chi2_kernel");
  return XRSD_OK;
}

int xrsd_objective_prepare(const xrsd_objective_t *obj, const double *d_q, int32_t n_q, const double *d_Iobs, const double *d_dI,
                           int64_t ld_obs, int32_t rows, double *d_fo, double *d_w, int64_t ld_out, void *stream) {
  if (rows <= 0 || n_q <= 0) return XRSD_OK;
  if (!obj || !d_q || !d_Iobs || !d_fo || !d_w) return fail(XRSD_EINVAL, "NULL pointer");
  if (obj->prepared) return fail(XRSD_EINVAL, "objective is already marked prepared");
  if (obj->has_dI && !d_dI) return fail(XRSD_EINVAL, "objective.has_dI set but d_dI is NULL");
  if ((rows > 1 && ld_obs < n_q) || ld_out < n_q) return fail(XRSD_EINVAL, "leading dimension < n_q");
  cudaError_t e = xrsd_launch_objective_prepare(obj, d_q, n_q, d_Iobs, obj->has_dI ? d_dI : nullptr, (long long)ld_obs, rows, d_fo,
                                                d_w, (long long)ld_out, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "objective_prepare_kernel");
  return XRSD_OK;
}

int xrsd_lm_trial_update(const xrsd_objective_t *obj, const double *d_q, int32_t n_q, const double *d_I, int64_t ld_I,
                         int32_t batch, const double *d_Iobs, const double *d_dI, int64_t ld_obs, double *d_params,
                         int32_t ld_params, const double *d_trial, double *d_chi2, double *d_chi2_trial, double *d_lambda,
                         uint8_t *d_active, uint8_t *d_accepted, uint8_t *d_needjac, uint8_t *d_reason, int32_t *d_n_accept,
                         const xrsd_tables_t *trial_tables, int32_t n_xtal, int32_t *d_flags, int32_t tag, double up,
                         double down, double ftol, double xtol, double lambda_max, void *stream) {
  if (batch <= 0) return XRSD_OK;
  if (!obj || !d_params || !d_trial || !d_chi2 || !d_chi2_trial || !d_lambda || !d_active || !d_accepted || !d_needjac ||
      !d_reason || !d_n_accept)
    return fail(XRSD_EINVAL, "NULL pointer");
  if (d_I) {
    if (!d_q || !d_Iobs || n_q <= 0) return fail(XRSD_EINVAL, "a trial pattern needs q, I_obs and n_q > 0");
    if ((obj->has_dI || obj->prepared) && !d_dI) return fail(XRSD_EINVAL, "objective.has_dI / prepared set but d_dI is NULL");
    if (ld_I < n_q || (ld_obs != 0 && ld_obs < n_q)) return fail(XRSD_EINVAL, "leading dimension < n_q");
  }
  if (ld_params < 1 || ld_params > XRSD_MAX_PARAMS) return fail(XRSD_EINVAL, "ld_params=%d", ld_params);
  if (trial_tables && (n_xtal < 1 || !trial_tables->status)) return fail(XRSD_EINVAL, "trial tables without status / n_xtal");
  if (!(up > 1.0) || !(down > 1.0)) return fail(XRSD_EINVAL, "up and down must be > 1");
  cudaError_t e = xrsd_launch_lm_trial_update(obj, d_q, n_q, d_I, (long long)ld_I, batch, d_Iobs,
                                              (obj->has_dI || obj->prepared) ? d_dI : nullptr, (long long)ld_obs, d_params, ld_params,
                                              d_trial, d_chi2, d_chi2_trial, d_lambda, d_active, d_accepted, d_needjac, d_reason,
                                              d_n_accept, trial_tables, n_xtal, d_flags, tag, up, down, ftol, xtol, lambda_max,
                                              (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "lm_trial_update_kernel");
  return XRSD_OK;
}

int xrsd_lm_accept(const uint8_t *d_accepted, int32_t batch, double *d_dst, int64_t dst_stride, const double *d_src,
                   int64_t src_stride, int32_t n_q, int32_t n_var, int32_t n_pops, const xrsd_tables_t *tables,
                   const xrsd_tables_t *trial_tables, int32_t n_xtal, void *stream) {
  if (batch <= 0) return XRSD_OK;
  if (!d_accepted) return fail(XRSD_EINVAL, "NULL pointer");
  if ((d_dst == nullptr) != (d_src == nullptr)) return fail(XRSD_EINVAL, "slab source and destination go together");
  if (d_dst) {
    if (n_q < 1 || n_var < 1 || n_pops < 1 || n_pops > XRSD_MAX_POPS) return fail(XRSD_EINVAL, "n_q %d, n_var %d, n_pops %d", n_q, n_var, n_pops);
    if (dst_stride < (int64_t)(n_var + n_pops) * n_q || src_stride < (int64_t)(1 + n_pops) * n_q)
      return fail(XRSD_EINVAL, "slab strides too small for the slab layout");
  }
  if ((tables == nullptr) != (trial_tables == nullptr)) return fail(XRSD_EINVAL, "tables and trial tables go together");
  if (tables) {
    if (n_xtal < 1) return fail(XRSD_EINVAL, "n_xtal=%d", n_xtal);
    if (tables->cap_shell != trial_tables->cap_shell || tables->n_pairs_max != trial_tables->n_pairs_max)
      return fail(XRSD_EINVAL, "table sets of different capacities");
  }
  cudaError_t e = xrsd_launch_lm_accept(d_accepted, batch, d_dst, (long long)dst_stride, d_src, (long long)src_stride, n_q, n_var,
                                        n_pops, tables, trial_tables, n_xtal, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "lm_accept_kernel");
  return XRSD_OK;
}

int xrsd_lm_iteration(const xrsd_lm_state_t *st, int32_t base_ready, void *stream) {
  if (!st || !st->layout || !st->obj) return fail(XRSD_EINVAL, "NULL state / layout / objective");
  if (!st->obj->prepared) return fail(XRSD_EINVAL, "xrsd_lm_iteration needs a prepared objective");
  if (!st->d_ws_base || !st->d_ws_trial || !st->d_flags || !st->d_chi2_init) return fail(XRSD_EINVAL, "NULL workspace / flags");
  const int n_x = st->layout->n_xtal;
  if (n_x > 0 && (!st->tables || !st->trial_tables)) return fail(XRSD_EINVAL, "crystalline layout without tables");
  int rc = xrsd_jacobian_batch(st->layout, st->obj, st->d_q, st->n_q, st->d_params, st->ld_params, st->batch, st->tables, st->d_fo,
                               st->d_w, st->ld_obs, st->jac_idx, st->n_free, st->ties, st->tie_coef, st->n_tie, st->rel_step,
                               base_ready ? XRSD_JAC_BASE_READY : XRSD_JAC_FULL, st->d_JTJ, st->d_JTr, st->d_chi2, st->d_ws_base,
                               st->d_needjac, stream);
  if (rc) return rc;
  if (!base_ready) {
    cudaError_t e = cudaMemcpyAsync(st->d_chi2_init, st->d_chi2, (size_t)st->batch * 8, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "chi2_init copy");
  }
  rc = xrsd_lm_propose(st->d_params, st->ld_params, st->batch, st->free_idx, st->n_free, st->ties, st->tie_coef, st->n_tie, st->d_JTJ,
                       st->d_JTr, st->d_lambda, st->d_lo, st->d_hi, st->d_active, st->d_trial, st->max_rel_step, st->d_flags + 2, stream);
  if (rc) return rc;
  if (n_x > 0) {
    rc = xrsd_reflection_tables(st->layout, st->d_trial, st->ld_params, st->batch, st->trial_tables, st->loop_cells, st->d_active, stream);
    if (rc) return rc;
  }
  rc = xrsd_jacobian_batch(st->layout, st->obj, st->d_q, st->n_q, st->d_trial, st->ld_params, st->batch, st->trial_tables, nullptr,
                           nullptr, 0, nullptr, 0, st->ties, st->tie_coef, st->n_tie, st->rel_step, XRSD_JAC_PATTERNS_ONLY, nullptr,
                           nullptr, nullptr, st->d_ws_trial, st->d_active, stream);
  if (rc) return rc;
  const int64_t ld_trial = (int64_t)(1 + st->n_pops) * st->n_q;
  rc = xrsd_lm_trial_update(st->obj, st->d_q, st->n_q, (const double *)st->d_ws_trial, ld_trial, st->batch, st->d_fo, st->d_w, st->ld_obs,
                            st->d_params, st->ld_params, st->d_trial, st->d_chi2, st->d_chi2_trial, st->d_lambda, st->d_active,
                            st->d_accepted, st->d_needjac, st->d_reason, st->d_n_accept, n_x > 0 ? st->trial_tables : nullptr, n_x,
                            st->d_flags, 0, st->up, st->down, st->ftol, st->xtol, st->lambda_max, stream);
  if (rc) return rc;
  return xrsd_lm_accept(st->d_accepted, st->batch, (double *)st->d_ws_base, (int64_t)(st->n_var + st->n_pops) * st->n_q,
                        (const double *)st->d_ws_trial, ld_trial, st->n_q, st->n_var, st->n_pops, n_x > 0 ? st->tables : nullptr,
                        n_x > 0 ? st->trial_tables : nullptr, n_x, stream);
}


int xrsd_copy_rows_masked(double *d_dst, int64_t dst_stride, const double *d_src, int64_t src_stride, int64_t row_len,
                          int32_t n_rows, const uint8_t *d_mask, void *stream) {
  if (n_rows < 0 || row_len < 0) return fail(XRSD_EINVAL, "n_rows %d, row_len %lld", n_rows, (long long)row_len);
  if (n_rows == 0 || row_len == 0) return XRSD_OK;
  if (!d_dst || !d_src || !d_mask) return fail(XRSD_EINVAL, "NULL pointer");
  if (dst_stride < row_len) return fail(XRSD_EINVAL, "dst_stride %lld < row_len %lld: rows would overlap", (long long)dst_stride, (long long)row_len);
  cudaError_t e = xrsd_launch_copy_rows_masked(d_dst, (long long)dst_stride, d_src, (long long)src_stride, (long long)row_len,
                                               n_rows, d_mask, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "copy_rows_masked_kernel");
  return XRSD_OK;
}

int xrsd_peak_profile(int32_t kind, double hwhm_or_hwhm_g, double hwhm_l, const double *d_x, int32_t n, double *d_out,
                      void *stream) {
  if (kind < XRSD_PROFILE_VOIGT || kind > XRSD_PROFILE_LORENTZIAN) return fail(XRSD_EINVAL, "profile kind %d", kind);
  if (n > 0 && (!d_x || !d_out)) return fail(XRSD_EINVAL, "NULL pointer");
  cudaError_t e = xrsd_launch_profile(kind, hwhm_or_hwhm_g, hwhm_l, d_x, n, d_out, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "profile_kernel");
  return XRSD_OK;
}

int xrsd_feature_batch(const double *d_q, int32_t n_q, const double *d_I, int64_t ld_I, int32_t batch, double *d_features,
                       int64_t ld_features, void *stream) {
  if (batch < 0) return fail(XRSD_EINVAL, "batch %d", batch);
  if (batch == 0) return XRSD_OK;
  if (!d_q || !d_I || !d_features) return fail(XRSD_EINVAL, "NULL pointer");
  if (n_q < 8 || n_q > XRSD_FEATURE_MAX_Q_GLOBAL)
    return fail(XRSD_EINVAL, "n_q %d outside [8, %d]", n_q, XRSD_FEATURE_MAX_Q_GLOBAL);
  if (ld_I < n_q || ld_features < XRSD_N_FEATURES) return fail(XRSD_EINVAL, "leading dimension too small");
  cudaError_t e = xrsd_launch_features(d_q, n_q, d_I, ld_I, batch, d_features, ld_features, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "feature_kernel");
  return XRSD_OK;
}

int xrsd_probe_fp64_peak(double *tflops, double *elapsed_ms, void *stream) {
  double *d_out = nullptr;
  cudaError_t e = cudaMalloc(&d_out, 8);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
  cudaStream_t s = (cudaStream_t)stream;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int blocks = sms * 8, iters = 1 << 15;
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  xrsd_launch_fp64_probe(d_out, blocks, 1024, s);   // warm-up
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(a, s);
    xrsd_launch_fp64_probe(d_out, blocks, iters, s);
    cudaEventRecord(b, s);
    e = cudaEventSynchronize(b);
    if (e != cudaSuccess) break;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    best = std::min(best, ms);
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(d_out);
  if (e != cudaSuccess) return cuda_fail(e, "fp64 probe");
  const double flops = 2.0 * 8.0 * (double)iters * 256.0 * (double)blocks;
  if (tflops) *tflops = flops / (best * 1e-3) / 1e12;
  if (elapsed_ms) *elapsed_ms = best;
  return XRSD_OK;
}

// ------------------------------------------------------------------ host-buffer convenience path
namespace {
// One workspace per (calling thread, device): device scratch, a mapped pinned staging buffer and a non-blocking
// stream of its own, so that the threads of a caller evaluate side by side (no process-wide lock, nothing on the
// legacy default stream) and a process that drives several GPUs gets a workspace on each.  A thread's workspaces are
// released by xrsd_release_workspace() or when the thread ends.
constexpr int kMaxDevices = 32;
struct Workspace {
  int device = -1;
  cudaStream_t stream = nullptr;
  void *base = nullptr;
  size_t bytes = 0;
  double *pinned = nullptr;
  double *pinned_dev = nullptr;   // the same buffer as the device addresses it (mapped pinned memory), or nullptr
  size_t pinned_bytes = 0;
  std::vector<double> q_host;     // the grid currently resident at offset 0 of the workspace (skips its upload)
  bool q_valid = false;
  void release() {
    if (device < 0) return;
    int cur = -1;
    const bool switched = cudaGetDevice(&cur) == cudaSuccess && cur != device && cudaSetDevice(device) == cudaSuccess;
    q_valid = false;
    if (pinned) cudaFreeHost(pinned);
    pinned = nullptr;
    pinned_dev = nullptr;
    pinned_bytes = 0;
    if (base) cudaFree(base);
    base = nullptr;
    bytes = 0;
    if (stream) cudaStreamDestroy(stream);
    stream = nullptr;
    if (switched) cudaSetDevice(cur);
    cudaGetLastError();           // at process exit the runtime may already be unloading: nothing to report
    device = -1;
  }
  ~Workspace() { release(); }
};
struct ThreadWorkspaces { Workspace dev[kMaxDevices]; };
thread_local ThreadWorkspaces t_ws;
thread_local Workspace *ws_cur = nullptr;

int ws_select() {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  if (dev < 0 || dev >= kMaxDevices) return fail(XRSD_EINVAL, "device ordinal %d beyond %d", dev, kMaxDevices - 1);
  Workspace &w = t_ws.dev[dev];
  if (w.device != dev) {
    w.release();
    e = cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { w.stream = nullptr; return cuda_fail(e, "cudaStreamCreate"); }
    w.device = dev;
  }
  ws_cur = &w;
  return XRSD_OK;
}

int ws_reserve(size_t bytes) {
  if (bytes <= ws_cur->bytes) return XRSD_OK;
  if (ws_cur->base) cudaFree(ws_cur->base);
  ws_cur->base = nullptr;
  ws_cur->bytes = 0;
  ws_cur->q_valid = false;
  cudaError_t e = cudaMalloc(&ws_cur->base, bytes);
  if (e != cudaSuccess) return fail(XRSD_ENOMEM, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
  ws_cur->bytes = bytes;
  return XRSD_OK;
}
size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// small calls stage their parameters and their result through one pinned buffer (pageable copies of a few KB cost
// ~10 us each in the driver); returns nullptr when the call is too large for it or the allocation fails
constexpr size_t kPinnedMax = 4u << 20;
double *pinned_reserve(size_t bytes) {
  if (bytes > kPinnedMax) return nullptr;
  if (bytes <= ws_cur->pinned_bytes) return ws_cur->pinned;
  if (ws_cur->pinned) cudaFreeHost(ws_cur->pinned);
  ws_cur->pinned = nullptr;
  ws_cur->pinned_bytes = 0;
  const size_t want = std::max(bytes, (size_t)256 << 10);
  ws_cur->pinned_dev = nullptr;
  if (cudaHostAlloc((void **)&ws_cur->pinned, want, cudaHostAllocMapped) != cudaSuccess) { cudaGetLastError(); ws_cur->pinned = nullptr; return nullptr; }
  ws_cur->pinned_bytes = want;
  if (cudaHostGetDevicePointer((void **)&ws_cur->pinned_dev, ws_cur->pinned, 0) != cudaSuccess) { cudaGetLastError(); ws_cur->pinned_dev = nullptr; }
  return ws_cur->pinned;
}
// Small non-crystalline calls skip both copies: the kernel reads the parameter rows from the mapped pinned buffer
// and writes I(q) straight into it (staged mode writes every value once, coalesced: posted PCIe writes), so the
// call is launch + kernel + one synchronisation.  Crystalline layouts accumulate in place in the output row
// (direct mode) and keep the device buffer.  XRSD_HOST_ZEROCOPY=0 turns it off (A/B timing).
constexpr size_t kZeroCopyMax = 1u << 20;
bool zero_copy_enabled() {
  static const bool on = [] { const char *v = getenv("XRSD_HOST_ZEROCOPY"); return !(v && v[0] == '0'); }();
  return on;
}
}  // namespace

double xrsd_uniform_step(const double *q, int32_t n) {
  // step of a linspace-like grid: every point within 4 ulp of q[0] + i*step and strictly ascending; else 0
  if (!q || n < 3) return 0.0;
  const double step = (q[n - 1] - q[0]) / (double)(n - 1);
  if (!(step > 0.0)) return 0.0;
  const double tol = 4.0 * 2.220446049250313e-16 * std::max(fabs(q[0]), fabs(q[n - 1]));
  // counted, not branched: without an early exit the host compiler vectorises both passes (this call sits on the
  // single-pattern latency path: 4.5 us -> well under 1 us for 2000 points)
  const double q0 = q[0];
  int bad = 0;
  for (int32_t i = 0; i < n; ++i) bad += !(fabs(q[i] - (q0 + (double)i * step)) <= tol);
  for (int32_t i = 1; i < n; ++i) bad += !(q[i] > q[i - 1]);
  return bad ? 0.0 : step;
}

void xrsd_release_workspace(void) {
  for (int d = 0; d < kMaxDevices; ++d) t_ws.dev[d].release();      // the calling thread's workspaces, every device
  ws_cur = nullptr;
}

int xrsd_intensity_batch_host(const xrsd_layout_t *layout, const double *q, int32_t n_q, const double *params,
                              int32_t ld_params, int32_t batch, double *I_out, int64_t ld_I) {
  int rc = check_layout(layout);
  if (rc) return rc;
  if (batch <= 0 || n_q <= 0) return XRSD_OK;
  if (!q || !params || !I_out) return fail(XRSD_EINVAL, "NULL host pointer");
  if (ld_I < n_q) return fail(XRSD_EINVAL, "ld_I < n_q");
  rc = ws_select();
  if (rc) return rc;
  const int n_x = layout->n_xtal;
  const size_t n_tab = (size_t)batch * n_x;
  const int cap_shell = 1024, n_pairs = XRSD_MAX_PAIRS;
  const int max_cells = n_x ? host_max_cells(layout, params, ld_params, batch) : 1;
  size_t off = 0;
  const size_t o_q = off; off += align256((size_t)n_q * 8);
  const size_t o_p = off; off += align256((size_t)batch * ld_params * 8);
  const size_t o_I = off; off += align256((size_t)batch * n_q * 8);
  const size_t o_ns = off; off += align256(n_tab * 4);
  const size_t o_nr = off; off += align256(n_tab * 4);
  const size_t o_na = off; off += align256(n_tab * 4);
  const size_t o_st = off; off += align256(n_tab * 4);
  const size_t o_sh = off; off += align256(n_tab * cap_shell * 4 * 4);
  const size_t o_sg = off; off += align256(n_tab * cap_shell * n_pairs * 8);
  rc = ws_reserve(off);
  if (rc) return rc;
  char *base = (char *)ws_cur->base;
  cudaStream_t s = ws_cur->stream;
  // repeated calls on the same grid (the usual case: one measured pattern, many evaluations) skip the upload of q
  cudaError_t e = cudaSuccess;
  const bool same_q = ws_cur->q_valid && ws_cur->q_host.size() == (size_t)n_q && memcmp(ws_cur->q_host.data(), q, (size_t)n_q * 8) == 0;
  if (!same_q) {
    e = cudaMemcpyAsync(base + o_q, q, (size_t)n_q * 8, cudaMemcpyHostToDevice, s);
    ws_cur->q_host.assign(q, q + n_q);
    ws_cur->q_valid = (e == cudaSuccess);
  }
  const size_t p_bytes = (size_t)batch * ld_params * 8, I_bytes = (size_t)batch * n_q * 8;
  double *pin = pinned_reserve(align256(p_bytes) + I_bytes);
  if (pin) memcpy(pin, params, p_bytes);
  const bool zero_copy = pin && ws_cur->pinned_dev && n_x == 0 && I_bytes <= kZeroCopyMax && zero_copy_enabled();
  if (e == cudaSuccess && !zero_copy)
    e = cudaMemcpyAsync(base + o_p, pin ? (const void *)pin : (const void *)params, p_bytes, cudaMemcpyHostToDevice, s);
  if (e != cudaSuccess) { ws_cur->q_valid = false; return cuda_fail(e, "H2D copy"); }
  const double *d_par = zero_copy ? ws_cur->pinned_dev : (const double *)(base + o_p);
  double *d_out = zero_copy ? ws_cur->pinned_dev + align256(p_bytes) / 8 : (double *)(base + o_I);
  xrsd_tables_t T = {};
  if (n_x) {
    T.cap_shell = cap_shell; T.cap_refl = 0; T.n_pairs_max = n_pairs;
    T.n_shell = (int32_t *)(base + o_ns); T.n_refl = (int32_t *)(base + o_nr); T.n_all = (int32_t *)(base + o_na);
    T.status = (int32_t *)(base + o_st); T.shell_hkl = (int32_t *)(base + o_sh); T.shell_geo = (double *)(base + o_sg);
  }
  // tables -> intensity -> copies are queued back to back; the table status comes back with the result and is looked
  // at after the one synchronisation (a capacity overflow, rare, repeats the sequence with global table scratch)
  void *big = nullptr;
  std::vector<int32_t> st(n_tab), na(n_tab);
  for (int attempt = 0; attempt < 2; ++attempt) {
    if (n_x) {
      rc = xrsd_reflection_tables(layout, (const double *)(base + o_p), ld_params, batch, &T, max_cells, nullptr, s);
      if (rc) break;
    }
    rc = xrsd_intensity_batch(layout, (const double *)(base + o_q), n_q, d_par, ld_params, batch, n_x ? &T : nullptr, d_out,
                              n_q, s);
    if (rc) break;
    if (n_x) {
      e = cudaMemcpyAsync(st.data(), T.status, n_tab * 4, cudaMemcpyDeviceToHost, s);
      if (e == cudaSuccess) e = cudaMemcpyAsync(na.data(), T.n_all, n_tab * 4, cudaMemcpyDeviceToHost, s);
      if (e != cudaSuccess) { rc = cuda_fail(e, "table status copy"); break; }
    }
    if (zero_copy) {
      e = cudaStreamSynchronize(s);       // the kernel has written the pinned result itself
    } else if (pin) {
      double *pin_I = pin + align256(p_bytes) / 8;
      e = cudaMemcpyAsync(pin_I, base + o_I, I_bytes, cudaMemcpyDeviceToHost, s);
      if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    } else {
      if (ld_I == n_q)
        e = cudaMemcpyAsync(I_out, base + o_I, I_bytes, cudaMemcpyDeviceToHost, s);
      else
        e = cudaMemcpy2DAsync(I_out, (size_t)ld_I * 8, base + o_I, (size_t)n_q * 8, (size_t)n_q * 8, batch, cudaMemcpyDeviceToHost, s);
      if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    }
    if (e != cudaSuccess) { rc = cuda_fail(e, "D2H copy / kernel execution"); break; }
    int worst = 0, n_max = 0;
    for (size_t i = 0; i < n_tab; ++i) { worst |= st[i]; n_max = std::max(n_max, na[i]); }
    if (worst == 0) {
      if (pin) {
        const double *pin_I = pin + align256(p_bytes) / 8;
        for (int32_t b = 0; b < batch; ++b) memcpy(I_out + (size_t)b * ld_I, pin_I + (size_t)b * n_q, (size_t)n_q * 8);
      }
      break;
    }
    if (worst == XRSD_TABLE_LIST_OVERFLOW && attempt == 0) {
      // large cell without usable symmetry: give the builder global scratch for its sort / shell arrays
      int64_t cap_big = 8192;
      while (cap_big < n_max) cap_big <<= 1;
      e = cudaMalloc(&big, (size_t)(n_tab * xrsd_table_scratch_bytes(cap_big)));
      if (e != cudaSuccess) { rc = fail(XRSD_ENOMEM, "cudaMalloc(table scratch): %s", cudaGetErrorString(e)); break; }
      T.big_scratch = big;
      T.cap_big = cap_big;
      T.cap_list_hint = 4096;
      continue;
    }
    rc = fail(XRSD_ECAPACITY, "reflection table: status %d (capacity exceeded)", worst);
    break;
  }
  if (big) { cudaStreamSynchronize(s); cudaFree(big); }
  if (rc) return rc;
  return XRSD_OK;
}

}  // extern "C"
