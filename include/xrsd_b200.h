/* xrsd_b200.h -- C ABI of the B200-native scattering-intensity engine.
 *
 * Drop-in boundary for xrsdkit's intensity path.  The reference has no FFI; its
 * narrowest seams are Python functions.  Each entry point below names the reference
 * interface it replaces (paths relative to the reference root):
 *
 *   xrsd_reflection_tables   <- scattering/__init__.py:240-297 (hkl enumeration, |q_hkl|)
 *                               + scattering/symmetries.py:44-92 (symmetrize_points)
 *                               + the geometric part of scattering/__init__.py:307-327
 *   xrsd_intensity_batch     <- system/__init__.py:112-130 (System.compute_intensity)
 *                               = system/noise.py:50-63 + scattering/__init__.py:12-66,
 *                               69-161, 299-333 + structure_factors.py:3-47
 *                               + form_factors.py:8-28 + tools/peak_math.py:24-47
 *   xrsd_residual_batch      <- system/__init__.py:134-187 (System.evaluate_residual)
 *                               + tools/__init__.py:104-125 (compute_chi2)
 *   xrsd_jacobian_batch /
 *   xrsd_lm_*                <- the per-evaluation work of system/__init__.py:220-278 (fit);
 *                               the reference drives Nelder-Mead through lmfit, these
 *                               provide the batched Levenberg-Marquardt replacement
 *   xrsd_*_host              <- the same calls for callers that own HOST buffers
 *                               (copies happen inside)
 *
 * Conventions: plain pointers and sizes, no C++/torch types.  "d_" arguments are device
 * pointers on the current device; `stream` is a cudaStream_t passed as void* (NULL =
 * default stream).  All floating point is IEEE double.  Functions return 0 on success
 * or a negative XRSD_E* code; xrsd_last_error() gives the message for the calling
 * thread.  Numerical degeneracies are not errors: NaN/inf propagate exactly as in the
 * reference (e.g. underflowing Gaussian normalisation).
 */
#ifndef XRSD_B200_H
#define XRSD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XRSD_ABI_VERSION 5

#define XRSD_MAX_POPS 6      /* noise model + up to 5 populations per system layout   */
#define XRSD_MAX_SPECIES 4   /* species (atoms) per crystalline population             */
#define XRSD_MAX_PAIRS 10    /* XRSD_MAX_SPECIES*(XRSD_MAX_SPECIES+1)/2                */
#define XRSD_MAX_OPS 9       /* symmetry operations per point group (m-3m has 9)       */
#define XRSD_MAX_PARAMS 64   /* flat parameter slots per system                        */
#define XRSD_MAX_FREE 16     /* free (fitted) parameters per system                    */
#define XRSD_MAX_TIES 48     /* constraint terms (linear terms and expression-program words) per system */
/* Constraint terms as the `ties` (int32 pairs) / `tie_coef` (double pairs) arguments carry them, applied in order:
 *   (dst, src, scale, offset), dst >= 0, src >= 0      p[dst]  = scale * p[src] + offset
 *   (-(dst + 1), src, scale, offset)                   p[dst] += scale * p[src] + offset      (further terms of a sum)
 *   (dst, -1, scale, offset)                           p[dst]  = scale * (top of the expression stack) + offset; stack cleared
 *   (-(XRSD_TIE_WORD + op), arg, value, 0)             a word of a postfix expression program over a stack of
 *                                                      XRSD_TIE_STACK values: PUSHC pushes `value`, PUSHP pushes p[arg], the
 *                                                      binary words replace the two top values (a below b) by a (op) b,
 *                                                      the unary ones the top value.  <- lmfit / asteval constraint_expr */
#define XRSD_TIE_WORD 1000
#define XRSD_TIE_OP0 10      /* internal: program words are kept as XRSD_TIE_OP0 + op */
#define XRSD_TIE_STACK 8
enum { XRSD_TIE_PUSHC = 0, XRSD_TIE_PUSHP = 1, XRSD_TIE_ADD = 2, XRSD_TIE_SUB = 3, XRSD_TIE_MUL = 4, XRSD_TIE_DIV = 5,
       XRSD_TIE_POW = 6, XRSD_TIE_MIN = 7, XRSD_TIE_MAX = 8, XRSD_TIE_NEG = 9, XRSD_TIE_SQRT = 10, XRSD_TIE_EXP = 11,
       XRSD_TIE_LOG = 12, XRSD_TIE_LOG10 = 13, XRSD_TIE_SIN = 14, XRSD_TIE_COS = 15, XRSD_TIE_TAN = 16, XRSD_TIE_ABS = 17,
       XRSD_TIE_ASIN = 18, XRSD_TIE_ACOS = 19, XRSD_TIE_ATAN = 20 };

/* xrsd_pop_t.kind */
enum { XRSD_POP_NONE = 0, XRSD_POP_NOISE_FLAT = 1, XRSD_POP_NOISE_LOW_Q = 2,
       XRSD_POP_NONCRYSTALLINE = 3, XRSD_POP_CRYSTALLINE = 4 };
/* xrsd_pop_t.form (non-crystalline) and species_form[] (crystalline) */
enum { XRSD_FORM_ATOMIC = 0, XRSD_FORM_SPHERE = 1, XRSD_FORM_SPHERE_R_NORMAL = 2, XRSD_FORM_GUINIER_POROD = 3 };
/* xrsd_pop_t.profile */
enum { XRSD_PROFILE_VOIGT = 0, XRSD_PROFILE_GAUSSIAN = 1, XRSD_PROFILE_LORENTZIAN = 2 };
/* xrsd_pop_t.lattice: geometry classes of definitions.py:504-551 */
enum { XRSD_LAT_CUBIC = 0, XRSD_LAT_HCP = 1, XRSD_LAT_HEXAGONAL = 2, XRSD_LAT_ORTHORHOMBIC = 3,
       XRSD_LAT_TETRAGONAL = 4, XRSD_LAT_MONOCLINIC = 5, XRSD_LAT_RHOMBOHEDRAL = 6, XRSD_LAT_TRICLINIC = 7 };
/* symmetry operation ids (scattering/symmetries.py:10-21) */
enum { XRSD_OP_INVERSION = 0, XRSD_OP_MIRROR_X = 1, XRSD_OP_MIRROR_Y = 2, XRSD_OP_MIRROR_Z = 3,
       XRSD_OP_MIRROR_X_Y = 4, XRSD_OP_MIRROR_Y_Z = 5, XRSD_OP_MIRROR_Z_X = 6,
       XRSD_OP_MIRROR_NX_Y = 7, XRSD_OP_MIRROR_NY_Z = 8, XRSD_OP_MIRROR_NZ_X = 9 };

/* error codes */
enum { XRSD_OK = 0, XRSD_EINVAL = -1, XRSD_ECUDA = -2, XRSD_ECAPACITY = -3, XRSD_ENOMEM = -4 };

/* per-table status bits written by xrsd_reflection_tables */
enum { XRSD_TABLE_OK = 0, XRSD_TABLE_GRID_OVERFLOW = 1, XRSD_TABLE_LIST_OVERFLOW = 2,
       XRSD_TABLE_SHELL_OVERFLOW = 4, XRSD_TABLE_BAD_LATTICE = 8 };

/* One population (or the noise model) of a system layout.  p_* are slots into the
 * flat per-system parameter vector (-1 = not used).  Plain old data, 8-byte aligned. */
typedef struct xrsd_pop {
  int32_t kind;
  int32_t disordered;          /* 1: multiply by the hard-sphere S(q) (structure 'disordered') */
  int32_t form;
  int32_t profile;
  int32_t sf_radial;           /* structure_factor_mode == 'radial' */
  int32_t polz, lorentz, use_symmetry;
  int32_t lattice;
  int32_t n_sites;             /* basis sites of the conventional cell (definitions.py:462-502) */
  int32_t n_ops;
  int32_t ops[XRSD_MAX_OPS];
  int32_t n_species;
  int32_t species_form[XRSD_MAX_SPECIES];
  int32_t p_I0, p_r, p_sigma, p_rg, p_D, p_r_hard, p_v_fraction, p_flat_fraction;
  int32_t p_lat[6];            /* a, b, c, alpha, beta, gamma */
  int32_t p_hwhm, p_hwhm_g, p_hwhm_l;
  int32_t p_uvw[XRSD_MAX_SPECIES][3];
  int32_t table_slot;          /* index among the layout's crystalline populations */
  int32_t prune_extinct;       /* drop shells whose geometric weight is < 1e-20 of the largest */
  int32_t reserved0;
  double sampling_width, sampling_step;
  double q_min, q_max;         /* q_max already resolved (falsy -> q[-1], scattering/__init__.py:226) */
  double sites[8][3];
  double atom_Z[XRSD_MAX_SPECIES];
  double atom_a[XRSD_MAX_SPECIES][4];
  double atom_b[XRSD_MAX_SPECIES][4];
} xrsd_pop_t;

typedef struct xrsd_layout {
  int32_t n_pops;              /* pops[0] is always the noise model */
  int32_t n_params;
  int32_t n_xtal;              /* number of crystalline populations (= table slots) */
  int32_t q0_is_zero;          /* q[0] == 0: the only zero the reference special-cases */
  double wavelength;
  double q_step;               /* > 0: q[i] == q[0] + i*q_step to within a few ulp (a linspace grid), which lets the
                                  size-quadrature loop advance its starting phases by a rotation from one pass over
                                  the CTA's threads (256 q-points) to the next instead of calling sincos, and tells the crystalline
                                  branch that q is strictly ascending without testing it; 0: no assumption about q */
  xrsd_pop_t pops[XRSD_MAX_POPS];
} xrsd_layout_t;

/* Reflection-table storage.  A batch has n_tables = batch * layout.n_xtal tables, table index
 * t = b * n_xtal + slot; table t lives in storage ROW table_of[t] of the arrays below (row = own_base + t when
 * table_of is NULL).  All arrays are device memory owned by the caller.
 *
 * Shared rows.  For a cubic cell without free atomic coordinates everything in a table -- the reduced hkl list,
 * the multiplicities, the shells and their geometric weights -- depends on the cell edge only through WHICH
 * integer shells N = h^2+k^2+l^2 pass the reference's  G_min < |G| <= G_max  test (scattering/__init__.py:276),
 * so systems that agree on that set share one table: with n_shared > 0 the builder first derives each system's
 * key (the largest passing N, decided with a 1e-13 guard band around the float comparison; undecidable, keyed
 * beyond n_shared, or non-cubic systems keep their own row), claims row shared_base + key through shared_tag
 * (0 = never built) and builds every distinct table ONCE; the other systems only get the row index.  Rows
 * already built stay valid between calls as long as shared_tag is not cleared (the fit loop keeps them for the
 * whole fit).  Several table sets (e.g. the accepted and the trial point of a fit) may be views of the same
 * arrays with different own_base / table_of. */
typedef struct xrsd_tables {
  int32_t cap_shell;           /* shells per table the arrays below can hold            */
  int32_t cap_refl;            /* reduced reflections per table (0: do not export)      */
  int32_t n_pairs_max;         /* leading dimension of shell_geo (<= XRSD_MAX_PAIRS)    */
  int32_t cap_list_hint;       /* reduced reflections the builder keeps in shared memory (0 = auto: 1024 or
                                  4096; XRSD_TABLE_LIST_OVERFLOW asks for a retry with a larger value) */
  int32_t *n_shell;            /* [n_rows]                                               */
  int32_t *n_refl;             /* [n_rows] reduced reflections (before shell merging)    */
  int32_t *n_all;              /* [n_rows] reflections in the G_min..G_max shell         */
  int32_t *status;             /* [n_rows] XRSD_TABLE_* bits                             */
  int32_t *shell_hkl;          /* [n_rows][cap_shell][4]: representative h,k,l, sum of multiplicities */
  double *shell_geo;           /* [n_rows][cap_shell][n_pairs_max]: sum mult*Re(G_a conj G_b)          */
  int32_t *refl_hkl;           /* [n_rows][cap_refl][4]: h,k,l,mult in reference order (optional)       */
  void *big_scratch;           /* optional: [n_tables] regions of xrsd_table_scratch_bytes(cap_big) bytes in global
                                  memory, used by a table whose reduced list exceeds the shared-memory capacity
                                  (large cells without usable symmetry); cap_big is a power of two              */
  int64_t cap_big;
  int32_t *table_of;           /* [n_tables] storage row of each table, written by the builder; NULL: row = own_base + t */
  uint8_t *build_flag;         /* [n_tables] builder scratch (1 = this call builds the table's row); needed with table_of */
  int32_t own_base;            /* first of this set's own rows */
  int32_t shared_base;         /* first shared row */
  int32_t n_shared;            /* shared rows (0: nothing is shared) */
  int32_t n_grid_slots;        /* regions in grid_scratch */
  int32_t *shared_tag;         /* [n_shared] 0 = row not built; else 1 + the G_min key it was built for */
  void *grid_scratch;          /* optional: n_grid_slots regions of 2*cap_grid_cells bytes (global memory) for candidate
                                  boxes beyond the shared-memory budget (`max_cells` of xrsd_reflection_tables); a table
                                  that finds neither room is flagged XRSD_TABLE_GRID_OVERFLOW */
  int64_t cap_grid_cells;
  int32_t *grid_slot_counter;  /* one device int (set to zero by the call) handing out the regions */
} xrsd_tables_t;

/* objective flags of xrsd_residual_batch (System.fit_report, system/__init__.py:23-28) */
typedef struct xrsd_objective {
  int32_t error_weighted;
  int32_t logI_weighted;
  int32_t has_dI;              /* 0: dI defaults to sqrt(I) on the fit mask (system/__init__.py:168-171) */
  int32_t prepared;            /* 1: the d_Iobs / d_dI arguments are the streams written by xrsd_objective_prepare
                                  (f(I_obs) and the point weights, -1 on masked points) instead of raw I_obs / dI */
  double q_lo, q_hi;           /* q_range */
} xrsd_objective_t;

const char *xrsd_last_error(void);
int xrsd_abi_version(void);
/* sizeof() of the structs above as compiled, so a binding can verify its mirror */
int xrsd_struct_sizes(int32_t *sizeof_pop, int32_t *sizeof_layout, int32_t *sizeof_tables, int32_t *sizeof_objective);

/* Bytes of one table's region in xrsd_tables_t.big_scratch for a reduced list of up to cap_big reflections. */
int64_t xrsd_table_scratch_bytes(int64_t cap_big);

/* Dynamic shared memory the table builder needs for the largest candidate box of a
 * batch, given the per-axis half-widths n1,n2,n3 (scattering/__init__.py:266-268). */
int64_t xrsd_table_smem_bytes(int32_t n1, int32_t n2, int32_t n3);

/* Build reflection tables for every crystalline population of every system of the batch.
 * d_active (optional, [batch] bytes): systems with a zero byte are skipped -- used by the fit loop
 * for converged systems; the same convention applies to the residual and Jacobian calls below. */
int xrsd_reflection_tables(const xrsd_layout_t *layout, const double *d_params, int32_t ld_params,
                           int32_t batch, const xrsd_tables_t *tables, int32_t max_cells,
                           const uint8_t *d_active, void *stream);

/* I[b, :] for b < batch.  d_q[n_q], d_params[batch][ld_params], d_I[batch][ld_I] (device memory; every element of
 * the rows is overwritten -- layouts with crystalline populations accumulate the populations in place in d_I,
 * the others stage a span in shared memory and write it once).  <- System.compute_intensity, system/__init__.py:112-130 */
int xrsd_intensity_batch(const xrsd_layout_t *layout, const double *d_q, int32_t n_q,
                         const double *d_params, int32_t ld_params, int32_t batch,
                         const xrsd_tables_t *tables, double *d_I, int64_t ld_I, void *stream);

/* The optional single-precision path (BASELINE.json north_star: "<= 1e-5 relative error in the optional FP32 path").
 * Same arguments and FP64 output as xrsd_intensity_batch, for layouts WITHOUT crystalline populations (others:
 * XRSD_EINVAL).  The O(n_q n_r) loop over a size distribution's radii runs in FP32 wherever q r >= 2 for a whole
 * warp; everything done once per q-point (starting phases, small-argument radii, normalisation, hard-sphere factor,
 * the sum over populations) stays FP64.  The reference has no such path; it is validated against
 * xrsd_intensity_batch (tests/test_gpu_parity.py::test_single_precision_path). */
int xrsd_intensity_batch_f32(const xrsd_layout_t *layout, const double *d_q, int32_t n_q,
                             const double *d_params, int32_t ld_params, int32_t batch,
                             double *d_I, int64_t ld_I, void *stream);

/* Same launch, additionally writing the running totals d_cum[batch][layout.n_pops][n_q] after the
 * noise model (index 0) and after each population: differences of consecutive rows are the
 * per-population patterns.  <- the (2 + n_pop) separate evaluations of visualization/__init__.py:23-31
 * (plot decomposition) and the I0 rescaling of models/predict.py:372-378, in one pass. */
int xrsd_intensity_components_batch(const xrsd_layout_t *layout, const double *d_q, int32_t n_q,
                                    const double *d_params, int32_t ld_params, int32_t batch,
                                    const xrsd_tables_t *tables, double *d_I, int64_t ld_I, double *d_cum,
                                    void *stream);

/* chi2[b] of System.evaluate_residual for b < batch; d_Iobs / d_dI are [batch][ld_obs]
 * (ld_obs == 0 broadcasts one pattern to the whole batch).  I(q) is computed in the same
 * kernel (grid = batch x q-spans) and never written to memory; the spans of a system leave their
 * partial sums in d_workspace (xrsd_residual_workspace_bytes(batch, n_q) bytes; its ticket area must be
 * zero before the first call -- zeroing the whole workspace once is the simple way -- and is left
 * zero by the kernel) and the span that arrives last adds them up in span order, so the result
 * is bit-identical from run to run. */
int64_t xrsd_residual_workspace_bytes(int32_t batch, int32_t n_q);
int xrsd_residual_batch(const xrsd_layout_t *layout, const xrsd_objective_t *obj,
                        const double *d_q, int32_t n_q, const double *d_params, int32_t ld_params,
                        int32_t batch, const xrsd_tables_t *tables,
                        const double *d_Iobs, const double *d_dI, int64_t ld_obs,
                        double *d_chi2, void *d_workspace, const uint8_t *d_active, void *stream);

/* Forward-difference Jacobian of the weighted residual vector, reduced in-kernel to the
 * normal equations: JTJ[b][n_free][n_free], JTr[b][n_free], chi2[b].  free_idx[n_free] (HOST
 * array) are parameter slots (an entry -(slot+1) forces a finite-difference variant for an I0, whose derivative is
 * otherwise taken from the population's own pattern: needed where a system sits at I0 == 0); rel_step is the relative step (absolute for zero-valued
 * parameters).  Two launches: variants (grid = batch x evaluated variants x q-spans; an I0 is
 * linear and needs no variant) and a chunked reduction.  d_workspace holds
 * xrsd_jacobian_workspace_bytes() bytes; the partial-sum / ticket area behind the variant slabs
 * must be zero before the FIRST call with a given (n_q, batch, n_free) (the reduction leaves the
 * tickets zero; zeroing the whole workspace once is the simple way).  The q-chunks of a system are
 * added up in chunk order by the chunk that arrives last: JTJ, JTr and chi2 are bit-identical from
 * run to run.  ties[n_tie][2] (HOST, may be NULL) are
 * constraints (dst slot, src slot) with tie_coef[n_tie][2] = (scale, offset) (HOST; NULL = identity):
 * params[dst] = scale * params[src] + offset -- constraint_expr values that are affine in one other
 * parameter (xrsdkit's own fitted systems use the plain name, "particle__r"); a dst given as -(slot + 1)
 * ADDS its term to the slot instead (params[slot] += scale * params[src] + offset), so a linear combination of
 * several parameters is a plain term followed by accumulating ones; terms are applied in order and the tied
 * slot follows its sources in every variant.  Reflection tables
 * and the radius counts of size quadratures are held fixed inside one evaluation. */
int64_t xrsd_jacobian_workspace_bytes(int32_t n_q, int32_t batch, int32_t n_free);
int xrsd_jacobian_batch(const xrsd_layout_t *layout, const xrsd_objective_t *obj,
                        const double *d_q, int32_t n_q, const double *d_params, int32_t ld_params,
                        int32_t batch, const xrsd_tables_t *tables,
                        const double *d_Iobs, const double *d_dI, int64_t ld_obs,
                        const int32_t *free_idx, int32_t n_free, const int32_t *ties, const double *tie_coef,
                        int32_t n_tie, double rel_step, int32_t mode, double *d_JTJ, double *d_JTr, double *d_chi2,
                        void *d_workspace, const uint8_t *d_active, void *stream);
/* Number of pattern slabs per system the call above evaluates and keeps at the head of d_workspace:
 * [batch][n_variants + layout.n_pops][n_q] doubles -- slab 0 is the base pattern, slabs 1..n_variants-1 the
 * perturbed ones (one per non-linear free parameter), the last n_pops the base pattern's running totals after
 * each population.  `mode` of xrsd_jacobian_batch:
 *   XRSD_JAC_FULL           evaluate every slab, then reduce;
 *   XRSD_JAC_BASE_READY     trust slab 0 and the running totals already in the workspace (e.g. copied there from
 *                           an accepted trial evaluation), evaluate only the perturbed variants, then reduce;
 *   XRSD_JAC_PATTERNS_ONLY  evaluate every slab and stop: no reduction, d_JTJ / d_JTr / d_chi2 / d_Iobs / d_dI are
 *                           not touched (may be NULL).  With a single linear free parameter this is "pattern and
 *                           running totals of every system, honouring d_active". */
#define XRSD_JAC_FULL 0
#define XRSD_JAC_BASE_READY 1
#define XRSD_JAC_PATTERNS_ONLY 2
int32_t xrsd_jacobian_n_variants(const xrsd_layout_t *layout, const int32_t *free_idx, int32_t n_free,
                                 const int32_t *ties, int32_t n_tie);

/* One damped Gauss-Newton (Levenberg-Marquardt) proposal per system:
 * solve (JTJ + lambda*diag(JTJ)) delta = -JTr, limit every component to max_rel_step * |parameter| (a relative
 * trust region; 0 = no limit), project onto [lo, hi], write trial params.  d_iter (optional): a device counter the launch
 * increments -- the number of the iteration it opens, which xrsd_lm_trial_update(tag = 0) reads from d_flags[2]; with it
 * the launches of an iteration carry no host-side iteration number and can be replayed as a CUDA graph. */
int xrsd_lm_propose(const double *d_params, int32_t ld_params, int32_t batch,
                    const int32_t *free_idx, int32_t n_free, const int32_t *ties, const double *tie_coef,
                    int32_t n_tie, const double *d_JTJ, const double *d_JTr, const double *d_lambda,
                    const double *d_lo, const double *d_hi, const uint8_t *d_active,
                    double *d_trial, double max_rel_step, int32_t *d_iter, void *stream);
/* Accept/reject: where chi2_trial < chi2 copy trial -> params, lambda /= down, else lambda *= up;
 * clears `active` for systems whose relative improvement fell below ftol or that hit lambda_max. */
int xrsd_lm_update(double *d_params, int32_t ld_params, int32_t batch,
                   const double *d_trial, double *d_chi2, const double *d_chi2_trial,
                   double *d_lambda, uint8_t *d_active, int32_t *d_n_accept,
                   double up, double down, double ftol, double lambda_max, void *stream);

/* Once per fit: d_fo[r][i] = f(I_obs) (log in log mode) and d_w[r][i] = the point's weight of the reference's objective
 * (system/__init__.py:164-172), -1 where the point is outside the fit mask (I_obs <= 0 or q outside q_range), for
 * r < rows (rows = 1 for one shared pattern).  A copy of `obj` with prepared = 1 then makes xrsd_residual_batch,
 * xrsd_jacobian_batch, xrsd_pattern_chi2_batch and xrsd_lm_trial_update read these two streams in place of d_Iobs /
 * d_dI (same leading dimension for both), which saves a logarithm, a square root and the mask tests per point and call. */
int xrsd_objective_prepare(const xrsd_objective_t *obj, const double *d_q, int32_t n_q, const double *d_Iobs, const double *d_dI,
                           int64_t ld_obs, int32_t rows, double *d_fo, double *d_w, int64_t ld_out, void *stream);

/* One launch for the trial point of an LM iteration: chi2_trial[b] of the stored trial patterns d_I[batch][ld_I] (NULL:
 * take the values already in d_chi2_trial), then accept / reject per system -- where chi2_trial < chi2 the trial row
 * replaces the parameter row, chi2 and lambda /= down, else lambda *= up.  d_accepted[b] says which, d_needjac[b] which
 * systems need a new Jacobian (accepted and still running), d_reason[b] why a system left the loop: 1 = converged
 * (chi2 changed by <= ftol relative, or every parameter by <= xtol relative), 2 = lambda beyond lambda_max, 3 = objective
 * not finite; d_active[b] is cleared with it.  A trial point whose reflection tables (trial_tables, optional) carry a
 * non-zero status counts as chi2 = inf and ORs that status into d_flags[0]; a system that is still running after this
 * update raises d_flags[1] to `tag` (atomic max; tag = 0: to d_flags[2], the counter xrsd_lm_propose keeps), so a caller
 * reads "anything left to do?" from these words without a reduction launch (d_flags: three int32, optional). */
int xrsd_lm_trial_update(const xrsd_objective_t *obj, const double *d_q, int32_t n_q, const double *d_I, int64_t ld_I,
                         int32_t batch, const double *d_Iobs, const double *d_dI, int64_t ld_obs, double *d_params,
                         int32_t ld_params, const double *d_trial, double *d_chi2, double *d_chi2_trial, double *d_lambda,
                         uint8_t *d_active, uint8_t *d_accepted, uint8_t *d_needjac, uint8_t *d_reason, int32_t *d_n_accept,
                         const xrsd_tables_t *trial_tables, int32_t n_xtal, int32_t *d_flags, int32_t tag, double up, double down,
                         double ftol, double xtol, double lambda_max, void *stream);

/* One WHOLE iteration of the resident loop in one call -- what xrsdkit_b200/fitting.py issues per iteration, from C, so
 * that the host cost of an iteration is one foreign call instead of eight (an iteration of a 2000-system batch is about a
 * millisecond of GPU work):
 *   xrsd_jacobian_batch (mode = XRSD_JAC_FULL on the first iteration, which also copies chi2 -> d_chi2_init; BASE_READY
 *   afterwards; masked by d_needjac) -> xrsd_lm_propose (counts the iteration in d_flags[2]) -> xrsd_reflection_tables of
 *   the trial point (layouts with crystalline populations; box `loop_cells`) -> xrsd_jacobian_batch(PATTERNS_ONLY) into
 *   the trial workspace -> xrsd_lm_trial_update (tag = 0) -> xrsd_lm_accept.
 * All pointers except layout / obj / index arrays / tables are device memory; obj must be a PREPARED objective
 * (d_fo / d_w are the streams of xrsd_objective_prepare); d_ws_base / d_ws_trial are workspaces of
 * xrsd_jacobian_workspace_bytes(n_q, batch, n_free) and (n_q, batch, 0) bytes whose partial-sum areas start zeroed. */
typedef struct xrsd_lm_state {
  const xrsd_layout_t *layout;
  const xrsd_objective_t *obj;
  const double *d_q;
  int32_t n_q, ld_params;
  double *d_params, *d_trial;              /* [batch][ld_params] accepted / trial parameters */
  int32_t batch, n_free;
  const double *d_fo, *d_w;                /* prepared objective streams */
  int64_t ld_obs;
  const int32_t *free_idx;                 /* [n_free] parameter slots (xrsd_lm_propose) */
  const int32_t *jac_idx;                  /* [n_free] the same, -(slot + 1) where a variant is forced (xrsd_jacobian_batch) */
  const int32_t *ties;
  const double *tie_coef;
  int32_t n_tie, n_var, n_pops, loop_cells;
  double rel_step, max_rel_step, up, down, ftol, xtol, lambda_max;
  double *d_JTJ, *d_JTr, *d_chi2, *d_chi2_trial, *d_chi2_init, *d_lambda;
  const double *d_lo, *d_hi;
  uint8_t *d_active, *d_accepted, *d_needjac, *d_reason;
  int32_t *d_n_accept, *d_flags;           /* d_flags: three int32, see xrsd_lm_trial_update */
  const xrsd_tables_t *tables, *trial_tables;
  void *d_ws_base, *d_ws_trial;
} xrsd_lm_state_t;
int xrsd_lm_iteration(const xrsd_lm_state_t *state, int32_t base_ready, void *stream);
int32_t xrsd_sizeof_lm_state(void);      /* sizeof(xrsd_lm_state_t) as compiled, for bindings */

/* Accepted systems (d_accepted[b] != 0) take over the trial evaluation: the 1 + n_pops slabs of the trial workspace
 * (pattern, running totals; see xrsd_jacobian_n_variants) go to slab 0 and slabs n_var.. of the base workspace
 * (d_dst / d_src may both be NULL), and the system's reflection tables become the trial tables -- a row index for
 * shared tables, a row copy otherwise (tables / trial_tables may both be NULL; they must be views of equal capacity). */
int xrsd_lm_accept(const uint8_t *d_accepted, int32_t batch, double *d_dst, int64_t dst_stride, const double *d_src,
                   int64_t src_stride, int32_t n_q, int32_t n_var, int32_t n_pops, const xrsd_tables_t *tables,
                   const xrsd_tables_t *trial_tables, int32_t n_xtal, void *stream);

/* chi2[b] of System.evaluate_residual for patterns that are already in device memory: d_I[batch][ld_I] against
 * d_Iobs / d_dI [batch][ld_obs] (ld_obs == 0 broadcasts), same objective and masks as xrsd_residual_batch. */
int xrsd_pattern_chi2_batch(const xrsd_objective_t *obj, const double *d_q, int32_t n_q, const double *d_I, int64_t ld_I,
                            int32_t batch, const double *d_Iobs, const double *d_dI, int64_t ld_obs, double *d_chi2,
                            const uint8_t *d_active, void *stream);

/* dst[r*dst_stride + i] = src[r*src_stride + i], i < row_len, for every row r < n_rows with d_mask[r] != 0
 * (strides and row_len in doubles).  An accepted LM step moves the trial pattern and its running totals over the
 * base rows of a resident Jacobian workspace with it (see xrsd_jacobian_n_variants for the slab layout). */
int xrsd_copy_rows_masked(double *d_dst, int64_t dst_stride, const double *d_src, int64_t src_stride, int64_t row_len,
                          int32_t n_rows, const uint8_t *d_mask, void *stream);

/* Host-buffer convenience call (what a non-Python caller binds): allocates / reuses an
 * internal device workspace, copies q and params in, builds tables when the layout has
 * crystalline populations, computes, copies I back.  Blocking.  Small calls without crystalline
 * populations make no copies at all: the kernel reads the parameters from and writes I(q) to mapped
 * pinned memory.
 * Threading: the d_* entry points above keep no state between calls (xrsd_last_error is per thread)
 * and may be called concurrently on different streams; this call keeps one workspace (device scratch, mapped
 * pinned staging buffer, a non-blocking stream of its own) per calling thread and device, selected by the
 * thread's current device (cudaSetDevice): threads evaluate side by side without a lock, and a process may drive
 * several GPUs.  xrsd_release_workspace frees the calling thread's workspaces; they also go when the thread ends. */
int xrsd_intensity_batch_host(const xrsd_layout_t *layout, const double *q, int32_t n_q,
                              const double *params, int32_t ld_params, int32_t batch,
                              double *I_out, int64_t ld_I);
void xrsd_release_workspace(void);

/* Host helper: the step of q if it is a linspace-like grid (every point within 4 ulp of q[0] + i*step and strictly
 * ascending), else 0 -- the value xrsd_layout_t.q_step expects. */
double xrsd_uniform_step(const double *q, int32_t n);

/* Peak profiles on their own: out[i] = profile(x[i]) for kind = XRSD_PROFILE_* with
 * (hwhm_g, hwhm_l) for voigt or (hwhm, unused) otherwise.  <- tools/peak_math.py:24-47
 * (gaussian, lorentzian, voigt; the reference's voigt calls scipy.special.wofz). */
int xrsd_peak_profile(int32_t kind, double hwhm_or_hwhm_g, double hwhm_l, const double *d_x, int32_t n,
                      double *d_out, void *stream);

/* Pattern featuriser: the 21 scale-free features of a pattern (q, I), in the order of
 * XRSD_N_FEATURES / the reference's profile_keys.  <- tools/profiler.py:64-221 (profile_pattern),
 * tools/peak_math.py:87-123 (humpness), tools/__init__.py:83-102 (pearson); the reference calls it
 * after every fit (system/__init__.py:276).  q[n_q] is shared by the batch; pattern b is
 * d_I + b*ld_I; features of pattern b go to d_features + b*ld_features (>= 21 doubles).
 * Up to XRSD_FEATURE_MAX_Q points the pattern is held in shared memory (26 bytes per point); longer patterns (up to
 * XRSD_FEATURE_MAX_Q_GLOBAL) use a stream-ordered global-memory working set inside the call -- same results, slower.
 * n_q >= 8.  Degenerate patterns give NaN features where the reference raises or warns. */
#define XRSD_N_FEATURES 21
#define XRSD_FEATURE_MAX_Q 8192
#define XRSD_FEATURE_MAX_Q_GLOBAL (1 << 22)
int xrsd_feature_batch(const double *d_q, int32_t n_q, const double *d_I, int64_t ld_I, int32_t batch,
                       double *d_features, int64_t ld_features, void *stream);

/* FP64 pipe probe used by bench.py for the roofline denominator: runs a DFMA-bound
 * kernel and returns the measured TFLOP/s (2 flop per DFMA) on the current device. */
int xrsd_probe_fp64_peak(double *tflops, double *elapsed_ms, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* XRSD_B200_H */
