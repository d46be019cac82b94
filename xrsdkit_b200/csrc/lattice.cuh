// lattice.cuh -- cell vectors, reciprocal vectors and the 10 Cartesian symmetry
// operations, as the reference defines them (definitions.py:504-570,
// scattering/symmetries.py:10-21).  FP64, evaluated per system on the device.
#pragma once
#include <math.h>
#include "../../include/xrsd_b200.h"

namespace xrsd {

struct Cell {
  double a[3][3];   // rows: a1, a2, a3 (Cartesian)
  double b[3][3];   // rows: b1, b2, b3 (crystallographic reciprocal vectors, no 2 pi)
};

__device__ __forceinline__ double deg2rad(double d) { return d * M_PI / 180.0; }

// definitions.py:504-551.  p = {a, b, c, alpha, beta, gamma}; unused entries are ignored.
__device__ inline void cell_vectors(int lattice, const double *p, double a[3][3]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) a[i][j] = 0.0;
  const double A = p[0], B = p[1], C = p[2];
  switch (lattice) {
    case XRSD_LAT_CUBIC:
      a[0][0] = A; a[1][1] = A; a[2][2] = A;
      break;
    case XRSD_LAT_HCP:
      a[0][0] = A; a[1][0] = 0.5 * A; a[1][1] = sqrt(3.0) / 2.0 * A; a[2][2] = sqrt(8.0 / 3.0) * A;
      break;
    case XRSD_LAT_HEXAGONAL:
      a[0][0] = A; a[1][0] = 0.5 * A; a[1][1] = sqrt(3.0) / 2.0 * A; a[2][2] = C;
      break;
    case XRSD_LAT_ORTHORHOMBIC:
      a[0][0] = A; a[1][1] = B; a[2][2] = C;
      break;
    case XRSD_LAT_TETRAGONAL:
      a[0][0] = A; a[1][1] = A; a[2][2] = C;
      break;
    case XRSD_LAT_MONOCLINIC: {
      const double br = deg2rad(p[4]);
      a[0][0] = A; a[1][1] = B; a[2][0] = C * cos(br); a[2][2] = C * sin(br);
    } break;
    case XRSD_LAT_RHOMBOHEDRAL: {
      const double ar = deg2rad(p[3]);
      const double ca = cos(ar), sa = sin(ar);
      const double omega = A * A * A * sqrt(1.0 - ca * ca - ca * ca - ca * ca + 2.0 * ca * ca * ca);
      a[0][0] = A;
      a[1][0] = A * ca; a[1][1] = A * sa;
      a[2][0] = A * ca; a[2][1] = A * (ca - ca * ca) / sa; a[2][2] = omega / (A * A * sa);
    } break;
    case XRSD_LAT_TRICLINIC: {
      const double ar = deg2rad(p[3]), br = deg2rad(p[4]), gr = deg2rad(p[5]);
      const double ca = cos(ar), cb = cos(br), cg = cos(gr), sg = sin(gr);
      const double omega = A * B * C * sqrt(1.0 - ca * ca - cb * cb - cg * cg + 2.0 * ca * cb * cg);
      a[0][0] = A;
      a[1][0] = B * cg; a[1][1] = B * sg;
      a[2][0] = C * cb; a[2][1] = C * (ca - cb * cg) / sg; a[2][2] = omega / (A * B * sg);
    } break;
    default: break;
  }
}

__device__ __forceinline__ void cross3(const double *u, const double *v, double *w) {
  w[0] = u[1] * v[2] - u[2] * v[1];
  w[1] = u[2] * v[0] - u[0] * v[2];
  w[2] = u[0] * v[1] - u[1] * v[0];
}

// definitions.py:553-570 (crystallographic=True)
__device__ inline void reciprocal_vectors(const double a[3][3], double b[3][3]) {
  double x1[3], x2[3], x3[3];
  cross3(a[1], a[2], x1);
  cross3(a[2], a[0], x2);
  cross3(a[0], a[1], x3);
  const double vol = a[0][0] * x1[0] + a[0][1] * x1[1] + a[0][2] * x1[2];
  for (int j = 0; j < 3; ++j) {
    b[0][j] = x1[j] / vol;
    b[1][j] = x2[j] / vol;
    b[2][j] = x3[j] / vol;
  }
}

__device__ inline void make_cell(const xrsd_pop_t &pop, const double *params, Cell *cell) {
  double p[6];
  for (int i = 0; i < 6; ++i) p[i] = (pop.p_lat[i] >= 0) ? params[pop.p_lat[i]] : 0.0;
  cell_vectors(pop.lattice, p, cell->a);
  reciprocal_vectors(cell->a, cell->b);
}

// |h b1 + k b2 + l b3| as scattering/__init__.py:276,287 computes it (dot, then 2-norm)
__device__ __forceinline__ double g_norm(const double b[3][3], double h, double k, double l) {
  const double gx = h * b[0][0] + k * b[1][0] + l * b[2][0];
  const double gy = h * b[0][1] + k * b[1][1] + l * b[2][1];
  const double gz = h * b[0][2] + k * b[1][2] + l * b[2][2];
  return sqrt(gx * gx + gy * gy + gz * gz);
}

// The symmetry operations are signed permutation matrices: row m of op `id` has a single
// non-zero entry sgn at column perm[m] (scattering/symmetries.py:10-21).
__device__ __forceinline__ void sym_op(int id, int perm[3], double sgn[3]) {
  perm[0] = 0; perm[1] = 1; perm[2] = 2;
  sgn[0] = 1.0; sgn[1] = 1.0; sgn[2] = 1.0;
  switch (id) {
    case XRSD_OP_INVERSION: sgn[0] = -1.0; sgn[1] = -1.0; sgn[2] = -1.0; break;
    case XRSD_OP_MIRROR_X: sgn[0] = -1.0; break;
    case XRSD_OP_MIRROR_Y: sgn[1] = -1.0; break;
    case XRSD_OP_MIRROR_Z: sgn[2] = -1.0; break;
    case XRSD_OP_MIRROR_X_Y: perm[0] = 1; perm[1] = 0; sgn[0] = -1.0; sgn[1] = -1.0; break;   // [[0,-1,0],[-1,0,0],[0,0,1]]
    case XRSD_OP_MIRROR_Y_Z: perm[1] = 2; perm[2] = 1; sgn[1] = -1.0; sgn[2] = -1.0; break;   // [[1,0,0],[0,0,-1],[0,-1,0]]
    case XRSD_OP_MIRROR_Z_X: perm[0] = 2; perm[2] = 0; sgn[0] = -1.0; sgn[2] = -1.0; break;   // [[0,0,-1],[0,1,0],[-1,0,0]]
    case XRSD_OP_MIRROR_NX_Y: perm[0] = 1; perm[1] = 0; break;                                // [[0,1,0],[1,0,0],[0,0,1]]
    case XRSD_OP_MIRROR_NY_Z: perm[1] = 2; perm[2] = 1; break;                                // [[1,0,0],[0,0,1],[0,1,0]]
    case XRSD_OP_MIRROR_NZ_X: perm[0] = 2; perm[2] = 0; break;                                // [[0,0,1],[0,1,0],[1,0,0]]
    default: break;
  }
}

}  // namespace xrsd
