// table_kernel.cu -- reflection-table builder (one CTA per crystalline population of
// one system of the batch), everything in shared memory.
//
// Replaces, bit-exactly on the hkl / multiplicity lists:
//   scattering/__init__.py:248-278   hkl enumeration in the G_min < |G| <= G_max shell
//   scattering/symmetries.py:44-92   greedy, order-dependent mirror reduction
// and precomputes the parameter-independent part of scattering/__init__.py:307-331:
// reflections are merged into shells of equal |G| and each shell carries
// sum_hkl mult * Re(G_a conj G_b), G_a = sum_sites exp(2 pi i (site + xyz_a).hkl).
//
// The reference forms an N x N distance matrix per operation and takes the row-wise
// argmin.  Every operation is an involutive signed permutation, so the only candidate
// that can pass the `dist < symprec` test is the lattice point nearest to op.p_i, found
// here by mapping back to integer hkl (O(N) per operation instead of O(N^2)); the
// distance itself is then evaluated exactly as the reference does.  Points are kept on
// the dense (2n1-1)(2n2-1)(2n3-1) candidate box as uint16 multiplicities (0 = absent),
// which also preserves the reference's l-outer / k / h-inner ordering for free.
#include <cuda_runtime.h>
#include <stdint.h>
#include "lattice.cuh"
#include "launch_util.cuh"

namespace xrsd {

constexpr int TABLE_THREADS = 256;

__host__ __device__ inline long long xrsd_scratch_bytes(long long cap_big) {
  // keys f64[cap] + list_cell i32[cap] + sidx i32[cap] + seg_start i32[cap + 2] + msum i32[cap], 16-byte multiple
  return ((cap_big * 24 + 8) + 15) & ~15LL;
}

struct TableShared {
  Cell cell;
  double g_min, g_max;
  int n[3];
  int dim[3];          // 2n-1
  float rcp_d0, rcp_d01;   // 1/dim[0], 1/(dim[0] dim[1])
  int cells;
  int n_all, n_red, n_shell_raw, n_shell;
  int hmin[3], hmax[3];
  int status;
  int grid_slot;           // >= 0: the candidate grid lives in region grid_slot of T.grid_scratch
  int scan_carry;
  int alive_cnt[3];        // rotating counters of the alive lists (one barrier per operation)
  double max_geo;
  int warp_tmp[TABLE_THREADS / 32];
};

// c / d for 0 <= c < 2^24 with a float reciprocal and a +-1 correction (an integer division by a run-time
// divisor costs ~25 instructions; this was 30 % of the kernel's instructions)
__device__ __forceinline__ int fast_div(int c, int d, float rcp) {
  int qd = (int)((float)c * rcp);
  const int r = c - qd * d;
  if (r < 0) --qd;
  else if (r >= d) ++qd;
  return qd;
}

__device__ __forceinline__ void cell_to_hkl(const TableShared &S, int c, int &h, int &k, int &l) {
  const int d0 = S.dim[0], d01 = S.dim[0] * S.dim[1];
  const int il = fast_div(c, d01, S.rcp_d01);
  const int rem = c - il * d01;
  const int ik = fast_div(rem, d0, S.rcp_d0);
  h = rem - ik * d0 - (S.n[0] - 1);
  k = ik - (S.n[1] - 1);
  l = il - (S.n[2] - 1);
}

__device__ __forceinline__ int hkl_to_cell(const TableShared &S, int h, int k, int l) {
  const int ih = h + S.n[0] - 1, ik = k + S.n[1] - 1, il = l + S.n[2] - 1;
  if (ih < 0 || ih >= S.dim[0] || ik < 0 || ik >= S.dim[1] || il < 0 || il >= S.dim[2]) return -1;
  return (il * S.dim[1] + ik) * S.dim[0] + ih;
}

// Appends `value` to list[] for every lane with `take` set: one shared-memory atomic per warp instead of one per lane.
// Must be called by all lanes of the warp (converged).
__device__ __forceinline__ void warp_append(bool take, uint16_t value, uint16_t *list, int *counter) {
  const unsigned m = __ballot_sync(0xffffffffu, take);
  if (m == 0) return;
  const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(counter, __popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (take) list[base + __popc(m & ((1u << lane) - 1))] = value;
}

// block-wide exclusive scan of one flag per thread; returns this thread's offset and adds the
// block total to S.scan_carry (which callers read as the running base).
__device__ __forceinline__ int block_flag_scan(TableShared &S, bool flag, int &base_out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned bal = __ballot_sync(0xffffffffu, flag);
  const int within = __popc(bal & ((1u << lane) - 1u));
  if (lane == 0) S.warp_tmp[warp] = __popc(bal);
  __syncthreads();
  int wbase = 0, total = 0;
  for (int w = 0; w < TABLE_THREADS / 32; ++w) {
    const int c = S.warp_tmp[w];
    if (w < warp) wbase += c;
    total += c;
  }
  base_out = S.scan_carry;
  __syncthreads();
  if (threadIdx.x == 0) S.scan_carry += total;
  __syncthreads();
  return wbase + within;
}

__global__ void __launch_bounds__(TABLE_THREADS, 3)
reflection_table_kernel(const __grid_constant__ xrsd_layout_t L, const double *__restrict__ params, int ldp,
                        int batch, const xrsd_tables_t T, int max_cells, int cap_list,
                        const uint8_t *__restrict__ active) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ TableShared S;

  const int t = blockIdx.x;
  const int b = t / L.n_xtal;
  const int slot = t - b * L.n_xtal;
  if (b >= batch) return;
  if (active != nullptr && !active[b]) return;      // converged system of a fit: its table is left as it is
  // storage row of this table; with shared rows only the CTA that claimed a row (table_key_kernel) builds it
  if (T.table_of != nullptr && !T.build_flag[t]) return;
  const int row = T.table_of != nullptr ? T.table_of[t] : T.own_base + t;
  int ipop = -1;
  for (int i = 0; i < L.n_pops; ++i)
    if (L.pops[i].kind == XRSD_POP_CRYSTALLINE && L.pops[i].table_slot == slot) ipop = i;
  if (ipop < 0) return;
  const xrsd_pop_t &pop = L.pops[ipop];
  const double *prm = params + (size_t)b * ldp;
  const int tid = threadIdx.x;

  // shared-memory carve-up
  uint16_t *grid = reinterpret_cast<uint16_t *>(smem_raw);
  size_t off = ((size_t)max_cells * sizeof(uint16_t) + 15) & ~(size_t)15;
  double *keys = reinterpret_cast<double *>(smem_raw + off);
  off += (size_t)cap_list * sizeof(double);
  int *list_cell = reinterpret_cast<int *>(smem_raw + off);
  off += (size_t)cap_list * sizeof(int);
  int *sidx = reinterpret_cast<int *>(smem_raw + off);
  off += (size_t)cap_list * sizeof(int);
  int *seg_start = reinterpret_cast<int *>(smem_raw + off);   // [cap_list + 1]
  off += ((size_t)cap_list + 1) * sizeof(int);
  int *msum_s = reinterpret_cast<int *>(smem_raw + off);      // [cap_list]
  long long cap_eff = cap_list;                                // capacity of the five arrays above

  if (tid == 0) {
    make_cell(pop, prm, &S.cell);
    // scattering/__init__.py:248-268
    const double q_max = pop.q_max, q_min = pop.q_min;
    const double d_min = 2.0 * M_PI / q_max;
    S.g_max = 1.0 / d_min;
    S.g_min = 0.0;
    if (q_min > 0.0) {
      const double d_max = 2.0 * M_PI / q_min;
      S.g_min = 1.0 / d_max;
    }
    int st = 0;
    long long cells = 1;
    for (int i = 0; i < 3; ++i) {
      const double *a = S.cell.a[i];
      const double *bb = S.cell.b[i];
      const double absa = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
      const double bdota = bb[0] * a[0] + bb[1] * a[1] + bb[2] * a[2];
      const double nf = ceil(S.g_max * absa / bdota);
      if (!(nf >= 1.0) || !(nf < 1.0e4)) { st |= XRSD_TABLE_BAD_LATTICE; S.n[i] = 1; }
      else S.n[i] = (int)nf;
      S.dim[i] = 2 * S.n[i] - 1;
      cells *= S.dim[i];
    }
    S.grid_slot = -1;
    if (cells > (long long)max_cells) {
      // beyond the shared-memory budget: take one of the caller's global-memory regions, if there is one left
      int gs = -1;
      if (T.grid_scratch != nullptr && cells <= T.cap_grid_cells && cells < (1LL << 24)) {   // fast_div is exact below 2^24
        gs = atomicAdd(T.grid_slot_counter, 1);
        if (gs >= T.n_grid_slots) gs = -1;
      }
      if (gs < 0) { st |= XRSD_TABLE_GRID_OVERFLOW; cells = 0; }
      S.grid_slot = gs;
    }
    S.cells = (int)cells;
    S.rcp_d0 = 1.0f / (float)S.dim[0];
    S.rcp_d01 = 1.0f / ((float)S.dim[0] * (float)S.dim[1]);
    S.status = st;
    S.n_all = 0; S.n_red = 0; S.n_shell_raw = 0; S.n_shell = 0;
    for (int i = 0; i < 3; ++i) { S.hmin[i] = 1 << 30; S.hmax[i] = -(1 << 30); }
    S.scan_carry = 0;
    S.max_geo = 0.0;
  }
  __syncthreads();

  auto finish = [&](int n_shell) {
    if (tid == 0) {
      T.n_shell[row] = n_shell;
      T.n_refl[row] = S.n_red;
      T.n_all[row] = S.n_all;
      T.status[row] = S.status;
    }
  };
  if (S.status != 0) { finish(0); return; }
  if (S.grid_slot >= 0)
    grid = reinterpret_cast<uint16_t *>(T.grid_scratch) + (size_t)S.grid_slot * (size_t)T.cap_grid_cells;
  // a cubic cell orders its reflections by the integer N = h^2+k^2+l^2 (exact), every other cell by the float |G|:
  // ties inside a shell then fall in list order, which makes the whole table a function of the passing set alone
  const bool int_keys = pop.lattice == XRSD_LAT_CUBIC;

  // ---- 1. candidates in the shell G_min < |G| <= G_max (scattering/__init__.py:275-276)
  {
    int cnt = 0, lo[3] = {1 << 30, 1 << 30, 1 << 30}, hi[3] = {-(1 << 30), -(1 << 30), -(1 << 30)};
    for (int c = tid; c < S.cells; c += TABLE_THREADS) {
      int h, k, l;
      cell_to_hkl(S, c, h, k, l);
      const double g = g_norm(S.cell.b, (double)h, (double)k, (double)l);
      const bool in = (S.g_min < g) && (g <= S.g_max);
      grid[c] = in ? 1 : 0;
      if (in) {
        ++cnt;
        lo[0] = min(lo[0], h); hi[0] = max(hi[0], h);
        lo[1] = min(lo[1], k); hi[1] = max(hi[1], k);
        lo[2] = min(lo[2], l); hi[2] = max(hi[2], l);
      }
    }
    if (cnt) {
      atomicAdd(&S.n_all, cnt);
      for (int i = 0; i < 3; ++i) { atomicMin(&S.hmin[i], lo[i]); atomicMax(&S.hmax[i], hi[i]); }
    }
  }
  __syncthreads();
  if (S.n_all == 0) { finish(0); return; }   // scattering/__init__.py:277-278 -> zeros

  // ---- 2. greedy reduction, one operation after the other (scattering/symmetries.py:61-90)
  if (pop.use_symmetry && pop.n_ops > 0) {
    const int rk = S.hmax[1] - S.hmin[1] + 1, rl = S.hmax[2] - S.hmin[2] + 1;   // symmetries.py:54-55
    // visit only the points still alive: two (unordered) uint16 cell lists laid over the sort/shell
    // scratch (20 bytes per list slot, not yet in use) that are re-compacted after every operation,
    // so the later operations do not walk over points the earlier ones dropped
    uint16_t *alive[2] = {reinterpret_cast<uint16_t *>(keys), reinterpret_cast<uint16_t *>(keys) + 5 * cap_list};
    const bool use_list = S.cells <= 65535 && S.n_all <= 5 * cap_list;
    if (use_list) {
      if (tid < 3) S.alive_cnt[tid] = 0;
      __syncthreads();
      for (int c0 = 0; c0 < S.cells; c0 += TABLE_THREADS) {
        const int c = c0 + tid;
        warp_append(c < S.cells && grid[c] != 0, (uint16_t)c, alive[0], &S.alive_cnt[2]);
      }
      __syncthreads();
    }
    int n_visit = use_list ? S.alive_cnt[2] : S.cells;
    for (int iop = 0; iop < pop.n_ops; ++iop) {
      int perm[3];
      double sgn[3];
      sym_op(pop.ops[iop], perm, sgn);
      const uint16_t *cur = alive[iop & 1];
      uint16_t *nxt = alive[(iop + 1) & 1];
      // survivors of operation iop are counted in alive_cnt[iop % 3]; the counter of the next operation is
      // cleared here -- every thread read it (as n_visit) before the barrier that ended the previous operation
      int *cnt = &S.alive_cnt[iop % 3];
      if (tid == 0) S.alive_cnt[(iop + 1) % 3] = 0;
      for (int iv0 = 0; iv0 < n_visit; iv0 += TABLE_THREADS) {
        const int iv = iv0 + tid;
        bool keep = false;
        int c = 0;
        if (iv < n_visit) {
          c = use_list ? (int)cur[iv] : iv;
          const unsigned mi = grid[c];
          if (mi) {
            int h, k, l;
            cell_to_hkl(S, c, h, k, l);
            // lat_pts = all_hkl . rlat^T  (symmetries.py:52): p_m = b_m . (h,k,l)
            double p[3], s[3];
            for (int m = 0; m < 3; ++m)
              p[m] = (double)h * S.cell.b[m][0] + (double)k * S.cell.b[m][1] + (double)l * S.cell.b[m][2];
            for (int m = 0; m < 3; ++m) s[m] = sgn[m] * p[perm[m]];
            // nearest lattice point to op.p_i: hkl_j = rlat^-1 s = A^T s
            int hj[3];
            for (int n = 0; n < 3; ++n)
              hj[n] = (int)rint(S.cell.a[0][n] * s[0] + S.cell.a[1][n] * s[1] + S.cell.a[2][n] * s[2]);
            const int cj = hkl_to_cell(S, hj[0], hj[1], hj[2]);
            bool drop = false;
            if (cj >= 0 && cj != c && grid[cj]) {
              // distance |p_i - op.p_j| (symmetries.py:62-72)
              double pj[3], d2 = 0.0;
              for (int m = 0; m < 3; ++m)
                pj[m] = (double)hj[0] * S.cell.b[m][0] + (double)hj[1] * S.cell.b[m][1] + (double)hj[2] * S.cell.b[m][2];
              for (int m = 0; m < 3; ++m) {
                const double df = p[m] - sgn[m] * pj[perm[m]];
                d2 += df * df;
              }
              if (d2 < 1.0e-12) {                    // dist < 1e-6 (the distances met are ~1e-16 or >~ 1e-3)
                const int rank_i = (h * rk + k) * rl + l;
                const int rank_j = (hj[0] * rk + hj[1]) * rl + hj[2];
                drop = rank_i < rank_j;
              }
            }
            if (drop) {
              // j keeps the point and collects i's multiplicity.  j cannot be dropped by this same
              // operation (its image is i, of lower rank), and only i's thread writes grid[cj].
              grid[cj] = (uint16_t)(grid[cj] + mi);
              grid[c] = 0;
            } else {
              keep = true;
            }
          }
        }
        if (use_list) warp_append(keep, (uint16_t)c, nxt, cnt);
      }
      __syncthreads();
      if (use_list) n_visit = *cnt;
    }
  }

  // ---- 3. ordered compaction (l outer, k, h inner = cell order): each thread owns a contiguous
  //         run of cells, one block-wide scan of the per-thread counts gives the output offsets
  {
    int32_t *out_refl = (T.refl_hkl && T.cap_refl > 0) ? T.refl_hkl + (size_t)row * T.cap_refl * 4 : nullptr;
    const int run = (S.cells + TABLE_THREADS - 1) / TABLE_THREADS;
    const int c_lo = min(tid * run, S.cells), c_hi = min(c_lo + run, S.cells);
    int cnt = 0;
    for (int c = c_lo; c < c_hi; ++c) cnt += grid[c] != 0;
    // inclusive scan of cnt over the block
    __shared__ int scan_sm[TABLE_THREADS];
    scan_sm[tid] = cnt;
    __syncthreads();
    for (int o = 1; o < TABLE_THREADS; o <<= 1) {
      const int v = (tid >= o) ? scan_sm[tid - o] : 0;
      __syncthreads();
      scan_sm[tid] += v;
      __syncthreads();
    }
    int pos = scan_sm[tid] - cnt;
    if (tid == TABLE_THREADS - 1) {
      S.n_red = scan_sm[tid];
      if (S.n_red > cap_list && !(T.big_scratch != nullptr && (long long)S.n_red <= T.cap_big))
        S.status |= XRSD_TABLE_LIST_OVERFLOW;
    }
    __syncthreads();
    if (S.status != 0) { finish(0); return; }
    if (S.n_red > cap_list) {
      // a reduced list too long for shared memory (large cell, no usable symmetry): the same arrays in this
      // table's region of the caller's global scratch -- slow, but correct and rare
      unsigned char *g = reinterpret_cast<unsigned char *>(T.big_scratch) + (size_t)t * (size_t)xrsd_scratch_bytes(T.cap_big);
      keys = reinterpret_cast<double *>(g);
      g += (size_t)T.cap_big * sizeof(double);
      list_cell = reinterpret_cast<int *>(g);
      g += (size_t)T.cap_big * sizeof(int);
      sidx = reinterpret_cast<int *>(g);
      g += (size_t)T.cap_big * sizeof(int);
      seg_start = reinterpret_cast<int *>(g);
      g += ((size_t)T.cap_big + 2) * sizeof(int);
      msum_s = reinterpret_cast<int *>(g);
      cap_eff = T.cap_big;
    }
    for (int c = c_lo; c < c_hi; ++c) {
      const unsigned m = grid[c];
      if (!m) continue;
      list_cell[pos] = c;
      if (out_refl && pos < T.cap_refl) {
        int h, k, l;
        cell_to_hkl(S, c, h, k, l);
        out_refl[pos * 4 + 0] = h; out_refl[pos * 4 + 1] = k; out_refl[pos * 4 + 2] = l; out_refl[pos * 4 + 3] = (int)m;
      }
      ++pos;
    }
    __syncthreads();
  }
  const int n_red = S.n_red;

  // ---- 4. sort reduced reflections by |G| (bitonic, shared memory)
  int npow = 1;
  while (npow < n_red) npow <<= 1;
  for (int r = tid; r < npow; r += TABLE_THREADS) {
    if (r < n_red) {
      int h, k, l;
      cell_to_hkl(S, list_cell[r], h, k, l);
      keys[r] = int_keys ? (double)(h * h + k * k + l * l) : g_norm(S.cell.b, (double)h, (double)k, (double)l);
      sidx[r] = r;
    } else {
      keys[r] = INFINITY;
      sidx[r] = r;
    }
  }
  __syncthreads();
  if (n_red <= 2 * TABLE_THREADS) {
    // short lists (the symmetric case, cfg3: 198): rank sort -- every element counts the elements ordered before
    // it (n^2 broadcast reads, all warps busy, two barriers) instead of ~40 dependent bitonic stages
    double my_key[2];
    int my_rank[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int i = tid + e * TABLE_THREADS;
      my_rank[e] = -1;
      if (i < n_red) {
        const double ki = keys[i];
        int rank = 0;
        for (int j = 0; j < n_red; ++j) {
          const double kj = keys[j];
          rank += (kj < ki) || (kj == ki && j < i);
        }
        my_key[e] = ki;
        my_rank[e] = rank;
      }
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 2; ++e)
      if (my_rank[e] >= 0) { keys[my_rank[e]] = my_key[e]; sidx[my_rank[e]] = tid + e * TABLE_THREADS; }
    __syncthreads();
  } else {
    for (int size = 2; size <= npow; size <<= 1) {
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int i = tid; i < npow / 2; i += TABLE_THREADS) {
          const int lo = 2 * i - (i & (stride - 1));
          const int hi = lo + stride;
          const bool up = ((lo & size) == 0);
          const double ka = keys[lo], kb = keys[hi];
          const int ia = sidx[lo], ib = sidx[hi];
          const bool gt = (ka > kb) || (ka == kb && ia > ib);
          if (gt == up) { keys[lo] = kb; keys[hi] = ka; sidx[lo] = ib; sidx[hi] = ia; }
        }
        __syncthreads();
      }
    }
  }

  // ---- 5. shells: runs of (numerically) equal |G|
  for (int r0 = 0; r0 < n_red; r0 += TABLE_THREADS) {
    const int r = r0 + tid;
    bool head = false;
    if (r < n_red) head = (r == 0) || (int_keys ? keys[r] != keys[r - 1] : keys[r] - keys[r - 1] > 8.9e-16 * keys[r]);
    int base;
    const int pos = block_flag_scan(S, head, base) + base;
    if (head && pos < cap_eff) seg_start[pos] = r;
  }
  if (tid == 0) {
    S.n_shell_raw = S.scan_carry;
    S.scan_carry = 0;
    const int npair = pop.n_species * (pop.n_species + 1) / 2;
    // shells are staged in shared memory (over the sort keys) and only the kept ones go to the output
    // arrays; if they do not fit there, the raw list itself must fit the output capacity
    if ((long long)S.n_shell_raw * npair > cap_eff && S.n_shell_raw > T.cap_shell) S.status |= XRSD_TABLE_SHELL_OVERFLOW;
    seg_start[S.n_shell_raw] = n_red;
  }
  __syncthreads();
  if (S.status != 0) { finish(0); return; }
  const int n_raw = S.n_shell_raw;
  const int n_sp = pop.n_species;
  const int n_pairs = n_sp * (n_sp + 1) / 2;
  int32_t *o_hkl = T.shell_hkl + (size_t)row * T.cap_shell * 4;
  double *o_geo = T.shell_geo + (size_t)row * T.cap_shell * T.n_pairs_max;

  // per-shell geometric weights: sum_members mult * Re(G_a conj G_b)  (scattering/__init__.py:312-331)
  const bool staged = (long long)n_raw * n_pairs <= cap_eff;
  double *geo_s = keys;                         // the sort keys are no longer needed
  double my_max = 0.0;
  for (int s = tid; s < n_raw; s += TABLE_THREADS) {
    double acc[XRSD_MAX_PAIRS];
    for (int i = 0; i < XRSD_MAX_PAIRS; ++i) acc[i] = 0.0;
    int msum = 0;
    int h0 = 0, k0 = 0, l0 = 0;
    const int r_lo = seg_start[s], r_hi = seg_start[s + 1];
    // members of a shell in list order (the reference's l, k, h order), whatever the rounding of their |G| was:
    // the representative hkl and the summation order below do not depend on the cell edge (this thread owns the segment)
    for (int r = r_lo + 1; r < r_hi; ++r) {
      const int v = sidx[r];
      int m = r - 1;
      while (m >= r_lo && sidx[m] > v) { sidx[m + 1] = sidx[m]; --m; }
      sidx[m + 1] = v;
    }
    for (int r = r_lo; r < r_hi; ++r) {
      const int c = list_cell[sidx[r]];
      int h, k, l;
      cell_to_hkl(S, c, h, k, l);
      if (r == r_lo) { h0 = h; k0 = k; l0 = l; }
      const double mult = (double)grid[c];
      msum += (int)grid[c];
      double gre[XRSD_MAX_SPECIES], gim[XRSD_MAX_SPECIES];
      for (int a = 0; a < n_sp; ++a) {
        double xyz[3] = {0.0, 0.0, 0.0};
        for (int j = 0; j < 3; ++j)
          if (pop.p_uvw[a][j] >= 0) xyz[j] = prm[pop.p_uvw[a][j]];
        double re = 0.0, im = 0.0;
        for (int site = 0; site < pop.n_sites; ++site) {
          const double gdr = (pop.sites[site][0] + xyz[0]) * h + (pop.sites[site][1] + xyz[1]) * k +
                             (pop.sites[site][2] + xyz[2]) * l;
          double sn, cs;
          sincos(2.0 * M_PI * gdr, &sn, &cs);     // np.exp(2j*np.pi*g_dot_r)
          re += cs; im += sn;
        }
        gre[a] = re; gim[a] = im;
      }
      int ip = 0;
      for (int a = 0; a < n_sp; ++a)
        for (int bq = a; bq < n_sp; ++bq, ++ip) acc[ip] += mult * (gre[a] * gre[bq] + gim[a] * gim[bq]);
    }
    if (staged) {
      msum_s[s] = msum;
    } else {
      o_hkl[s * 4 + 0] = h0; o_hkl[s * 4 + 1] = k0; o_hkl[s * 4 + 2] = l0; o_hkl[s * 4 + 3] = msum;
    }
    for (int ip = 0; ip < n_pairs; ++ip) {
      if (!staged) o_geo[(size_t)s * T.n_pairs_max + ip] = acc[ip];
      my_max = fmax(my_max, fabs(acc[ip]));
    }
    if (staged) {
      // every thread first finishes reading keys-independent data above; geo_s aliases keys, which
      // nobody reads any more after the shell boundaries were found
      for (int ip = 0; ip < n_pairs; ++ip) geo_s[(size_t)s * n_pairs + ip] = acc[ip];
    }
  }
  if (!pop.prune_extinct && !staged) { finish(n_raw); return; }

  // ---- 6. keep shells with a pair weight above 1e-20 of the largest (all of them when pruning is
  //         off) and write them, compacted, to the output arrays
  for (int o = 16; o > 0; o >>= 1) my_max = fmax(my_max, __shfl_xor_sync(0xffffffffu, my_max, o));
  __shared__ double wmax[TABLE_THREADS / 32];
  if ((tid & 31) == 0) wmax[tid >> 5] = my_max;
  __syncthreads();
  if (tid == 0) {
    double m = 0.0;
    for (int w = 0; w < TABLE_THREADS / 32; ++w) m = fmax(m, wmax[w]);
    S.max_geo = m;
  }
  __syncthreads();
  const double thr = pop.prune_extinct ? 1.0e-20 * S.max_geo : -1.0;
  if (staged) {
    // count first so that a capacity overflow leaves the output untouched
    int cnt = 0;
    for (int s = tid; s < n_raw; s += TABLE_THREADS) {
      bool keep = false;
      for (int ip = 0; ip < n_pairs; ++ip) keep = keep || (fabs(geo_s[(size_t)s * n_pairs + ip]) > thr);
      cnt += keep;
    }
    if (cnt) atomicAdd(&S.n_shell, cnt);
    __syncthreads();
    if (S.n_shell > T.cap_shell) {
      if (tid == 0) S.status |= XRSD_TABLE_SHELL_OVERFLOW;
      __syncthreads();
      finish(0);
      return;
    }
  }
  for (int s0 = 0; s0 < n_raw; s0 += TABLE_THREADS) {
    const int s = s0 + tid;
    bool keep = false;
    int hk[4] = {0, 0, 0, 0};
    double geo[XRSD_MAX_PAIRS];
    if (s < n_raw) {
      if (staged) {
        int h, k, l;
        cell_to_hkl(S, list_cell[sidx[seg_start[s]]], h, k, l);
        hk[0] = h; hk[1] = k; hk[2] = l; hk[3] = msum_s[s];
      } else {
        for (int j = 0; j < 4; ++j) hk[j] = o_hkl[s * 4 + j];
      }
      for (int ip = 0; ip < n_pairs; ++ip) {
        geo[ip] = staged ? geo_s[(size_t)s * n_pairs + ip] : o_geo[(size_t)s * T.n_pairs_max + ip];
        keep = keep || (fabs(geo[ip]) > thr);
      }
    }
    int base;
    const int pos = block_flag_scan(S, keep, base) + base;   // has __syncthreads: all reads precede writes
    if (keep) {
      for (int j = 0; j < 4; ++j) o_hkl[pos * 4 + j] = hk[j];
      for (int ip = 0; ip < n_pairs; ++ip) o_geo[(size_t)pos * T.n_pairs_max + ip] = geo[ip];
    }
    __syncthreads();
  }
  finish(S.scan_carry);
}

// One thread per table: which storage row does it use, and does this call have to build it?  (See xrsd_tables_t.)
__global__ void table_key_kernel(const __grid_constant__ xrsd_layout_t L, const double *__restrict__ params, int ldp,
                                 int batch, const xrsd_tables_t T, const uint8_t *__restrict__ active) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= batch * L.n_xtal) return;
  const int b = t / L.n_xtal, slot = t - b * L.n_xtal;
  if (active != nullptr && !active[b]) return;
  int ipop = -1;
  for (int i = 0; i < L.n_pops; ++i)
    if (L.pops[i].kind == XRSD_POP_CRYSTALLINE && L.pops[i].table_slot == slot) ipop = i;
  if (ipop < 0) return;
  const xrsd_pop_t &pop = L.pops[ipop];
  int row = T.own_base + t;
  uint8_t build = 1;
  bool shareable = T.n_shared > 0 && T.shared_tag != nullptr && pop.lattice == XRSD_LAT_CUBIC;
  for (int a = 0; a < pop.n_species; ++a)
    for (int j = 0; j < 3; ++j) shareable = shareable && pop.p_uvw[a][j] < 0;     // free coordinates enter the weights
  if (shareable) {
    Cell cell;
    make_cell(pop, params + (size_t)b * ldp, &cell);
    const double bi = cell.b[0][0];
    const bool diag = bi > 0.0 && cell.b[1][1] == bi && cell.b[2][2] == bi && cell.b[0][1] == 0.0 && cell.b[0][2] == 0.0 &&
                      cell.b[1][0] == 0.0 && cell.b[1][2] == 0.0 && cell.b[2][0] == 0.0 && cell.b[2][1] == 0.0;
    // the builder's own limits (scattering/__init__.py:248-268)
    const double g_max = 1.0 / (2.0 * M_PI / pop.q_max);
    const double g_min = pop.q_min > 0.0 ? 1.0 / (2.0 * M_PI / pop.q_min) : 0.0;
    // |G| of (h,k,l) is bi*sqrt(N) to within a few ulp, however it is rounded: N passes `|G| <= g_max` for certain when
    // N <= (g_max/bi)^2 (1 - 1e-13) and fails for certain above (1 + 1e-13); an integer inside the band is undecidable
    const double x = g_max / bi, y = g_min / bi;
    const double nf = x * x, mf = y * y;
    const double n_lo = floor(nf * (1.0 - 1.0e-13)), n_hi = floor(nf * (1.0 + 1.0e-13));
    const double m_lo = floor(mf * (1.0 - 1.0e-13)), m_hi = floor(mf * (1.0 + 1.0e-13));
    if (diag && n_lo == n_hi && m_lo == m_hi && n_lo >= 0.0 && n_lo < (double)T.n_shared && m_lo < 2.0e9) {
      const int key = (int)n_lo, tag = (int)m_lo + 1;
      const int old = atomicCAS(&T.shared_tag[key], 0, tag);
      if (old == 0 || old == tag) {
        row = T.shared_base + key;
        build = old == 0;
      }
    }
  }
  T.table_of[t] = row;
  T.build_flag[t] = build;
}

}  // namespace xrsd

extern "C" int64_t xrsd_table_scratch_bytes(int64_t cap_big) { return (int64_t)xrsd::xrsd_scratch_bytes(cap_big); }

extern "C" int64_t xrsd_table_smem_layout(int32_t max_cells, int32_t cap_list) {
  size_t off = ((size_t)max_cells * sizeof(uint16_t) + 15) & ~(size_t)15;
  off += (size_t)cap_list * sizeof(double);
  off += (size_t)cap_list * sizeof(int) * 2;
  off += ((size_t)cap_list + 1) * sizeof(int);
  off += (size_t)cap_list * sizeof(int);
  return (int64_t)((off + 15) & ~(size_t)15);
}

extern "C" cudaError_t xrsd_launch_tables(const xrsd_layout_t *layout, const double *d_params, int ldp, int batch,
                                          const xrsd_tables_t *tables, int max_cells, int cap_list,
                                          const uint8_t *d_active, cudaStream_t stream) {
  const size_t smem = (size_t)xrsd_table_smem_layout(max_cells, cap_list);
  cudaError_t e = xrsd::allow_max_dynamic_smem(xrsd::reflection_table_kernel);
  if (e != cudaSuccess) return e;
  const int n_tables = batch * layout->n_xtal;
  if (n_tables <= 0) return cudaSuccess;
  if (tables->grid_scratch != nullptr) {
    e = cudaMemsetAsync(tables->grid_slot_counter, 0, sizeof(int32_t), stream);
    if (e != cudaSuccess) return e;
  }
  if (tables->table_of != nullptr) {
    xrsd::table_key_kernel<<<(n_tables + 127) / 128, 128, 0, stream>>>(*layout, d_params, ldp, batch, *tables, d_active);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  xrsd::reflection_table_kernel<<<n_tables, xrsd::TABLE_THREADS, smem, stream>>>(*layout, d_params, ldp, batch, *tables,
                                                                                 max_cells, cap_list, d_active);
  return cudaGetLastError();
}
