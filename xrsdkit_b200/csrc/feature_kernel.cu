// Pattern featuriser: the 21 scale-free numbers xrsdkit derives from a 1-D pattern (q, I).
//   reference: xrsdkit/tools/profiler.py:64-221 (profile_pattern), tools/peak_math.py:87-123
//   (humpness), tools/__init__.py:83-102 (pearson); called after every fit
//   (system/__init__.py:276) and when datasets are assembled (tools/ymltools.py:99).
//
// One CTA per pattern.  The pattern lives in shared memory for the whole evaluation:
//   I_s[n]   the intensities
//   A_s[n]   log I of the positive points (compacted), later standardised in place;
//            before that, scratch for the Fourier transform
//   B_s[n]   standardised q of the positive points; before that, twiddle factors
//   idx_s[n] positions of the positive points (uint16)
// i.e. 26 bytes per q point, 208 KB at the 8192-point limit.  Every feature is a reduction over
// the pattern (sums, arg-extrema, windowed correlations), so HBM traffic is one read of I per
// pattern and 168 bytes out; the arithmetic is dominated by the windowed correlations
// (n x 101).
#include <cuda_runtime.h>
#include "launch_util.cuh"
#include <math.h>
#include <stdint.h>

#include "../../include/xrsd_b200.h"

namespace {

constexpr int MAX_WARPS = 32;
constexpr int HUMP_W = 50;   // half window of the hump / trough score (peak_math.py:87, w=50)

struct RedScratch {
  double v[MAX_WARPS][16];
  double bv[MAX_WARPS];
  int bi[MAX_WARPS];
  int scan[MAX_WARPS];
};

// block-wide sums of K per-thread partials; result broadcast to every thread (fixed order => deterministic)
template <int K>
__device__ __forceinline__ void block_sums(double (&x)[K], RedScratch &R) {
  static_assert(K <= 16, "too many simultaneous sums");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x[k] += __shfl_xor_sync(0xffffffffu, x[k], o);
  }
  __syncthreads();                       // previous users of R are done
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) R.v[warp][k] = x[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double s = 0.0;
    for (int w = 0; w < nw; ++w) s += R.v[w][k];
    x[k] = s;
  }
}

// numpy argmax semantics: first occurrence of the maximum, a NaN beats everything (first NaN wins)
__device__ __forceinline__ bool beats(double a, int ia, double b, int ib) {
  if (ib < 0) return ia >= 0;
  if (ia < 0) return false;
  const bool na = isnan(a), nb = isnan(b);
  if (na || nb) return na && (!nb || ia < ib);
  return a > b || (a == b && ia < ib);
}

__device__ __forceinline__ void block_argmax(double &v, int &i, RedScratch &R) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (beats(ov, oi, v, i)) { v = ov; i = oi; }
  }
  __syncthreads();
  if (lane == 0) { R.bv[warp] = v; R.bi[warp] = i; }
  __syncthreads();
  v = R.bv[0]; i = R.bi[0];
  for (int w = 1; w < nw; ++w)
    if (beats(R.bv[w], R.bi[w], v, i)) { v = R.bv[w]; i = R.bi[w]; }
}

// Quadratic least squares through the points of a window, by orthogonal polynomials about the window mean:
//   y ~ c + b1 t + a (t^2 - alpha t - beta),  t = x - xbar.
// Three block-wide passes give (a, slope at t=0, xbar); same fit as np.polyfit(x, y, 2) in exact arithmetic
// (profiler.py:181-184) without forming the ill-conditioned Vandermonde matrix.
struct Parabola { double a, bt, xbar; };

__device__ __forceinline__ double vertex_of(const Parabola &p) { return p.xbar - p.bt / (2.0 * p.a); }

}  // namespace

// IdxT: uint16 positions while the pattern fits in shared memory (n <= XRSD_FEATURE_MAX_Q); longer patterns keep the
// same four arrays in a per-CTA region of global memory (`work`, L2-resident for the sizes met in practice) with 32-bit
// positions -- same code, same results, slower.
template <typename IdxT>
__global__ void __launch_bounds__(512, 2)     // 64 registers: two CTAs per SM up to 4096 q-points (measured +19 %)
feature_kernel(const double *__restrict__ q, int n, const double *__restrict__ I_all, int64_t ld_I, int batch,
               double *__restrict__ feat, int64_t ld_f, double *__restrict__ work, size_t work_stride) {
  extern __shared__ __align__(16) double smem[];
  const size_t n_al = ((size_t)n + 1) & ~(size_t)1;      // regions start on 16-byte boundaries (double2 views)
  double *base = work != nullptr ? work + (size_t)blockIdx.x * work_stride : smem;
  double *I_s = base, *A_s = base + n_al, *B_s = base + 2 * n_al;
  IdxT *idx_s = reinterpret_cast<IdxT *>(base + 3 * n_al);
  __shared__ RedScratch R;
  __shared__ int m_sh;
  const int tid = threadIdx.x, NT = blockDim.x;
  const bool pow2 = (n & (n - 1)) == 0 && n >= 4;

  for (int b = blockIdx.x; b < batch; b += gridDim.x) {
    const double *I = I_all + (size_t)b * ld_I;
    __syncthreads();
    for (int i = tid; i < n; i += NT) I_s[i] = I[i];
    __syncthreads();

    // ---- pass 1: plain sums, extrema, trapezoids, turning-point sum (profiler.py:91-104, 124-131, 140-149)
    const double q_first = q[0], q_last = q[n - 1];
    const double q_low = __dadd_rn(q_first, __dmul_rn(0.1, __dsub_rn(q_last, q_first)));
    double s1[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    // 0 sum I, 1 sum q, 2 sum q^2, 3 sum e^q, 4 sum e^-q, 5 sum I(low q), 6 count(low q), 7 sum dq*It, 8 sum qc*dq*It, 9 turning sum
    double vmax = 0.0, vmin = 0.0;
    int imax = -1, imin = -1, npos = 0;
    for (int i = tid; i < n; i += NT) {
      const double Ii = I_s[i], qi = q[i];
      s1[0] += Ii; s1[1] += qi; s1[2] += qi * qi; s1[3] += exp(qi); s1[4] += exp(-qi);
      if (qi < q_low) { s1[5] += Ii; s1[6] += 1.0; }
      if (beats(Ii, i, vmax, imax)) { vmax = Ii; imax = i; }
      if (beats(-Ii, i, vmin, imin)) { vmin = -Ii; imin = i; }
      npos += (Ii > 0.0);
      if (i + 1 < n) {
        const double In = I_s[i + 1], qn = q[i + 1];
        const double dq = qn - qi, It = 0.5 * (In + Ii);
        s1[7] += dq * It;
        s1[8] += 0.5 * (qn + qi) * dq * It;
        const double d = In - Ii;
        if (i == 0 || __dmul_rn(d, Ii - I_s[i - 1]) < 0.0) s1[9] += fabs(d);
      }
    }
    block_sums(s1, R);
    block_argmax(vmax, imax, R);
    block_argmax(vmin, imin, R);
    const double I_max = vmax, I_min = -vmin, nn = (double)n;
    const double I_mean = s1[0] / nn, q_mean = s1[1] / nn;
    const double q_at_max = q[imax];
    const double near_lo = 0.9 * q_at_max, near_hi = 1.1 * q_at_max;

    // ---- pass 2: centred sums (std, Pearson correlations; profiler.py:95-99, 152-156), mean near the maximum (:119-120)
    double s2[11] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    // 0 dI^2, 1 dq^2, 2 near sum, 3 near count, 4 dq dI, 5 d(q^2) dI, 6 d(q^2)^2, 7 d(e^q) dI, 8 d(e^q)^2, 9 d(e^-q) dI, 10 d(e^-q)^2
    const double mq2 = s1[2] / nn, meq = s1[3] / nn, mieq = s1[4] / nn;
    for (int i = tid; i < n; i += NT) {
      const double Ii = I_s[i], qi = q[i];
      const double dI = Ii - I_mean, dq = qi - q_mean;
      s2[0] += dI * dI; s2[1] += dq * dq; s2[4] += dq * dI;
      if (qi > near_lo && qi < near_hi) { s2[2] += Ii; s2[3] += 1.0; }
      const double d2 = qi * qi - mq2, de = exp(qi) - meq, di = exp(-qi) - mieq;
      s2[5] += d2 * dI; s2[6] += d2 * d2; s2[7] += de * dI; s2[8] += de * de; s2[9] += di * dI; s2[10] += di * di;
    }
    block_sums(s2, R);
    const double I_std = sqrt(s2[0] / nn), q_std = sqrt(s2[1] / nn);
    const double sdI = sqrt(s2[0]);

    // ---- Fourier amplitude centroid over the positive frequencies (profiler.py:158-172); A_s/B_s are scratch here
    double fsum[2] = {0.0, 0.0};       // sum dr*At, sum rc*dr*At
    {
      const int K = (n - 1) / 2;       // frequencies k = 1..K have fftfreq > 0
      const double rstep = 1.0 / nn;   // numpy.fft.fftfreq: k * (1/n)
      if (pow2) {
        // real transform through a half-length complex FFT: z_j = I_2j + i I_2j+1 (radix 2, decimation in time)
        const int N = n >> 1, logN = 31 - __clz(N);
        double2 *z = reinterpret_cast<double2 *>(A_s);     // N complex
        double2 *tw = reinterpret_cast<double2 *>(B_s);    // N/2 complex: exp(-2 pi i j / N)
        for (int j = tid; j < N; j += NT) {
          const int r = (int)(__brev((unsigned)j) >> (32 - logN));
          z[r] = make_double2(I_s[2 * j], I_s[2 * j + 1]);
          if (j < (N >> 1)) {
            double s, c;
            sincospi(-2.0 * (double)j / (double)N, &s, &c);
            tw[j] = make_double2(c, s);
          }
        }
        __syncthreads();
        for (int st = 0; st < logN; ++st) {
          const int half = 1 << st, tstride = N >> (st + 1);
          for (int t = tid; t < (N >> 1); t += NT) {
            const int grp = t >> st, pos = t & (half - 1);
            const int i0 = (grp << (st + 1)) + pos, i1 = i0 + half;
            const double2 w = tw[pos * tstride], u = z[i0], v = z[i1];
            const double xr = v.x * w.x - v.y * w.y, xi = v.x * w.y + v.y * w.x;
            z[i0] = make_double2(u.x + xr, u.y + xi);
            z[i1] = make_double2(u.x - xr, u.y - xi);
          }
          __syncthreads();
        }
        // F_k = (Z_k + conj Z_{N-k})/2 - (i/2) e^{-2 pi i k/n} (Z_k - conj Z_{N-k});  amplitudes of k and k+1 per term
        auto amp = [&](int k) {
          const double2 a = z[k], c = z[(N - k) & (N - 1)];
          const double er = 0.5 * (a.x + c.x), ei = 0.5 * (a.y - c.y);     // even part
          const double orr = 0.5 * (a.x - c.x), oi = 0.5 * (a.y + c.y);    // (Z_k - conj Z_{N-k})/2
          double s, cc;
          sincospi(-2.0 * (double)k / nn, &s, &cc);
          // -i * (cc + i s) * (orr + i oi) = (s*orr + cc*oi) + i (s*oi - cc*orr)
          return hypot(er + (s * orr + cc * oi), ei + (s * oi - cc * orr));
        };
        for (int k = 1 + tid; k < K; k += NT) {
          const double r0 = (double)k * rstep, r1 = (double)(k + 1) * rstep;
          const double dr = r1 - r0, At = 0.5 * (amp(k + 1) + amp(k));
          fsum[0] += dr * At;
          fsum[1] += 0.5 * (r1 + r0) * dr * At;
        }
      } else {
        // any other length: direct transform with an exact table of the n-th roots of unity
        double2 *w = reinterpret_cast<double2 *>(A_s);     // n complex over A_s and B_s
        for (int j = tid; j < n; j += NT) {
          double s, c;
          sincospi(-2.0 * (double)j / nn, &s, &c);
          w[j] = make_double2(c, s);
        }
        __syncthreads();
        auto amp = [&](int k) {
          double re = 0.0, im = 0.0;
          int mm = 0;
          for (int j = 0; j < n; ++j) {
            const double2 ww = w[mm];
            const double v = I_s[j];
            re = fma(v, ww.x, re);
            im = fma(v, ww.y, im);
            mm += k;
            if (mm >= n) mm -= n;
          }
          return hypot(re, im);
        };
        // contiguous k ranges per thread so that each amplitude is computed once (plus one per range boundary)
        const int per = (K - 1 + NT - 1) / NT, k_lo = 1 + tid * per, k_hi = min(K, k_lo + per);
        double a0 = (k_lo < k_hi) ? amp(k_lo) : 0.0;
        for (int k = k_lo; k < k_hi; ++k) {
          const double a1 = amp(k + 1);
          const double r0 = (double)k * rstep, r1 = (double)(k + 1) * rstep;
          const double dr = r1 - r0, At = 0.5 * (a1 + a0);
          fsum[0] += dr * At;
          fsum[1] += 0.5 * (r1 + r0) * dr * At;
          a0 = a1;
        }
      }
      block_sums(fsum, R);
    }
    __syncthreads();

    // ---- positive points, in order (profiler.py:105-110): idx_s, A_s = log I, B_s = standardised q
    {
      // per-thread contiguous runs keep the order; block scan over run counts
      const int per = (n + NT - 1) / NT, lo = min(n, tid * per), hi = min(n, lo + per);
      int cnt = 0;
      for (int i = lo; i < hi; ++i) cnt += (I_s[i] > 0.0);
      const int lane = tid & 31, warp = tid >> 5, nw = (NT + 31) >> 5;
      int inc = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      if (lane == 31) R.scan[warp] = inc;
      __syncthreads();
      int base = 0;
      for (int w = 0; w < warp; ++w) base += R.scan[w];
      if (tid == NT - 1) m_sh = base + inc;
      int pos = base + inc - cnt;
      for (int i = lo; i < hi; ++i)
        if (I_s[i] > 0.0) idx_s[pos++] = (IdxT)i;
      (void)nw;
      __syncthreads();
    }
    const int m = m_sh;
    const double mm_d = (double)m;
    for (int k = tid; k < m; k += NT) {
      const int i = idx_s[k];
      A_s[k] = log(I_s[i]);
      B_s[k] = (q[i] - q_mean) / q_std;
    }
    __syncthreads();

    // ---- log-intensity sums (profiler.py:111-117, 132-137, 143-150)
    double s3[4] = {0, 0, 0, 0};   // sum L, sum dq*Lt, sum qc*dq*Lt, turning sum
    double lmax = 0.0, lmin = 0.0;
    int ilmax = -1, ilmin = -1;
    for (int k = tid; k < m; k += NT) {
      const double Lk = A_s[k];
      s3[0] += Lk;
      if (beats(Lk, k, lmax, ilmax)) { lmax = Lk; ilmax = k; }
      if (beats(-Lk, k, lmin, ilmin)) { lmin = -Lk; ilmin = k; }
      if (k + 1 < m) {
        const double Ln = A_s[k + 1], qk = q[idx_s[k]], qn = q[idx_s[k + 1]];
        const double dq = qn - qk, Lt = 0.5 * (Ln + Lk);
        s3[1] += dq * Lt;
        s3[2] += 0.5 * (qn + qk) * dq * Lt;
        const double d = Ln - Lk;
        if (k == 0 || __dmul_rn(d, Lk - A_s[k - 1]) < 0.0) s3[3] += fabs(d);
      }
    }
    block_sums(s3, R);
    block_argmax(lmax, ilmax, R);
    block_argmax(lmin, ilmin, R);
    const double L_max = lmax, L_min = -lmin, L_mean = s3[0] / mm_d;
    double s4[1] = {0.0};
    for (int k = tid; k < m; k += NT) { const double d = A_s[k] - L_mean; s4[0] += d * d; }
    block_sums(s4, R);
    const double L_std = sqrt(s4[0] / mm_d);
    __syncthreads();
    for (int k = tid; k < m; k += NT) A_s[k] = (A_s[k] - L_mean) / L_std;       // standardised log I
    __syncthreads();

    // ---- hump / trough scores (peak_math.py:104-114): windowed Pearson r of y against (x - x_k)^2
    double hbest = 0.0, tbest = 0.0;
    int hidx = -1, tidx = -1;
    for (int k = tid; k < m; k += NT) {
      const int lo = max(0, k - HUMP_W), hi = min(m, k + HUMP_W + 1);
      const double xk = B_s[k], cntw = (double)(hi - lo);
      // one pass over the window with the values taken relative to the centre point (y - y_k, (x - x_k)^2): the
      // centred sums follow from the raw ones with a cancellation factor of order one
      const double yk = A_s[k];
      double sy = 0.0, sx = 0.0, syy = 0.0, sxx = 0.0, sxy = 0.0;
      for (int j = lo; j < hi; ++j) {
        const double dx = B_s[j] - xk, x2 = dx * dx, yy = A_s[j] - yk;
        sy += yy; sx += x2;
        syy = fma(yy, yy, syy); sxx = fma(x2, x2, sxx); sxy = fma(yy, x2, sxy);
      }
      syy -= sy * sy / cntw; sxx -= sx * sx / cntw; sxy -= sy * sx / cntw;
      if (syy < 0.0) syy = 0.0;
      const double r = sxy / (sqrt(syy) * sqrt(sxx));
      const double hump = -1.0 * A_s[k] * r, trough = sqrt(syy / cntw) * r;
      if (beats(hump, k, hbest, hidx)) { hbest = hump; hidx = k; }
      if (beats(trough, k, tbest, tidx)) { tbest = trough; tidx = k; }
    }
    block_argmax(hbest, hidx, R);
    block_argmax(tbest, tidx, R);

    // ---- parabolas through the points within 0.1 std(q) of the best hump / trough (profiler.py:176-196)
    Parabola P[4];   // hump on I, trough on I, hump on log I, trough on log I
    {
      const double half = __dmul_rn(0.1, q_std);
      const double qh = (hidx >= 0) ? q[idx_s[hidx]] : 0.0, qt = (tidx >= 0) ? q[idx_s[tidx]] : 0.0;
      const double h_lo = __dsub_rn(qh, half), h_hi = __dadd_rn(qh, half);
      const double t_lo = __dsub_rn(qt, half), t_hi = __dadd_rn(qt, half);
      double a1[4] = {0, 0, 0, 0};            // count, sum x  for the hump window, then for the trough window
      for (int k = tid; k < m; k += NT) {
        const double qk = q[idx_s[k]], x = B_s[k];
        if (qk > h_lo && qk < h_hi) { a1[0] += 1.0; a1[1] += x; }
        if (qk > t_lo && qk < t_hi) { a1[2] += 1.0; a1[3] += x; }
      }
      block_sums(a1, R);
      const double xb[2] = {a1[1] / a1[0], a1[3] / a1[2]}, cw[2] = {a1[0], a1[2]};
      double a2[12];                          // per window: sum t^2, t^3, yI, yI t, yL, yL t   (t = x - mean x)
#pragma unroll
      for (int i = 0; i < 12; ++i) a2[i] = 0.0;
      for (int k = tid; k < m; k += NT) {
        const double qk = q[idx_s[k]], x = B_s[k];
        const double yI = (I_s[idx_s[k]] - I_mean) / I_std, yL = A_s[k];
        if (qk > h_lo && qk < h_hi) {
          const double t = x - xb[0];
          a2[0] += t * t; a2[1] += t * t * t; a2[2] += yI; a2[3] += yI * t; a2[4] += yL; a2[5] += yL * t;
        }
        if (qk > t_lo && qk < t_hi) {
          const double t = x - xb[1];
          a2[6] += t * t; a2[7] += t * t * t; a2[8] += yI; a2[9] += yI * t; a2[10] += yL; a2[11] += yL * t;
        }
      }
      block_sums(a2, R);
      const double alpha[2] = {a2[1] / a2[0], a2[7] / a2[6]}, beta[2] = {a2[0] / cw[0], a2[6] / cw[1]};
      double a3[6] = {0, 0, 0, 0, 0, 0};      // per window: sum p2^2, sum yI p2, sum yL p2,  p2 = t^2 - alpha t - beta
      for (int k = tid; k < m; k += NT) {
        const double qk = q[idx_s[k]], x = B_s[k];
        const double yI = (I_s[idx_s[k]] - I_mean) / I_std, yL = A_s[k];
        if (qk > h_lo && qk < h_hi) {
          const double t = x - xb[0], p2 = t * t - alpha[0] * t - beta[0];
          a3[0] += p2 * p2; a3[1] += yI * p2; a3[2] += yL * p2;
        }
        if (qk > t_lo && qk < t_hi) {
          const double t = x - xb[1], p2 = t * t - alpha[1] * t - beta[1];
          a3[3] += p2 * p2; a3[4] += yI * p2; a3[5] += yL * p2;
        }
      }
      block_sums(a3, R);
#pragma unroll
      for (int w = 0; w < 2; ++w) {
#pragma unroll
        for (int v = 0; v < 2; ++v) {         // v = 0: standardised I, v = 1: standardised log I
          Parabola &p = P[2 * v + w];
          p.a = a3[3 * w + 1 + v] / a3[3 * w];
          p.bt = a2[6 * w + 3 + 2 * v] / a2[6 * w] - p.a * alpha[w];   // slope at t = 0
          p.xbar = xb[w];
        }
      }
    }

    if (tid == 0) {
      double *f = feat + (size_t)b * ld_f;
      f[0] = I_max / I_mean;                                  // Imax_over_Imean
      f[1] = (s1[5] / s1[6]) / I_mean;                        // Ilowq_over_Imean
      f[2] = I_max / (s2[2] / s2[3]);                         // Imax_sharpness
      f[3] = s1[9] / (I_max - I_min);                         // I_fluctuation
      f[4] = s3[3] / (L_max - L_min);                         // logI_fluctuation
      f[5] = L_max / L_std;                                   // logI_max_over_std
      f[6] = fsum[1] / fsum[0];                               // r_fftIcentroid
      f[7] = s1[8] / s1[7];                                   // q_Icentroid
      f[8] = s3[2] / s3[1];                                   // q_logIcentroid
      f[9] = s2[4] / (sqrt(s2[1]) * sdI);                     // pearson_q
      f[10] = s2[5] / (sqrt(s2[6]) * sdI);                    // pearson_q2
      f[11] = s2[7] / (sqrt(s2[8]) * sdI);                    // pearson_expq
      f[12] = s2[9] / (sqrt(s2[10]) * sdI);                   // pearson_invexpq
      f[13] = vertex_of(P[0]) * q_std + q_mean;               // q_best_hump
      f[14] = vertex_of(P[1]) * q_std + q_mean;               // q_best_trough
      f[15] = fabs(1.0 / P[0].a) * q_std;                     // best_hump_qwidth
      f[16] = fabs(1.0 / P[1].a) * q_std;                     // best_trough_qwidth
      f[17] = vertex_of(P[2]) * q_std + q_mean;               // q_best_hump_log
      f[18] = vertex_of(P[3]) * q_std + q_mean;               // q_best_trough_log
      f[19] = fabs(1.0 / P[2].a) * q_std;                     // best_hump_qwidth_log
      f[20] = fabs(1.0 / P[3].a) * q_std;                     // best_trough_qwidth_log
    }
  }
}

extern "C" size_t xrsd_feature_smem_bytes_impl(int n_q) { return (((size_t)n_q + 1) & ~(size_t)1) * 24 + (size_t)n_q * 2 + 16; }

extern "C" cudaError_t xrsd_launch_features(const double *d_q, int n_q, const double *d_I, int64_t ld_I, int batch,
                                 double *d_feat, int64_t ld_f, cudaStream_t stream) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (n_q <= XRSD_FEATURE_MAX_Q) {
    const size_t smem = xrsd_feature_smem_bytes_impl(n_q);
    // opt in on every launch: the attribute is per device and static shared memory counts against the 48 KB default too
    cudaError_t e = xrsd::allow_max_dynamic_smem(feature_kernel<unsigned short>);
    if (e != cudaSuccess) return e;
    const int threads = n_q >= 2048 ? 512 : 256;
    const int grid = batch < sms * 16 ? batch : sms * 16;
    feature_kernel<unsigned short><<<grid, threads, smem, stream>>>(d_q, n_q, d_I, ld_I, batch, d_feat, ld_f, nullptr, 0);
    return cudaGetLastError();
  }
  // patterns beyond the shared-memory limit: the working set of every CTA in a stream-ordered global allocation
  const size_t n_al = ((size_t)n_q + 1) & ~(size_t)1;
  const size_t stride = 3 * n_al + (((size_t)n_q * sizeof(unsigned) + 15) / 16) * 2;       // doubles, 16-byte multiple
  const int grid = batch < sms * 2 ? batch : sms * 2;
  double *work = nullptr;
  cudaError_t e = cudaMallocAsync((void **)&work, stride * sizeof(double) * (size_t)grid, stream);
  if (e != cudaSuccess) return e;
  feature_kernel<unsigned><<<grid, 512, 0, stream>>>(d_q, n_q, d_I, ld_I, batch, d_feat, ld_f, work, stride);
  e = cudaGetLastError();
  const cudaError_t e2 = cudaFreeAsync(work, stream);
  return e != cudaSuccess ? e : e2;
}
