// K3 -- per-base event statistics (nanoraw/resquiggle.py:60-70): mean and (population) standard
// deviation of the normalised signal inside each new segment, plus start / length.
//
// One CTA per read, one thread per base.  The normalised value of a sample depends only on its
// int16 DAC code, so the float64 division of nanoraw_helper.py:248 is done once per distinct code
// into a shared-memory table covering the unclipped range; codes outside it (outliers) are
// computed directly.  Sums follow numpy's add.reduce order (8-accumulator blocks of <= 128, pairwise
// above; -0.0 seed below 8 elements), so means and sds come out bit-identical to numpy's, although
// the contract only asks for 1e-9.
#include "nrq_common.cuh"

namespace nrq {

constexpr int kEvThreads = 256;
constexpr int kLut = 1024;

struct XAt {
  const int16_t* sig;
  const double* lut;
  int lut_lo;
  NormCoef coef;
  __device__ __forceinline__ double operator()(int i) const {
    int v = (int)sig[i];
    unsigned k = (unsigned)(v - lut_lo);
    return k < (unsigned)kLut ? lut[k] : norm_value(v, coef);
  }
};

struct SqDevAt {
  XAt x;
  double mean;
  __device__ __forceinline__ double operator()(int i) const {
    double d = __dsub_rn(x(i), mean);
    return __dmul_rn(d, d);
  }
};

// numpy DOUBLE_pairwise_sum for n <= 128
template <class F>
__device__ __forceinline__ double pw_block(const F& f, int off, int n) {
  if (n < 8) {
    double r = -0.0;
    for (int i = 0; i < n; ++i) r = __dadd_rn(r, f(off + i));
    return r;
  }
  double r0 = f(off), r1 = f(off + 1), r2 = f(off + 2), r3 = f(off + 3);
  double r4 = f(off + 4), r5 = f(off + 5), r6 = f(off + 6), r7 = f(off + 7);
  int i = 8;
  for (; i < n - (n & 7); i += 8) {
    r0 = __dadd_rn(r0, f(off + i));     r1 = __dadd_rn(r1, f(off + i + 1));
    r2 = __dadd_rn(r2, f(off + i + 2)); r3 = __dadd_rn(r3, f(off + i + 3));
    r4 = __dadd_rn(r4, f(off + i + 4)); r5 = __dadd_rn(r5, f(off + i + 5));
    r6 = __dadd_rn(r6, f(off + i + 6)); r7 = __dadd_rn(r7, f(off + i + 7));
  }
  double res = __dadd_rn(__dadd_rn(__dadd_rn(r0, r1), __dadd_rn(r2, r3)),
                         __dadd_rn(__dadd_rn(r4, r5), __dadd_rn(r6, r7)));
  for (; i < n; ++i) res = __dadd_rn(res, f(off + i));
  return res;
}

// n > 128: numpy splits at n/2 rounded down to a multiple of 8, recursively
template <class F>
__device__ __noinline__ double pw_large(const F& f, int off0, int n0) {
  int s_off[40], s_n[40], s_stage[40];
  double s_left[40];
  int sp = 0;
  double ret = 0.0;
  s_off[0] = off0; s_n[0] = n0; s_stage[0] = 0; sp = 1;
  while (sp > 0) {
    int k = sp - 1;
    int n = s_n[k], off = s_off[k];
    if (n <= 128) { ret = pw_block(f, off, n); --sp; continue; }
    int n2 = n / 2;
    n2 -= n2 & 7;
    if (s_stage[k] == 0) {
      s_stage[k] = 1;
      s_off[sp] = off; s_n[sp] = n2; s_stage[sp] = 0; ++sp;
    } else if (s_stage[k] == 1) {
      s_left[k] = ret;
      s_stage[k] = 2;
      s_off[sp] = off + n2; s_n[sp] = n - n2; s_stage[sp] = 0; ++sp;
    } else {
      ret = __dadd_rn(s_left[k], ret);
      --sp;
    }
  }
  return ret;
}

template <class F>
__device__ __forceinline__ double pw_sum(const F& f, int off, int n) {
  return n <= 128 ? pw_block(f, off, n) : pw_large(f, off, n);
}

__global__ void __launch_bounds__(kEvThreads)
k_event_table(nrq_batch b, nrq_params p, const double* __restrict__ sv, const int32_t* __restrict__ n_bases,
              const int32_t* __restrict__ new_segs, const int32_t* __restrict__ status,
              double* __restrict__ norm_mean, double* __restrict__ norm_stdev,
              uint32_t* __restrict__ ev_start, uint32_t* __restrict__ ev_length) {
  __shared__ double lut[kLut];
  const int r = blockIdx.x;
  if (status[r] != NRQ_ST_OK) return;
  const int tid = threadIdx.x;
  ReadGeom g = read_geom(b, r);
  NormCoef coef = load_coef(sv, r);
  const int nb = n_bases[r];
  const int32_t* segs = new_segs + b.align_off[r] + r;
  const size_t off = (size_t)b.align_off[r];

  // table of normalised values over the codes that survive clipping (or around a pivot sample)
  int lut_lo = (int)g.sig[g.n_norm / 2] - kLut / 2;
  if (coef.clip) {
    double a = coef.shift + coef.lo * coef.scale, c = coef.shift + coef.hi * coef.scale;
    double lo = fmin(a, c), hi = fmax(a, c);
    if (isfinite(lo) && isfinite(hi) && hi - lo < (double)(kLut - 4) && lo > -40000.0 && hi < 40000.0)
      lut_lo = (int)floor(lo) - 2;
  }
  for (int k = tid; k < kLut; k += kEvThreads) lut[k] = norm_value(lut_lo + k, coef);
  __syncthreads();

  XAt x{g.sig, lut, lut_lo, coef};
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  for (int bi = tid; bi < nb; bi += kEvThreads) {
    int s_lo = segs[bi], s_hi = segs[bi + 1];
    // np.split(norm_signal, new_segs[1:-1]): first piece starts at 0, last runs to the end (:64)
    int a0 = bi == 0 ? 0 : s_lo;
    int a1 = bi == nb - 1 ? g.n_norm : s_hi;
    int cnt = a1 - a0;
    double mean = qnan, sd = qnan;
    if (cnt > 0) {
      mean = __ddiv_rn(pw_sum(x, a0, cnt), (double)cnt);
      if (p.compute_sd) {
        SqDevAt sq{x, mean};
        sd = __dsqrt_rn(__ddiv_rn(pw_sum(sq, a0, cnt), (double)cnt));
      }
    }
    norm_mean[off + bi] = mean;
    norm_stdev[off + bi] = sd;
    ev_start[off + bi] = (uint32_t)s_lo;
    ev_length[off + bi] = (uint32_t)(s_hi - s_lo);
  }
}

}  // namespace nrq
