// K1 -- the order statistics of normalize_raw_signal (nanoraw/nanoraw_helper.py:240-256).
//
// One CTA per read, eight CTAs per SM (the kernel is a latency-bound stream of shared-memory atomics: occupancy
// is what buys bandwidth, so the histogram window is kept to 8 KB).  Fast path: a single streaming pass builds a
// shared-memory histogram of the int16 DAC values over a window of kWin values centred on a pivot (median of
// three samples); both medians of the
// 'median' normalisation are then exact integer selects on that histogram (SURVEY.md section 8a
// exactness notes: shift is in Z/2, |raw-shift| in Z/2, median(x) == +0.0, mad from the same two
// middle order statistics).  Reads whose needed ranks fall outside the window take the general
// path: exact selection by bisection on the key domain with one counting pass per step.
//
// K1b -- the normalised / winsorised signal itself (nanoraw_helper.py:248, 257-262) as float64.
#include "nrq_common.cuh"

namespace nrq {

#ifndef NRQ_NORM_THREADS
#define NRQ_NORM_THREADS 256
#endif
constexpr int kNormThreads = NRQ_NORM_THREADS;
#ifndef NRQ_NORM_WIN
#define NRQ_NORM_WIN 2048
#endif
#ifndef NRQ_NORM_MINB
#define NRQ_NORM_MINB 8
#endif
constexpr int kWin = NRQ_NORM_WIN;  // histogram window, in DAC values (4 bytes of shared memory each)

// ---------------------------------------------------------------- block primitives
__device__ __forceinline__ long long block_sum(long long v, long long* sh) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane_id() == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  long long t = 0;
#pragma unroll
  for (int w = 0; w < kNormThreads / 32; ++w) t += sh[w];
  return t;
}

// exclusive scan over the kNormThreads per-thread values; returns (exclusive prefix, total)
__device__ __forceinline__ void block_excl_scan(unsigned v, unsigned* sh, unsigned& excl, unsigned& total) {
  unsigned inc = v;
  for (int o = 1; o < 32; o <<= 1) {
    unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
    if ((int)lane_id() >= o) inc += t;
  }
  __syncthreads();
  if (lane_id() == 31) sh[threadIdx.x >> 5] = inc;
  __syncthreads();
  unsigned base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kNormThreads / 32; ++w) {
    unsigned s = sh[w];
    if (w < (int)(threadIdx.x >> 5)) base += s;
    tot += s;
  }
  excl = base + inc - v;
  total = tot;
}

// visit every sample of the read once with 128-bit loads where alignment allows
template <class F>
__device__ __forceinline__ void for_each_sample(const int16_t* __restrict__ sig, int n, F f) {
  const int tid = threadIdx.x;
  int head = (int)(((16u - (unsigned)((uintptr_t)sig & 15u)) & 15u) >> 1);
  if (head > n) head = n;
  for (int i = tid; i < head; i += kNormThreads) f((int)sig[i]);
  const int4* v = reinterpret_cast<const int4*>(sig + head);
  int nv = (n - head) >> 3;
  auto eight = [&](const int4& q) {
    f((int)(short)(q.x & 0xffff)); f((int)(short)(q.x >> 16));
    f((int)(short)(q.y & 0xffff)); f((int)(short)(q.y >> 16));
    f((int)(short)(q.z & 0xffff)); f((int)(short)(q.z >> 16));
    f((int)(short)(q.w & 0xffff)); f((int)(short)(q.w >> 16));
  };
  int i = tid;
  // four 128-bit loads in flight per thread before any of them is consumed
  for (; i + 3 * kNormThreads < nv; i += 4 * kNormThreads) {
    int4 q0 = __ldg(v + i), q1 = __ldg(v + i + kNormThreads);
    int4 q2 = __ldg(v + i + 2 * kNormThreads), q3 = __ldg(v + i + 3 * kNormThreads);
    eight(q0); eight(q1); eight(q2); eight(q3);
  }
  for (; i < nv; i += kNormThreads) eight(__ldg(v + i));
  for (int i = head + (nv << 3) + tid; i < n; i += kNormThreads) f((int)sig[i]);
}

// Find the bins holding ranks q1 <= q2 (0-based) of a binned multiset; cnt(i) = count in bin i,
// nbins <= 32 * kNormThreads.  Every thread returns the same answer.  false if rank q2 is not covered.
template <class C>
__device__ __forceinline__ bool block_select2(C cnt, int nbins, unsigned q1, unsigned q2,
                                              unsigned* sh_scan, int* sh_res, int& i1, int& i2) {
  const int t = threadIdx.x;
  unsigned local = 0;
#pragma unroll 4
  for (int j = 0; j < 32; ++j) {
    int i = 32 * t + ((j + t) & 31);  // rotated: conflict-free shared-memory reads
    if (i < nbins) local += cnt(i);
  }
  unsigned excl, total;
  block_excl_scan(local, sh_scan, excl, total);
  if (total <= q2) return false;  // uniform
  if (q1 >= excl && q1 < excl + local) {
    unsigned c = excl;
    int i = 32 * t;
    for (;; ++i) { c += cnt(i); if (c > q1) break; }
    sh_res[0] = i;
  }
  if (q2 >= excl && q2 < excl + local) {
    unsigned c = excl;
    int i = 32 * t;
    for (;; ++i) { c += cnt(i); if (c > q2) break; }
    sh_res[1] = i;
  }
  __syncthreads();
  i1 = sh_res[0];
  i2 = sh_res[1];
  __syncthreads();
  return true;
}

// exact k-th smallest key by bisection on [klo, khi]; one counting pass over the read per step
template <class K>
__device__ __forceinline__ unsigned long long block_bisect(const int16_t* sig, int n, K key, long long k,
                                                           unsigned long long klo, unsigned long long khi,
                                                           long long* sh) {
  while (klo < khi) {
    unsigned long long mid = klo + ((khi - klo) >> 1);
    long long c = 0;
    for_each_sample(sig, n, [&](int v) { c += (key(v) <= mid) ? 1 : 0; });
    c = block_sum(c, sh);
    if (c > k) khi = mid; else klo = mid + 1;
  }
  return klo;
}

#ifndef NRQ_NORM_BULK
#define NRQ_NORM_BULK 0
#endif
#if NRQ_NORM_BULK
// Experiment (profiles/README.md, "bulk-async staging"): the read streamed into shared memory by one thread with
// 1-D bulk copies (cp.async.bulk + mbarrier complete_tx, the TMA engine's non-tensor form) in a two-deep ring of
// kBulkTile bytes, the other threads taking their samples from there.  The loads leave the issue slots (one
// UBLKCP per tile instead of a LDG.128 per eight samples and thread); the shared-memory atomics stay.
constexpr int kBulkTile = 4096;  // bytes per stage = 2048 samples = 8 samples per thread
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void bar_wait(unsigned long long* bar, unsigned parity) {
  unsigned done = 0;
  while (!done)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(smem_addr(bar)), "r"(parity) : "memory");
}
// visit every sample once: head and tail (to 16-byte alignment / tile size) by plain loads, the middle through the ring
template <class F>
__device__ __forceinline__ void for_each_sample_bulk(const int16_t* __restrict__ sig, int n, int4* ring,
                                                     unsigned long long* bars, F f) {
  const int tid = threadIdx.x;
  int head = (int)(((16u - (unsigned)((uintptr_t)sig & 15u)) & 15u) >> 1);
  if (head > n) head = n;
  for (int i = tid; i < head; i += kNormThreads) f((int)sig[i]);
  const int16_t* mid = sig + head;
  const int tiles = (n - head) / (kBulkTile / 2);
  constexpr int kTileVec = kBulkTile / 16;  // int4 per tile
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bars)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bars + 1)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    if (tiles > 0) bulk_load(ring, mid, kBulkTile, bars);
    if (tiles > 1) bulk_load(ring + kTileVec, mid + kBulkTile / 2, kBulkTile, bars + 1);
  }
  for (int t = 0; t < tiles; ++t) {
    const int s = t & 1;
    bar_wait(bars + s, (unsigned)((t >> 1) & 1));
    for (int i = tid; i < kTileVec; i += kNormThreads) {
      const int4 q = ring[s * kTileVec + i];
      f((int)(short)(q.x & 0xffff)); f((int)(short)(q.x >> 16));
      f((int)(short)(q.y & 0xffff)); f((int)(short)(q.y >> 16));
      f((int)(short)(q.z & 0xffff)); f((int)(short)(q.z >> 16));
      f((int)(short)(q.w & 0xffff)); f((int)(short)(q.w >> 16));
    }
    __syncthreads();  // everybody is done with stage s
    if (tid == 0 && t + 2 < tiles) bulk_load(ring + s * kTileVec, mid + (size_t)(t + 2) * (kBulkTile / 2), kBulkTile, bars + s);
  }
  for (int i = head + tiles * (kBulkTile / 2) + tid; i < n; i += kNormThreads) f((int)sig[i]);
  __syncthreads();
  if (tid == 0) {
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_addr(bars)));
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_addr(bars + 1)));
  }
}
#endif

__device__ __forceinline__ unsigned long long dbits(double x) {
  return (unsigned long long)__double_as_longlong(x);
}

// ---------------------------------------------------------------- K1
__global__ void __launch_bounds__(kNormThreads, NRQ_NORM_MINB)
k_norm_stats(nrq_batch b, nrq_params p, int32_t* __restrict__ status, double* __restrict__ sv_out) {
  // hist[-1] counts the samples below the window, hist[kWin] those above it
  __shared__ unsigned hist_all[kWin + 2];
  unsigned* const hist = hist_all + 1;
  __shared__ long long sh_red[kNormThreads / 32];
  __shared__ unsigned sh_scan[kNormThreads / 32];
  __shared__ int sh_res[2];
  __shared__ double sh_d[4];
  __shared__ int sh_flag;
#if NRQ_NORM_BULK
  __shared__ __align__(128) int4 bulk_ring[2 * kBulkTile / 16];
  __shared__ __align__(8) unsigned long long bulk_bars[2];
#endif

  const int r = blockIdx.x;
  const int tid = threadIdx.x;
  double* sv = sv_out + 4 * (size_t)r;
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  if (status[r] != NRQ_ST_OK) {
    if (tid < 4) sv[tid] = qnan;
    return;
  }
  ReadGeom g = read_geom(b, r);
  const int n = g.n_norm;
  if (n <= 0 || g.n_obs <= 0) {
    if (tid < 4) sv[tid] = qnan;
    if (tid == 0) status[r] = NRQ_ST_BAD_INPUT;
    return;
  }
  const int16_t* sig = g.sig;
  const long long k1 = (n - 1) / 2, k2 = n / 2;  // np.median: mean of the two middle elements
  const bool force_general = (p.debug_flags & 1) != 0;
  const bool need_median_norm = p.norm_type == NRQ_NORM_MEDIAN;
  const bool need_outlier = p.use_outlier != 0;  // takes precedence over given limits (nanoraw_helper.py:252)
  const bool given_limits = p.given_limits && !need_outlier && b.scale_values_in != nullptr;
  const bool need_hist = need_median_norm || need_outlier;

  // ---- fast path: windowed histogram
  int lo_w = 0;
  long long below = 0, above = 0;
  bool win_ok = false;
  if (need_hist && !force_general) {
    for (int i = tid; i < kWin + 2; i += kNormThreads) hist_all[i] = 0;
    {  // pivot: median of three samples, so that a single spike cannot misplace the window
      int a = sig[n / 4], c = sig[n / 2], e = sig[(3 * (long long)n) / 4];
      int piv = max(min(a, c), min(max(a, c), e));
      lo_w = piv - kWin / 2;
    }
    __syncthreads();
    // one clamped, branch-free shared-memory increment per sample; the table's shared address is formed once
    // (left to the compiler it is re-derived per sample, and the two-sided test costs four branch instructions)
    unsigned hist_sa = (unsigned)__cvta_generic_to_shared(hist_all);
    asm volatile("" : "+r"(hist_sa));
    auto count_sample = [&](int v) {
      const int slot = min(max(v - lo_w, -1), kWin) + 1;
      asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(hist_sa + 4u * (unsigned)slot) : "memory");
    };
#if NRQ_NORM_BULK
    for_each_sample_bulk(sig, n, bulk_ring, bulk_bars, count_sample);
#else
    for_each_sample(sig, n, count_sample);
#endif
    __syncthreads();
    below = hist[-1];
    above = hist[kWin];
    win_ok = (below <= k1) && (k2 < n - above);
  }

  int m1 = 0, m2 = 0;  // raw order statistics at ranks k1, k2
  bool have_m = false;
  if (need_hist) {
    if (win_ok) {
      int i1, i2;
      have_m = block_select2([&](int i) { return hist[i]; }, kWin, (unsigned)(k1 - below),
                             (unsigned)(k2 - below), sh_scan, sh_res, i1, i2);
      m1 = lo_w + i1;
      m2 = lo_w + i2;
    }
    if (!have_m) {
      auto key = [](int v) { return (unsigned long long)(v + 32768); };
      m1 = (int)block_bisect(sig, n, key, k1, 0ull, 65535ull, sh_red) - 32768;
      m2 = (k2 == k1) ? m1 : (int)block_bisect(sig, n, key, k2, 0ull, 65535ull, sh_red) - 32768;
    }
  }

  double shift, scale, lower = qnan, upper = qnan;
  int fail = NRQ_ST_OK;
  if (need_median_norm) {
    // shift = median(raw) = (m1 + m2) / 2 exactly          (nanoraw_helper.py:241)
    const int shift2 = m1 + m2;
    shift = (double)shift2 * 0.5;
    // scale = median(|raw - shift|): select on dd = |2 raw - 2 shift|   (nanoraw_helper.py:242)
    long long U = -1, V = -1;
    if (win_ok && have_m) {
      const int odd = shift2 & 1;
      // floor / ceil of shift (shift2 may be negative)
      const int c_lo_f = (shift2 >= 0) ? shift2 / 2 : -((1 - shift2) / 2);
      const int c_hi_f = c_lo_f + odd;
      int jmax = min(c_lo_f - lo_w, lo_w + kWin - 1 - c_hi_f);
      if (jmax >= 0) {
        int i1, i2;
        auto folded = [&](int j) -> unsigned {
          unsigned a = hist[c_lo_f - lo_w - j];
          unsigned c = hist[c_hi_f - lo_w + j];
          return (odd || j > 0) ? a + c : a;
        };
        if (block_select2(folded, jmax + 1, (unsigned)k1, (unsigned)k2, sh_scan, sh_res, i1, i2)) {
          U = 2ll * i1 + odd;
          V = 2ll * i2 + odd;
        }
      }
    }
    if (U < 0) {
      auto key = [shift2](int v) { long long d = 2ll * v - shift2; return (unsigned long long)(d < 0 ? -d : d); };
      U = (long long)block_bisect(sig, n, key, k1, 0ull, 131072ull, sh_red);
      V = (k2 == k1) ? U : (long long)block_bisect(sig, n, key, k2, 0ull, 131072ull, sh_red);
    }
    scale = (double)(U + V) * 0.25;
    if (U + V == 0) {
      // x = (raw - shift) / 0 under np.seterr(all='raise') (resquiggle.py:8): 'divide by zero' if any
      // sample differs from shift, otherwise only 0/0 -> 'invalid value'
      long long neq = 0;
      for_each_sample(sig, n, [&](int v) { neq += (2 * v != shift2) ? 1 : 0; });
      neq = block_sum(neq, sh_red);
      fail = neq > 0 ? NRQ_ST_FPE_DIVIDE : NRQ_ST_FPE_INVALID;
    } else if (need_outlier) {
      // med = median(x) == +0.0; mad from the same two order statistics (nanoraw_helper.py:253-254)
      double xu = __ddiv_rn((double)U * 0.5, scale);
      double xv = __ddiv_rn((double)V * 0.5, scale);
      double mad = (k1 == k2) ? xu : __dmul_rn(__dadd_rn(xu, xv), 0.5);
      double thr = __dmul_rn(mad, p.outlier_thresh);
      lower = __dsub_rn(0.0, thr);   // :255
      upper = __dadd_rn(0.0, thr);   // :256
    } else if (given_limits) {
      lower = b.scale_values_in[4 * (size_t)r + 2];
      upper = b.scale_values_in[4 * (size_t)r + 3];
    }
  } else {
    shift = b.scale_values_in[4 * (size_t)r + 0];
    scale = b.scale_values_in[4 * (size_t)r + 1];
    if (given_limits) {
      lower = b.scale_values_in[4 * (size_t)r + 2];
      upper = b.scale_values_in[4 * (size_t)r + 3];
    }
    if (scale == 0.0) {
      long long neq = 0;
      for_each_sample(sig, n, [&](int v) { neq += ((double)v != shift) ? 1 : 0; });
      neq = block_sum(neq, sh_red);
      fail = neq > 0 ? NRQ_ST_FPE_DIVIDE : NRQ_ST_FPE_INVALID;
    } else if (need_outlier) {
      auto X = [shift, scale](int v) { return __ddiv_rn(__dsub_rn((double)v, shift), scale); };
      const double xa = X(m1), xb = X(m2);
      // sorted x: ascending in raw if scale > 0, descending otherwise; the middle pair is the same
      const double med = (k1 == k2) ? xa
                         : __dmul_rn(scale > 0 ? __dadd_rn(xa, xb) : __dadd_rn(xb, xa), 0.5);
      auto dev = [&](int v) { return fabs(__dsub_rn(X(v), med)); };
      double D1 = -1.0, D2 = -1.0;
      bool walked = false;
      if (win_ok && have_m) {
        // |x - med| is non-decreasing moving away from the middle on either side: merge the two
        // directions bin by bin until both ranks are covered (serial; a few hundred bins).
        if (tid == 0) {
          int l = m1 - lo_w, h = m1 - lo_w + 1;
          long long cum = 0;
          int okw = 1;
          double d1 = -1.0, d2 = -1.0;
          while (d2 < 0.0) {
            while (l >= 0 && hist[l] == 0) --l;
            while (h < kWin && hist[h] == 0) ++h;
            bool hasl = l >= 0, hash = h < kWin;
            if ((!hasl && below > 0) || (!hash && above > 0) || (!hasl && !hash)) { okw = 0; break; }
            double dl = hasl ? dev(lo_w + l) : 0.0, dh = hash ? dev(lo_w + h) : 0.0;
            bool take_l = hasl && (!hash || dl <= dh);
            double d = take_l ? dl : dh;
            cum += take_l ? hist[l] : hist[h];
            if (take_l) --l; else ++h;
            if (d1 < 0.0 && cum > k1) d1 = d;
            if (cum > k2) d2 = d;
          }
          sh_d[0] = d1; sh_d[1] = d2; sh_flag = okw;
        }
        __syncthreads();
        walked = sh_flag != 0;
        D1 = sh_d[0]; D2 = sh_d[1];
        __syncthreads();
      }
      if (!walked) {
        auto key = [&](int v) { return dbits(dev(v)); };
        D1 = __longlong_as_double((long long)block_bisect(sig, n, key, k1, 0ull, 0x7ff0000000000000ull, sh_red));
        D2 = (k1 == k2) ? D1
             : __longlong_as_double((long long)block_bisect(sig, n, key, k2, 0ull, 0x7ff0000000000000ull, sh_red));
      }
      double mad = (k1 == k2) ? D1 : __dmul_rn(__dadd_rn(D1, D2), 0.5);
      double thr = __dmul_rn(mad, p.outlier_thresh);
      lower = __dsub_rn(med, thr);
      upper = __dadd_rn(med, thr);
    }
  }

  if (tid == 0) {
    sv[0] = shift; sv[1] = scale; sv[2] = lower; sv[3] = upper;
    if (fail != NRQ_ST_OK) status[r] = fail;
    else if (g.n_obs > g.n_norm) status[r] = NRQ_ST_PAST_END;  // starts[-1] > len(norm_signal), resquiggle.py:379-381
  }
}

// ---------------------------------------------------------------- K1b
__global__ void __launch_bounds__(256)
k_normalize_signal(nrq_batch b, const double* __restrict__ sv, const int32_t* __restrict__ status,
                   const int64_t* __restrict__ norm_off, double* __restrict__ out) {
  const int r = blockIdx.x;
  int st = status[r];
  if (st != NRQ_ST_OK && st != NRQ_ST_PAST_END) return;
  ReadGeom g = read_geom(b, r);
  NormCoef c = load_coef(sv, r);
  double* o = out + norm_off[r];
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < g.n_norm; i += gridDim.y * blockDim.x)
    o[i] = norm_value((int)g.sig[i], c);
}

}  // namespace nrq
