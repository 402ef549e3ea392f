// K3 -- per-base event statistics (nanoraw/resquiggle.py:60-70): mean and (population) standard
// deviation of the normalised signal inside each new segment, plus start / length.
//
// Two kernels.  k_event_table (default): one warp streams one read once, 256 samples per step, and
// gets every base's sums from exact integer prefix sums of the DAC codes (see below).
// k_event_table_exact (debug_flags bit 1) replays numpy's float64 summation order and is
// bit-identical to numpy; it is what the integer kernel is validated against.
//
// k_event_table_exact: one CTA per read, one thread per base.  The normalised value of a sample depends only on its
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
  const double* xsig;  // pre-normalised signal or nullptr
  const double* lut;
  int lut_lo;
  NormCoef coef;
  __device__ __forceinline__ double operator()(int i) const {
    if (xsig) return xsig[i];
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
k_event_table_exact(nrq_batch b, nrq_params p, const double* __restrict__ sv, const int32_t* __restrict__ n_bases,
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
  int lut_lo = (g.xsig ? 0 : (int)g.sig[g.n_norm / 2]) - kLut / 2;
  if (coef.clip) {
    double a = coef.shift + coef.lo * coef.scale, c = coef.shift + coef.hi * coef.scale;
    double lo = fmin(a, c), hi = fmax(a, c);
    if (isfinite(lo) && isfinite(hi) && hi - lo < (double)(kLut - 4) && lo > -40000.0 && hi < 40000.0)
      lut_lo = (int)floor(lo) - 2;
  }
  for (int k = tid; k < kLut; k += kEvThreads) lut[k] = norm_value(lut_lo + k, coef);
  __syncthreads();

  XAt x{g.sig, g.xsig, lut, lut_lo, coef};
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

// ---------------------------------------------------------------------------------------------
// k_event_table: integer-domain statistics.
//
// x_i = clip((raw_i - shift) / scale) depends on the DAC code only, and is affine in it between the
// clip limits.  With d_i = raw_i - pivot for unclipped samples,
//     sum_u x   = (S1 + n_u (pivot - shift)) / scale                      S1 = sum d_i   (int64, exact)
//     sum_u (x - mean_u)^2 = (n_u S2 - S1^2) / (n_u scale^2)              S2 = sum d_i^2 (uint64, exact)
// and clipped samples contribute n_lo / n_hi copies of the limits.  Per-base S1, S2, n_lo, n_hi are
// differences of running prefix sums taken at the segment boundaries, so the read is streamed once,
// coalesced, with no per-base loops.  The numerator n_u S2 - S1^2 is formed in 128-bit integers:
// no cancellation.  Results differ from numpy's (which sums the individually rounded x_i) by a few
// ulps -- far inside the 1e-9 contract -- and are checked against the exact kernel in tests.
#ifndef NRQ_EV_WARPS
#define NRQ_EV_WARPS 2
#endif
constexpr int kEvWarps = NRQ_EV_WARPS;   // reads per CTA: one warp streams one read; small CTAs, so that few warps idle while the longest read of a CTA finishes (8 / 4 / 2 warps: 13.8 / 13.6 / 13.5 ms)
constexpr int kSub = 32 * 8;  // samples per step: one 128-bit load per lane

struct EvClass {
  int u_lo, u_hi;  // DAC codes in [u_lo, u_hi] are not clipped
  int pivot;
  double below_val, above_val;  // value taken by codes below u_lo / above u_hi
};

__device__ __forceinline__ void ev_accum(int v, bool valid, int u_lo, int u_hi, int pivot, int& s1,
                                         unsigned long long& s2, unsigned& cnt) {
  bool unc = valid && v >= u_lo && v <= u_hi;
  int d = unc ? v - pivot : 0;
  s1 += d;
  s2 += (unsigned long long)((long long)d * (long long)d);
  cnt += (valid && v < u_lo ? 1u : 0u) + (valid && v > u_hi ? 65536u : 0u);
}

__device__ __forceinline__ int unpack16(const int4& q, int j) {
  int w = j < 2 ? q.x : (j < 4 ? q.y : (j < 6 ? q.z : q.w));
  return (int)(short)((j & 1) ? (w >> 16) : (w & 0xffff));
}

// 1/y for y >= 1: the hardware's 20-bit double reciprocal + two Newton steps (a few ulps)
__device__ __forceinline__ double ev_rcp(double y) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(y));
  r = fma(r, fma(-y, r, 1.0), r);
  r = fma(r, fma(-y, r, 1.0), r);
  return r;
}

__device__ __forceinline__ double ev_sqrt(double y) {
  // 2^-100 <= y < 2^100 (one integer compare on the exponent word; zero, NaN and negatives fall out too):
  // the float seed needs no denormal handling
  if ((unsigned)(__double2hiint(y) - 0x39b00000) >= (unsigned)(0x46300000 - 0x39b00000)) return sqrt(y);
  float seed;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(seed) : "f"((float)y));
  double r = (double)seed;
  const double h = 0.5 * y;
  r = r * fma(-h * r, r, 1.5);
  r = r * fma(-h * r, r, 1.5);
  double s = y * r;
  return fma(fma(-s, s, y), 0.5 * r, s);
}


// sqrt of an integer-valued y in [0, 2^64): the seed is the hardware's 20-bit double rsqrt of y + 2^-60, which is y
// itself for y >= 1 and keeps y = 0 finite, so that 0 comes out as 0 without a branch.  One Newton step takes the
// seed to 2^-39, the closing step on s = y r squares that again: within an ulp of sqrt (checked against an 18-bit
// seed of either sign of error)
__device__ __forceinline__ double ev_sqrt_int(double y) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(y + 0x1p-60));
  const double h = 0.5 * y;
  r = r * fma(-h * r, r, 1.5);
  const double s = y * r;
  return fma(fma(-s, s, y), 0.5 * r, s);
}

struct EvRead {       // per-read constants of the statistics
  double c_off;       // pivot - shift
  double c0;          // (pivot - shift) / scale
  double inv_scale, inv_scale2, inv_scale_abs;
  double below_val, above_val;
  const double* rcp_tab;  // 1/n for n < 64
  int compute_sd;
};

// mean / sd of one base from its integer sums: S1 = sum d, S2 = sum d^2 over the nu unclipped samples,
// nl / nh samples clipped to below_val / above_val
__device__ __forceinline__ void ev_base_stats(const EvRead& R, int n, long long S1, unsigned long long S2,
                                              int nl, int nh, double& mean, double& sd) {
  const int nu = n - nl - nh;
  const double dnu = (double)nu;
  const double rn = n < 64 ? R.rcp_tab[n] : ev_rcp((double)n);
  const double sum_u = ((double)S1 + R.c_off * dnu) * R.inv_scale;
  // every path gives a base without clipped samples the same bits: (S1 / n + pivot - shift) / scale, as the rounds do
  if (nl | nh) mean = (sum_u + ((double)nl * R.below_val + (double)nh * R.above_val)) * rn;
  else mean = fma((double)S1 * rn, R.inv_scale, R.c0);
  if (!R.compute_sd) return;
  double ss = 0.0;
  if (nu > 0) {
    double dnum;
    const unsigned long long aS1 = (unsigned long long)(S1 < 0 ? -S1 : S1);
    if ((S2 >> 31) == 0ull && (aS1 >> 31) == 0ull) {
      dnum = (double)((unsigned long long)nu * S2 - aS1 * aS1);
    } else {
      unsigned __int128 num = (unsigned __int128)(unsigned long long)nu * S2 -
                              (unsigned __int128)aS1 * (unsigned __int128)aS1;
      dnum = (double)(unsigned long long)(num >> 64) * 18446744073709551616.0 +
             (double)(unsigned long long)num;
    }
    const double rnu = (nl | nh) ? (nu < 64 ? R.rcp_tab[nu] : ev_rcp(dnu)) : rn;
    ss = dnum * rnu * R.inv_scale2;
    if (nl | nh) {
      double du = sum_u * rnu - mean;
      ss += dnu * du * du;
    }
  }
  if (nl) { double dl = R.below_val - mean; ss += (double)nl * dl * dl; }
  if (nh) { double dh = R.above_val - mean; ss += (double)nh * dh * dh; }
  sd = ev_sqrt(ss * rn);
}

__device__ __forceinline__ EvClass ev_classify(const NormCoef& coef, int pivot_sample) {
  EvClass k;
  k.u_lo = -32768; k.u_hi = 32767;
  k.below_val = 0.0; k.above_val = 0.0;
  if (coef.clip) {
    // x is monotone in the DAC code, so each clip limit is one integer threshold.  Start from the
    // real-valued estimate, then settle it with the exact per-code evaluation (nanoraw_helper.py:248).
    const bool inc = coef.scale > 0.0;
    auto xval = [&](int v) { return __ddiv_rn(__dsub_rn((double)v, coef.shift), coef.scale); };
    // codes clipped on the low-code side / high-code side (which limit that is depends on the sign of scale)
    auto low_side = [&](int v) { double x = xval(v); return inc ? (x < coef.lo) : (x > coef.hi); };
    auto high_side = [&](int v) { double x = xval(v); return inc ? (x > coef.hi) : (x < coef.lo); };
    double e0 = coef.shift + (inc ? coef.lo : coef.hi) * coef.scale;
    double e1 = coef.shift + (inc ? coef.hi : coef.lo) * coef.scale;
    int lo = (int)fmin(fmax(floor(e0), -32768.0), 32768.0);
    int hi = (int)fmin(fmax(ceil(e1), -32769.0), 32767.0);
    while (lo > -32768 && !low_side(lo - 1)) --lo;   // u_lo = smallest code not clipped on the low side
    while (lo <= 32767 && low_side(lo)) ++lo;
    while (hi < 32767 && !high_side(hi + 1)) ++hi;   // u_hi = largest code not clipped on the high side
    while (hi >= -32768 && high_side(hi)) --hi;
    k.u_lo = lo; k.u_hi = hi;
    k.below_val = inc ? coef.lo : coef.hi;
    k.above_val = inc ? coef.hi : coef.lo;
  }
  long long mid = ((long long)max(k.u_lo, -32768) + (long long)min(k.u_hi, 32767)) / 2;
  k.pivot = coef.clip ? (int)mid : pivot_sample;
  // a short unclipped range is summed as d = code - u_lo >= 0: no select per sample, 32-bit sums per step
  if (coef.clip && k.u_hi >= k.u_lo && k.u_hi - k.u_lo <= 2047) k.pivot = k.u_lo;
  return k;
}

#ifndef NRQ_EV_MINB
#define NRQ_EV_MINB 24
#endif

// one step of the inclusive warp scan of (a, b): the shuffle's own predicate says whether lane - o exists, so
// no lane compare is needed
__device__ __forceinline__ void ev_scan_step(int& a, unsigned& b, int o) {
  asm volatile(
      "{ .reg .pred p; .reg .b32 t0, t1;\n"
      "shfl.sync.up.b32 t0|p, %0, %2, 0, 0xffffffff;\n"
      "shfl.sync.up.b32 t1, %1, %2, 0, 0xffffffff;\n"
      "@p add.s32 %0, %0, t0;\n"
      "@p add.u32 %1, %1, t1; }"
      : "+r"(a), "+r"(b) : "r"(o));
}

// lane-local exclusive prefixes (sum d, sum d^2) of one step: lane t's eight pairs, padded to 80 bytes so that
// the four 128-bit stores of a warp are bank-conflict free
struct __align__(16) EvLoc {
  int2 p[8];
  int2 pad[2];
};

__device__ __forceinline__ int2 ev_loc_at(const EvLoc* locw, int sl) {  // pair of slot sl = 8 t + j
  return *reinterpret_cast<const int2*>(reinterpret_cast<const char*>(locw) + sl * 8 + (sl >> 3) * 16);
}

// Deferred rounds.  A round of boundaries costs the same whether it resolves 32 of them or 3, and a 256-sample step
// holds 29 on average: one step in five needed a second, nearly empty round.  While a read is in its plain stretch
// (narrow range, interior steps, no clipped sample in the step or waiting in the current base) the prefixes of the
// last two steps are therefore kept (a ring of 512 slots) and a round is only run when 32 boundaries wait, or when
// some wait from the step before; what is left over is resolved together with the next step's.  Ring layout: lane t
// of step s owns 48 bytes at ((s & 1) * 32 + t) * 48: eight 32-bit prefixes of d^2, then eight 16-bit prefixes of d
// (at most 7 * 2047); its running prefix pair (low 32 bits of the sums since the read began) sits in a second array.
// All arithmetic of such a round is modulo 2^32, exact while a base has at most 1 024 samples (1 024 * 2047^2 <
// 2^32), which the distance test at the head of the step guarantees.
#ifndef NRQ_EV_DEFER
#define NRQ_EV_DEFER 1
#endif

constexpr int kEvArena = 3584;        // bytes of shared memory per warp
constexpr int kEvRingG = 64 * 48;     // offset of the ring's running prefixes
constexpr int kEvDeferReach = 768;    // a step defers only if the last resolved boundary lies at most this far behind it

// the ring is addressed through one 32-bit shared-window address per warp (made opaque, so that it stays in a
// register: left to itself the compiler re-forms the generic address, five instructions, at every use)
template <int OFF>
__device__ __forceinline__ unsigned ev_lds32(unsigned sa) {
  unsigned v;
  asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(sa), "n"(OFF) : "memory");
  return v;
}
template <int OFF>
__device__ __forceinline__ unsigned ev_lds16(unsigned sa) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1+%2];" : "=h"(v) : "r"(sa), "n"(OFF) : "memory");
  return (unsigned)v;
}
template <int OFF>
__device__ __forceinline__ uint2 ev_lds64(unsigned sa) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2+%3];" : "=r"(v.x), "=r"(v.y) : "r"(sa), "n"(OFF) : "memory");
  return v;
}
template <int OFF>
__device__ __forceinline__ void ev_sts128(unsigned sa, unsigned x, unsigned y, unsigned z, unsigned w) {
  asm volatile("st.shared.v4.u32 [%0+%1], {%2, %3, %4, %5};" ::"r"(sa), "n"(OFF), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
template <int OFF>
__device__ __forceinline__ void ev_sts64(unsigned sa, unsigned x, unsigned y) {
  asm volatile("st.shared.v2.u32 [%0+%1], {%2, %3};" ::"r"(sa), "n"(OFF), "r"(x), "r"(y) : "memory");
}

#ifndef NRQ_EV_AHEAD
#define NRQ_EV_AHEAD 2
#endif
constexpr int kEvAhead = NRQ_EV_AHEAD;  // steps between a prefetch and the load it serves
constexpr int kEvLongRead = 262144;     // samples from which a read is streamed in parts (when the launch has any)
constexpr int kEvPartSamples = 65536;   // ... of about this many samples each
constexpr int kNoPos = 0x3fffffff;
  // "no such boundary": never inside a step

// PARTS: long reads are streamed in gridDim.y parts (nrq_params.long_read_parts).  (compute_sd as a second template
// parameter takes a test and a branch out of every round and costs more than that in spills: 11.77 against 11.68 ms)
template <bool PARTS>
__global__ void __launch_bounds__(kEvWarps * 32, NRQ_EV_MINB)
k_event_table(nrq_batch b, nrq_params p, const double* __restrict__ sv, const int32_t* __restrict__ n_bases,
              const int32_t* __restrict__ new_segs, const int32_t* __restrict__ status,
              double* __restrict__ norm_mean, double* __restrict__ norm_stdev,
              uint32_t* __restrict__ ev_start, uint32_t* __restrict__ ev_length) {
  __shared__ double rcp_tab[64];
  // per warp: lane-local exclusive prefixes of (d, d^2) (EvLoc[32], 2 560 bytes) and of the packed clip counters
  // ([2][32][4] words, 1 024 bytes) of the current step -- or, while rounds are deferred (below), the prefixes of the
  // last TWO steps in a compact form (64 x 48 bytes) and their lanes' running prefixes (64 x 8 bytes)
  __shared__ __align__(16) unsigned char arena[kEvWarps][kEvArena];
  for (int i = threadIdx.x; i < 64; i += kEvWarps * 32) rcp_tab[i] = i ? 1.0 / (double)i : 0.0;
  __syncthreads();
  // broadcast from lane 0: tells the compiler that the warp index and everything loaded through it are
  // warp-uniform, so the shuffles and votes below need no convergence barriers around them
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int r = blockIdx.x * kEvWarps + warp;
  if (r >= b.n_reads || status[r] != NRQ_ST_OK) return;
  const int lane = (int)lane_id();
  ReadGeom g = read_geom(b, r);
  const NormCoef coef = load_coef(sv, r);
  int nb = n_bases[r];
  const int32_t* segs = new_segs + b.align_off[r] + r;
  size_t off = (size_t)b.align_off[r];
  int n_norm = g.n_norm;
  const EvClass k = ev_classify(coef, (int)g.sig[n_norm / 2]);  // every lane: a handful of divisions
  // ---- long reads (BASELINE configs[3]: ~1 M samples): one warp would stream the read for milliseconds while
  // the device idles, so the launch carries gridDim.y parts and a read of more than kEvLongRead samples is cut at
  // base boundaries into that many (at most) runs of bases, one per warp.  A part is a read of its own from then
  // on: its samples start at seg0 (the part's first boundary), positions below are relative to that.  The
  // classification above (pivot, clip codes) is the whole read's, so every part does the same arithmetic.
  int seg0 = 0;  // (stays a compile-time 0 without PARTS)
  if (PARTS) {
    const int parts = n_norm > kEvLongRead ? min((int)gridDim.y, (n_norm + kEvPartSamples - 1) / kEvPartSamples) : 1;
    const int part = (int)blockIdx.y;
    if (part >= parts) return;
    if (parts > 1) {
      auto first_base_from = [&](long long s) {  // first boundary index in [0, nb] whose position is >= s
        int lo = 0, hi = nb;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if ((long long)__ldg(segs + mid) < s) lo = mid + 1; else hi = mid;
        }
        return lo;
      };
      const int b0 = part == 0 ? 0 : first_base_from((long long)n_norm * part / parts);
      const int b1 = part + 1 == parts ? nb : first_base_from((long long)n_norm * (part + 1) / parts);
      if (b0 >= b1) return;
      seg0 = b0 == 0 ? 0 : __ldg(segs + b0);
      const int seg1 = b1 == nb ? n_norm : __ldg(segs + b1);
      g.sig += seg0;
      n_norm = seg1 - seg0;
      nb = b1 - b0;
      segs += b0;
      off += (size_t)b0;
    }
  }
  EvLoc* const locw = reinterpret_cast<EvLoc*>(arena[warp]);
  unsigned (*const locc)[32][4] = reinterpret_cast<unsigned (*)[32][4]>(arena[warp] + sizeof(EvLoc) * 32);
  unsigned ring_sa = (unsigned)__cvta_generic_to_shared(arena[warp]);
  asm volatile("" : "+r"(ring_sa));
  double* const o_mean = norm_mean + off;
  double* const o_sd = norm_stdev + off;
  uint32_t* const o_start = ev_start + off;
  uint32_t* const o_length = ev_length + off;

  const int u_lo = k.u_lo, u_hi = k.u_hi, pivot = k.pivot;
  EvRead R;
  R.c_off = __dsub_rn((double)pivot, coef.shift);
  R.inv_scale = 1.0 / coef.scale;
  R.c0 = R.c_off * R.inv_scale;
  R.inv_scale2 = R.inv_scale * R.inv_scale;
  R.inv_scale_abs = fabs(R.inv_scale);
  R.below_val = k.below_val;
  R.above_val = k.above_val;
  R.rcp_tab = rcp_tab;
  R.compute_sd = p.compute_sd;
  // d^2 summed over one 256-sample step fits 32 bits when |d| <= 2048
  const unsigned range = (unsigned)(u_hi - u_lo);
  const bool narrow = coef.clip && u_hi >= u_lo && range <= 4094u;
  const bool tiny = narrow && pivot == u_lo;  // ev_classify: unclipped range <= 2047 codes, d = code - u_lo >= 0
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);

  // steps are laid out on 16-byte boundaries of the raw buffer: sample i sits at slot i + a
  const int a = (int)(((uintptr_t)g.sig >> 1) & 7);
  const int16_t* base_al = g.sig - a;
  const int16_t* raw_end = b.raw + b.raw_off[b.n_reads];
  // running prefix at the start of the current step, and prefix at the last resolved boundary
  long long run1 = 0, prev1 = 0;
  unsigned long long run2 = 0, prev2 = 0;
  int runl = 0, runh = 0, prevl = 0, prevh = 0;
  int cur = 0;      // next boundary to resolve (0 .. nb)
  // Boundary positions are kept as SLOTS (position + a, the coordinate the steps are laid out in): every round
  // compares them with step limits and indexes the step's prefixes by them.
  const int adj = a - seg0;  // slot of a boundary = new_segs value + adj
  // the lane's own view of the boundaries, pinned in a vector register pair: as a warp-uniform base the compiler
  // keeps it on the uniform datapath and re-forms the address from there (seven instructions a load).  The same
  // for the samples or the four output columns costs registers the kernel does not have at 48 warps per SM
  // (measured: 12.4-12.9 ms against 11.8; at 48 registers / 42 warps 12.6 ms without any pinning).
  const int32_t* segs_l = segs + lane;
  asm volatile("" : "+l"(segs_l));
  const int16_t* const base_l = base_al + 8 * lane;
  int prevpos = a;  // slot of boundary cur - 1
  // slot of boundary cur + lane: np.split(norm_signal, new_segs[1:-1]) -- the first piece starts at 0, the
  // last runs to the end (:64).  Loaded as soon as cur is known, a step ahead of its use.
  auto load_pos = [&](int c) {  // c > 0 (cur only grows, and boundary 0 is the first one resolved)
    const unsigned bidx = (unsigned)(c + lane);
    int q = bidx == (unsigned)nb ? n_norm + a : kNoPos;
    if (bidx < (unsigned)nb) q = __ldg(segs_l + (unsigned)c) + adj;
    return q;
  };
  int pos = load_pos(0);
  if (lane == 0) pos = a;  // boundary 0 sits at sample 0 whatever new_segs[0] says
  // interior steps (all 256 samples belong to the read) of a read with a narrow unclipped range: [fast_lo, fast_hi)
  const int fast_lo = narrow ? ((a + kSub - 1) / kSub) * kSub : 0x7fffffff;
  const int fast_hi = n_norm + a - kSub + 1;

  bool dpend = false;  // boundaries of the step before may still wait in the ring
  // no clipped sample waits in the current base and boundary 1 is behind us: settled at the end of every step that
  // is not deferred (a deferred step changes none of it)
  bool calm = false;
  unsigned pe1 = 0u, pe2 = 0u;  // while they may: low words of the prefixes at the last resolved boundary
  for (int tile0 = 0; tile0 <= n_norm + a; tile0 += kSub) {
    const int slot = tile0 + 8 * lane;
    const bool fast = tile0 >= fast_lo && tile0 < fast_hi;  // interior step of a read with a narrow unclipped range
    int4 q = make_int4(0, 0, 0, 0);
    bool ds = false;  // this step's rounds may be deferred
    if (fast) {
      q = __ldg(reinterpret_cast<const int4*>(base_l + (unsigned)tile0));
      // the samples two steps ahead start their trip from HBM now (no register held for them)
      if (slot + kEvAhead * kSub + 8 <= n_norm)
        asm volatile("prefetch.global.L1 [%0];" ::"l"(base_l + (unsigned)tile0 + kEvAhead * kSub));
      if (NRQ_EV_DEFER && tiny) {  // optimistic: no clipped sample in the step, d = code - u_lo needs no select
        int l1[8];
        unsigned l2[8];
        int s1 = 0;
        unsigned s2 = 0;
        bool oor = false;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const unsigned d = (unsigned)(unpack16(q, j) - u_lo);
          oor |= d > range;
          l1[j] = s1;
          l2[j] = s2;
          s1 += (int)d;
          s2 += d * d;
        }
        ds = !__any_sync(0xffffffffu, oor) && calm && tile0 - prevpos <= kEvDeferReach;
        if (ds) {
          const unsigned rl = (unsigned)((tile0 >> 3) & 32) + (unsigned)lane;  // the lane's place in the ring
          const unsigned mine = ring_sa + rl * 48u;
          ev_sts128<0>(mine, l2[0], l2[1], l2[2], l2[3]);
          ev_sts128<16>(mine, l2[4], l2[5], l2[6], l2[7]);
          ev_sts128<32>(mine, (unsigned)l1[0] | ((unsigned)l1[1] << 16), (unsigned)l1[2] | ((unsigned)l1[3] << 16),
                        (unsigned)l1[4] | ((unsigned)l1[5] << 16), (unsigned)l1[6] | ((unsigned)l1[7] << 16));
          int i1 = s1;
          unsigned i2 = s2;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) ev_scan_step(i1, i2, o);
          ev_sts64<kEvRingG>(ring_sa + rl * 8u, (unsigned)run1 + (unsigned)(i1 - s1), (unsigned)run2 + (i2 - s2));
          if (!dpend) { pe1 = (unsigned)prev1; pe2 = (unsigned)prev2; }
          // (the rounds below work on the ring and the low words alone: the running sums can move on now)
          run1 += __shfl_sync(0xffffffffu, i1, 31);
          run2 += __shfl_sync(0xffffffffu, i2, 31);
          __syncwarp();
        }
      }
    }
    // ================= rounds over the ring: this step's (deferred where possible), or what the step before left =====
    if (ds || dpend) {
      const int limit = ds ? tile0 + kSub : tile0;
      for (;;) {
        const int gs = pos;
        const bool inb = gs < limit;
        const unsigned bal = __ballot_sync(0xffffffffu, inb);
        if (bal == 0u) break;
        // fewer than 32 to resolve: they wait for the next step's, unless some (lane 0's is the oldest) wait already
        if (ds && bal != 0xffffffffu && __shfl_sync(0xffffffffu, gs, 0) >= tile0) break;
        const int cntb = __popc(bal);
        const unsigned sl = (unsigned)gs & 511u;
        const unsigned t8 = sl & 0x1f8u;  // 8 x (place in the ring)
        const unsigned L2 = ev_lds32<0>(ring_sa + sl * 4u + t8 * 2u);
        const unsigned L1 = ev_lds16<32>(ring_sa + sl * 2u + t8 * 4u);
        const uint2 G = ev_lds64<kEvRingG>(ring_sa + t8);
        const unsigned E1 = G.x + L1;
        const unsigned E2 = G.y + L2;
        unsigned pE1 = __shfl_up_sync(0xffffffffu, E1, 1);
        unsigned pE2 = __shfl_up_sync(0xffffffffu, E2, 1);
        int bpos = __shfl_up_sync(0xffffffffu, pos, 1);
        if (lane == 0) { pE1 = pe1; pE2 = pe2; bpos = prevpos; }
        if (inb) {
          const int n = pos - bpos;
          const int S1 = (int)(E1 - pE1);  // 0 <= S1 <= S2 < 2^32: d is a non-negative integer
          const unsigned S2 = E2 - pE2;
          // (an empty base -- never from nrq_resegment, whose reads have increasing boundaries -- divides by zero:
          // 1 / 0 = inf, the Newton step makes it NaN, and NaN is what np.mean / np.std of nothing give)
          const double dn = (double)max(n, 0);
          const double rn = ev_rcp(dn);
          // mean = (S1 / n + pivot - shift) / scale
          const double mean = fma((double)S1 * rn, R.inv_scale, R.c0);
          double sd = qnan;
          if (R.compute_sd) {
            const unsigned long long num =
                (unsigned long long)(unsigned)n * (unsigned long long)S2 - (unsigned long long)((long long)S1 * (long long)S1);
            sd = ev_sqrt_int((double)num) * (rn * R.inv_scale_abs);  // sqrt(num / n^2 / scale^2)
          }
          const unsigned o = (unsigned)(cur + lane - 1);  // 32-bit index on per-read base pointers
          o_mean[o] = mean;
          o_sd[o] = sd;
          o_start[o] = (uint32_t)(bpos - adj);
          o_length[o] = (uint32_t)n;
        }
        pe1 = __shfl_sync(0xffffffffu, E1, cntb - 1);
        pe2 = __shfl_sync(0xffffffffu, E2, cntb - 1);
        prevpos = __shfl_sync(0xffffffffu, pos, cntb - 1);
        cur += cntb;
        pos = load_pos(cur);
      }
      if (ds) {
        dpend = true;
        continue;
      }
      // back to 64-bit prefixes: the last resolved boundary lies before this step's first sample, at most 1 024
      // samples before it, so the distances are the differences of the low words, taken as unsigned
      dpend = false;
      prev1 = run1 - (long long)(unsigned long long)((unsigned)run1 - pe1);
      prev2 = run2 - (unsigned long long)((unsigned)run2 - pe2);
    }
    // ================= fast step: interior step of a read with a narrow unclipped range =================
    if (fast) {
      int l1[8];
      unsigned l2[8];
      int s1 = 0;
      unsigned s2 = 0, cnt = 0;
      bool anyc = true;
      if (!NRQ_EV_DEFER && tiny) {  // optimistic: no clipped sample in the step, d = code - u_lo needs no select
        bool oor = false;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const unsigned d = (unsigned)(unpack16(q, j) - u_lo);
          oor |= d > range;
          l1[j] = s1;
          l2[j] = s2;
          s1 += (int)d;
          s2 += d * d;
        }
        anyc = __any_sync(0xffffffffu, oor);
      }
      if (anyc) {
        unsigned out_of_range = 0;
        s1 = 0;
        s2 = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int v = unpack16(q, j);
          const bool unc = (unsigned)(v - u_lo) <= range;
          out_of_range |= unc ? 0u : 1u;
          const int d = unc ? v - pivot : 0;
          l1[j] = s1;
          l2[j] = s2;
          s1 += d;
          s2 += (unsigned)(d * d);
        }
        // clipped samples are rare: only steps that contain one carry the (below, above) counters
        anyc = __any_sync(0xffffffffu, out_of_range);
        if (anyc) {
          unsigned lc[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int v = unpack16(q, j);
            lc[j] = cnt;
            cnt += (v < u_lo ? 1u : 0u) + (v > u_hi ? 65536u : 0u);
          }
          *reinterpret_cast<uint4*>(locc[0][lane]) = make_uint4(lc[0], lc[1], lc[2], lc[3]);
          *reinterpret_cast<uint4*>(locc[1][lane]) = make_uint4(lc[4], lc[5], lc[6], lc[7]);
        }
      }
      {
        int4* dst = reinterpret_cast<int4*>(locw[lane].p);
        dst[0] = make_int4(l1[0], (int)l2[0], l1[1], (int)l2[1]);
        dst[1] = make_int4(l1[2], (int)l2[2], l1[3], (int)l2[3]);
        dst[2] = make_int4(l1[4], (int)l2[4], l1[5], (int)l2[5]);
        dst[3] = make_int4(l1[6], (int)l2[6], l1[7], (int)l2[7]);
      }
      int i1 = s1;
      unsigned i2 = s2, ic = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        ev_scan_step(i1, i2, o);
        if (anyc) {
          unsigned tc = __shfl_up_sync(0xffffffffu, ic, o);
          if (lane >= o) ic += tc;
        }
      }
      const int ex1 = i1 - s1;
      const unsigned ex2 = i2 - s2, exc = ic - cnt;
      const int tot1 = __shfl_sync(0xffffffffu, i1, 31);
      const unsigned tot2 = __shfl_sync(0xffffffffu, i2, 31);
      const unsigned totc = anyc ? __shfl_sync(0xffffffffu, ic, 31) : 0u;
      __syncwarp();
      const int tile_end = tile0 + kSub;
      for (;;) {
        const bool inb = pos < tile_end;
        const unsigned bal = __ballot_sync(0xffffffffu, inb);
        const int cntb = __popc(bal);
        if (cntb == 0) break;
        const int sl = (pos - tile0) & (kSub - 1);
        const int t = sl >> 3;
        const int2 L = ev_loc_at(locw, sl);
        const int E1 = __shfl_sync(0xffffffffu, ex1, t) + L.x;
        const unsigned E2 = __shfl_sync(0xffffffffu, ex2, t) + (unsigned)L.y;
        int pE1 = __shfl_up_sync(0xffffffffu, E1, 1);
        unsigned pE2 = __shfl_up_sync(0xffffffffu, E2, 1);
        int bpos = __shfl_up_sync(0xffffffffu, pos, 1);
        // sums carried in from earlier steps for the base that ends at the first boundary of this round
        const long long c1 = run1 - prev1;
        const unsigned long long c2 = run2 - prev2;
        // the 32-bit rounds below work modulo 2^32, so the carried sums may also be "negative": in the second round
        // of a step the last resolved boundary lies inside the step, ahead of the running prefix (|c2| < 2^30 then)
        const bool c_small = ((c2 + (1ull << 30)) >> 31) == 0ull;
        // ---- round of a step with clipped samples (one step in eight): bases 1 .. nb - 2 only, carried sums fit 32
        // bits; the clipped samples of each base are differences of the packed (below, above) counter prefixes, and
        // only a base that holds one takes the general statistics
        if (anyc && cur >= 2 && cur + 32 < nb && c_small) {
          const unsigned EC = __shfl_sync(0xffffffffu, exc, t) + locc[(sl & 7) >> 2][t][sl & 3];
          const unsigned pEC = __shfl_up_sync(0xffffffffu, EC, 1);
          int nl = (int)(EC & 0xffffu), nh = (int)(EC >> 16);
          if (lane == 0) {  // the base may have begun in an earlier step
            pE1 = -(int)c1; pE2 = 0u - (unsigned)c2; bpos = prevpos;
            nl += runl - prevl;
            nh += runh - prevh;
          } else {
            nl -= (int)(pEC & 0xffffu);
            nh -= (int)(pEC >> 16);
          }
          if (inb) {
            const int n = pos - bpos;
            const int S1 = E1 - pE1;
            const unsigned S2 = E2 - pE2;
            double mean = qnan, sd = qnan;
            if (n > 0) {
              if ((nl | nh) == 0) {
                const double dn = (double)n;
                const double rn = ev_rcp(dn);
                mean = fma((double)S1 * rn, R.inv_scale, R.c0);
                if (R.compute_sd) {
                  const unsigned long long num = (unsigned long long)(unsigned)n * (unsigned long long)S2 -
                                                 (unsigned long long)((long long)S1 * (long long)S1);
                  sd = ev_sqrt_int((double)num) * (rn * R.inv_scale_abs);  // sqrt(num / n^2 / scale^2)
                }
              } else {
                ev_base_stats(R, n, (long long)S1, (unsigned long long)S2, nl, nh, mean, sd);
              }
            }
            const unsigned o = (unsigned)(cur + lane - 1);
            o_mean[o] = mean;
            o_sd[o] = sd;
            o_start[o] = (uint32_t)(bpos - adj);
            o_length[o] = (uint32_t)n;
          }
          const unsigned last = __shfl_sync(0xffffffffu, EC, cntb - 1);
          prev1 = run1 + __shfl_sync(0xffffffffu, E1, cntb - 1);
          prev2 = run2 + __shfl_sync(0xffffffffu, E2, cntb - 1);
          prevl = runl + (int)(last & 0xffffu);
          prevh = runh + (int)(last >> 16);
          prevpos = __shfl_sync(0xffffffffu, pos, cntb - 1);
          cur += cntb;
          pos = load_pos(cur);
          if (cntb < 32) break;
          continue;
        }
        // ---- common round: no clipped sample in play, bases 1 .. nb - 2 only, carried sums fit 32 bits
        if (!anyc && cur >= 2 && cur + 32 < nb && c_small && runl == prevl && runh == prevh) {
          if (lane == 0) { pE1 = -(int)c1; pE2 = 0u - (unsigned)c2; bpos = prevpos; }
          if (inb) {
            const int n = pos - bpos;
            const int S1 = E1 - pE1;        // |S1| <= S2 < 2^32: d is an integer
            const unsigned S2 = E2 - pE2;
            const double dn = (double)n;
            const double rn = ev_rcp(dn);
            double mean = fma((double)S1 * rn, R.inv_scale, R.c0), sd = qnan;
            if (R.compute_sd) {
              const unsigned long long num =
                  (unsigned long long)(unsigned)n * (unsigned long long)S2 - (unsigned long long)((long long)S1 * (long long)S1);
              sd = ev_sqrt_int((double)num) * (rn * R.inv_scale_abs);  // sqrt(num / n^2 / scale^2)
            }
            if (n <= 0) { mean = qnan; sd = qnan; }
            const unsigned o = (unsigned)(cur + lane - 1);  // 32-bit index on per-read base pointers
            o_mean[o] = mean;
            o_sd[o] = sd;
            o_start[o] = (uint32_t)(bpos - adj);
            o_length[o] = (uint32_t)n;
          }
          prev1 = run1 + __shfl_sync(0xffffffffu, E1, cntb - 1);
          prev2 = run2 + __shfl_sync(0xffffffffu, E2, cntb - 1);
          prevpos = __shfl_sync(0xffffffffu, pos, cntb - 1);
          cur += cntb;
          pos = load_pos(cur);
          if (cntb < 32) break;
          continue;
        }
        // ---- any other round of a fast step
        const int bidx = cur + lane;
        int el = runl, eh = runh, pl = runl, ph = runh;
        if (anyc) {
          const unsigned EC = __shfl_sync(0xffffffffu, exc, t) + locc[(sl & 7) >> 2][t][sl & 3];
          el += (int)(EC & 0xffffu);
          eh += (int)(EC >> 16);
          pl = __shfl_up_sync(0xffffffffu, el, 1);
          ph = __shfl_up_sync(0xffffffffu, eh, 1);
        }
        if (inb && bidx >= 1) {
          const int kbase = bidx - 1;
          const int s_hi = segs[bidx] - seg0;
          const int s_lo = (lane == 0) ? segs[kbase] - seg0 : (kbase == 0 ? segs[0] - seg0 : bpos - a);
          const int pos0 = (kbase == 0) ? 0 : s_lo;
          const int n = pos - a - pos0;
          double mean = qnan, sd = qnan;
          if (n > 0) {
            long long S1;
            unsigned long long S2;
            if (lane == 0) {  // the base may have begun in an earlier step
              S1 = c1 + E1;
              S2 = c2 + E2;
              pl = prevl;
              ph = prevh;
            } else {
              S1 = (long long)(E1 - pE1);
              S2 = (unsigned long long)(E2 - pE2);
            }
            ev_base_stats(R, n, S1, S2, el - pl, eh - ph, mean, sd);
          }
          norm_mean[off + kbase] = mean;
          norm_stdev[off + kbase] = sd;
          ev_start[off + kbase] = (uint32_t)(s_lo + seg0);
          ev_length[off + kbase] = (uint32_t)(s_hi - s_lo);
        }
        prev1 = run1 + __shfl_sync(0xffffffffu, E1, cntb - 1);
        prev2 = run2 + __shfl_sync(0xffffffffu, E2, cntb - 1);
        prevl = anyc ? __shfl_sync(0xffffffffu, el, cntb - 1) : runl;
        prevh = anyc ? __shfl_sync(0xffffffffu, eh, cntb - 1) : runh;
        prevpos = __shfl_sync(0xffffffffu, pos, cntb - 1);
        cur += cntb;
        pos = load_pos(cur);
        if (cntb < 32) break;
      }
      run1 += tot1;
      run2 += tot2;
      runl += (int)(totc & 0xffffu);
      runh += (int)(totc >> 16);
      calm = cur >= 2 && runl == prevl && runh == prevh;
      __syncwarp();
      continue;
    }
    // ================= general step =================
    // ---- load 8 samples, lane totals
    if (slot < n_norm + a) {
      const int16_t* ptr = base_al + slot;
      if (ptr >= b.raw && ptr + 8 <= raw_end) q = __ldg(reinterpret_cast<const int4*>(ptr));
      else {
        int w[4] = {0, 0, 0, 0};
        for (int j = 0; j < 8; ++j)
          if (ptr + j >= b.raw && ptr + j < raw_end) w[j >> 1] |= ((int)(unsigned short)ptr[j]) << (16 * (j & 1));
        q = make_int4(w[0], w[1], w[2], w[3]);
      }
    }
    int s1 = 0;
    unsigned long long s2 = 0;
    unsigned cnt = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int i = slot + j - a;
      ev_accum(unpack16(q, j), i >= 0 && i < n_norm, u_lo, u_hi, pivot, s1, s2, cnt);
    }
    // ---- warp scan of (s1, s2, cnt): exclusive prefixes ex*, step totals tot*
    int i1 = s1;
    unsigned long long i2 = s2;
    unsigned ic = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t1 = __shfl_up_sync(0xffffffffu, i1, o);
      unsigned long long t2 = __shfl_up_sync(0xffffffffu, i2, o);
      unsigned tc = __shfl_up_sync(0xffffffffu, ic, o);
      if (lane >= o) { i1 += t1; i2 += t2; ic += tc; }
    }
    const int ex1 = i1 - s1;
    const unsigned long long ex2 = i2 - s2;
    const unsigned exc = ic - cnt;
    const int tot1 = __shfl_sync(0xffffffffu, i1, 31);
    const unsigned long long tot2 = __shfl_sync(0xffffffffu, i2, 31);
    const unsigned totc = __shfl_sync(0xffffffffu, ic, 31);

    // ---- boundaries that fall inside this step, 32 at a time
    const int tile_end = tile0 + kSub;
    for (;;) {
      const int bidx = cur + lane;
      const bool inb = pos < tile_end;
      const unsigned bal = __ballot_sync(0xffffffffu, inb);
      const int cntb = __popc(bal);
      if (cntb == 0) break;
      int sl = inb ? pos - tile0 : 0;  // slot inside the step
      if (sl < 0) sl = 0;
      const int t = sl >> 3, jn = sl & 7;
      int4 qq;
      qq.x = __shfl_sync(0xffffffffu, q.x, t);
      qq.y = __shfl_sync(0xffffffffu, q.y, t);
      qq.z = __shfl_sync(0xffffffffu, q.z, t);
      qq.w = __shfl_sync(0xffffffffu, q.w, t);
      const int tex1 = __shfl_sync(0xffffffffu, ex1, t);
      const unsigned long long tex2 = __shfl_sync(0xffffffffu, ex2, t);
      const unsigned texc = __shfl_sync(0xffffffffu, exc, t);
      int p1 = 0;
      unsigned long long p2 = 0;
      unsigned pc = 0;
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        int i = tile0 + 8 * t + j - a;
        ev_accum(unpack16(qq, j), j < jn && i >= 0 && i < n_norm, u_lo, u_hi, pivot, p1, p2, pc);
      }
      const unsigned cc = texc + pc;
      const long long e1 = run1 + tex1 + p1;
      const unsigned long long e2 = run2 + tex2 + p2;
      const int el = runl + (int)(cc & 0xffffu), eh = runh + (int)(cc >> 16);
      // prefix at the previous boundary: lane - 1, or the carry for lane 0
      long long b1 = __shfl_up_sync(0xffffffffu, e1, 1);
      unsigned long long b2 = __shfl_up_sync(0xffffffffu, e2, 1);
      int bl = __shfl_up_sync(0xffffffffu, el, 1), bh = __shfl_up_sync(0xffffffffu, eh, 1);
      int bpos = __shfl_up_sync(0xffffffffu, pos, 1);
      if (lane == 0) { b1 = prev1; b2 = prev2; bl = prevl; bh = prevh; }
      if (inb && bidx >= 1) {
        const int kbase = bidx - 1;
        const int s_hi = segs[bidx] - seg0;
        const int s_lo = (lane == 0) ? segs[kbase] - seg0 : (kbase == 0 ? segs[0] - seg0 : bpos - a);
        const int pos0 = (kbase == 0) ? 0 : s_lo;
        const int n = pos - a - pos0;
        double mean = qnan, sd = qnan;
        if (n > 0) ev_base_stats(R, n, e1 - b1, e2 - b2, el - bl, eh - bh, mean, sd);
        norm_mean[off + kbase] = mean;
        norm_stdev[off + kbase] = sd;
        ev_start[off + kbase] = (uint32_t)(s_lo + seg0);
        ev_length[off + kbase] = (uint32_t)(s_hi - s_lo);
      }
      // carry: prefix at the last boundary resolved in this round
      prev1 = __shfl_sync(0xffffffffu, e1, cntb - 1);
      prev2 = __shfl_sync(0xffffffffu, e2, cntb - 1);
      prevl = __shfl_sync(0xffffffffu, el, cntb - 1);
      prevh = __shfl_sync(0xffffffffu, eh, cntb - 1);
      prevpos = __shfl_sync(0xffffffffu, pos, cntb - 1);
      cur += cntb;
      pos = load_pos(cur);
      if (cntb < 32) break;
    }
    run1 += tot1;
    run2 += tot2;
    runl += (int)(totc & 0xffffu);
    runh += (int)(totc >> 16);
    calm = cur >= 2 && runl == prevl && runh == prevh;
  }
}

}  // namespace nrq
