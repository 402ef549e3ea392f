// K2 -- get_indel_groups + splice (nanoraw/resquiggle.py:203-316, 362-387).
//
// A read is a sequential, data-dependent state machine over its indels (a failed change-point
// search grows the region, which may swallow finished groups), so parallelism is across reads:
// one warp owns one read at a time and pulls the next from a global work counter.  Inside a
// candidate region the warp cooperates: lanes normalise the region's samples into shared memory,
// lane 0 runs the strictly sequential float64 cumulative sum (the reference restarts np.cumsum at
// every candidate region, so every rounding step has to be replayed in order), all lanes form the
// running differences, and the change-points are chosen by repeated masked arg-max under the
// total order (value desc, index desc) == greedy over argsort(kind='stable')[::-1].
//
// Finished change-points are written straight into their final new_segs slots: an untouched
// boundary starts[j] lands at j + D, D = sum of indel diffs of the groups before it, so a group
// (start, stop, K) owns new_segs[start + D + 1 .. start + D + K].
#include "nrq_common.cuh"

namespace nrq {

constexpr int kSegWarps = 8;
#ifndef NRQ_SEG_CAP
#define NRQ_SEG_CAP 254
#endif
#ifndef NRQ_SEG_LUT
#define NRQ_SEG_LUT 256
#endif
constexpr int kRegionCap = NRQ_SEG_CAP;  // region samples held in shared memory per warp (p99 of real regions ~200)
constexpr int kSegLut = NRQ_SEG_LUT;     // DAC codes covered by the per-warp table of normalised values

struct SegCtx {
  const int32_t* t;   // starts_rel_to_read
  int last;           // Br (largest index into t)
  const int16_t* sig;
  const double* xsig;  // pre-normalised signal (API-compat path) or nullptr
  NormCoef coef;
  const int4* ind;    // (start, end, diff, diff_prefix)
  int n_ind;
  int4* groups;       // finished groups: (start, stop, first indel, max indel end)
  int ng;
  int top_stop;       // groups[ng - 1].y, or -1: what a new group is checked against (:220, :267)
  int32_t* out;       // new_segs
  int m;              // min_seg_len
  int limit;          // cpts limit or < 0
  double* buf_sm;
  double* buf_big;
  int big_cap;
  const double* lut;  // normalised value of DAC code lut_lo + k
  unsigned* picks;    // this warp's picks of one selection, in pick order (shared memory)
  int lut_lo;
  const int2* spec;     // K2a's outcome per group head (nrq_speculate.cu) or nullptr
  const unsigned char* cpt_bytes;  // K2a's change-points of this read, as byte offsets from the region start
};

// extend_group, resquiggle.py:203-216
__device__ __forceinline__ int extend_group(const SegCtx& c, int lo, int hi, int maxend,
                                            int& start, int& stop, int& K, int& t_start, int& t_stop) {
  int4 a = c.ind[lo], z = c.ind[hi];
  start = a.x;
  stop = maxend;
  if (start < 0 || stop > c.last) return NRQ_ST_BAD_ALIGNMENT;
  K = (z.w + z.z - a.w) + stop - start - 1;
  for (;;) {
    t_start = c.t[start];
    t_stop = c.t[stop];
    if (t_stop - t_start >= (K + 2) * c.m) return NRQ_ST_OK;
    if (start == 0 && stop == c.last) return NRQ_ST_UNSATISFIABLE;
    K += (start > 0) + (stop < c.last);
    start = max(0, start - 1);
    stop = min(c.last, stop + 1);
  }
}

// the merge loop shared by extend_and_join (:220-225) and extend_for_cpts (:267-272)
__device__ __forceinline__ int swallow_previous(SegCtx& c, int& lo, int hi, int& maxend,
                                                int& start, int& stop, int& K, int& t_start, int& t_stop) {
  while (start <= c.top_stop) {  // top_stop: stop of the last finished group, -1 when there is none
    int4 prev = c.groups[c.ng - 1];
    --c.ng;
    c.top_stop = c.ng > 0 ? c.groups[c.ng - 1].y : -1;
    lo = prev.z;
    maxend = max(maxend, prev.w);
    int st = extend_group(c, lo, hi, maxend, start, stop, K, t_start, t_stop);
    if (st != NRQ_ST_OK) return st;
  }
  return NRQ_ST_OK;
}

// normalised value of a DAC code: shared-memory table over the codes that survive clipping, exact
// division (the same arithmetic that filled the table) for anything outside it
__device__ __forceinline__ double seg_x(const SegCtx& c, int v) {
  unsigned k = (unsigned)(v - c.lut_lo);
  return k < (unsigned)kSegLut ? c.lut[k] : norm_value(v, c.coef);
}

// running_diffs[i] = |2 cs[i+m] - cs[i] - cs[i+2m]| (:241-243) as an ordered key:
// bits + 2 (0 = blacklisted, 1 = picked).  cs[i] lives at cs[i].
__device__ __forceinline__ unsigned long long rd_key(const double* cs, int i, int m) {
  double a = cs[i], bq = cs[i + m], cq = cs[i + 2 * m];
  double rd = fabs(__dsub_rn(__dsub_rn(__dmul_rn(2.0, bq), a), cq));
  return (unsigned long long)__double_as_longlong(rd) + 2ull;
}

// lexicographic arg-max of (key, index) over the warp; returns index + 1, or 0 if no lane has a candidate
__device__ __forceinline__ unsigned warp_argmax(unsigned long long best, int bi) {
  unsigned hi = (unsigned)(best >> 32), lo = (unsigned)best;
  unsigned mh = __reduce_max_sync(kFull, hi);
  bool c1 = hi == mh;
  unsigned ml = __reduce_max_sync(kFull, c1 ? lo : 0u);
  bool c2 = c1 && lo == ml;
  return __reduce_max_sync(kFull, c2 ? (unsigned)(bi + 1) : 0u);
}

// greedy selection (:244-254) with the keys held in registers: lane owns i = lane + 32 e, e < E.  A key is the
// bit pattern of the running difference split in two words, the high one + 1 so that 0 there means "not a
// candidate" (out of range or blacklisted) whatever the low word holds.  Each round's pick goes to `picks`
// (shared memory, one store by lane 0); the sorted output is formed once at the end from a bitmap of the picks.
// 1: K change-points written in ascending order; 0: None.
template <int E>
__device__ __forceinline__ int select_cpts_regs(const SegCtx& c, const double* cs, int nrd, int K, int t0,
                                                int outpos, unsigned* picks) {
  const int lane = (int)lane_id();
  const int m = c.m;
  unsigned kh[E], kl[E];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int i = lane + 32 * e;
    kh[e] = 0u;
    kl[e] = 0u;
    if (i < nrd) {
      const unsigned long long k = rd_key(cs, i, m) - 2ull;  // the bit pattern itself
      kh[e] = (unsigned)(k >> 32) + 1u;
      kl[e] = (unsigned)k;
    }
  }
  for (int round = 0; round < K; ++round) {
    // this lane's best candidate; strict compares from the top, so the larger index wins ties
    unsigned bh = kh[E - 1], bl = kl[E - 1];
    int bi1 = lane + 32 * (E - 1) + 1;
#pragma unroll
    for (int e = E - 2; e >= 0; --e) {
      const bool g = kh[e] > bh || (kh[e] == bh && kl[e] > bl);
      bh = g ? kh[e] : bh;
      bl = g ? kl[e] : bl;
      bi1 = g ? lane + 32 * e + 1 : bi1;
    }
    const unsigned mh = __reduce_max_sync(kFull, bh);
    if (mh == 0u) return 0;  // fewer than K placeable: None (:253-254)
    const bool c1 = bh == mh;
    const unsigned ml = __reduce_max_sync(kFull, c1 ? bl : 0u);
    const bool c2 = c1 && bl == ml;
    const int p = (int)__reduce_max_sync(kFull, c2 ? (unsigned)bi1 : 0u) - 1;
    if (lane == 0) picks[round] = (unsigned)p;
#pragma unroll
    for (int e = 0; e < E; ++e)  // blacklist p-m+1 .. p+m (:249-250)
      if ((unsigned)(lane + 32 * e - (p - m + 1)) < (unsigned)(2 * m)) kh[e] = 0u;
  }
  __syncwarp();
  // sorted(cpt + min_seg_len) + align_segs[group_start]  (:255, :275): bit i of the bitmap = position i picked
  unsigned word[E];
#pragma unroll
  for (int e = 0; e < E; ++e) word[e] = 0u;
#pragma unroll 1
  for (int base = 0; base < K; base += 32) {  // K <= 64: at most two passes
    const unsigned pr = base + lane < K ? picks[base + lane] : 0xffffffffu;
#pragma unroll
    for (int e = 0; e < E; ++e)
      word[e] |= __reduce_or_sync(kFull, (pr >> 5) == (unsigned)e ? 1u << (pr & 31u) : 0u);
  }
  int cnt = 0;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    if ((word[e] >> lane) & 1u)
      c.out[outpos + cnt + __popc(word[e] & ((1u << lane) - 1u))] = t0 + lane + 32 * e + m;
    cnt += __popc(word[e]);
  }
  return 1;
}

// same, keys in memory (regions too large for registers / shared memory)
__device__ int select_cpts_mem(const SegCtx& c, double* cs, int nrd, int K, int t0, int outpos) {
  const unsigned lane = lane_id();
  const int m = c.m;
  unsigned long long* key = reinterpret_cast<unsigned long long*>(cs);
  for (int base = 0; base < nrd; base += 32) {
    int i = base + (int)lane;
    unsigned long long k = i < nrd ? rd_key(cs, i, m) : 0ull;
    __syncwarp();
    if (i < nrd) key[i] = k;
  }
  __syncwarp();
  for (int round = 0; round < K; ++round) {
    unsigned long long best = 0;
    int bi = -1;
    for (int i = lane; i < nrd; i += 32) {
      unsigned long long k = key[i];
      if (k >= 2ull && k >= best) { best = k; bi = i; }
    }
    unsigned mi = warp_argmax(best, bi);
    if (mi == 0u) return 0;
    int p = (int)mi - 1;
    if ((int)lane < 2 * m) {
      int q = p - m + 1 + (int)lane;
      if (q >= 0 && q < nrd) {
        if (q == p) key[q] = 1ull;
        else if (key[q] != 1ull) key[q] = 0ull;
      }
    }
    __syncwarp();
  }
  int cnt = 0;
  for (int base = 0; base < nrd; base += 32) {
    int i = base + (int)lane;
    bool f = i < nrd && key[i] == 1ull;
    unsigned bal = __ballot_sync(kFull, f);
    if (f) c.out[outpos + cnt + __popc(bal & ((1u << lane) - 1u))] = t0 + i + m;
    cnt += __popc(bal);
  }
  return 1;
}

// get_cpts, resquiggle.py:227-255.  1: change-points written to out[outpos..outpos+K); 0: None;
// < 0: -status.
__device__ __forceinline__ int get_cpts(const SegCtx& c, int t0, int t_stop, int K, int outpos) {
  if (c.limit >= 0 && K > c.limit) return -NRQ_ST_CPTS_LIMIT;
  const unsigned lane = lane_id();
  const int m = c.m;
  const int n = t_stop - t0;
  const int nrd = n - 2 * m + 1;  // len(running_diffs)
  if (nrd <= 0) return 0;
  const int16_t* s = c.sig + t0;
  if (n <= kRegionCap && K <= 64) {  // (64: capacity of the per-warp pick list)
    // shared memory: buf[1] = cs[0] = 0, buf[2 + j] = x[j] then cs[j + 1]; pairs are 16-byte aligned
    double* buf = c.buf_sm;
    if (c.xsig) for (int j = lane; j < n; j += 32) buf[2 + j] = c.xsig[t0 + j];
    else for (int j = lane; j < n; j += 32) buf[2 + j] = seg_x(c, (int)s[j]);
    if (lane == 0) buf[1] = 0.0;
    __syncwarp();
    if (lane == 0) {  // np.cumsum (:239): strictly left to right
      double acc = 0.0;
      int j = 0;
      for (; j + 7 < n; j += 8) {
        double2* q = reinterpret_cast<double2*>(buf + 2 + j);
        double2 v0 = q[0], v1 = q[1], v2 = q[2], v3 = q[3];
        acc = __dadd_rn(acc, v0.x); v0.x = acc;
        acc = __dadd_rn(acc, v0.y); v0.y = acc;
        acc = __dadd_rn(acc, v1.x); v1.x = acc;
        acc = __dadd_rn(acc, v1.y); v1.y = acc;
        acc = __dadd_rn(acc, v2.x); v2.x = acc;
        acc = __dadd_rn(acc, v2.y); v2.y = acc;
        acc = __dadd_rn(acc, v3.x); v3.x = acc;
        acc = __dadd_rn(acc, v3.y); v3.y = acc;
        q[0] = v0; q[1] = v1; q[2] = v2; q[3] = v3;
      }
      for (; j + 1 < n; j += 2) {
        double2 v = *reinterpret_cast<double2*>(buf + 2 + j);
        acc = __dadd_rn(acc, v.x); v.x = acc;
        acc = __dadd_rn(acc, v.y); v.y = acc;
        *reinterpret_cast<double2*>(buf + 2 + j) = v;
      }
      if (j < n) buf[2 + j] = __dadd_rn(acc, buf[2 + j]);
    }
    __syncwarp();
    const double* cs = buf + 1;
    int res;
    // keys per lane by region size.  Every instantiation is a copy of the selection loop in a kernel whose
    // instruction footprint matters (see drive_until_push): NRQ_SEG_ESET picks which sizes get their own
#ifndef NRQ_SEG_ESET
#define NRQ_SEG_ESET 1248
#endif
#if NRQ_SEG_ESET == 1248
    if (nrd <= 32) res = select_cpts_regs<1>(c, cs, nrd, K, t0, outpos, c.picks);
    else if (nrd <= 64) res = select_cpts_regs<2>(c, cs, nrd, K, t0, outpos, c.picks);
    else if (kRegionCap <= 135 || nrd <= 128) res = select_cpts_regs<4>(c, cs, nrd, K, t0, outpos, c.picks);
    else res = select_cpts_regs<8>(c, cs, nrd, K, t0, outpos, c.picks);
#elif NRQ_SEG_ESET == 248
    if (nrd <= 64) res = select_cpts_regs<2>(c, cs, nrd, K, t0, outpos, c.picks);
    else if (kRegionCap <= 135 || nrd <= 128) res = select_cpts_regs<4>(c, cs, nrd, K, t0, outpos, c.picks);
    else res = select_cpts_regs<8>(c, cs, nrd, K, t0, outpos, c.picks);
#elif NRQ_SEG_ESET == 28
    if (nrd <= 64) res = select_cpts_regs<2>(c, cs, nrd, K, t0, outpos, c.picks);
    else res = select_cpts_regs<8>(c, cs, nrd, K, t0, outpos, c.picks);
#else
    if (nrd <= 128) res = select_cpts_regs<4>(c, cs, nrd, K, t0, outpos, c.picks);
    else res = select_cpts_regs<8>(c, cs, nrd, K, t0, outpos, c.picks);
#endif
    __syncwarp();
    return res;
  }
  if (n > c.big_cap) return -NRQ_ST_REGION_TOO_LARGE;
  double* buf = c.buf_big;  // buf[0] = cs[0], buf[j + 1] = x[j] then cs[j + 1]
  if (c.xsig) for (int j = lane; j < n; j += 32) buf[j + 1] = c.xsig[t0 + j];
  else for (int j = lane; j < n; j += 32) buf[j + 1] = seg_x(c, (int)s[j]);
  if (lane == 0) buf[0] = 0.0;
  __syncwarp();
  if (lane == 0) {
    double acc = 0.0;
    for (int j = 1; j <= n; ++j) { acc = __dadd_rn(acc, buf[j]); buf[j] = acc; }
  }
  __syncwarp();
  return select_cpts_mem(c, buf, nrd, K, t0, outpos);
}

// extend_for_cpts (:256-276) for a group whose extend_group result is already known, preceded by the merge
// loop of extend_and_join (:220-225)
// known_none: K2a searched exactly this (start, stop, K) and that many of its retries, and found None every time;
// honoured only up to the first merge with a finished group (what follows a merge is a different search)
__device__ __forceinline__ int search_group(SegCtx& c, int& lo, int hi, int& maxend, int& start, int& stop,
                                            int K, int t_start, int t_stop, int known_none = 0) {
  if (start <= c.top_stop) known_none = 0;
  int st = swallow_previous(c, lo, hi, maxend, start, stop, K, t_start, t_stop);
  if (st != NRQ_ST_OK) return st;
  for (;;) {
    int res = 0;
    if (known_none > 0) {
      if (c.limit >= 0 && K > c.limit) return NRQ_ST_CPTS_LIMIT;
      --known_none;
    } else {
      res = get_cpts(c, t_start, t_stop, K, start + c.ind[lo].w + 1);
    }
    if (res > 0) return NRQ_ST_OK;
    if (res < 0) return -res;
    if (start == 0 && stop == c.last) return NRQ_ST_UNSATISFIABLE;
    K += (start > 0) + (stop < c.last);
    start = max(0, start - 1);
    stop = min(c.last, stop + 1);
    t_start = c.t[start];
    t_stop = c.t[stop];
    if (start <= c.top_stop) known_none = 0;
    st = swallow_previous(c, lo, hi, maxend, start, stop, K, t_start, t_stop);
    if (st != NRQ_ST_OK) return st;
  }
}

__device__ __forceinline__ long long global_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// splice (:362-372) + validation (:373-387), cooperative over one warp: finished change-points already sit in
// their new_segs slots; copy the untouched boundaries between groups and check the result
__device__ int splice_and_validate(const SegCtx& c, int n_bases, int n_norm, int& region_samples) {
  const unsigned lane = lane_id();
  int dtotal = 0;
  if (c.n_ind > 0) {
    int4 z = c.ind[c.n_ind - 1];
    dtotal = z.w + z.z;
  }
  // untouched boundaries before each group: one group per lane (the stretches are a few boundaries long);
  // a stretch too long for one lane is copied by the whole warp
  constexpr int kLaneCopy = 24;
  int reg = 0;
  for (int g0 = 0; g0 < c.ng; g0 += 32) {
    const int g = g0 + (int)lane;
    int from = 0, to = -1, d = 0;
    if (g < c.ng) {
      const int4 grp = c.groups[g];
      d = c.ind[grp.z].w;
      from = g > 0 ? c.groups[g - 1].y : 0;
      to = grp.x;
      reg += c.t[grp.y] - c.t[grp.x];
    }
    const bool wide = to - from >= kLaneCopy;
    if (!wide) {  // loads first (read-only path, four in flight), then the stores: a lane's copy is latency, not work
#pragma unroll 4
      for (int j = from; j <= to; ++j) c.out[j + d] = __ldg(c.t + j);
    }
    unsigned todo = __ballot_sync(kFull, wide);
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const int f = __shfl_sync(kFull, from, src), z = __shfl_sync(kFull, to, src), dd = __shfl_sync(kFull, d, src);
#pragma unroll 4
      for (int j = f + (int)lane; j <= z; j += 32) c.out[j + dd] = __ldg(c.t + j);
    }
  }
  region_samples += __reduce_add_sync(kFull, reg);
#pragma unroll 4
  for (int j = (c.ng > 0 ? c.groups[c.ng - 1].y : 0) + (int)lane; j <= c.last; j += 32) c.out[j + dtotal] = __ldg(c.t + j);
  __syncwarp();
  const int total = c.last + 1 + dtotal;
  int zero_len = 0;
#pragma unroll 4
  for (int i = lane; i + 1 < total; i += 32) zero_len |= (c.out[i + 1] - c.out[i] < 1);
  zero_len = __any_sync(kFull, zero_len);
  if (zero_len) return NRQ_ST_ZERO_LEN;
  if (c.out[0] < 0) return NRQ_ST_NEG_START;
  if (c.out[total - 1] > n_norm) return NRQ_ST_PAST_END;
  if (total != n_bases + 1) return NRQ_ST_SEQ_MISMATCH;
  return NRQ_ST_OK;
}

__device__ __forceinline__ void push_group(SegCtx& c, int start, int stop, int lo, int maxend) {
  if (lane_id() == 0) c.groups[c.ng] = make_int4(start, stop, lo, maxend);
  __syncwarp();
  ++c.ng;
  c.top_stop = stop;
}

// The driver loop of get_indel_groups (:286-314), indel by indel, from a group (lo .. hi, maxend) that is still
// open and indel k as the next one to look at, until that group has been finished and pushed.  On return k is the
// first indel of the next group (n_ind when the read's last group was pushed).  `planned`: the group is complete
// (the next indel does not touch it) and its extend_group result (p_start, p_stop, p_K) is at hand, as are
// `known_none` searches K2a has already lost.  This is the ONE call site of the change-point search in the kernel:
// every inlined copy carries the whole search with it, and with the first searches done by K2a the walk is control
// flow spread over the whole kernel -- its instruction footprint is what the caches have to hold (r2c: 4.3
// 'no instruction' stall cycles per issued instruction with two copies).
__device__ __forceinline__ int drive_until_push(SegCtx& c, int& k, int lo, int hi, int maxend, bool planned,
                                                int p_start, int p_stop, int p_K, int p_t_start, int p_t_stop,
                                                int p_nxt_x, int p_nxt_y, int known_none) {
  int start, stop;
  for (;;) {
    const bool more = k < c.n_ind;
    int nxt_x = p_nxt_x, nxt_y = p_nxt_y;  // (a planned group's next indel is in the planner's registers)
    if (more && !planned) {
      const int4 nxt = c.ind[k];
      nxt_x = nxt.x;
      nxt_y = nxt.y;
      if (maxend >= nxt_x) {  // :291-292
        hi = k++;
        maxend = max(maxend, nxt_y);
        continue;
      }
    }
    // extend_and_join (:217-226) followed by extend_for_cpts (:256-276); :307-314 for the read's last group
    int K, t_start, t_stop, st = NRQ_ST_OK;
    if (planned) {
      start = p_start; stop = p_stop; K = p_K;
      t_start = p_t_start;
      t_stop = p_t_stop;
    } else {
      st = extend_group(c, lo, hi, maxend, start, stop, K, t_start, t_stop);
      known_none = 0;
    }
    if (st != NRQ_ST_OK) return st;
    st = search_group(c, lo, hi, maxend, start, stop, K, t_start, t_stop, known_none);
    if (st != NRQ_ST_OK) return st;
    planned = false;
    if (more && stop >= nxt_x) {  // :299-300
      hi = k++;
      maxend = max(maxend, nxt_y);
      continue;
    }
    push_group(c, start, stop, lo, maxend);  // :302-304
    return NRQ_ST_OK;
  }
}

constexpr int kPlanGeneric = 15;       // record status: take the indel-by-indel path for this group
constexpr int kPlanKBits = 22;         // a record packs K (22 bits), the group's last lane (5) and a status (4)

// get_indel_groups (:278-316).  The grouping of indels by touching padded spans (:291-292) and the first
// extend_group of each group (:203-216) do not depend on how earlier groups fared, so they are worked out for the
// next 32 indels at once, one lane per indel: a running maximum of the indel ends (warp scan) marks where groups
// break, and the lane of each group's first indel grows the group until it can hold its change-points.  The plans
// stay in those lanes' registers.  The leading groups that K2a has settled are finished at once; the first group
// that is not goes through what the reference does for a group (merge with finished groups it runs into, search,
// retry, carry on indel by indel if it ends up reaching the next indel (:299-300) or does not fit the 32-indel
// window: drive_until_push), and planning resumes behind it.
__device__ int resegment_read(SegCtx& c, long long timeout_ns, int n_bases, int n_norm, int& region_samples) {
  const int lane = (int)lane_id();
  const long long t_begin = timeout_ns > 0 ? global_ns() : 0;
  int k = 0;  // first indel that is not part of a finished group
  while (k < c.n_ind) {
    if (timeout_ns > 0 && global_ns() - t_begin > timeout_ns) return NRQ_ST_TIMEOUT;  // :288-289
    // ---- plan the groups among indels k .. k + 31
    const bool valid = k + lane < c.n_ind;
    int4 my = make_int4(0x7fffffff, 0, 0, 0);  // past the last indel: always starts a group, i.e. ends the one before
    if (valid) my = c.ind[k + lane];
    int M = my.y;  // running maximum of the ends: equals the open group's own maximum wherever it matters
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(kFull, M, o);
      if (lane >= o) M = max(M, t);
    }
    const int m_prev = __shfl_up_sync(kFull, M, 1);
    const bool brk = lane == 0 || m_prev < my.x;
    const unsigned heads = __ballot_sync(kFull, brk);
    const unsigned after = lane == 31 ? 0u : heads >> (lane + 1);
    const bool complete = after != 0u;                       // the group's end lies inside the window
    const int e = complete ? lane + __ffs(after) - 1 : lane;  // lane of the group's last indel
    const int gmax = __shfl_sync(kFull, M, e);
    const int zsum = __shfl_sync(kFull, my.w + my.z, e);
    const bool planned = brk && valid && complete;
    int2 spw = make_int2(0, 0);
    if (c.spec && valid) spw = c.spec[k + lane];
    bool sp_found = false, sp_none = false;
    // this lane's plan, if it heads a planned group: extend_group's result (q_*: start, stop, K, the boundaries
    // t[start] / t[stop], a status) and where K2a's retries ended, if it found change-points (p_*).  The plans stay
    // in the head lanes' registers; the walk below fetches one with shuffles.
    int q_start = 0, q_stop = 0, q_K = 0, q_ts = 0, q_te = 0, q_st = NRQ_ST_OK;
    int p_start = 0, p_stop = 0, p_K = 0;
    if (planned) {  // extend_group (:203-216) for the group lane .. e
      int start = my.x, stop = gmax, st = NRQ_ST_OK;
      int K = (zsum - my.w) + stop - start - 1;
      if (start < 0 || stop > c.last) st = NRQ_ST_BAD_ALIGNMENT;
      while (st == NRQ_ST_OK) {
        q_ts = c.t[start];
        q_te = c.t[stop];
        if (q_te - q_ts >= (K + 2) * c.m) break;
        if (start == 0 && stop == c.last) { st = NRQ_ST_UNSATISFIABLE; break; }
        K += (start > 0) + (stop < c.last);
        start = max(0, start - 1);
        stop = min(c.last, stop + 1);
      }
      if (K < 0 || K >= (1 << kPlanKBits)) { st = kPlanGeneric; K = 0; }
      q_start = start; q_stop = stop; q_K = K; q_st = st;
      p_start = start; p_stop = stop; p_K = K;
      // K2a's result counts only if it started from this very plan
      if (st == NRQ_ST_OK && K >= 1 && K <= kSpecMaxK && stop - start <= 255 && e - lane <= kSpecMaxLen) {
        sp_found = spw.x == spec_word(kSpecFound, K, start, stop);
        sp_none = spw.x == spec_word(kSpecNone, K, start, stop);
      }
      if (sp_found) {  // where its retries (:262-266) ended: one base more on either side per retry
        const int nret = spec_retries(spw.y);
        p_K = K + min(nret, start) + min(nret, c.last - stop);
        p_start = max(0, start - nret);
        p_stop = min(c.last, stop + nret);
      }
    }
    const unsigned found_mask = __ballot_sync(kFull, sp_found), none_mask = __ballot_sync(kFull, sp_none);
    unsigned todo = __ballot_sync(kFull, planned);
    __syncwarp();
    int next_k = k;
    bool drive = todo == 0u;  // more than 32 indels in one group
    // ---- the leading planned groups that K2a has already searched successfully, that begin behind the group before
    // them (:220: no merge) and end before the next indel (:299: not left open) are finished as they stand: all of
    // them at once, one lane per group (stack entry, change-points copied into their new_segs slots).
    if (c.spec) {
      const unsigned before = todo & ((1u << lane) - 1u);
      const int prev_lane = before ? 31 - __clz(before) : 0;
      const int prev_sh = __shfl_sync(kFull, p_stop, prev_lane);
      const int prev_stop = before ? prev_sh : c.top_stop;
      const int nxt_x = __shfl_sync(kFull, my.x, (e + 1) & 31);  // a planned group ends inside the window
      const bool clean = sp_found && p_start > prev_stop && p_stop < nxt_x;
      const unsigned bad = todo & ~__ballot_sync(kFull, clean);
      const unsigned prefix = bad ? todo & ((1u << (__ffs(bad) - 1)) - 1u) : todo;
      if (prefix) {
        if ((prefix >> lane) & 1u) {
          c.groups[c.ng + __popc(prefix & ((1u << lane) - 1u))] = make_int4(p_start, p_stop, k + lane, gmax);
          const int t0 = c.t[p_start];
          int32_t* o = c.out + p_start + my.w + 1;
          const unsigned char* src = c.cpt_bytes + spec_offset(spw.y);
#pragma unroll 4
          for (int i = 0; i < p_K; ++i) o[i] = t0 + (int)__ldg(src + i);
        }
        const int last_lane = 31 - __clz(prefix);
        c.ng += __popc(prefix);
        c.top_stop = __shfl_sync(kFull, p_stop, last_lane);
        next_k = k + __shfl_sync(kFull, e, last_lane) + 1;
        todo &= ~prefix;
        __syncwarp();
      }
    }
    // ---- the next group that is not finished yet: taken from its plan if there is one (a speculated find that is
    // not clean, a speculated None, or no speculation at all), else indel by indel -- one call site below
    bool go = drive, g_planned = false;
    int g_lo = k, g_hi = k, g_maxend = 0, g_k = k + 1, g_start = 0, g_stop = 0, g_K = 0, g_known = 0;
    if (drive) g_maxend = c.ind[k].y;
    int g_ts = 0, g_te = 0, g_nx = 0, g_ny = 0;
    if (todo && !go) {  // (one group per window goes this way: whatever follows is planned again)
      const int s = __ffs(todo) - 1;
      const int lo = k + s;
      const int hi = k + __shfl_sync(kFull, e, s);
      const int st = __shfl_sync(kFull, q_st, s);
      int start = __shfl_sync(kFull, q_start, s), stop = __shfl_sync(kFull, q_stop, s);
      const int maxend = __shfl_sync(kFull, gmax, s);
      const int d_before = __shfl_sync(kFull, my.w, s);      // sum of the diffs of the indels before the group
      const int nl = (hi + 1 - k) & 31;                      // the next indel sits in the window (group complete)
      const int nxt_x = __shfl_sync(kFull, my.x, nl), nxt_y = __shfl_sync(kFull, my.y, nl);
      g_lo = lo; g_hi = hi; g_maxend = maxend; g_k = hi + 1;
      if (st == kPlanGeneric) {  // a K that does not fit the plan's fields: same group, from scratch
        go = true;
      } else if (st != NRQ_ST_OK) {
        return st;
      } else {
      const int sp_y = __shfl_sync(kFull, spw.y, s);
      const int f_start = __shfl_sync(kFull, p_start, s), f_stop = __shfl_sync(kFull, p_stop, s);
      const int f_K = __shfl_sync(kFull, p_K, s);
      const bool found = (found_mask >> s) & 1u;
      if (!(found && f_start > c.top_stop)) {
        // to be searched; searches K2a is known to have lost: all of a chain that ended in None, the retries
        // before a find
        go = g_planned = true;
        g_start = start; g_stop = stop; g_K = __shfl_sync(kFull, q_K, s);
        g_ts = __shfl_sync(kFull, q_ts, s);
        g_te = __shfl_sync(kFull, q_te, s);
        g_nx = nxt_x; g_ny = nxt_y;
        g_known = found ? spec_retries(sp_y) : (((none_mask >> s) & 1u) ? spec_retries(sp_y) + 1 : 0);
      } else {
      // searched by K2a to the end, and no finished group in the way (but it reaches the next indel, or it
      // follows a group that was not clean)
      start = f_start;
      stop = f_stop;
      const int t0 = c.t[start];
      int32_t* o = c.out + start + d_before + 1;
      const unsigned char* src = c.cpt_bytes + spec_offset(sp_y);
      for (int i = lane; i < f_K; i += 32) o[i] = t0 + (int)src[i];
      if (hi + 1 < c.n_ind && stop >= nxt_x) {  // :299-300: the finished group reaches the next indel and stays open
        go = true;                              // plans made behind this point may no longer hold
        g_hi = hi + 1; g_maxend = max(maxend, nxt_y); g_k = hi + 2;
      } else {
        push_group(c, start, stop, lo, maxend);
        next_k = hi + 1;
      }
      }
      }
    }
    if (go) {
      const int st = drive_until_push(c, g_k, g_lo, g_hi, g_maxend, g_planned, g_start, g_stop, g_K, g_ts, g_te,
                                      g_nx, g_ny, g_known);
      if (st != NRQ_ST_OK) return st;
      next_k = g_k;
    }
    k = next_k;
  }
  return splice_and_validate(c, n_bases, n_norm, region_samples);
}

// f-4: the running-difference trace plot_correction draws (plot_commands.py:225-229) for one window of an
// already normalised signal: sequential float64 cumsum, then |2 cs[i+m] - cs[i] - cs[i+2m]|.  `cs` is
// caller-provided scratch of n + 1 doubles.  One CTA; the window is a few hundred samples.
__global__ void k_running_diffs(const double* __restrict__ sig, int n, int m, double* __restrict__ cs,
                                double* __restrict__ out) {
  if (threadIdx.x == 0) {
    double acc = 0.0;
    cs[0] = 0.0;
    for (int j = 0; j < n; ++j) { acc = __dadd_rn(acc, sig[j]); cs[j + 1] = acc; }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n - 2 * m + 1; i += blockDim.x)
    out[i] = fabs(__dsub_rn(__dsub_rn(__dmul_rn(2.0, cs[i + m]), cs[i]), cs[i + 2 * m]));
}

#ifndef NRQ_SEG_MINB
#define NRQ_SEG_MINB 6
#endif
__global__ void __launch_bounds__(kSegWarps * 32, NRQ_SEG_MINB)
k_resegment(nrq_batch b, nrq_params p, const double* __restrict__ sv, const int64_t* __restrict__ indel_off,
            const int4* __restrict__ indels, const int32_t* __restrict__ n_bases, int4* __restrict__ group_scratch,
            double* __restrict__ big_scratch, int32_t* __restrict__ work_counter, int32_t* __restrict__ status,
            int32_t* __restrict__ new_segs, int32_t* __restrict__ region_samples,
            int32_t* __restrict__ n_groups, const int2* __restrict__ spec,
            const unsigned char* __restrict__ spec_bytes) {
  struct WarpSm {
    double region[kRegionCap + 2];
    double lut[kSegLut];
    unsigned picks[64];  // picks are at least min_seg_len apart and the region holds <= 254 samples
  };
  __shared__ __align__(16) WarpSm sm[kSegWarps];
  const int warp = threadIdx.x >> 5;
  const unsigned lane = lane_id();
  const int gwarp = blockIdx.x * kSegWarps + warp;
  // the warp's shared-memory byte offset, made opaque so that it stays in one register: left to itself the
  // compiler re-derives it from threadIdx at every use (8 % of the executed instructions, ncu r1m)
  unsigned sm_off = (unsigned)warp * (unsigned)sizeof(WarpSm);
  asm volatile("" : "+r"(sm_off));
  WarpSm& my = *reinterpret_cast<WarpSm*>(reinterpret_cast<char*>(sm) + sm_off);
  for (;;) {
    int r = 0;
    if (lane == 0) r = atomicAdd(work_counter, 1);
    r = __shfl_sync(kFull, r, 0);
    if (r >= b.n_reads) break;
    if (status[r] != NRQ_ST_OK) continue;
    ReadGeom g = read_geom(b, r);
    SegCtx c;
    c.t = g.starts;
    c.last = g.n_starts - 1;
    c.sig = g.sig;
    c.xsig = g.xsig;
    c.coef = load_coef(sv, r);
    c.ind = indels + indel_off[r];
    c.n_ind = (int)(indel_off[r + 1] - indel_off[r]);
    c.groups = group_scratch + indel_off[r];
    c.ng = 0;
    c.top_stop = -1;
    c.out = new_segs + b.align_off[r] + r;
    c.spec = spec ? spec + indel_off[r] : nullptr;
    c.cpt_bytes = spec ? spec_bytes + (size_t)kSpecCptBytesPerIndel * (size_t)indel_off[r] : nullptr;
    // the two pointers the state machine dereferences most are pinned in vector registers; as warp-uniform
    // values the compiler otherwise re-forms their addresses on the uniform datapath at every use (4 instructions)
    asm volatile("" : "+l"(c.t));
    asm volatile("" : "+l"(c.ind));
    c.m = p.min_seg_len;
    c.limit = p.num_cpts_limit;
    c.buf_sm = my.region;
    // table of normalised values over the codes that survive clipping (or around a pivot sample)
    c.lut = my.lut;
    c.picks = my.picks;
    c.lut_lo = (g.xsig ? 0 : (int)g.sig[g.n_norm / 2]) - kSegLut / 2;
    if (c.coef.clip) {
      double e0 = c.coef.shift + c.coef.lo * c.coef.scale, e1 = c.coef.shift + c.coef.hi * c.coef.scale;
      double lo = fmin(e0, e1), hi = fmax(e0, e1);
      if (isfinite(lo) && isfinite(hi) && lo > -40000.0 && hi < 40000.0) {
        if (hi - lo < (double)(kSegLut - 4)) c.lut_lo = (int)floor(lo) - 2;         // every unclipped code
        else c.lut_lo = (int)floor(0.5 * (lo + hi)) - kSegLut / 2;                  // the middle of them
      }
    }
    __syncwarp();
    for (int k = lane; k < kSegLut; k += 32) my.lut[k] = norm_value(c.lut_lo + k, c.coef);
    __syncwarp();
    c.big_cap = p.big_region_cap;
    c.buf_big = big_scratch + (size_t)gwarp * (size_t)(p.big_region_cap + 1);
    int reg = 0;
    int st = resegment_read(c, p.timeout_ns, n_bases[r], g.n_norm, reg);
    if (lane == 0 && st != NRQ_ST_OK) status[r] = st;
    if (lane == 0 && region_samples) region_samples[r] = reg;
    if (lane == 0 && n_groups) n_groups[r] = c.ng;
    __syncwarp();
  }
}

}  // namespace nrq
