// K2a -- speculative first evaluation of every indel group (resquiggle.py:203-216, 227-255).
//
// Which indels chain into a group (:291-292), the group's first extend_group (:203-216) and the first change-point
// search over the resulting region (:227-255) do not depend on how earlier groups of the read fared; only merges
// with finished groups (:220-225, :267-272), retries after a failed search (:262-266) and groups that grow into the
// next indel (:299-300) do.  About 85 % of all get_cpts calls are such first searches.  This kernel runs them for
// the whole read at once, ONE REGION PER THREAD:
//
//   plan     one thread per indel, 256 indels per tile: a block-wide running maximum of the indel ends marks the
//            group heads, the head's thread grows the group until it can hold its change-points and files a task
//            (region start, samples, K) in a shared-memory pool;
//   sort     counting sort of the pool by region size, so that the 32 regions a warp takes cost about the same;
//   search   every lane walks its own region: 128-bit loads of the DAC codes, normalised value from a per-read
//            table, strictly sequential float64 cumulative sum with the last nine sums in registers, running
//            differences |2 cs[i+4] - cs[i] - cs[i+8]| as ordered 64-bit keys into the warp's slab of shared
//            memory (key i of lane l at i * lanes + l: conflict-free), then K rounds of arg-max over the lane's own
//            keys under (value desc, index desc) with the +-window cleared after each pick.  No cross-lane traffic,
//            no shuffles; lanes differ only in trip counts.
//
// A found set of change-points is parked in the read's byte scratch and spec[first indel of the group] says what was
// searched, how it ended and where the change-points lie.  k_resegment (K2b) then walks the read as the reference
// does; where its own plan matches a speculated one and the group does not run into a finished group, it copies the
// result into the group's new_segs slots instead of searching.  Anything not speculated here (groups of more than 31 indels, regions
// of more than 254 samples, pool overflow, min_seg_len != 4) is simply searched there: speculation is optional.
#include "nrq_common.cuh"

namespace nrq {

#ifndef NRQ_SPEC_WARPS
#define NRQ_SPEC_WARPS 8
#endif
#ifndef NRQ_SPEC_POOL
#define NRQ_SPEC_POOL 768
#endif
#ifndef NRQ_SPEC_SLAB
#define NRQ_SPEC_SLAB 1424
#endif
#ifndef NRQ_SPEC_MINB
#define NRQ_SPEC_MINB 2
#endif
constexpr int kSpecWarps = NRQ_SPEC_WARPS;
constexpr int kSpecThreads = kSpecWarps * 32;  // = indels planned per tile
constexpr int kSpecLut = 1024;                 // DAC codes covered by the per-read table of normalised values
constexpr int kSpecPool = NRQ_SPEC_POOL;       // tasks held before the pool is run
constexpr int kSpecSlab = NRQ_SPEC_SLAB;       // 64-bit keys per warp
constexpr int kSpecMaxN = 254;                 // largest region searched here (samples)
constexpr int kSpecMaxK = 61;
constexpr int kSpecMaxLen = 30;                // indels after the head: K2b plans 32-indel windows
constexpr int kSpecM = 4;                      // min_seg_len this kernel is written for

// spec entry of a group head: .x = the group's first plan (so that K2b can tell its own plan is the same one) and
// the outcome, .y = how many retries (:262-266: one base more on either side, K grown to match) came before it and
// where the found change-points lie in the read's byte scratch (one byte each: offset from the region's first
// sample).  A retry chain is followed as long as nothing only the walk can know gets in the way; it assumes that no
// finished group is run into (:267-272), which K2b checks against the final region before it takes the result.  Speculated results never go to new_segs directly: the regions of neighbouring speculated groups may
// overlap, and a group K2b ends up merging would otherwise leave its marks in a neighbour's slots.
constexpr int kSpecFound = 1, kSpecNone = 2;
// Longest chain of retries followed here.  0: a search that comes back None is left to the walk.  Measured (r2e,
// profiles/README.md): with chains followed to the end (24) the walk gets 14 % faster and this kernel 3.3 times
// slower -- a round of retries is a tenth of the tasks of the round before, so most warps of the CTA wait at its
// barriers.  The mechanism stays for a pool that holds the tasks of several reads.
#ifndef NRQ_SPEC_RETRY
#define NRQ_SPEC_RETRY 0
#endif
constexpr int kSpecMaxRetry = NRQ_SPEC_RETRY;
// .y of an entry: byte offset of the change-points (24 bits) | retries before the outcome << 24
__host__ __device__ __forceinline__ int spec_retries(int y) { return (int)((unsigned)y >> 24); }
__host__ __device__ __forceinline__ int spec_offset(int y) { return y & 0xffffff; }
constexpr int kSpecCptBytesPerIndel = 8;  // byte scratch per read: this many bytes per indel
__host__ __device__ __forceinline__ int spec_word(int flag, int K, int start, int stop) {
  return flag | (K << 2) | ((stop - start) << 8) | ((start & 0xffff) << 16);
}

struct __align__(8) SpecTask {
  int t0;       // starts[group start]
  unsigned pk;  // samples (8 bits) | K << 8 (6) | indels after the head << 14 (5) | (stop - start) << 19 (8)
  int start;    // group start (read base index)
  int lo;       // the group's first indel
#if NRQ_SPEC_RETRY > 0
  int word0;    // spec_word(0, ...) of the group's first plan
  int nret;     // retries that led to this task
#endif
};
// (without retry chains a task IS its group's first plan: 16 bytes instead of 24, and the 6 KB that saves in the
// pool go to the slabs -- more lanes per batch)
__device__ __forceinline__ int spec_task_word0(const SpecTask& tk) {
#if NRQ_SPEC_RETRY > 0
  return tk.word0;
#else
  return spec_word(0, (int)((tk.pk >> 8) & 63u), tk.start, tk.start + (int)((tk.pk >> 19) & 255u));
#endif
}
__device__ __forceinline__ int spec_task_nret(const SpecTask& tk) {
#if NRQ_SPEC_RETRY > 0
  return tk.nret;
#else
  return 0;
#endif
}

struct SpecSm {
  double lut[kSpecLut + 2];  // [kSpecLut] = +0.0: what a sample outside the region adds
  long long slab[kSpecWarps][kSpecSlab];
  SpecTask task[kSpecPool];
  unsigned short order[kSpecPool];
  // (the sort's histogram and the two planning arrays -- int hist[256], mi[kSpecThreads]: running maximum of the
  // indel ends, inclusive, zs[kSpecThreads]: sum of diffs up to and including this indel -- live in the first
  // warp's slab: planning, sort and search are separated by block barriers, and the 3 KB buy every warp 48 keys)
  unsigned heads[kSpecWarps];
  int wmax[kSpecWarps];
  int n_tasks, cursor, read, cpt_alloc, n_retry;
};

// Sort bin of a task: tasks of one batch run in lock-step, so they should need the same number of selection rounds
// (K, major) and about the same number of samples (four classes of samples per change-point, minor); bin 0 first.
#ifndef NRQ_SPEC_BINS
#define NRQ_SPEC_BINS 2
#endif
__device__ __forceinline__ int spec_bin(unsigned pk, int off) {  // off: samples of the region's first 16-byte word before it
  const int n = (int)(pk & 255u), K = (int)((pk >> 8) & 63u);
  const int cls = min(3, max(0, (n - 4 * (K + 2)) / (2 * (K + 2))));
#if NRQ_SPEC_BINS == 2
  // K <= 15 (nearly all tasks): one bin per (K, key blocks of the region), so that the lanes of a batch run the same
  // number of cumulative-sum steps and scan the same number of block maxima; above that K in fours with the classes
  if (K > 15) return ((kSpecMaxK - K) >> 2) * 4 + (3 - cls);                     // 0 .. 47
  const int nb = min(12, (n - 2 * kSpecM + 1 + 7) >> 3);                          // 1 .. 12
  return 48 + (15 - K) * 12 + (12 - nb);                                          // 48 .. 227
#elif NRQ_SPEC_BINS == 3
  // ... or per (K, 16-byte words the lane reads): a word costs more than a key block
  if (K > 15) return ((kSpecMaxK - K) >> 2) * 4 + (3 - cls);
  const int nw = min(13, max(2, (n + off + 7) >> 3));                             // 2 .. 13
  return 48 + (15 - K) * 12 + (13 - nw);
#else
  return (kSpecMaxK - K) * 4 + (3 - cls);
#endif
}

// a code outside a table that does not span the clip limits: the division itself (rare; kept out of line)
__device__ __noinline__ double spec_norm_slow(int code, const NormCoef& coef) { return norm_value(code, coef); }

// eight DAC codes as one 16-byte word; a word that sticks out of the raw array is put together sample by sample
__device__ __forceinline__ int4 spec_load_word(const int4* q, const int16_t* raw_lo, const int16_t* raw_hi) {
  const int16_t* qs = reinterpret_cast<const int16_t*>(q);
  if (qs >= raw_lo && qs + 8 <= raw_hi) return __ldg(q);
  int h[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) h[u] = (qs + u >= raw_lo && qs + u < raw_hi) ? (int)qs[u] & 0xffff : 0;
  return make_int4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16));
}

// one batch: lanes 0 .. L-1 each search the region of task order[i0 + lane].  FULL: the table spans the clip limits,
// so a code outside it takes the table's first / last entry.
template <bool FULL>
__device__ __forceinline__ void spec_run_batch(SpecSm& sm, int warp, int lane, int i0, int L, const int16_t* sig,
                                               const int16_t* raw_lo, const int16_t* raw_hi, const NormCoef& coef,
                                               int lut_lo, int2* sp, unsigned char* cpt_bytes, int cpt_cap,
                                               const int32_t* t, int last, int limit, int nt) {
  const bool act = lane < L;
  SpecTask tk;
  tk.t0 = 0; tk.pk = 0u; tk.start = 0; tk.lo = 0;
#if NRQ_SPEC_RETRY > 0
  tk.word0 = 0; tk.nret = 0;
#endif
  if (act) tk = sm.task[sm.order[i0 + lane]];
  const int n = (int)(tk.pk & 255u);
  const int K = (int)((tk.pk >> 8) & 63u);
  const int nrd = act ? n - 2 * kSpecM + 1 : 0;  // len(running_diffs)
  const int nb_max = ((int)__reduce_max_sync(kFull, (unsigned)nrd) + 7) >> 3;  // key blocks of the largest region
  const int nrd_max = 8 * nb_max;
  const int nb_own = (nrd + 7) >> 3;  // this lane's key blocks
  const int col = act ? lane : 0;
  long long* slab = sm.slab[warp];       // key i of this lane at slab[i * L + col], i < 8 nb_max
  int* bmaxh = reinterpret_cast<int*>(slab + nrd_max * L);  // largest high word of block b at bmaxh[b * L + col]
  // bitmap of the lane's picks, one bit per running difference: word w at pick[w * L + col], behind the block maxima
  // (a fixed 256 bits per lane took 8 KB of the CTA's shared memory, i.e. three to four lanes of every batch)
  unsigned* pick = reinterpret_cast<unsigned*>(bmaxh + nb_max * L);
  const int n_pw = (nb_max + 3) >> 2;
  __syncwarp();  // the previous batch's columns are laid out differently
  if (act)
    for (int w = 0; w < n_pw; ++w) pick[w * L + col] = 0u;
  if (act) {  // the tail of the lane's last block is never a candidate
    for (int j = nrd; j < 8 * nb_own; ++j) slab[j * L + col] = -1ll;
  }

  // ---- cumulative sum and keys.  The lane reads whole 16-byte words; samples of the first word that lie before the
  // region count as +0.0 (cs stays +0.0, exactly), so that the position inside the word fixes which of the eight
  // registers holds which sum: w[g & 7] = cs'[g], cs'[g + 1] = cs'[g] + x'[g], region sample j = x'[j + off].
  // A key is the bit pattern of |running difference| (>= 0 as a signed integer); -1 = cleared.
  {
    const int16_t* s = sig + tk.t0;
    const uintptr_t a = reinterpret_cast<uintptr_t>(s);
    const int off = (int)((a >> 1) & 7u);
    const int4* wp = reinterpret_cast<const int4*>(a & ~(uintptr_t)15);
    const int nv = act ? n + off : 0;
    const int iters = (int)__reduce_max_sync(kFull, (unsigned)((nv + 7) >> 3));
    double w[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) w[k] = 0.0;
    double acc = 0.0;
    // every word of the lane's region starts its trip from HBM now (no register held for it); the loop below then
    // finds them in L1
    for (int it = 2; it < iters; ++it)
      if (8 * it < nv) asm volatile("prefetch.global.L1 [%0];" ::"l"(wp + it));
    int4 nxt = make_int4(0, 0, 0, 0);
    if (nv > 0) nxt = spec_load_word(wp, raw_lo, raw_hi);
    int rel = -off;                                  // region index of the word's first sample
    long long* kq = slab + col + (-off - 7) * L;     // &key[rel - 7]
    const int L8 = L;
    for (int it = 0; it < iters; ++it) {
      const int4 v = nxt;
      if (8 * (it + 1) < nv) nxt = spec_load_word(wp + it + 1, raw_lo, raw_hi);  // next word under way
      const int cw[4] = {v.x, v.y, v.z, v.w};
      double x[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int code = (u & 1) ? (cw[u >> 1] >> 16) : (int)(short)(cw[u >> 1] & 0xffff);
        const int k = code - lut_lo;
        const bool in = (unsigned)(rel + u) < (unsigned)n;
        const int idx = in ? min(max(k, 0), kSpecLut - 1) : kSpecLut;
        x[u] = sm.lut[idx];
        if (!FULL) {
          if (in && (unsigned)k >= (unsigned)kSpecLut) x[u] = spec_norm_slow(code, coef);
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const double c = __dadd_rn(acc, x[u]);  // cs'[g + 1], g = 8 it + u
        const int k8 = (u + 1) & 7;
        const double A = w[k8], B = w[(k8 + 4) & 7];  // cs'[g - 7], cs'[g - 3]
        const double rd = __dsub_rn(__dsub_rn(__dadd_rn(B, B), A), c);  // 2 B == B + B exactly
        const long long key = (long long)(((unsigned long long)((unsigned)__double2hiint(rd) & 0x7fffffffu) << 32) |
                                          (unsigned)__double2loint(rd));
        if ((unsigned)(rel + u - 7) < (unsigned)nrd) kq[u * L8] = key;  // running_diffs[rel + u - 7]
        w[k8] = c;
        acc = c;
      }
      rel += 8;
      kq += 8 * L8;
    }
  }

  // ---- K rounds of arg-max over the lane's own keys: larger key first, larger index among equal keys (:244-254 under
  // argsort(kind='stable')[::-1]); the pick's window i-3 .. i+4 is cleared (:249-250).  Per block of eight keys the
  // slab also holds the largest HIGH WORD of its keys (a 32-bit maximum costs one instruction per key, a 64-bit one
  // four): a round finds the largest high word over the blocks, then the exact arg-max among the keys of the
  // block(s) that hold it -- nearly always one block, more only where keys of different blocks agree in their
  // upper 32 bits -- and brings the maxima of the (at most two) blocks the cleared window touches up to date.
  auto block_max_hi = [&](int b) {  // (little endian: the high word is the second int of a key; cleared key: -1)
    const int* q = reinterpret_cast<const int*>(slab + (8 * b) * L + col) + 1;
    int v = q[0];
#pragma unroll
    for (int u = 1; u < 8; ++u) v = max(v, q[2 * u * L]);
    return v;
  };
  for (int b = 0; b < nb_max; ++b) {
    int v = -1;  // blocks beyond the lane's own region: nothing there
    if (b < nb_own) v = block_max_hi(b);
    if (act) bmaxh[b * L + col] = v;
  }
  const int Kmax = (int)__reduce_max_sync(kFull, (unsigned)K);
  bool fail = false;
  for (int round = 0; round < Kmax; ++round) {
    if (round < K && !fail) {
      int best_hi = -1;      // every real key has a high word >= 0, a cleared one -1
      unsigned tied = 0u;    // blocks whose maximum is best_hi
      const int* q = bmaxh + col;
      for (int b = 0; b < nb_max; ++b, q += L) {
        const int h = q[0];
        tied = h > best_hi ? 1u << b : (h == best_hi ? tied | (1u << b) : tied);
        best_hi = max(best_hi, h);
      }
      if (best_hi < 0) {
        fail = true;  // fewer than K placeable: None (:253-254)
      } else {
        long long best = -1ll;
        int bi = 0;
        for (unsigned tb = tied; tb; tb &= tb - 1u) {  // ascending blocks, ascending keys, '>=': the larger index wins
          const int b = __ffs(tb) - 1;
          const long long* kq = slab + (8 * b) * L + col;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const long long kv = kq[u * L];
            if (kv >= best) { best = kv; bi = 8 * b + u; }
          }
        }
        pick[(bi >> 5) * L + col] |= 1u << (bi & 31);
#pragma unroll
        for (int d = -kSpecM + 1; d <= kSpecM; ++d) {
          const int i = bi + d;
          if ((unsigned)i < (unsigned)(8 * nb_own)) slab[i * L + col] = -1ll;
        }
        const int b_lo = max(bi - kSpecM + 1, 0) >> 3, b_hi = min(bi + kSpecM, 8 * nb_own - 1) >> 3;
        bmaxh[b_lo * L + col] = block_max_hi(b_lo);
        if (b_hi != b_lo) bmaxh[b_hi * L + col] = block_max_hi(b_hi);
      }
    }
    if (!__any_sync(kFull, round + 1 < K && !fail)) break;
  }

  // ---- found: sorted(cpt + min_seg_len) (:255) as byte offsets from the region's first sample, into the read's
  // scratch.  None: the retry of extend_for_cpts (:262-266) becomes a task of the next round.
  if (act) {
    const int start = tk.start, stop = start + (int)((tk.pk >> 19) & 255u);
    int2 entry = make_int2(0, 0);  // "not speculated"
    if (!fail) {
      const int base = atomicAdd(&sm.cpt_alloc, K);
      if (base + K <= min(cpt_cap, 1 << 24)) {
        entry = make_int2(spec_task_word0(tk) | kSpecFound, base | (spec_task_nret(tk) << 24));
        unsigned char* o = cpt_bytes + base;
        int c = 0;
        for (int w = 0; w * 32 < nrd; ++w) {
          unsigned word = pick[w * L + col];
          while (word) {
            const int bpos = __ffs(word) - 1;
            word &= word - 1;
            o[c++] = (unsigned char)(w * 32 + bpos + kSpecM);
          }
        }
      }
    } else {
      bool again = spec_task_nret(tk) < kSpecMaxRetry && !(start == 0 && stop == last);
      SpecTask nx = tk;
      if (again) {
        const int K2 = K + (start > 0) + (stop < last);
        const int s2 = max(0, start - 1), e2 = min(last, stop + 1);
        const int ts = t[s2], n2 = t[e2] - ts;
        again = K2 <= kSpecMaxK && n2 <= kSpecMaxN && e2 - s2 <= 255 && (limit < 0 || K2 <= limit);
        nx.t0 = ts;
        nx.pk = (unsigned)n2 | ((unsigned)K2 << 8) | (tk.pk & (31u << 14)) | ((unsigned)(e2 - s2) << 19);
        nx.start = s2;
#if NRQ_SPEC_RETRY > 0
        nx.nret = tk.nret + 1;
#endif
      }
      if (again) {
        const int slot = nt + atomicAdd(&sm.n_retry, 1);
        again = slot < kSpecPool;
        if (again) sm.task[slot] = nx;
      }
      if (again) entry.x = -1;  // outcome still to come
      else entry = make_int2(spec_task_word0(tk) | kSpecNone, spec_task_nret(tk) << 24);  // nret + 1 searches without merges: all None
    }
    if (entry.x != -1) sp[tk.lo] = entry;
  }
}

__global__ void __launch_bounds__(kSpecThreads, NRQ_SPEC_MINB)
k_speculate(nrq_batch b, nrq_params p, const double* __restrict__ sv, const int64_t* __restrict__ indel_off,
            const int4* __restrict__ indels, int32_t* __restrict__ work_counter,
            const int32_t* __restrict__ status, int2* __restrict__ spec, unsigned char* __restrict__ spec_bytes) {
  extern __shared__ __align__(16) unsigned char spec_smem[];
  SpecSm& sm = *reinterpret_cast<SpecSm*>(spec_smem);
  const int tid = (int)threadIdx.x;
  const int lane = (int)lane_id();
  const int warp = __shfl_sync(kFull, tid >> 5, 0);  // provably warp-uniform: no convergence barriers around shuffles
  const int16_t* raw_lo = b.raw;
  const int16_t* raw_hi = b.raw + b.raw_off[b.n_reads];
  const int m = p.min_seg_len;
  int* const hist_ = reinterpret_cast<int*>(sm.slab[0]);
  int* const mi_ = hist_ + 256;
  int* const zs_ = mi_ + kSpecThreads;
  static_assert((256 + 2 * kSpecThreads) * sizeof(int) <= kSpecSlab * sizeof(long long), "planning arrays in a slab");
  for (;;) {
    __syncthreads();
    if (tid == 0) {
      sm.read = atomicAdd(work_counter, 1);
      sm.n_tasks = 0;
      sm.cpt_alloc = 0;
    }
    __syncthreads();
    const int r = sm.read;
    if (r >= b.n_reads) break;
    if (status[r] != NRQ_ST_OK) continue;
    const ReadGeom g = read_geom(b, r);
    const int last = g.n_starts - 1;
    const int sig_al = (int)((reinterpret_cast<uintptr_t>(g.sig) >> 1) & 7u);
    const int32_t* t = g.starts;
    const NormCoef coef = load_coef(sv, r);
    const int64_t ioff = indel_off[r];
    const int4* ind = indels + ioff;
    const int n_ind = (int)(indel_off[r + 1] - ioff);
    int2* sp = spec + ioff;
    unsigned char* cpt_bytes = spec_bytes + (size_t)kSpecCptBytesPerIndel * (size_t)ioff;
    const int cpt_cap = kSpecCptBytesPerIndel * n_ind;

    // ---- table of normalised values: every code between the clip limits when they are close enough together
    int lut_lo = (int)g.sig[g.n_norm / 2] - kSpecLut / 2;
    bool lut_full = false;
    if (coef.clip) {
      const double e0 = coef.shift + coef.lo * coef.scale, e1 = coef.shift + coef.hi * coef.scale;
      const double lo = fmin(e0, e1), hi = fmax(e0, e1);
      if (isfinite(lo) && isfinite(hi) && lo > -40000.0 && hi < 40000.0) {
        if (hi - lo < (double)(kSpecLut - 4)) {
          lut_lo = (int)floor(lo) - 2;
          lut_full = true;
        } else {
          lut_lo = (int)floor(0.5 * (lo + hi)) - kSpecLut / 2;
        }
      }
    }
    for (int k = tid; k < kSpecLut; k += kSpecThreads) sm.lut[k] = norm_value(lut_lo + k, coef);
    if (tid < 2) sm.lut[kSpecLut + tid] = 0.0;
    __syncthreads();
    // codes outside a full table take the table's end values: both ends must already be clipped
    {
      const double e0 = sm.lut[0], e1 = sm.lut[kSpecLut - 1];
      lut_full = lut_full && (e0 == coef.lo || e0 == coef.hi) && (e1 == coef.lo || e1 == coef.hi) && e0 != e1;
    }

    int tile = 0;
    int carry = (int)0x80000000;  // running maximum of the ends of every indel before the tile
    for (;;) {
      const bool more = tile < n_ind;
      __syncthreads();
      const int nt_now = min(sm.n_tasks, kSpecPool);
      if ((!more && nt_now > 0) || nt_now > kSpecPool - 64) {
        // ---- run the pool: counting sort by cost, then warps take batches; searches that came back None leave their
        // retry in the pool's tail, which is the pool of the next round
        int nt = nt_now;
        for (;;) {
        for (int k = tid; k < 256; k += kSpecThreads) hist_[k] = 0;
        __syncthreads();
        for (int i = tid; i < nt; i += kSpecThreads)
          atomicAdd(&hist_[spec_bin(sm.task[i].pk, (sig_al + sm.task[i].t0) & 7)], 1);
        __syncthreads();
        if (warp == 0) {
          int v[8], s = 0;
#pragma unroll
          for (int q = 0; q < 8; ++q) { v[q] = hist_[8 * lane + q]; s += v[q]; }
          int incl = s;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += u;
          }
          int run = incl - s;
#pragma unroll
          for (int q = 0; q < 8; ++q) { hist_[8 * lane + q] = run; run += v[q]; }
        }
        __syncthreads();
        for (int i = tid; i < nt; i += kSpecThreads)
          sm.order[atomicAdd(&hist_[spec_bin(sm.task[i].pk, (sig_al + sm.task[i].t0) & 7)], 1)] = (unsigned short)i;
        if (tid == 0) { sm.cursor = 0; sm.n_retry = 0; }
        __syncthreads();
        for (;;) {
          // the next L tasks in sorted order: as many as the warp's slab holds (keys in blocks of 8 + a 32-bit maximum
          // per block, for the largest region among them)
          int i0 = 0, L = 0;
          for (;;) {
            if (lane == 0) i0 = *reinterpret_cast<volatile int*>(&sm.cursor);
            i0 = __shfl_sync(kFull, i0, 0);
            if (i0 >= nt) { L = 0; break; }
            int nrd_l = 0;
            if (i0 + lane < nt) nrd_l = (int)(sm.task[sm.order[i0 + lane]].pk & 255u) - 2 * kSpecM + 1;
            int pm = nrd_l;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              const int u = __shfl_up_sync(kFull, pm, o);
              if (lane >= o) pm = max(pm, u);
            }
            // per lane: 8 keys and half a key's worth of block maximum for every block of the largest region, and a
            // bit per key for the picks -- block maxima and pick words counted together in 4-byte units, rounded up
            // to whole keys over the lanes taken
            const int nbp = (pm + 7) >> 3;
            const int words = nbp + ((nbp + 3) >> 2);
            const bool fits = i0 + lane < nt && (lane + 1) * 8 * nbp + (((lane + 1) * words + 1) >> 1) <= kSpecSlab;
            L = __popc(__ballot_sync(kFull, fits));  // a prefix of the lanes: the need only grows with the lane
            int won = 0;
            if (lane == 0) won = atomicCAS(&sm.cursor, i0, i0 + L) == i0;
            if (__shfl_sync(kFull, won, 0)) break;
          }
          if (L == 0) break;
          if (lut_full)
            spec_run_batch<true>(sm, warp, lane, i0, L, g.sig, raw_lo, raw_hi, coef, lut_lo, sp, cpt_bytes, cpt_cap,
                                 t, last, p.num_cpts_limit, nt);
          else
            spec_run_batch<false>(sm, warp, lane, i0, L, g.sig, raw_lo, raw_hi, coef, lut_lo, sp, cpt_bytes, cpt_cap,
                                  t, last, p.num_cpts_limit, nt);
        }
        __syncthreads();
        const int n_again = min(sm.n_retry, kSpecPool - nt);  // (one retry per task at most: n_again <= nt)
        if (n_again <= 0) break;
        for (int i = tid; i < n_again; i += kSpecThreads) sm.task[i] = sm.task[nt + i];
        nt = n_again;
        __syncthreads();
        }
        if (tid == 0) sm.n_tasks = 0;
        __syncthreads();
      }
      if (!more) break;

      // ---- plan the groups whose first indel lies in this tile
      const int idx = tile + tid;
      const bool valid = idx < n_ind;
      int4 my = make_int4(0x7fffffff, (int)0x80000000, 0, 0);  // past the last indel: ends the group before it
      if (valid) {
        my = ind[idx];
        sp[idx] = make_int2(0, 0);  // "not speculated" unless a task of this launch says otherwise
        // planning is a chain of trips to memory (records, then the boundaries extend_group compares): the
        // boundaries around this indel and the records of the next tile start their trip now
        asm volatile("prefetch.global.L1 [%0];" ::"l"(t + max(my.x - 1, 0)));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(t + min(my.y + 1, last)));
        if (idx + kSpecThreads < n_ind) asm volatile("prefetch.global.L1 [%0];" ::"l"(ind + idx + kSpecThreads));
      }
      int M = my.y;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(kFull, M, o);
        if (lane >= o) M = max(M, u);
      }
      if (lane == 31) sm.wmax[warp] = M;
      __syncthreads();
      int pm = carry;
      for (int w2 = 0; w2 < warp; ++w2) pm = max(pm, sm.wmax[w2]);
      const int m_up = __shfl_up_sync(kFull, M, 1);
      const int m_prev = lane ? max(pm, m_up) : pm;
      M = max(M, pm);
      const bool brk = m_prev < my.x;  // no earlier indel reaches this one (:291-292)
      const unsigned hm = __ballot_sync(kFull, brk);
      if (lane == 0) sm.heads[warp] = hm;
      mi_[tid] = M;
      zs_[tid] = my.w + my.z;
      __syncthreads();
      int nh = -1;  // the next group's head inside the tile
      {
        const unsigned above = lane == 31 ? 0u : hm >> (lane + 1);
        if (above) {
          nh = tid + __ffs(above);
        } else {
          for (int w2 = warp + 1; w2 < kSpecWarps; ++w2) {
            const unsigned h2 = sm.heads[w2];
            if (h2) { nh = w2 * 32 + __ffs(h2) - 1; break; }
          }
        }
      }
      int lh = 0;  // the tile's last head: the next tile starts there (its group may end beyond this tile)
      for (int w2 = kSpecWarps - 1; w2 >= 0; --w2) {
        const unsigned h2 = sm.heads[w2];
        if (h2) { lh = w2 * 32 + 31 - __clz(h2); break; }
      }
      if (brk && valid && nh >= 0) {  // extend_group (:203-216) for indels tid .. nh - 1
        const int e = nh - 1;
        const int gmax = mi_[e];
        int start = my.x, stop = gmax;
        int K = (zs_[e] - my.w) + stop - start - 1;
        bool ok = !(start < 0 || stop > last);
        int ts = 0, te = 0;
        while (ok) {
          ts = t[start];
          te = t[stop];
          if (te - ts >= (K + 2) * m) break;
          if (start == 0 && stop == last) { ok = false; break; }
          K += (start > 0) + (stop < last);
          start = max(0, start - 1);
          stop = min(last, stop + 1);
        }
        const int n = te - ts;
        ok = ok && e - tid <= kSpecMaxLen && K >= 1 && K <= kSpecMaxK && n <= kSpecMaxN && stop - start <= 255 &&
             (p.num_cpts_limit < 0 || K <= p.num_cpts_limit);
        if (ok) {
          const int slot = atomicAdd(&sm.n_tasks, 1);
          if (slot < kSpecPool) {
            SpecTask tk;
            tk.t0 = ts;
            tk.pk = (unsigned)n | ((unsigned)K << 8) | ((unsigned)(e - tid) << 14) | ((unsigned)(stop - start) << 19);
            tk.start = start;
            tk.lo = idx;
#if NRQ_SPEC_RETRY > 0
            tk.word0 = spec_word(0, K, start, stop);
            tk.nret = 0;
#endif
            sm.task[slot] = tk;
          }
        }
      }
      if (lh > 0) {
        carry = mi_[lh - 1];
        tile += lh;
      } else {  // one group fills the tile (or the tile lies inside one): left to K2b
        carry = mi_[kSpecThreads - 1];
        tile += kSpecThreads;
      }
    }
  }
}

}  // namespace nrq
