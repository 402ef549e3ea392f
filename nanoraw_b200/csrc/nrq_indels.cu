// K0 -- get_all_indels (nanoraw/resquiggle.py:142-201) for a batch of reads.
//
// One warp per read walks the two alignment rows.  Two launches: a counting pass (also yields Bg/Br and
// validates the alignment) and, after an exclusive scan of the counts, a fill pass that writes one
// (start, end, diff, diff_prefix) int4 per indel.
//
// Wide path (both row buffers 16-byte aligned, the normal case): 512 columns per step, each lane takes 16
// consecutive columns with one 128-bit load per row, turns them into 16-bit gap masks with byte-parallel
// integer arithmetic, finds run starts / ends with shifts on those masks, and gets its run, read-base and
// genome-base offsets from one packed warp scan.  Byte path (any alignment): 32 columns per step, one byte per
// lane, masks from ballots.
#include "nrq_common.cuh"

namespace nrq {

#ifndef NRQ_K0_WARPS
#define NRQ_K0_WARPS 8
#endif
constexpr int kWarpsPerCta = NRQ_K0_WARPS;  // one read per warp

struct ColMasks {
  unsigned mr, mg, valid;
};

__device__ __forceinline__ ColMasks load_cols(const uint8_t* __restrict__ ra,
                                              const uint8_t* __restrict__ ga, int base, int n) {
  int c = base + (int)lane_id();
  bool ok = c < n;
  bool rg = ok && ra[c] == NRQ_GAP;
  bool gg = ok && ga[c] == NRQ_GAP;
  ColMasks m;
  m.mr = __ballot_sync(0xffffffffu, rg);
  m.mg = __ballot_sync(0xffffffffu, gg);
  m.valid = __ballot_sync(0xffffffffu, ok);
  return m;
}


// ---------------------------------------------------------------- wide path helpers
// bit k = byte k of w is the gap character (exact for every byte value)
__device__ __forceinline__ unsigned gap_bits4(unsigned w) {
  const unsigned t = w ^ (0x01010101u * (unsigned)NRQ_GAP);
  unsigned y = (t & 0x7f7f7f7fu) + 0x7f7f7f7fu;
  y = ~(y | t | 0x7f7f7f7fu);               // 0x80 in every zero byte of t
  return ((y >> 7) * 0x01020408u) >> 24;    // gathers bits 0, 8, 16, 24 into bits 24..27
}

__device__ __forceinline__ unsigned gap_bits16(const uint4& q) {
  return gap_bits4(q.x) | (gap_bits4(q.y) << 4) | (gap_bits4(q.z) << 8) | (gap_bits4(q.w) << 12);
}

// 16 bytes at p (16-byte aligned); bytes outside [lo, hi) are never touched
__device__ __forceinline__ uint4 load16(const uint8_t* p, const uint8_t* lo, const uint8_t* hi) {
  if (p >= lo && p + 16 <= hi) return __ldg(reinterpret_cast<const uint4*>(p));
  unsigned w[4] = {0u, 0u, 0u, 0u};
  for (int j = 0; j < 16; ++j)
    if (p + j >= lo && p + j < hi) w[j >> 2] |= (unsigned)p[j] << (8 * (j & 3));
  return make_uint4(w[0], w[1], w[2], w[3]);
}

struct Cols16 {
  unsigned mr, mg, valid;  // 16-bit masks: bit i = column v0 + i - mis is a read gap / genome gap / exists
  uint4 g;                 // the genome row bytes
};

// lane's 16 columns of the step starting at virtual column vb (virtual column v = column + mis, so that
// v % 16 == 0 falls on a 16-byte boundary of both rows)
__device__ __forceinline__ Cols16 load_cols16(const uint8_t* ra, const uint8_t* ga, const uint8_t* ra_lo,
                                              const uint8_t* ra_hi, const uint8_t* ga_lo, const uint8_t* ga_hi,
                                              int mis, int n, int vb) {
  const int v0 = vb + 16 * (int)lane_id();
  Cols16 c;
  const int first = max(mis - v0, 0), last = min(n + mis - v0, 16);  // valid bits [first, last)
  c.valid = last > first ? (((1u << last) - 1u) & ~((1u << first) - 1u)) : 0u;
  c.mr = c.mg = 0u;
  c.g = make_uint4(0u, 0u, 0u, 0u);
  if (c.valid) {
    const uint4 r = load16(ra - mis + v0, ra_lo, ra_hi);
    c.g = load16(ga - mis + v0, ga_lo, ga_hi);
    c.mr = gap_bits16(r) & c.valid;
    c.mg = gap_bits16(c.g) & c.valid;
  }
  return c;
}

template <bool WIDE>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
k_align_count(nrq_batch b, int32_t* __restrict__ status, int32_t* __restrict__ n_indels,
              int32_t* __restrict__ n_bases) {
  // lane-0 broadcast: marks r (and what is loaded through it) warp-uniform for the compiler
  int r = blockIdx.x * kWarpsPerCta + __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  if (r >= b.n_reads) return;
  int64_t a0 = b.align_off[r];
  int n = (int)(b.align_off[r + 1] - a0);
  const uint8_t* ra = b.read_align + a0;
  const uint8_t* ga = b.genome_align + a0;
  unsigned carry_r = 0, carry_g = 0;
  int runs = 0, gbases = 0, rbases = 0, bad = 0;
  if (WIDE) {
    const int64_t total = b.align_off[b.n_reads];
    const int mis = (int)(a0 & 15);
    const int lane = (int)lane_id();
    unsigned carry = 0;  // bit 0 / 1: the column before this step is a read / genome gap
    for (int vb = 0; vb < n + mis; vb += 512) {
      const Cols16 c = load_cols16(ra, ga, b.read_align, b.read_align + total, b.genome_align,
                                   b.genome_align + total, mis, n, vb);
      const unsigned tail = (c.mr >> 15) | ((c.mg >> 15) << 1);
      unsigned before = __shfl_up_sync(0xffffffffu, tail, 1);
      if (lane == 0) before = carry;
      carry = __shfl_sync(0xffffffffu, tail, 31);
      const unsigned sr = c.mr & ~((c.mr << 1) | (before & 1u));
      const unsigned sg = c.mg & ~((c.mg << 1) | (before >> 1));
      runs += __popc(sr & 0xffffu) + __popc(sg & 0xffffu);
      gbases += __popc(~c.mg & c.valid);
      rbases += __popc(~c.mr & c.valid);
      bad |= (c.mr & c.mg) != 0u;
      const int v0 = vb + 16 * lane;
      const unsigned any = c.mr | c.mg;
      if ((unsigned)(mis - v0) < 16u) bad |= (any >> (mis - v0)) & 1u;                  // must start on a match column
      if ((unsigned)(n - 1 + mis - v0) < 16u) bad |= (any >> (n - 1 + mis - v0)) & 1u;  // and end on one
    }
    runs = __reduce_add_sync(0xffffffffu, runs);
    gbases = __reduce_add_sync(0xffffffffu, gbases);
    rbases = __reduce_add_sync(0xffffffffu, rbases);
    bad = __any_sync(0xffffffffu, bad);
  } else
  for (int base = 0; base < n; base += 32) {
    ColMasks m = load_cols(ra, ga, base, n);
    unsigned sr = m.mr & ~((m.mr << 1) | carry_r);
    unsigned sg = m.mg & ~((m.mg << 1) | carry_g);
    runs += __popc(sr) + __popc(sg);
    gbases += __popc(~m.mg & m.valid);
    rbases += __popc(~m.mr & m.valid);
    bad |= (m.mr & m.mg) != 0;
    if (base == 0) bad |= ((m.mr | m.mg) & 1u) != 0;            // must start on a match column
    if (base + 32 >= n) bad |= (((m.mr | m.mg) >> ((n - 1 - base) & 31)) & 1u) != 0;  // and end on one
    carry_r = m.mr >> 31;
    carry_g = m.mg >> 31;
  }
  if (lane_id() == 0) {
    int n_starts = (int)(b.starts_off[r + 1] - b.starts_off[r]);
    int st = NRQ_ST_OK;
    if (n <= 0 || n_starts < 2) st = NRQ_ST_BAD_INPUT;
    else if (bad) st = NRQ_ST_BAD_ALIGNMENT;
    else if (rbases + 1 != n_starts) st = NRQ_ST_BAD_INPUT;
    n_indels[r] = (st == NRQ_ST_OK) ? runs : 0;
    n_bases[r] = gbases;
    status[r] = st;
  }
}

// exclusive scan of int32 counts into int64 offsets; one CTA (R is at most a few million)
__global__ void __launch_bounds__(1024) k_scan_counts(const int32_t* __restrict__ cnt, int n,
                                                     int64_t* __restrict__ off) {
  __shared__ long long part[1024];
  int t = threadIdx.x;
  int per = (n + 1023) / 1024;
  int lo = t * per, hi = min(n, lo + per);
  long long s = 0;
  for (int i = lo; i < hi; ++i) s += cnt[i];
  part[t] = s;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {
    long long v = (t >= d) ? part[t - d] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  long long run = part[t] - s;
  for (int i = lo; i < hi; ++i) {
    off[i] = run;
    run += cnt[i];
  }
  if (t == 1023) off[n] = part[1023];
}

__device__ __forceinline__ int pymod(int a, int n) {
  int m = a % n;
  return m < 0 ? m + n : m;
}

template <bool WIDE>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
k_align_fill(nrq_batch b, int32_t* status, const int64_t* __restrict__ indel_off,
             int4* __restrict__ indels, int64_t indel_capacity, uint8_t* __restrict__ ev_base) {
  // lane-0 broadcast: marks r (and what is loaded through it) warp-uniform for the compiler
  int r = blockIdx.x * kWarpsPerCta + __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  if (r >= b.n_reads || status[r] != NRQ_ST_OK) return;
  __shared__ __align__(16) uint8_t stage[kWarpsPerCta][512 + 48];  // one step's genome bases, see below
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const unsigned lane = lane_id();
  const unsigned lt = (1u << lane) - 1u;
  int64_t a0 = b.align_off[r];
  int n = (int)(b.align_off[r + 1] - a0);
  const uint8_t* ra = b.read_align + a0;
  const uint8_t* ga = b.genome_align + a0;
  int4* out = indels + indel_off[r];
  int n_ind = (int)(indel_off[r + 1] - indel_off[r]);
  if (indel_off[r + 1] > indel_capacity) {  // caller's indel buffer is too small for this read
    if (lane == 0) status[r] = NRQ_ST_WORKSPACE;
    return;
  }
  uint8_t* bases = ev_base ? ev_base + a0 : nullptr;  // align_seq: genome row without gaps (resquiggle.py:384)

  // pass 1: (run start column, run end column, read bases before the run, gap-in-read?)
  unsigned carry_r = 0, carry_g = 0;
  int idx_base = 0, rbase = 0, gbase = 0;
  if (WIDE) {
    const int64_t total = b.align_off[b.n_reads];
    const int mis = (int)(a0 & 15);
    unsigned carry = 0;
    for (int vb = 0; vb < n + mis; vb += 512) {
      const Cols16 c = load_cols16(ra, ga, b.read_align, b.read_align + total, b.genome_align,
                                   b.genome_align + total, mis, n, vb);
      const unsigned tail = (c.mr >> 15) | ((c.mg >> 15) << 1);
      unsigned before = __shfl_up_sync(0xffffffffu, tail, 1);
      if (lane == 0) before = carry;
      carry = __shfl_sync(0xffffffffu, tail, 31);
      const unsigned prev_r = ((c.mr << 1) | (before & 1u)) & 0xffffu, prev_g = ((c.mg << 1) | (before >> 1)) & 0xffffu;
      const unsigned sr = c.mr & ~prev_r, sg = c.mg & ~prev_g;
      const unsigned sany = sr | sg;
      const unsigned eany = ((~c.mr & prev_r) | (~c.mg & prev_g)) & c.valid;
      const unsigned rmask = ~c.mr & c.valid, gmask = ~c.mg & c.valid;
      // one scan for three lane offsets: runs, read bases, genome bases (each at most 512 per step)
      const unsigned mine = (unsigned)__popc(sany) | ((unsigned)__popc(rmask) << 10) | ((unsigned)__popc(gmask) << 20);
      unsigned inc = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, inc, o);
        if ((int)lane >= o) inc += v;
      }
      const unsigned tot = __shfl_sync(0xffffffffu, inc, 31);
      const unsigned ex = inc - mine;
      const int run0 = idx_base + (int)(ex & 1023u), r0 = rbase + (int)((ex >> 10) & 1023u);
      const int col0 = vb + 16 * (int)lane - mis;
      int k = 0;
      for (unsigned todo = sany; todo; todo &= todo - 1u, ++k) {
        const int i = __ffs(todo) - 1;
        // one 16-byte store; the end column (.y) comes from whichever lane sees the run end, possibly a later step
        out[run0 + k] = make_int4(col0 + i, 0, r0 + __popc(rmask & ((1u << i) - 1u)), (int)((sr >> i) & 1u));
      }
      __syncwarp();  // a run may start in one lane and end in the next: starts before ends
      for (unsigned todo = eany; todo; todo &= todo - 1u) {
        const int i = __ffs(todo) - 1;
        out[run0 + __popc(sany & ((1u << i) - 1u)) - 1].y = col0 + i;
      }
      if (bases) {
        // genome bases of this step, gaps squeezed out: each lane packs its own into shared memory, placed so that
        // 16-byte chunks of the staging buffer line up with 16-byte chunks of the output, then the warp copies
        // whole chunks with 128-bit stores (scattered byte stores to global cost a sector each)
        uint8_t* const dst0 = bases + gbase;
        const int lo = (int)((uintptr_t)dst0 & 15u), hi = lo + (int)(tot >> 20);
        uint8_t* sh = stage[warp] + lo + (int)(ex >> 20);
        const unsigned w[4] = {c.g.x, c.g.y, c.g.z, c.g.w};
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if ((gmask >> i) & 1u) *sh++ = (uint8_t)(w[i >> 2] >> (8 * (i & 3)));
        __syncwarp();
        const int first_full = (lo + 15) >> 4, last_full = hi >> 4;  // whole chunks [first_full, last_full)
        for (int ch = first_full + (int)lane; ch < last_full; ch += 32)
          *reinterpret_cast<uint4*>(dst0 - lo + 16 * ch) = *reinterpret_cast<const uint4*>(stage[warp] + 16 * ch);
        const int head_end = min(16 * first_full, hi);
        int j = lo + (int)lane;
        if (j < head_end) dst0[j - lo] = stage[warp][j];
        j = max(16 * last_full, head_end) + (int)lane;
        if (j < hi) dst0[j - lo] = stage[warp][j];
        __syncwarp();
      }
      idx_base += (int)(tot & 1023u);
      rbase += (int)((tot >> 10) & 1023u);
      gbase += (int)(tot >> 20);
    }
  } else
  for (int base = 0; base < n; base += 32) {
    ColMasks m = load_cols(ra, ga, base, n);
    if (bases) {
      unsigned gm = ~m.mg & m.valid;
      if ((gm >> lane) & 1u) bases[gbase + __popc(gm & lt)] = ga[base + (int)lane];
      gbase += __popc(gm);
    }
    unsigned prev_r = (m.mr << 1) | carry_r, prev_g = (m.mg << 1) | carry_g;
    unsigned sr = m.mr & ~prev_r, sg = m.mg & ~prev_g;
    unsigned er = ~m.mr & prev_r & m.valid, eg = ~m.mg & prev_g & m.valid;
    unsigned sany = sr | sg;
    int before = idx_base + __popc(sany & lt);
    int c = base + (int)lane;
    if ((sany >> lane) & 1u) {
      out[before].x = c;
      out[before].z = rbase + __popc(~m.mr & m.valid & lt);
      out[before].w = (int)((sr >> lane) & 1u);
    }
    if (((er | eg) >> lane) & 1u) out[before - 1].y = c;
    idx_base += __popc(sany);
    rbase += __popc(~m.mr & m.valid);
    carry_r = m.mr >> 31;
    carry_g = m.mg >> 31;
  }
  if (n_ind == 0) return;
  __syncwarp();

  // pass 2: ambiguity extension (resquiggle.py:187-195) and read coordinates, 32 indels at a time
  int prev_end_carry = 0;  // end column of the indel before this chunk
  int dpre = 0;
  for (int base = 0; base < n_ind; base += 32) {
    int k = base + (int)lane;
    bool ok = k < n_ind;
    int4 me = ok ? out[k] : make_int4(0, 0, 0, 0);
    int up_e = __shfl_up_sync(0xffffffffu, me.y, 1);
    int prev_end = (lane == 0) ? prev_end_carry : up_e;
    int next_start = n;
    if (ok && k + 1 < n_ind) next_start = out[k + 1].x;
    int last_e = __shfl_sync(0xffffffffu, me.y, 31);
    __syncwarp();
    int start = 0, end = 0, diff = 0;
    if (ok) {
      int s = me.x, e = me.y, crl = me.z, ins = me.w;
      int len = e - s;
      const uint8_t* seq = (ins ? ga : ra) + s;  // :160-164
      int n_before = s - prev_end, n_after = next_start - e;
      // seq[d % len] and seq[u mod len] walk the run cyclically: indices kept by stepping, no division
      int d = 0, jd = 0;
      while (d < n_after - 1 && seq[jd] == ga[e + d]) {                   // :188-190
        ++d;
        if (++jd == len) jd = 0;
      }
      int u = -1, ju = len - 1;
      while (-u <= n_before - 1 && seq[ju] == ga[s + u]) {                // :191-193
        --u;
        if (--ju < 0) ju = len - 1;
      }
      start = crl + u;
      end = (ins ? crl + 1 : crl + len + 1) + d;                          // :178-179
      diff = ins ? len : -len;                                             // :180
    }
    // exclusive prefix of diff
    int inc = diff;
    for (int o = 1; o < 32; o <<= 1) {
      int v = __shfl_up_sync(0xffffffffu, inc, o);
      if ((int)lane >= o) inc += v;
    }
    if (ok) out[k] = make_int4(start, end, diff, dpre + inc - diff);
    dpre += __shfl_sync(0xffffffffu, inc, 31);
    prev_end_carry = last_e;
    __syncwarp();
  }
}

}  // namespace nrq
