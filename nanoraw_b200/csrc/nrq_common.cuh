// Shared device helpers for libnrq (sm_100a).  All float64 arithmetic that feeds a
// boundary decision goes through the explicit round-to-nearest intrinsics so that no
// FMA contraction can change a bit relative to numpy.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/nrq.h"

#define NRQ_GAP 45  // '-'

namespace nrq {
constexpr unsigned kFull = 0xffffffffu;  // all lanes of a warp
}

struct NormCoef {
  double shift, scale, lo, hi;
  int clip;
};

__device__ __forceinline__ NormCoef load_coef(const double* __restrict__ sv, int r) {
  NormCoef c;
  c.shift = sv[4 * r + 0];
  c.scale = sv[4 * r + 1];
  c.lo = sv[4 * r + 2];
  c.hi = sv[4 * r + 3];
  c.clip = !(isnan(c.lo) || isnan(c.hi));
  return c;
}

// nanoraw_helper.py:248 then :257-262 for one sample
__device__ __forceinline__ double norm_value(int raw, const NormCoef& c) {
  double x = __ddiv_rn(__dsub_rn((double)raw, c.shift), c.scale);
  if (c.clip) x = (x > c.hi) ? c.hi : ((x < c.lo) ? c.lo : x);
  return x;
}

__device__ __forceinline__ unsigned lane_id() {
  unsigned l;
  asm("mov.u32 %0, %%laneid;" : "=r"(l));  // one S2R where the compiler re-derives it, instead of tid & 31
  return l;
}

// per-read geometry shared by every kernel
struct ReadGeom {
  const int16_t* sig;   // first sample of the read inside raw (all_raw_signal[read_start:])
  const int32_t* starts;
  int n_starts;         // Br + 1
  int n_obs;            // starts[-1] = read_obs_len
  int n_norm;           // len(norm_signal) = min(n_obs, samples available)
  const double* xsig;   // pre-normalised float64 signal (nrq_batch.norm_signal) or nullptr
};

__device__ __forceinline__ ReadGeom read_geom(const nrq_batch& b, int r) {
  ReadGeom g;
  int64_t r0 = b.raw_off[r], r1 = b.raw_off[r + 1];
  int64_t s0 = b.starts_off[r], s1 = b.starts_off[r + 1];
  int rs = b.read_start[r];
  g.sig = b.raw + r0 + rs;
  g.starts = b.starts + s0;
  g.n_starts = (int)(s1 - s0);
  g.n_obs = g.n_starts > 0 ? g.starts[g.n_starts - 1] : 0;
  int64_t avail = (r1 - r0) - rs;
  if (avail < 0) avail = 0;
  g.n_norm = (int)(g.n_obs < avail ? g.n_obs : avail);
  g.xsig = nullptr;
  if (b.norm_signal) {
    g.xsig = b.norm_signal + b.norm_off[r];
    g.n_norm = (int)(b.norm_off[r + 1] - b.norm_off[r]);
  }
  return g;
}
