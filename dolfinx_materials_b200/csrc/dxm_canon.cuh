// Canonical fp64 building blocks.  The whole library is compiled with -fmad=false: the compiler never contracts
// a product into an addition on its own.  Every expression below is a sequence of correctly rounded IEEE-754
// operations in a fixed order -- + - * / sqrt rint and the EXPLICIT fused multiply-add fma_c / fms_c / fnma_c, placed by
// hand at the same positions as in the CPU oracle (oracle/canon.py, oracle/c/dxm_canon.h) -- which is what makes the
// kernels bit-comparable with the oracle while issuing one DFMA where the un-fused form needed DMUL + DADD.
// -DDXM_UNFUSED turns the three helpers back into two roundings (the round-1 arithmetic; A/B builds only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>

namespace dxm {

constexpr double kLog2e = 1.4426950408889634;
constexpr double kLn2Hi = 6.93147180369123816490e-01;
constexpr double kLn2Lo = 1.90821492927058770002e-10;
constexpr double kExpClamp = 700.0;

#define DXM_HD __host__ __device__ __forceinline__

// a*b + c, a*b - c, c - a*b with ONE rounding (DFMA on the device, fma() of <cmath> on the host)
#ifdef DXM_UNFUSED
DXM_HD double fma_c(double a, double b, double c) { return a * b + c; }
DXM_HD double fms_c(double a, double b, double c) { return a * b - c; }
DXM_HD double fnma_c(double a, double b, double c) { return c - a * b; }
#else
DXM_HD double fma_c(double a, double b, double c) { return fma(a, b, c); }
DXM_HD double fms_c(double a, double b, double c) { return fma(a, b, -c); }
DXM_HD double fnma_c(double a, double b, double c) { return fma(-a, b, c); }
#endif

// degree-13 Horner polynomial of exp on the reduced argument (13 fused steps)
DXM_HD double exp_poly(double r) {
  double y = 1.0 / 6227020800.0;
  y = fma_c(y, r, 1.0 / 479001600.0);
  y = fma_c(y, r, 1.0 / 39916800.0);
  y = fma_c(y, r, 1.0 / 3628800.0);
  y = fma_c(y, r, 1.0 / 362880.0);
  y = fma_c(y, r, 1.0 / 40320.0);
  y = fma_c(y, r, 1.0 / 5040.0);
  y = fma_c(y, r, 1.0 / 720.0);
  y = fma_c(y, r, 1.0 / 120.0);
  y = fma_c(y, r, 1.0 / 24.0);
  y = fma_c(y, r, 1.0 / 6.0);
  y = fma_c(y, r, 0.5);
  y = fma_c(y, r, 1.0);
  y = fma_c(y, r, 1.0);
  return y;
}

// exp(x): Cody-Waite reduction, degree-13 Horner (fused steps), exact 2^k scaling.
__device__ __forceinline__ double exp_c(double x) {
  const bool inr = (x >= -kExpClamp) && (x <= kExpClamp);
  const double xs = inr ? x : 0.0;
  const double k = rint(xs * kLog2e);
  const double r = fnma_c(k, kLn2Lo, fnma_c(k, kLn2Hi, xs));
  double y = exp_poly(r);
  // |k| <= 1010 and y in [0.7, 1.5]: 2^k is a normal double and the product is exact
  const int ki = (int)k;
  y = y * __hiloint2double((ki + 1023) << 20, 0);
  if (x < -kExpClamp) y = 0.0;
  if (x > kExpClamp) y = __longlong_as_double(0x7ff0000000000000LL);
  if (x != x) y = x;
  return y;
}

// exp_c for code that is compiled for the host as well (the per-point routines the CPU tests execute against the
// oracle, tests/*_host_check.cu): on the device it IS exp_c; on the host the same operations, with ldexp doing the
// exact 2^k scaling.
DXM_HD double exp_hd(double x) {
#ifdef __CUDA_ARCH__
  return exp_c(x);
#else
  if (x != x) return x;
  if (x < -kExpClamp) return 0.0;
  if (x > kExpClamp) return INFINITY;
  const double k = rint(x * kLog2e);
  const double r = fnma_c(k, kLn2Lo, fnma_c(k, kLn2Hi, x));
  return ldexp(exp_poly(r), (int)k);
#endif
}

// streaming (evict-first) scalar access for code that also compiles for the host
DXM_HD double ld_stream(const double* p) {
#ifdef __CUDA_ARCH__
  return __ldcs(p);
#else
  return *p;
#endif
}
DXM_HD void st_stream(double* p, double v) {
#ifdef __CUDA_ARCH__
  __stcs(p, v);
#else
  *p = v;
#endif
}

// warp vote of the local Newton loops; a host caller (CPU test harness) is a warp of one lane
#ifdef __CUDA_ARCH__
#define DXM_ANY_SYNC(mask, pred) __any_sync(mask, pred)
#else
#define DXM_ANY_SYNC(mask, pred) (pred)
#endif

// Resident layout of a symmetric 6x6 tangent: the 21 entries (j <= i) of the upper triangle, row-major
// (the order the small-strain kernel forms them in).  sym6_packed(c) maps a full row-major index c = j*6+i to it.
constexpr int kSym6Rows = 21;
__host__ __device__ __forceinline__ constexpr int sym6_packed(int c) {
  const int a = c / 6, b = c % 6;
  const int j = a < b ? a : b, i = a < b ? b : a;
  return j * 6 - (j * (j - 1)) / 2 + (i - j);
}

// streaming (evict-first) vector loads / stores of PPT consecutive points
template <int PPT>
__device__ __forceinline__ void ldv(const double* __restrict__ p, double (&v)[PPT]);
template <>
__device__ __forceinline__ void ldv<1>(const double* __restrict__ p, double (&v)[1]) {
  v[0] = __ldcs(p);
}
template <>
__device__ __forceinline__ void ldv<2>(const double* __restrict__ p, double (&v)[2]) {
  const double2 t = __ldcs(reinterpret_cast<const double2*>(p));
  v[0] = t.x;
  v[1] = t.y;
}
template <int PPT>
__device__ __forceinline__ void stv(double* __restrict__ p, const double (&v)[PPT]);
template <>
__device__ __forceinline__ void stv<1>(double* __restrict__ p, const double (&v)[1]) {
  __stcs(p, v[0]);
}
template <>
__device__ __forceinline__ void stv<2>(double* __restrict__ p, const double (&v)[2]) {
  __stcs(reinterpret_cast<double2*>(p), make_double2(v[0], v[1]));
}

// ---- per-call statistics, reduced and published by the update kernel itself -------------------------------------------
// Every CTA reduces its threads' statistics (warp shuffles -> shared memory) and adds them to one of 32 spread
// accumulation slots.  The LAST CTA of the call's last launch (ticket counter) then folds the 32 slots into one record,
// publishes it -- straight into page-locked mapped host memory, where the host reads it with plain loads, or into a
// device record that an in-stream NCCL all-gather picks up in a multi-GPU run -- and clears slots and ticket for the next
// call.  No memset, no device-to-host copy, no stream synchronisation on the host side of a call.
//
// The record is five self-validating 8-byte words: the top 16 bits of each carry the call's sequence number, so the
// words may land in any order and the reader simply waits until all five carry the tag it expects -- no system-scope
// fence in the kernel's epilogue (measured: the fenced version cost 6 of the 9.4 us of a one-point launch,
// profiles/r02d_small_launches*.csv).
constexpr int kStatSlots = 32;
struct StatSlot {
  unsigned long long n_plastic, n_fail, max_iter, max_resid_bits;
};
struct StatBlock {
  StatSlot slot[kStatSlots];
  unsigned int ticket, pad;
};
constexpr int kStatWords = 5;
struct StatRecord {  // 64 bytes
  unsigned long long w[kStatWords];  // tag<<48 | {n_plastic, n_fail, max_iter<<32 | resid_hi, resid_lo, n_points}
  unsigned long long pad[8 - kStatWords];
};
// Multi-GPU, one node: every rank maps every other rank's exchange buffer (cudaIpc, NVLink / NVSwitch peer memory).  The
// CTA that publishes a call's record stores it into slot [xslot][my rank] of EVERY rank's buffer, waits until the records
// of all ranks have arrived in its own buffer, folds them (SUM counts and points, MAX iterations / residual) and writes
// the global record to the host -- the collective is part of the update kernel's epilogue: no NCCL call, no extra launch.
constexpr int kXchgMaxRanks = 16;
constexpr int kXchgSlots = 256;  // handles with global statistics per process
struct StatXchg {
  // peer[r]: rank r's buffer [kXchgSlots][2][nranks], peer-mapped; peer[rank] is local.  Two record sets per slot, used
  // alternately by call parity: a fast rank's record of call k+1 can then never overwrite its record of call k while a
  // slower rank is still reading it (k+2 needs the slower rank's k+1, which it sends only after finishing k).
  StatRecord* peer[kXchgMaxRanks];
  int nranks, rank;
};
struct StatSink {
  StatBlock* blk;
  StatRecord* out;  // mapped host memory, or the device record a multi-GPU run all-gathers with NCCL before publishing
  unsigned long long seq, n_points;
  int finalize;  // 1: this launch is the last one of the call; 2: ... and also its only one (clean slots)
  const StatXchg* xchg;  // non-null: exchange the record over peer memory (then `out` is the mapped host record)
  int xslot;
};
constexpr unsigned long long kStatTimeoutBit = 1ull << 47;  // set in n_fail when a peer's record never arrived

struct PointStats {
  unsigned n_plastic = 0, n_fail = 0, max_iter = 0;
  double max_resid = 0.0;
};

constexpr unsigned long long kStatMask48 = (1ull << 48) - 1ull;
__host__ __device__ __forceinline__ void stat_encode(unsigned long long (&w)[kStatWords], unsigned long long np,
                                                     unsigned long long nf, unsigned long long mi, unsigned long long rb,
                                                     unsigned long long npts, unsigned long long seq) {
  const unsigned long long tag = (seq & 0xffffull) << 48;
  w[0] = tag | (np & kStatMask48);
  w[1] = tag | (nf & kStatMask48);
  w[2] = tag | ((mi & 0xffffull) << 32) | (rb >> 32);
  w[3] = tag | (rb & 0xffffffffull);
  w[4] = tag | (npts & kStatMask48);
}
__host__ __device__ __forceinline__ void stat_decode(const unsigned long long (&w)[kStatWords], unsigned long long& np,
                                                     unsigned long long& nf, unsigned long long& mi, unsigned long long& rb,
                                                     unsigned long long& npts) {
  np = w[0] & kStatMask48;
  nf = w[1] & kStatMask48;
  mi = (w[2] >> 32) & 0xffffull;
  rb = ((w[2] & 0xffffffffull) << 32) | (w[3] & 0xffffffffull);
  npts = w[4] & kStatMask48;
}

__device__ __forceinline__ void publish_record(StatRecord* r, unsigned long long np, unsigned long long nf,
                                               unsigned long long mi, unsigned long long rb, unsigned long long npts,
                                               unsigned long long seq) {
  unsigned long long w[kStatWords];
  stat_encode(w, np, nf, mi, rb, npts, seq);
#pragma unroll
  for (int i = 0; i < kStatWords; ++i) *reinterpret_cast<volatile unsigned long long*>(&r->w[i]) = w[i];
}

// The publishing warp (all 32 lanes, each holding the call's totals): local record, or exchange + fold over peer memory.
__device__ __forceinline__ void finish_record(const StatSink& sink, unsigned long long a, unsigned long long b,
                                              unsigned long long c, unsigned long long d, const int l) {
  if (!sink.xchg) {
    if (l == 0) publish_record(sink.out, a, b, c, d, sink.n_points, sink.seq);
    return;
  }
  const StatXchg& x = *sink.xchg;
  const int nr = x.nranks;
  const size_t set = ((size_t)sink.xslot * 2 + (size_t)(sink.seq & 1ull)) * nr;
  // lane r stores this rank's record into rank r's buffer (self-validating words: no fence, any arrival order) ...
  if (l < nr) publish_record(x.peer[l] + set + x.rank, a, b, c, d, sink.n_points, sink.seq);
  // ... and waits for rank r's record in this rank's own buffer (bounded: ~1 s, then the call is flagged)
  unsigned long long ra = 0, rb = 0, rc = 0, rd = 0, rn = 0;
  bool timed_out = false;
  if (l < nr) {
    const volatile unsigned long long* rec = (x.peer[x.rank] + set + l)->w;
    const unsigned long long tag = sink.seq & 0xffffull;
    unsigned long long w[kStatWords];
    bool ready = false;
    for (unsigned spin = 0; spin < 4000000u && !ready; ++spin) {
      ready = true;
#pragma unroll
      for (int i = 0; i < kStatWords; ++i) {
        w[i] = rec[i];
        ready = ready && (w[i] >> 48) == tag;
      }
    }
    if (ready)
      stat_decode(w, ra, rb, rc, rd, rn);
    else
      timed_out = true;
  }
  const bool any_timeout = __any_sync(0xffffffffu, timed_out);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ra += __shfl_xor_sync(0xffffffffu, ra, o);
    rb += __shfl_xor_sync(0xffffffffu, rb, o);
    rn += __shfl_xor_sync(0xffffffffu, rn, o);
    const unsigned long long c2 = __shfl_xor_sync(0xffffffffu, rc, o);
    rc = c2 > rc ? c2 : rc;
    const unsigned long long d2 = __shfl_xor_sync(0xffffffffu, rd, o);
    rd = d2 > rd ? d2 : rd;
  }
  if (l == 0) publish_record(sink.out, ra, any_timeout ? (rb | kStatTimeoutBit) : rb, rc, rd, rn, sink.seq);
}

__device__ __forceinline__ void block_reduce_stats(const PointStats& s, const StatSink& sink) {
  unsigned np = __reduce_add_sync(0xffffffffu, s.n_plastic);
  unsigned nf = __reduce_add_sync(0xffffffffu, s.n_fail);
  unsigned mi = __reduce_max_sync(0xffffffffu, s.max_iter);
  // residuals are >= 0 (or NaN-free by construction): order of doubles == order of their bits
  unsigned long long rb = (unsigned long long)__double_as_longlong(s.max_resid);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long other = __shfl_xor_sync(0xffffffffu, rb, o);
    rb = other > rb ? other : rb;
  }
  __shared__ unsigned long long sh[4][32];
  __shared__ int s_last;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) {
    sh[0][w] = np;
    sh[1][w] = nf;
    sh[2][w] = mi;
    sh[3][w] = rb;
  }
  __syncthreads();
  if (w == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    unsigned long long a = l < nw ? sh[0][l] : 0ull, b = l < nw ? sh[1][l] : 0ull,
                       c = l < nw ? sh[2][l] : 0ull, d = l < nw ? sh[3][l] : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
      unsigned long long c2 = __shfl_xor_sync(0xffffffffu, c, o);
      c = c2 > c ? c2 : c;
      unsigned long long d2 = __shfl_xor_sync(0xffffffffu, d, o);
      d = d2 > d ? d2 : d;
    }
    if (sink.finalize == 2 && gridDim.x == 1) {
      // the call is this one CTA: its totals are the record
      finish_record(sink, a, b, c, d, l);
      if (l == 0) s_last = 0;
    } else if (l == 0) {
      int last = 0;
      {
        StatSlot* s2 = sink.blk->slot + (blockIdx.x % kStatSlots);
        if (a) atomicAdd(&s2->n_plastic, a);
        if (b) atomicAdd(&s2->n_fail, b);
        if (c) atomicMax(&s2->max_iter, c);
        if (d) atomicMax(&s2->max_resid_bits, d);
        if (sink.finalize) {
          __threadfence();  // this CTA's slot updates before its ticket
          last = atomicAdd(&sink.blk->ticket, 1u) == gridDim.x - 1;
        }
      }
      s_last = last;
    }
  }
  __syncthreads();
  if (s_last && w == 0) {
    __threadfence();  // every other CTA's slot updates precede its ticket, which precedes ours
    StatSlot* sl = sink.blk->slot + l;
    // read through L2 (the slots were updated by atomics from every SM), then clear for the next call
    unsigned long long a = __ldcg(&sl->n_plastic), b = __ldcg(&sl->n_fail), c = __ldcg(&sl->max_iter),
                       d = __ldcg(&sl->max_resid_bits);
    sl->n_plastic = 0ull;
    sl->n_fail = 0ull;
    sl->max_iter = 0ull;
    sl->max_resid_bits = 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
      unsigned long long c2 = __shfl_xor_sync(0xffffffffu, c, o);
      c = c2 > c ? c2 : c;
      unsigned long long d2 = __shfl_xor_sync(0xffffffffu, d, o);
      d = d2 > d ? d2 : d;
    }
    if (l == 0) sink.blk->ticket = 0u;
    finish_record(sink, a, b, c, d, l);
  }
}

}  // namespace dxm
