// libdxm_cuda.so -- C ABI (include/dxm.h) over the sm_100a constitutive-update kernels.
// Host side only: handle / buffer management, the chunked host<->device pipeline, statistics.
// (FE-side entry points: dxm_fe_api.cu; measurement support: dxm_peaks.cu.)
#include "dxm_internal.cuh"
#include "dxm_fefp.cuh"
#include "dxm_host_mirror.hpp"
#include "dxm_hosford.cuh"
#include "dxm_layout.cuh"
#include "dxm_small_strain.cuh"

using namespace dxm;

namespace dxm_detail {
thread_local std::string g_err;
std::atomic<long long> g_launches{0};
}  // namespace dxm_detail

namespace {
constexpr int64_t kChunk = 1 << 19;  // points per pipeline chunk on the host path
const char* kPropNames[kNProp] = {"E", "nu", "sig0", "H", "sigu", "b"};
constexpr double kHosLowPlastic = 0.25;  // Hosford: 168-register build below this plastic fraction (previous call)
// host arrays of at most this many points skip the copy engines: the transposition kernels read / write a mapped
// page-locked staging block directly (no cudaMemcpyAsync, no event hand-offs between three streams: one stream, one wait)
constexpr int64_t kSmallHostPoints = 2048;
constexpr int64_t kAutoTimingPoints = 1 << 18;  // kernel_ms events by default only where two event records are noise
constexpr bool kHostMirrorDefault = false;  // A/B on the B200 box: profiles/ (DXM_HOST_MIRROR overrides)
}  // namespace

namespace {

const Field* find_field(const dxm_handle* h, const char* name) {
  for (const Field& f : h->fields)
    if (std::strcmp(f.name, name) == 0) return &f;
  return nullptr;
}

double* field_ptr(dxm_handle* h, int gen, const Field* f, bool for_read) {
  int g = gen == 0 ? h->i0 : 1 - h->i0;
  if (gen == 1 && for_read && !h->s1_valid) g = h->i0;
  return h->gen[g] + (int64_t)f->row * h->ld;
}

// Grid sizing.  Measured on B200 (profiles/r01_layout_experiment.txt, r01_grid_sweep.json): for these
// streaming kernels a persistent grid (resident CTAs x tile-stride loop) caps at ~5.65 TB/s because all
// CTAs walk the 74 SoA streams in lock-step; handing the hardware block scheduler many small CTAs
// (a few 256-point tiles each) desynchronises them and reaches ~6.5 TB/s.  One tile per CTA is slower
// again (per-CTA statistics epilogue).  Default: kTilesPerCta tiles per CTA.
// DXM_TPB=<t> overrides the tiles per CTA, DXM_GRID=<k> forces k CTAs per SM (persistent style).
constexpr int kTilesPerCta = 4;
int grid_for(const void* kernel, int block, size_t smem, int num_sms, int64_t ntile) {
  (void)kernel;
  (void)block;
  (void)smem;
  static const int mult = [] {
    const char* e = std::getenv("DXM_GRID");
    return e ? std::atoi(e) : 0;
  }();
  static const int tpb = [] {
    const char* e = std::getenv("DXM_TPB");
    const int v = e ? std::atoi(e) : kTilesPerCta;
    return v > 0 ? v : kTilesPerCta;
  }();
  int64_t g = mult > 0 ? (int64_t)mult * num_sms : (ntile + tpb - 1) / tpb;
  if (g > ntile) g = ntile;
  if (g > 0x7fffffff) g = 0x7fffffff;
  if (g < 1) g = 1;
  return (int)g;
}

int launch_aos_to_soa(dxm_handle* h, cudaStream_t st, const double* src, int64_t rs, int c0,
                      double* dst, int64_t d0, int64_t count, int D) {
  if (count <= 0) return 0;
  const size_t smem = (size_t)kTile * (D | 1) * sizeof(double);
  if (smem > 48 * 1024) CK(cudaFuncSetAttribute(aos_to_soa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t ntile = (count + kTile - 1) / kTile;
  const int grid = (int)std::min<int64_t>(ntile, (int64_t)h->num_sms * 8);
  aos_to_soa_kernel<<<grid, 256, smem, st>>>(src, rs, c0, dst, h->ld, d0, count, D);
  LAUNCH_CHECK();
  return 0;
}

int launch_soa_to_aos(dxm_handle* h, cudaStream_t st, const double* src, int64_t s0, double* dst,
                      int64_t rs, int c0, int64_t count, int D, int sym6 = 0) {
  if (count <= 0) return 0;
  const size_t smem = (size_t)kTile * (D | 1) * sizeof(double);
  if (smem > 48 * 1024) CK(cudaFuncSetAttribute(soa_to_aos_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t ntile = (count + kTile - 1) / kTile;
  const int grid = (int)std::min<int64_t>(ntile, (int64_t)h->num_sms * 4);
  soa_to_aos_kernel<<<grid, 256, smem, st>>>(src, h->ld, s0, dst, rs, c0, count, D, sym6);
  LAUNCH_CHECK();
  return 0;
}

int ensure_staging(dxm_handle* h) {
  if (h->chunk) return 0;
  // ~8 pipeline stages for mid-size batches (H2D | update | D2H overlap), 32 Ki .. 512 Ki points per chunk
  int64_t c = (h->n + 7) / 8;
  c = std::max<int64_t>(c, 32768);
  c = std::min<int64_t>(c, kChunk);
  c = (c + 255) & ~int64_t(255);
  h->chunk = std::min<int64_t>(c, (h->n + 1) & ~int64_t(1));
  if (h->chunk < 2) h->chunk = 2;
  const int64_t nout = h->nflux + h->nisv + h->nct;
  for (int s = 0; s < 2; ++s) {
    CK(cudaMalloc(&h->d_in[s], sizeof(double) * h->chunk * h->ngrad));
    CK(cudaMalloc(&h->d_out[s], sizeof(double) * h->chunk * nout));
  }
  return 0;
}

// DXM_HOST_MIRROR=0|1: send the symmetric tangent over PCIe packed (21 of 36 entries) and mirror it into the
// caller's (n, 36) array with host threads (dxm_host_mirror.hpp) instead of expanding it on the device first.
bool host_mirror_enabled() {
  const char* e = std::getenv("DXM_HOST_MIRROR");
  return e ? std::atoi(e) != 0 : kHostMirrorDefault;
}

int ensure_mirror_ring(dxm_handle* h) {
  if (h->h_ctp[0]) return 0;
  for (int s = 0; s < dxm_handle::kRing; ++s) {
    CK(cudaMallocHost(&h->h_ctp[s], sizeof(double) * h->chunk * h->nct_store));
    CK(cudaEventCreateWithFlags(&h->ev_ct[s], cudaEventDisableTiming));
  }
  return 0;
}

template <int HARD, bool PERPOINT, int PPT, bool DIAG, int MINB = 2, bool COMPACT = false>
int launch_small_strain(dxm_handle* h, const SmallStrainArgs& a) {
  const void* k = (const void*)dxm_small_strain_kernel<HARD, PERPOINT, PPT, DIAG, MINB, COMPACT>;
  // Small batches (cfg1's meshes, per-rank shards of a strongly scaled solve) are latency-bound: spread them over the SMs
  // in 64-point CTAs of one tile each instead of a few 4-tile CTAs (1e4 points: 10 CTAs x 4 sequential tiles = 26.7 us
  // -> 157 CTAs x 1 tile, profiles/r02d_latency.json)
  const bool small = !COMPACT && a.count <= (int64_t)h->num_sms * 1024;
  const int block = small ? 64 : 256;
  const int64_t ntile = (a.count + (int64_t)block * PPT - 1) / ((int64_t)block * PPT);
  const int grid = small ? (int)ntile : grid_for(k, block, 0, h->num_sms, ntile);
  dxm_small_strain_kernel<HARD, PERPOINT, PPT, DIAG, MINB, COMPACT><<<grid, block, 0, h->stream>>>(a);
  LAUNCH_CHECK();
  return 0;
}

template <int HARD, bool PERPOINT>
int dispatch_small_strain2(dxm_handle* h, const SmallStrainArgs& a) {
  // tuning knob (DXM_MINB: resident blocks per SM the register allocation targets) for the
  // production kernel only; see profiles/ for the sweep that picked the defaults
  if (HARD == HARD_GENERAL && !PERPOINT && !h->diag && (h->minb == 1 || h->minb == 4)) {
    if (h->ppt == 2) {
      if (h->minb == 1) return launch_small_strain<HARD_GENERAL, false, 2, false, 1>(h, a);
    } else {
      if (h->minb == 1) return launch_small_strain<HARD_GENERAL, false, 1, false, 1>(h, a);
      if (h->minb == 4) return launch_small_strain<HARD_GENERAL, false, 1, false, 4>(h, a);
    }
  }
  if (HARD == HARD_GENERAL && !PERPOINT && h->compact == 1 && h->ppt == 1) {
    // block-level compaction of the plastic points' Newton solves (DXM_COMPACT=1; A/B in profiles/)
    return h->diag ? launch_small_strain<HARD_GENERAL, false, 1, true, 2, true>(h, a)
                   : launch_small_strain<HARD_GENERAL, false, 1, false, 2, true>(h, a);
  }
  if (h->ppt == 2) {
    return h->diag ? launch_small_strain<HARD, PERPOINT, 2, true>(h, a)
                   : launch_small_strain<HARD, PERPOINT, 2, false>(h, a);
  }
  // Resident CTAs per SM the register allocation targets.  With the packed tangent the uniform Voce / elastic
  // kernels fit 3 CTAs (80 registers, <= 16 B of spills): 13.7 vs 12.1 G points/s at 2 CTAs for every batch size
  // from 1e6 to 1e8 (profiles/r01e_sweep_minb_vs_n.json); the linear-hardening and per-point-property variants
  // spill 40-64 B at 80 registers and stay at 2 (cfg1 11.5 vs 10.6, cfg4 12.6 vs 12.0 G points/s).
  const int minb = h->minb > 0 ? h->minb : ((!PERPOINT && HARD != HARD_LINEAR) ? 3 : 2);
  if (minb == 2)
    return h->diag ? launch_small_strain<HARD, PERPOINT, 1, true, 2>(h, a)
                   : launch_small_strain<HARD, PERPOINT, 1, false, 2>(h, a);
  return h->diag ? launch_small_strain<HARD, PERPOINT, 1, true, 3>(h, a)
                 : launch_small_strain<HARD, PERPOINT, 1, false, 3>(h, a);
}

double prop(const dxm_handle* h, int i) {
  if (i == 4 && !h->set[4]) return h->uni[2];  // sigu defaults to sig0 (no saturation term)
  return h->uni[i];
}

// where the launch's statistics go; `finalize`: this is the call's last launch, its last CTA folds and publishes
// (2: it is also the call's only launch)
StatSink stat_sink(const dxm_handle* h, int finalize) {
  StatSink k{};
  k.blk = h->d_statblk;
  const bool global = h->global_stats && dxm_comm::size() > 1;
  const bool p2p = global && h->xslot >= 0 && h->xgen == dxm_comm::generation() && dxm_comm::xchg();
  // multi-GPU: over peer memory inside the kernel's epilogue, else all-gathered with NCCL and published by a second kernel
  k.out = (global && !p2p) ? h->d_rec : h->h_rec;
  k.xchg = p2p ? dxm_comm::xchg() : nullptr;
  k.xslot = h->xslot;
  k.seq = h->seq;
  k.n_points = (unsigned long long)h->last.n_points;
  k.finalize = finalize;
  return k;
}

// launches the constitutive kernel for points [start, start+count) on h->stream
int launch_update(dxm_handle* h, int64_t start, int64_t count, double dt, int finalize) {
  (void)dt;  // rate-independent behaviours (the reference's FE path never passes dt: quadrature_map.py:321)
  if (count <= 0) return 0;
  double* s0 = h->gen[h->i0];
  double* s1 = h->gen[1 - h->i0];
  const int64_t ld = h->ld;
  if (h->behaviour == DXM_FEFP_VOCE) {
    FeFpArgs a{};
    a.F = s1;
    a.P = s1 + 9 * ld;
    a.p = s1 + 18 * ld;
    a.be = s1 + 19 * ld;
    a.ct = h->ct;
    a.F_old = s0;
    a.p_old = s0 + 18 * ld;
    a.be_old = s0 + 19 * ld;
    a.ld = ld;
    a.start = start;
    a.count = count;
    a.perpoint = h->perpoint;
    const double E = prop(h, 0), nu = prop(h, 1);
    a.E = E;
    a.mu = E / 2 / (1 + nu);
    a.kappa = E / (3 * (1 - 2 * nu));
    a.sig0 = prop(h, 2);
    a.H = prop(h, 3);
    const double d = prop(h, 4) - prop(h, 2);
    a.dsu = std::isfinite(d) ? d : 0.0;
    a.b = prop(h, 5);
    for (int i = 0; i < kNProp; ++i) a.pp[i] = h->pp ? h->pp + (int64_t)i * ld : nullptr;
    a.stats = stat_sink(h, finalize);
    a.vote = h->vote;
    a.d_flag = h->d_flag;
    a.d_iter = h->d_iter;
    a.d_resid = h->d_resid;
    a.d_fail = h->d_fail;
    // block-level compaction of the plastic points' local solves pays only when few points are plastic
    // (profiles/r01d_compaction_ab.json: +5 % at 8 % plastic, -3 % at 60-80 %): auto mode keys on the
    // plastic fraction of the previous call
    bool compact = h->compact == 1;
    if (h->compact < 0 && h->prev_points > 0) {
      const double frac = (double)h->prev_plastic / (double)h->prev_points;
      compact = frac > 0.005 && frac < 0.2;
    }
    return launch_fefp(a, h->diag, compact && !h->perpoint, h->num_sms, h->stream, &g_launches, &g_err);
  }
  SmallStrainArgs a{};
  a.eps = s1;
  a.sig = s1 + 6 * ld;
  a.p = s1 + 12 * ld;
  a.epsp = s1 + 13 * ld;
  a.ct = h->ct;
  a.eps_old = s0;
  a.sig_old = s0 + 6 * ld;
  a.p_old = s0 + 12 * ld;
  a.epsp_old = s0 + 13 * ld;
  a.ld = ld;
  a.start = start;
  a.count = count;
  const double E = prop(h, 0), nu = prop(h, 1);
  a.lam = E * nu / (1 + nu) / (1 - 2 * nu);
  a.mu = E / 2 / (1 + nu);
  a.sig0 = prop(h, 2);
  a.H = prop(h, 3);
  const double d = prop(h, 4) - prop(h, 2);
  a.dsu = std::isfinite(d) ? d : 0.0;
  a.b = prop(h, 5);
  if (h->pp) {
    a.pE = h->pp;
    a.pnu = h->pp + ld;
    a.psig0 = h->pp + 2 * ld;
    a.pH = h->pp + 3 * ld;
    a.psigu = h->pp + 4 * ld;
    a.pb = h->pp + 5 * ld;
  }
  a.table = h->table;
  a.ntab = h->ntab;
  a.stats = stat_sink(h, finalize);
  a.vote = h->vote;
  a.d_flag = h->d_flag;
  a.d_iter = h->d_iter;
  a.d_resid = h->d_resid;
  a.d_fail = h->d_fail;
  if (h->behaviour == DXM_HOSFORD_LINEAR) {
    a.hos_a = h->hos_a;
    a.hos_bound = hosford_bound(h->hos_a);
    // The fused kernel (every thread runs the whole routine on its own point) is the default at every plastic fraction
    // since the eigen-decomposition became non-iterative: 6.5-6.9 G points/s at 3 % plastic, 4.9 G at 80 %
    // (profiles/r02t_hosford_ab.json).  The tiled kernel (stream a tile, pack the candidates into full warps) won below
    // ~30 % plastic while the local solve cost twice as much; now it only ties at 3 % (6.7 G) and loses 20-30 % above
    // 10 %, so it stays as an opt-in: DXM_HOS_SPLIT=1 (0 forces fused).  Register target: 3 resident CTAs per SM
    // (168 registers) while few points are plastic (previous call < 25 %), 4 (128) otherwise -- worth 2-5 % either way.
    // (A warp-private queue -- every heavy pass a full warp, no block barrier -- was also measured in round 2: same
    // times as the tiled kernel to 2-5 %.  Not kept.)
    const char* e = std::getenv("DXM_HOS_SPLIT");
    const bool tiled = e ? std::atoi(e) != 0 : false;
    const char* mb = std::getenv("DXM_HOS_MINB");  // A/B knob, read per call like DXM_HOS_SPLIT
    int minb = mb ? std::atoi(mb) : 0;
    if (!minb && !tiled && h->prev_points > 0 && (double)h->prev_plastic < kHosLowPlastic * (double)h->prev_points)
      minb = 3;
    // sigu is only ever set for a hardening law with a saturation term (VoceHardening); otherwise it follows sig0
    HosLaunch cfg{h->num_sms, h->stream, h->set[4], tiled, kTilesPerCta, minb};
    int launches = 0;
    const int rc = launch_hosford(a, cfg, &launches);
    g_launches.fetch_add(launches);
    return rc;
  }
  if (h->behaviour == DXM_J2_TABLE) {
    // 2 CTAs/SM: the segment walk keeps a few more values live than the closed forms
    if (h->perpoint)
      return h->diag ? launch_small_strain<HARD_TABLE, true, 1, true, 2>(h, a)
                     : launch_small_strain<HARD_TABLE, true, 1, false, 2>(h, a);
    return h->diag ? launch_small_strain<HARD_TABLE, false, 1, true, 2>(h, a)
                   : launch_small_strain<HARD_TABLE, false, 1, false, 2>(h, a);
  }
  if (h->perpoint) {
    switch (h->behaviour) {
      case DXM_ELASTIC: return dispatch_small_strain2<HARD_NONE, true>(h, a);
      case DXM_J2_LINEAR: return dispatch_small_strain2<HARD_LINEAR, true>(h, a);
      default: return dispatch_small_strain2<HARD_GENERAL, true>(h, a);
    }
  }
  switch (h->behaviour) {
    case DXM_ELASTIC: return dispatch_small_strain2<HARD_NONE, false>(h, a);
    case DXM_J2_LINEAR: return dispatch_small_strain2<HARD_LINEAR, false>(h, a);
    default: return dispatch_small_strain2<HARD_GENERAL, false>(h, a);
  }
}

int next_event_pair(dxm_handle* h, cudaEvent_t** pair) {
  if ((size_t)h->n_ev_used + 2 > h->ev_k.size()) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    h->ev_k.push_back(a);
    h->ev_k.push_back(b);
  }
  *pair = &h->ev_k[h->n_ev_used];
  h->n_ev_used += 2;
  return 0;
}

// multi-GPU: folds the all-gathered per-rank records (SUM of counts and points, MAX of iterations and residual) into
// the mapped host record.  One warp; launched on the handle's stream right after the collective.
__global__ void stats_publish_kernel(const StatRecord* __restrict__ gathered, int nranks, StatRecord* host_rec,
                                     unsigned long long seq) {
  const int l = threadIdx.x;
  unsigned long long a = 0, b = 0, c = 0, d = 0, n = 0;
  for (int r = l; r < nranks; r += 32) {
    unsigned long long w[kStatWords], ra, rb_, rc, rd, rn;
    for (int i = 0; i < kStatWords; ++i) w[i] = gathered[r].w[i];
    stat_decode(w, ra, rb_, rc, rd, rn);
    a += ra;
    b += rb_;
    c = rc > c ? rc : c;
    d = rd > d ? rd : d;
    n += rn;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
    n += __shfl_xor_sync(0xffffffffu, n, o);
    unsigned long long c2 = __shfl_xor_sync(0xffffffffu, c, o);
    c = c2 > c ? c2 : c;
    unsigned long long d2 = __shfl_xor_sync(0xffffffffu, d, o);
    d = d2 > d ? d2 : d;
  }
  if (l == 0) publish_record(host_rec, a, b, c, d, n, seq);
}

// one kernel launch of the call; `finalize` marks the last one (it publishes the statistics, and the in-stream
// all-gather + publish of a multi-GPU run follow it)
int timed_update(dxm_handle* h, int64_t start, int64_t count, double dt, int finalize) {
  const bool timing = h->timing > 0 || (h->timing < 0 && h->last.n_points >= kAutoTimingPoints);
  cudaEvent_t* ev = nullptr;
  if (timing) {
    if (next_event_pair(h, &ev)) return -1;
    CK(cudaEventRecord(ev[0], h->stream));
  }
  if (launch_update(h, start, count, dt, finalize)) return -1;
  if (timing) CK(cudaEventRecord(ev[1], h->stream));
  if (finalize) {
    h->finalize_launched = true;
    if (h->global_stats && dxm_comm::size() > 1 &&
        !(h->xslot >= 0 && h->xgen == dxm_comm::generation() && dxm_comm::xchg())) {
      if (dxm_comm::all_gather(h->d_rec, h->d_gather, sizeof(StatRecord), h->stream)) return -1;
      stats_publish_kernel<<<1, 32, 0, h->stream>>>(h->d_gather, dxm_comm::size(), h->h_rec, h->seq);
      LAUNCH_CHECK();
    }
  }
  return 0;
}

// Waits for the call's published record: a spin on one word of page-locked memory the kernel writes last -- no stream
// synchronisation, no copy.  The stream is queried now and then so that a failed launch cannot hang the caller.
int finish_stats(dxm_handle* h) {
  if (!h->stats_pending) return 0;
  h->stats_pending = false;
  if (!h->finalize_launched) {  // the call broke off before its last launch: nothing will be published
    cudaStreamSynchronize(h->stream);
    cudaMemsetAsync(h->d_statblk, 0, sizeof(StatBlock), h->stream);
    return fail("dxm: the previous integrate call did not complete; its statistics are lost");
  }
  const unsigned long long tag = h->seq & 0xffffull;
  volatile unsigned long long* rec = h->h_rec->w;
  unsigned long long w[kStatWords];
  for (unsigned spin = 0;; ++spin) {
    bool ready = true;
    for (int i = 0; i < kStatWords; ++i) {
      w[i] = rec[i];
      ready = ready && (w[i] >> 48) == tag;
    }
    if (ready) break;
    if ((spin & 0x3ff) == 0x3ff) {
      const cudaError_t q = cudaStreamQuery(h->stream);
      if (q == cudaSuccess) {
        ready = true;
        for (int i = 0; i < kStatWords; ++i) {
          w[i] = rec[i];
          ready = ready && (w[i] >> 48) == tag;
        }
        if (ready) break;
        return fail("dxm: the update kernel finished without publishing its statistics");
      }
      if (q != cudaErrorNotReady) CK(q);
    }
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }
  dxm_stats s{};
  unsigned long long np, nf, mi, rb, npts;
  stat_decode(w, np, nf, mi, rb, npts);
  s.n_points = (int64_t)npts;
  s.n_plastic = (int64_t)np;
  if (nf & kStatTimeoutBit)
    return fail("dxm: a peer's statistics record never arrived (did every rank call integrate on this material?)");
  s.n_fail = (int64_t)nf;
  s.max_iter = (int64_t)mi;
  std::memcpy(&s.max_residual, &rb, sizeof(double));
  double ms = 0;
  if (h->n_ev_used > 0) {
    // the record is published by the kernel's last CTA: the end event completes within a microsecond -- spin on it
    // rather than sleep in cudaEventSynchronize
    for (;;) {
      const cudaError_t q = cudaEventQuery(h->ev_k[h->n_ev_used - 1]);
      if (q == cudaSuccess) break;
      if (q != cudaErrorNotReady) CK(q);
    }
    for (int i = 0; i + 1 < h->n_ev_used; i += 2) {
      float t = 0;
      CK(cudaEventElapsedTime(&t, h->ev_k[i], h->ev_k[i + 1]));
      ms += t;
    }
  }
  s.kernel_ms = ms;
  h->last = s;
  // the launch heuristics (FeFp compaction, Hosford fused / tiled) key on the plastic fraction of the previous call
  h->prev_plastic = s.n_plastic;
  h->prev_points = s.n_points;
  return 0;
}


}  // namespace

// ---- DLPack (ABI of dlpack.h v0.8; redeclared here, no external headers) -----------------------
extern "C" {
typedef struct {
  int32_t device_type;
  int32_t device_id;
} DxmDLDevice;
typedef struct {
  uint8_t code;
  uint8_t bits;
  uint16_t lanes;
} DxmDLDataType;
typedef struct {
  void* data;
  DxmDLDevice device;
  int32_t ndim;
  DxmDLDataType dtype;
  int64_t* shape;
  int64_t* strides;
  uint64_t byte_offset;
} DxmDLTensor;
typedef struct DxmDLManagedTensor {
  DxmDLTensor dl_tensor;
  void* manager_ctx;
  void (*deleter)(struct DxmDLManagedTensor* self);
} DxmDLManagedTensor;
}

namespace {
struct DlCtx {
  dxm_handle* h;
  int64_t shape[2];
  int64_t strides[2];
};

void release_handle(dxm_handle* h);

void dxm_dl_deleter(DxmDLManagedTensor* self) {
  DlCtx* ctx = static_cast<DlCtx*>(self->manager_ctx);
  release_handle(ctx->h);
  delete ctx;
  delete self;
}

void free_handle(dxm_handle* h) {
  cudaSetDevice(h->device);
  if (h->own_stream) cudaStreamSynchronize(h->own_stream);
  for (int g = 0; g < 2; ++g) cudaFree(h->gen[g]);
  cudaFree(h->ct);
  cudaFree(h->pp);
  cudaFree(h->table);
  cudaFree(h->d_statblk);
  cudaFree(h->d_rec);
  cudaFree(h->d_gather);
  cudaFreeHost(h->h_rec);
  if (h->h_small) cudaFreeHost(h->h_small);
  for (int s = 0; s < 2; ++s) {
    cudaFree(h->d_in[s]);
    cudaFree(h->d_out[s]);
    if (h->ev_in[s]) cudaEventDestroy(h->ev_in[s]);
    if (h->ev_in_free[s]) cudaEventDestroy(h->ev_in_free[s]);
    if (h->ev_packed[s]) cudaEventDestroy(h->ev_packed[s]);
    if (h->ev_out_free[s]) cudaEventDestroy(h->ev_out_free[s]);
  }
  for (int s = 0; s < dxm_handle::kRing; ++s) {
    if (h->h_ctp[s]) cudaFreeHost(h->h_ctp[s]);
    if (h->ev_ct[s]) cudaEventDestroy(h->ev_ct[s]);
  }
  for (cudaEvent_t e : h->ev_k) cudaEventDestroy(e);
  cudaFree(h->d_flag);
  cudaFree(h->d_fail);
  cudaFree(h->d_iter);
  cudaFree(h->d_resid);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  if (h->s_in) cudaStreamDestroy(h->s_in);
  if (h->s_out) cudaStreamDestroy(h->s_out);
  delete h;
}

void release_handle(dxm_handle* h) {
  if (h->refs.fetch_sub(1) == 1) free_handle(h);
}
}  // namespace

// ================================================================================================
extern "C" {

const char* dxm_last_error(void) { return g_err.c_str(); }
const char* dxm_version(void) { return "dxm-b200 0.2 (sm_100a, fp64, hand-placed fma, -fmad=false)"; }
int64_t dxm_launch_count(void) { return g_launches.load(); }

int dxm_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int dxm_create(int behaviour, int device, int64_t n, dxm_handle** out) {
  if (!out) return fail("dxm_create: out is NULL");
  *out = nullptr;
  if (n <= 0) return fail("dxm_create: n must be positive");
  if (behaviour < DXM_ELASTIC || behaviour > DXM_HOSFORD_LINEAR)
    return fail("dxm_create: unknown behaviour " + std::to_string(behaviour));
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev)
    return fail("dxm_create: no CUDA device " + std::to_string(device) +
                " (this library has no CPU fallback)");
  CK(cudaSetDevice(device));
  cudaDeviceProp prop{};
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(std::string("dxm_create: device is sm_") + std::to_string(prop.major) +
                std::to_string(prop.minor) + ", this library is built for sm_100a only");
  dxm_handle* h = new dxm_handle();
  h->behaviour = behaviour;
  h->device = device;
  h->n = n;
  h->ld = (n + 63) & ~int64_t(63);
  h->num_sms = prop.multiProcessorCount;
  const char* env = std::getenv("DXM_PPT");
  // scalar 8-byte accesses measured faster than double2 on B200 (profiles/r01_variant_sweep.json)
  h->ppt = (env && std::atoi(env) == 2) ? 2 : 1;
  env = std::getenv("DXM_MINB");
  h->minb = env ? std::atoi(env) : 0;  // 0 = per-variant default (dispatch_small_strain2)
  env = std::getenv("DXM_VOTE");
  h->vote = env ? std::atoi(env) : 1;
  env = std::getenv("DXM_COMPACT");
  h->compact = env ? std::atoi(env) : -1;
  if (behaviour == DXM_FEFP_VOCE) {
    h->ngrad = 9;
    h->nflux = 9;
    h->nisv = 7;
    h->fields = {{"F", 0, 9}, {"PK1", 9, 9}, {"p", 18, 1}, {"be_bar", 19, 6}};
  } else {
    h->ngrad = 6;
    h->nflux = 6;
    h->nisv = 7;
    h->fields = {{"strain", 0, 6}, {"stress", 6, 6}, {"p", 12, 1}, {"epsp", 13, 6}};
  }
  h->nrows = h->ngrad + h->nflux + h->nisv;
  h->nct = h->nflux * h->ngrad;
  // the small-strain tangent is symmetric: the kernels store its 21 unique entries once (sym6_packed)
  h->nct_store = behaviour == DXM_FEFP_VOCE ? h->nct : kSym6Rows;
  auto bail = [&](int) {
    std::string keep = g_err;
    free_handle(h);
    g_err = keep;
    return -1;
  };
#define CKH(call)                                                                    \
  do {                                                                               \
    cudaError_t e_ = (call);                                                         \
    if (e_ != cudaSuccess) {                                                         \
      g_err = std::string(#call) + ": " + cudaGetErrorString(e_) + " (dxm_create)"; \
      return bail(0);                                                                \
    }                                                                                \
  } while (0)
  const size_t gen_bytes = sizeof(double) * h->nrows * h->ld;
  for (int g = 0; g < 2; ++g) {
    CKH(cudaMalloc(&h->gen[g], gen_bytes));
    CKH(cudaMemset(h->gen[g], 0, gen_bytes));
  }
  CKH(cudaMalloc(&h->ct, sizeof(double) * h->nct_store * h->ld));
  CKH(cudaMemset(h->ct, 0, sizeof(double) * h->nct_store * h->ld));
  CKH(cudaMalloc(&h->d_statblk, sizeof(StatBlock)));
  CKH(cudaMemset(h->d_statblk, 0, sizeof(StatBlock)));
  CKH(cudaMalloc(&h->d_rec, sizeof(StatRecord)));
  CKH(cudaMemset(h->d_rec, 0, sizeof(StatRecord)));
  CKH(cudaHostAlloc(&h->h_rec, sizeof(StatRecord), cudaHostAllocMapped | cudaHostAllocPortable));
  std::memset(h->h_rec, 0, sizeof(StatRecord));
  h->seq = 0x100;  // tags start away from the zero-filled record
  env = std::getenv("DXM_TIMING");
  h->timing = env ? std::atoi(env) : -1;
  CKH(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
  CKH(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
  CKH(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
  h->stream = h->own_stream;
  for (int s = 0; s < 2; ++s) {
    CKH(cudaEventCreateWithFlags(&h->ev_in[s], cudaEventDisableTiming));
    CKH(cudaEventCreateWithFlags(&h->ev_in_free[s], cudaEventDisableTiming));
    CKH(cudaEventCreateWithFlags(&h->ev_packed[s], cudaEventDisableTiming));
    CKH(cudaEventCreateWithFlags(&h->ev_out_free[s], cudaEventDisableTiming));
  }
  if (behaviour == DXM_FEFP_VOCE) {
    // virgin finite-strain state: F = I, be_bar = I (jaxmat init_state; demo
    // finite_strain_elastoplasticity.py:181 sets be_bar to the identity explicitly)
    for (int g = 0; g < 2; ++g) {
      for (int r : {0, 1, 2, 19, 20, 21}) {
        fill_kernel<<<h->num_sms, 256, 0, h->stream>>>(h->gen[g] + (int64_t)r * h->ld, h->ld, 1.0);
        g_launches.fetch_add(1);
      }
    }
    CKH(cudaGetLastError());
    CKH(cudaStreamSynchronize(h->stream));
  }
#undef CKH
  *out = h;
  return 0;
}

int dxm_destroy(dxm_handle* h) {
  if (!h) return 0;
  release_handle(h);
  return 0;
}

int dxm_set_stream(dxm_handle* h, void* cuda_stream) {
  if (!h) return fail("dxm_set_stream: NULL handle");
  if (set_device(h)) return -1;
  CK(cudaStreamSynchronize(h->stream));
  h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
  return 0;
}

int64_t dxm_ld(const dxm_handle* h) { return h ? h->ld : -1; }
int64_t dxm_npoints(const dxm_handle* h) { return h ? h->n : -1; }

int dxm_field_dim(const dxm_handle* h, const char* field) {
  if (!h || !field) return -1;
  if (std::strcmp(field, "Ct") == 0) return h->nct;
  const Field* f = find_field(h, field);
  return f ? f->dim : -1;
}

int dxm_set_property(dxm_handle* h, const char* name, const double* v, int64_t count, int mem) {
  if (!h || !name || !v) return fail("dxm_set_property: NULL argument");
  if (std::strcmp(name, "a") == 0) {
    if (h->behaviour != DXM_HOSFORD_LINEAR) return fail("dxm_set_property: 'a' is a property of DXM_HOSFORD_LINEAR only");
    if (count != 1) return fail("dxm_set_property: the Hosford exponent 'a' is uniform (count must be 1)");
    double val;
    if (mem == DXM_MEM_HOST) {
      val = v[0];
    } else {
      if (set_device(h)) return -1;
      CK(cudaMemcpy(&val, v, sizeof(double), cudaMemcpyDeviceToHost));
    }
    const int ai = (int)val;
    if ((double)ai != val || ai < 2 || ai > 64 || (ai & 1))
      return fail("dxm_set_property: the Hosford exponent 'a' must be an even integer in [2, 64]");
    h->hos_a = ai;
    return 0;
  }
  int idx = -1;
  for (int i = 0; i < kNProp; ++i)
    if (std::strcmp(name, kPropNames[i]) == 0) idx = i;
  if (idx < 0) return fail(std::string("dxm_set_property: unknown property '") + name + "'");
  if (count != 1 && count != h->n)
    return fail(std::string("dxm_set_property: '") + name + "' needs 1 or n=" +
                std::to_string(h->n) + " values, got " + std::to_string(count));
  if (set_device(h)) return -1;
  if (count == 1 && !h->perpoint) {
    double val;
    if (mem == DXM_MEM_HOST)
      val = v[0];
    else
      CK(cudaMemcpy(&val, v, sizeof(double), cudaMemcpyDeviceToHost));
    h->uni[idx] = val;
    h->set[idx] = true;
    return 0;
  }
  // switch to (or stay in) per-point mode: all six rows live on the device
  if (!h->pp) {
    CK(cudaMalloc(&h->pp, sizeof(double) * kNProp * h->ld));
    for (int i = 0; i < kNProp; ++i) {
      fill_kernel<<<h->num_sms, 256, 0, h->stream>>>(h->pp + (int64_t)i * h->ld, h->ld, prop(h, i));
      LAUNCH_CHECK();
    }
    h->perpoint = true;
  }
  double* row = h->pp + (int64_t)idx * h->ld;
  if (count == 1) {
    double val;
    if (mem == DXM_MEM_HOST)
      val = v[0];
    else
      CK(cudaMemcpy(&val, v, sizeof(double), cudaMemcpyDeviceToHost));
    h->uni[idx] = val;
    fill_kernel<<<h->num_sms, 256, 0, h->stream>>>(row, h->ld, val);
    LAUNCH_CHECK();
  } else {
    CK(cudaMemcpyAsync(row, v, sizeof(double) * h->n,
                       mem == DXM_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice,
                       h->stream));
  }
  // an unset sigu follows sig0 (also per point)
  if (idx == 2 && !h->set[4]) {
    CK(cudaMemcpyAsync(h->pp + 4 * h->ld, h->pp + 2 * h->ld, sizeof(double) * h->ld,
                       cudaMemcpyDeviceToDevice, h->stream));
  }
  h->set[idx] = true;
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int dxm_set_hardening_table(dxm_handle* h, const double* p, const double* sig, int count) {
  if (!h || !p || !sig) return fail("dxm_set_hardening_table: NULL argument");
  if (h->behaviour != DXM_J2_TABLE) return fail("dxm_set_hardening_table: the handle is not a DXM_J2_TABLE behaviour");
  if (count < 2 || count > kMaxTable)
    return fail("dxm_set_hardening_table: between 2 and " + std::to_string(kMaxTable) + " points");
  if (p[0] != 0.0) return fail("dxm_set_hardening_table: p[0] must be 0");
  std::vector<double> t(3 * (size_t)count);
  for (int k = 0; k < count; ++k) {
    if (k > 0 && !(p[k] > p[k - 1])) return fail("dxm_set_hardening_table: p must increase strictly");
    if (!std::isfinite(p[k]) || !std::isfinite(sig[k])) return fail("dxm_set_hardening_table: non-finite entry");
    t[k] = p[k];
    t[count + k] = sig[k];
  }
  for (int k = 0; k + 1 < count; ++k) t[2 * count + k] = (sig[k + 1] - sig[k]) / (p[k + 1] - p[k]);
  t[3 * count - 1] = t[3 * count - 2];
  if (set_device(h)) return -1;
  CK(cudaStreamSynchronize(h->stream));
  if (!h->table) CK(cudaMalloc(&h->table, sizeof(double) * 3 * kMaxTable));
  CK(cudaMemcpy(h->table, t.data(), sizeof(double) * t.size(), cudaMemcpyHostToDevice));
  h->ntab = count;
  return 0;
}

int dxm_set_state(dxm_handle* h, int gen, const char* field, const double* v, int mem) {
  if (!h || !field || !v) return fail("dxm_set_state: NULL argument");
  if (gen != 0 && gen != 1) return fail("dxm_set_state: gen must be 0 or 1");
  const Field* f = find_field(h, field);
  if (!f) return fail(std::string("dxm_set_state: unknown field '") + field + "'");
  if (set_device(h)) return -1;
  if (gen == 1 && !h->s1_valid) {
    // materialise the alias before a partial write: flux and internal-state rows only -- the gradient rows of s1 are the
    // caller's to write even while s1 aliases s0 (dxm_device_ptr, GradientEvaluator) and must not be overwritten with
    // the previous step's
    const int64_t off = (int64_t)h->ngrad * h->ld;
    CK(cudaMemcpyAsync(h->gen[1 - h->i0] + off, h->gen[h->i0] + off, sizeof(double) * (h->nrows - h->ngrad) * h->ld,
                       cudaMemcpyDeviceToDevice, h->stream));
    h->s1_valid = true;
  }
  double* dst = field_ptr(h, gen, f, false);
  if (mem == DXM_MEM_RESIDENT) return fail("dxm_set_state: RESIDENT is not a source");
  if (mem == DXM_MEM_DEVICE) {
    if (launch_aos_to_soa(h, h->stream, v, f->dim, 0, dst, 0, h->n, f->dim)) return -1;
    CK(cudaStreamSynchronize(h->stream));
    return 0;
  }
  if (ensure_staging(h)) return -1;
  const int64_t rows_per = (h->chunk * (h->nflux + h->nisv + h->nct)) / f->dim;
  for (int64_t s = 0; s < h->n; s += rows_per) {
    const int64_t m = std::min(rows_per, h->n - s);
    CK(cudaMemcpyAsync(h->d_out[0], v + s * f->dim, sizeof(double) * m * f->dim,
                       cudaMemcpyHostToDevice, h->stream));
    if (launch_aos_to_soa(h, h->stream, h->d_out[0], f->dim, 0, dst, s, m, f->dim)) return -1;
    CK(cudaStreamSynchronize(h->stream));
  }
  return 0;
}

int dxm_get_state(dxm_handle* h, int gen, const char* field, double* out, int mem) {
  if (!h || !field || !out) return fail("dxm_get_state: NULL argument");
  if (gen != 0 && gen != 1) return fail("dxm_get_state: gen must be 0 or 1");
  if (set_device(h)) return -1;
  const double* src;
  int dim, sym6 = 0;
  if (std::strcmp(field, "Ct") == 0) {
    src = h->ct;
    dim = h->nct;
    sym6 = h->nct_store != h->nct;
  } else {
    const Field* f = find_field(h, field);
    if (!f) return fail(std::string("dxm_get_state: unknown field '") + field + "'");
    src = field_ptr(h, gen, f, true);
    dim = f->dim;
  }
  if (mem == DXM_MEM_RESIDENT) return fail("dxm_get_state: RESIDENT is not a destination");
  if (mem == DXM_MEM_DEVICE) {
    if (launch_soa_to_aos(h, h->stream, src, 0, out, dim, 0, h->n, dim, sym6)) return -1;
    CK(cudaStreamSynchronize(h->stream));
    return 0;
  }
  if (ensure_staging(h)) return -1;
  const int64_t rows_per = (h->chunk * (h->nflux + h->nisv + h->nct)) / dim;
  for (int64_t s = 0; s < h->n; s += rows_per) {
    const int64_t m = std::min(rows_per, h->n - s);
    if (launch_soa_to_aos(h, h->stream, src, s, h->d_out[0], dim, 0, m, dim, sym6)) return -1;
    CK(cudaMemcpyAsync(out + s * dim, h->d_out[0], sizeof(double) * m * dim, cudaMemcpyDeviceToHost,
                       h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  return 0;
}

int dxm_device_ptr(dxm_handle* h, int gen, const char* field, double** ptr) {
  if (!h || !field || !ptr) return fail("dxm_device_ptr: NULL argument");
  if (std::strcmp(field, "Ct") == 0) {
    *ptr = h->ct;
    return 0;
  }
  const Field* f = find_field(h, field);
  if (!f) return fail(std::string("dxm_device_ptr: unknown field '") + field + "'");
  // the gradient buffer of s1 is writable by the caller even while s1 aliases s0
  const bool is_grad = f->row == 0;
  *ptr = field_ptr(h, gen, f, !(gen == 1 && is_grad));
  return 0;
}

int dxm_export_dlpack(dxm_handle* h, int gen, const char* field, void** out) {
  if (!out) return fail("dxm_export_dlpack: out is NULL");
  double* p = nullptr;
  if (dxm_device_ptr(h, gen, field, &p)) return -1;
  // the resident tangent of the small-strain behaviours is packed: 21 rows (include/dxm.h)
  const int dim = std::strcmp(field, "Ct") == 0 ? h->nct_store : dxm_field_dim(h, field);
  DlCtx* ctx = new DlCtx{h, {dim, h->n}, {h->ld, 1}};
  DxmDLManagedTensor* t = new DxmDLManagedTensor();
  t->dl_tensor.data = p;
  t->dl_tensor.device = {2 /* kDLCUDA */, h->device};
  t->dl_tensor.ndim = 2;
  t->dl_tensor.dtype = {2 /* kDLFloat */, 64, 1};
  t->dl_tensor.shape = ctx->shape;
  t->dl_tensor.strides = ctx->strides;
  t->dl_tensor.byte_offset = 0;
  t->manager_ctx = ctx;
  t->deleter = dxm_dl_deleter;
  h->refs.fetch_add(1);
  *out = t;
  return 0;
}

int dxm_last_stats(dxm_handle* h, dxm_stats* stats) {
  if (!h || !stats) return fail("dxm_last_stats: NULL argument");
  if (set_device(h)) return -1;
  if (finish_stats(h)) return -1;
  *stats = h->last;
  return 0;
}

// points [start, start + count) of the handle; the host arrays hold `count` rows (dxm_integrate: the whole handle)
static int integrate_impl(dxm_handle* h, int64_t start, int64_t count, const double* grad, int mem, double dt,
                          double* flux, double* isv, double* ct, int out_mem, dxm_stats* stats) {
  if (!h) return fail("dxm_integrate: NULL handle");
  if (set_device(h)) return -1;
  if (!h->set[0] || !h->set[1]) return fail("dxm_integrate: properties E and nu must be set");
  if (h->behaviour == DXM_J2_TABLE) {
    if (!h->table) return fail("dxm_integrate: hardening table not set (dxm_set_hardening_table)");
  } else if (h->behaviour != DXM_ELASTIC && !h->set[2]) {
    return fail("dxm_integrate: property sig0 must be set");
  }
  if (mem != DXM_MEM_RESIDENT && !grad) return fail("dxm_integrate: grad is NULL");
  if (start < 0 || count <= 0 || start + count > h->n || (start & 1))
    return fail("dxm_integrate_range: need 0 <= start (even), count > 0, start + count <= n");
  const bool whole = start == 0 && count == h->n;
  if (!whole && mem != DXM_MEM_HOST && !(mem == DXM_MEM_RESIDENT && !flux && !isv && !ct))
    return fail("dxm_integrate_range: partial ranges take host arrays, or resident gradients without outputs");
  if (finish_stats(h)) return -1;  // drain a previous asynchronous call
  h->n_ev_used = 0;
  h->last = dxm_stats{};
  h->last.n_points = count;
  ++h->seq;
  h->finalize_launched = false;
  h->stats_pending = true;  // from here on the accumulation slots may be dirty: finish_stats() cleans up after a failure
  double* s1 = h->gen[1 - h->i0];
  const int64_t ld = h->ld, n = count;
  const int isv_row = h->ngrad + h->nflux;
  const bool any_out = flux || isv || ct;

  if (mem == DXM_MEM_RESIDENT || mem == DXM_MEM_DEVICE) {
    if (mem == DXM_MEM_DEVICE)
      if (launch_aos_to_soa(h, h->stream, grad, h->ngrad, 0, s1, 0, n, h->ngrad)) return -1;
    if (timed_update(h, start, n, dt, 2)) return -1;
    h->s1_valid = true;
    if (any_out) {
      if (out_mem == DXM_MEM_DEVICE) {
        if (flux && launch_soa_to_aos(h, h->stream, s1 + (int64_t)h->ngrad * ld, 0, flux, h->nflux,
                                      0, n, h->nflux))
          return -1;
        if (isv && launch_soa_to_aos(h, h->stream, s1 + (int64_t)isv_row * ld, 0, isv, h->nisv, 0,
                                     n, h->nisv))
          return -1;
        if (ct && launch_soa_to_aos(h, h->stream, h->ct, 0, ct, h->nct, 0, n, h->nct, h->nct_store != h->nct)) return -1;
      } else if (out_mem == DXM_MEM_HOST) {
        if (finish_stats(h)) return -1;
        const dxm_stats keep = h->last;
        if (flux && dxm_get_state(h, 1, h->fields[1].name, flux, DXM_MEM_HOST)) return -1;
        if (isv) {
          // isv = [p | second field] rows are contiguous in the SoA block: one (n, nisv) gather
          if (ensure_staging(h)) return -1;
          const int64_t rows_per = (h->chunk * (h->nflux + h->nisv + h->nct)) / h->nisv;
          for (int64_t s = 0; s < n; s += rows_per) {
            const int64_t m = std::min(rows_per, n - s);
            if (launch_soa_to_aos(h, h->stream, s1 + (int64_t)isv_row * ld, s, h->d_out[0], h->nisv,
                                  0, m, h->nisv))
              return -1;
            CK(cudaMemcpyAsync(isv + s * h->nisv, h->d_out[0], sizeof(double) * m * h->nisv,
                               cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
          }
        }
        if (ct && dxm_get_state(h, 1, "Ct", ct, DXM_MEM_HOST)) return -1;
        h->last = keep;
      }
    }
  } else if (mem == DXM_MEM_HOST) {
    if (any_out && out_mem != DXM_MEM_HOST)
      return fail("dxm_integrate: host gradients require host outputs");
    if (n <= kSmallHostPoints) {
      // small batches (the reference's own test meshes hold 1 ... 16 points): latency, not bandwidth.  The gradients are
      // copied by the CPU into a mapped page-locked block which the transposition kernel reads over the link, the packed
      // outputs are written straight into that block by the device, and one stream wait ends the call.
      const int nf = h->nflux, ni = h->nisv, nc = h->nct, ncs = h->nct_store;
      const int64_t K = kSmallHostPoints;
      if (!h->h_small) {
        CK(cudaHostAlloc(&h->h_small, sizeof(double) * K * (h->ngrad + nf + ni + nc),
                         cudaHostAllocMapped | cudaHostAllocPortable));
      }
      double* hin = h->h_small;
      double* hfl = hin + K * h->ngrad;
      double* his = hfl + K * nf;
      double* hct = his + K * ni;
      std::memcpy(hin, grad, sizeof(double) * n * h->ngrad);
      if (launch_aos_to_soa(h, h->stream, hin, h->ngrad, 0, s1, start, n, h->ngrad)) return -1;
      if (timed_update(h, start, n, dt, 2)) return -1;
      if (flux && launch_soa_to_aos(h, h->stream, s1 + (int64_t)h->ngrad * ld, start, hfl, nf, 0, n, nf)) return -1;
      if (isv && launch_soa_to_aos(h, h->stream, s1 + (int64_t)isv_row * ld, start, his, ni, 0, n, ni)) return -1;
      if (ct && launch_soa_to_aos(h, h->stream, h->ct, start, hct, nc, 0, n, nc, ncs != nc)) return -1;
      h->s1_valid = true;
      CK(cudaStreamSynchronize(h->stream));
      if (flux) std::memcpy(flux, hfl, sizeof(double) * n * nf);
      if (isv) std::memcpy(isv, his, sizeof(double) * n * ni);
      if (ct) std::memcpy(ct, hct, sizeof(double) * n * nc);
      if (!stats) return 0;
      if (finish_stats(h)) return -1;
      *stats = h->last;
      return (int)std::min<int64_t>(h->last.n_fail, 0x7fffffff);
    }
    if (ensure_staging(h)) return -1;
    // 3-stage pipeline over chunks: H2D (s_in) | transpose + update + pack (stream) | D2H (s_out)
    // (+ a 4th stage on the host when the packed tangent is mirrored there, dxm_host_mirror.hpp)
    const int64_t CH = h->chunk;
    const int nf = h->nflux, ni = h->nisv, nc = h->nct, ncs = h->nct_store;
    const bool mirror = ct && ncs != nc && host_mirror_enabled();
    if (mirror && ensure_mirror_ring(h)) return -1;
    struct Pending {
      int slot;
      int64_t s, m;
    };
    Pending pend[dxm_handle::kRing] = {};
    int npend = 0;
    auto drain_one = [&]() -> int {  // oldest in-flight packed chunk -> the caller's (n, 36) rows
      const Pending p = pend[0];
      for (int i = 1; i < npend; ++i) pend[i - 1] = pend[i];
      --npend;
      CK(cudaEventSynchronize(h->ev_ct[p.slot]));
      dxm_host::mirror_sym6(h->h_ctp[p.slot], ct + p.s * nc, p.m, 0);
      return 0;
    };
    // make the copy streams wait for whatever precedes on the compute stream
    CK(cudaEventRecord(h->ev_in_free[0], h->stream));
    CK(cudaEventRecord(h->ev_in_free[1], h->stream));
    CK(cudaEventRecord(h->ev_out_free[0], h->stream));
    CK(cudaEventRecord(h->ev_out_free[1], h->stream));
    int64_t c = 0;
    for (int64_t s = 0; s < n; s += CH, ++c) {
      const int64_t m = std::min(CH, n - s);
      const int b = (int)(c & 1);
      CK(cudaStreamWaitEvent(h->s_in, h->ev_in_free[b], 0));
      CK(cudaMemcpyAsync(h->d_in[b], grad + s * h->ngrad, sizeof(double) * m * h->ngrad,
                         cudaMemcpyHostToDevice, h->s_in));
      CK(cudaEventRecord(h->ev_in[b], h->s_in));
      CK(cudaStreamWaitEvent(h->stream, h->ev_in[b], 0));
      if (launch_aos_to_soa(h, h->stream, h->d_in[b], h->ngrad, 0, s1, start + s, m, h->ngrad)) return -1;
      CK(cudaEventRecord(h->ev_in_free[b], h->stream));
      if (timed_update(h, start + s, m, dt, s + CH >= n ? (n <= CH ? 2 : 1) : 0)) return -1;
      if (any_out) {
        CK(cudaStreamWaitEvent(h->stream, h->ev_out_free[b], 0));
        double* o = h->d_out[b];
        if (flux && launch_soa_to_aos(h, h->stream, s1 + (int64_t)h->ngrad * ld, start + s, o, nf, 0, m, nf))
          return -1;
        if (isv && launch_soa_to_aos(h, h->stream, s1 + (int64_t)isv_row * ld, start + s, o + CH * nf, ni, 0,
                                     m, ni))
          return -1;
        if (ct && (mirror ? launch_soa_to_aos(h, h->stream, h->ct, start + s, o + CH * (nf + ni), ncs, 0, m, ncs, 0)
                          : launch_soa_to_aos(h, h->stream, h->ct, start + s, o + CH * (nf + ni), nc, 0, m, nc, ncs != nc)))
          return -1;
        CK(cudaEventRecord(h->ev_packed[b], h->stream));
        CK(cudaStreamWaitEvent(h->s_out, h->ev_packed[b], 0));
        if (flux)
          CK(cudaMemcpyAsync(flux + s * nf, o, sizeof(double) * m * nf, cudaMemcpyDeviceToHost,
                             h->s_out));
        if (isv)
          CK(cudaMemcpyAsync(isv + s * ni, o + CH * nf, sizeof(double) * m * ni,
                             cudaMemcpyDeviceToHost, h->s_out));
        if (mirror) {
          // ring slot c % kRing was last used by chunk c - kRing, drained below before chunk c - 1 was left
          const int slot = (int)(c % dxm_handle::kRing);
          CK(cudaMemcpyAsync(h->h_ctp[slot], o + CH * (nf + ni), sizeof(double) * m * ncs,
                             cudaMemcpyDeviceToHost, h->s_out));
          CK(cudaEventRecord(h->ev_ct[slot], h->s_out));
          pend[npend++] = Pending{slot, s, m};
        } else if (ct) {
          CK(cudaMemcpyAsync(ct + s * nc, o + CH * (nf + ni), sizeof(double) * m * nc,
                             cudaMemcpyDeviceToHost, h->s_out));
        }
        CK(cudaEventRecord(h->ev_out_free[b], h->s_out));
        // keep kRing - 1 chunks queued on the device while the host mirrors the oldest one
        while (npend > dxm_handle::kRing - 1)
          if (drain_one()) return -1;
      }
    }
    while (npend > 0)
      if (drain_one()) return -1;
    h->s1_valid = true;
    CK(cudaStreamSynchronize(h->s_out));
    CK(cudaStreamSynchronize(h->stream));
  } else {
    return fail("dxm_integrate: bad mem kind");
  }
  if (!stats) return 0;  // asynchronous: fetch later with dxm_last_stats
  if (finish_stats(h)) return -1;
  *stats = h->last;
  return (int)std::min<int64_t>(h->last.n_fail, 0x7fffffff);
}

int dxm_integrate(dxm_handle* h, const double* grad, int mem, double dt, double* flux, double* isv,
                  double* ct, int out_mem, dxm_stats* stats) {
  return integrate_impl(h, 0, h ? h->n : 0, grad, mem, dt, flux, isv, ct, out_mem, stats);
}

int dxm_integrate_range(dxm_handle* h, int64_t start, int64_t count, const double* grad, int mem, double dt,
                        double* flux, double* isv, double* ct, int out_mem, dxm_stats* stats) {
  return integrate_impl(h, start, count, grad, mem, dt, flux, isv, ct, out_mem, stats);
}

int dxm_update(dxm_handle* h) {
  if (!h) return fail("dxm_update: NULL handle");
  if (h->s1_valid) {
    h->i0 = 1 - h->i0;  // s0 <- s1 by swapping generations; s1 now aliases s0 until rewritten
    h->s1_valid = false;
  }
  return 0;
}

int dxm_revert(dxm_handle* h) {
  if (!h) return fail("dxm_revert: NULL handle");
  h->s1_valid = false;  // s1 <- s0
  return 0;
}

int dxm_use_global_stats(dxm_handle* h, int on) {
  if (!h) return fail("dxm_use_global_stats: NULL handle");
  if (set_device(h)) return -1;
  if (finish_stats(h)) return -1;
  if (on) {
    if (dxm_comm::size() < 2) return fail("dxm_use_global_stats: no multi-rank communicator (dxm_comm_init)");
    if (!h->d_gather) CK(cudaMalloc(&h->d_gather, sizeof(StatRecord) * dxm_comm::size()));
    if ((h->xslot < 0 || h->xgen != dxm_comm::generation()) && dxm_comm::xchg()) {
      h->xslot = dxm_comm::xchg_slot();  // same order on every rank
      h->xgen = dxm_comm::generation();
    }
    if (h->d_gather && h->xgen != dxm_comm::generation()) {  // a new communicator may have another size
      cudaFree(h->d_gather);
      h->d_gather = nullptr;
      CK(cudaMalloc(&h->d_gather, sizeof(StatRecord) * dxm_comm::size()));
    }
  }
  h->global_stats = on != 0;
  return 0;
}

int dxm_enable_timing(dxm_handle* h, int mode) {
  if (!h) return fail("dxm_enable_timing: NULL handle");
  h->timing = mode;
  return 0;
}

int dxm_enable_diagnostics(dxm_handle* h, int on) {
  if (!h) return fail("dxm_enable_diagnostics: NULL handle");
  if (set_device(h)) return -1;
  if (on && !h->d_flag) {
    CK(cudaMalloc(&h->d_flag, h->ld));
    CK(cudaMalloc(&h->d_fail, h->ld));
    CK(cudaMalloc(&h->d_iter, sizeof(int32_t) * h->ld));
    CK(cudaMalloc(&h->d_resid, sizeof(double) * h->ld));
    CK(cudaMemset(h->d_flag, 0, h->ld));
    CK(cudaMemset(h->d_fail, 0, h->ld));
    CK(cudaMemset(h->d_iter, 0, sizeof(int32_t) * h->ld));
    CK(cudaMemset(h->d_resid, 0, sizeof(double) * h->ld));
  }
  h->diag = on != 0;
  return 0;
}

int dxm_get_diagnostics(dxm_handle* h, uint8_t* flag, int32_t* n_iter, double* resid,
                        uint8_t* failed) {
  if (!h) return fail("dxm_get_diagnostics: NULL handle");
  if (!h->d_flag) return fail("dxm_get_diagnostics: diagnostics were never enabled");
  if (set_device(h)) return -1;
  CK(cudaStreamSynchronize(h->stream));
  if (flag) CK(cudaMemcpy(flag, h->d_flag, h->n, cudaMemcpyDeviceToHost));
  if (failed) CK(cudaMemcpy(failed, h->d_fail, h->n, cudaMemcpyDeviceToHost));
  if (n_iter) CK(cudaMemcpy(n_iter, h->d_iter, sizeof(int32_t) * h->n, cudaMemcpyDeviceToHost));
  if (resid) CK(cudaMemcpy(resid, h->d_resid, sizeof(double) * h->n, cudaMemcpyDeviceToHost));
  return 0;
}

int dxm_synth_gradients(dxm_handle* h, int recipe, uint64_t seed, double amp, int k, int K,
                        int64_t start) {
  if (!h) return fail("dxm_synth_gradients: NULL handle");
  if (recipe != 0 && recipe != 1) return fail("dxm_synth_gradients: recipe must be 0 or 1");
  if ((recipe == 0) != (h->ngrad == 6))
    return fail("dxm_synth_gradients: recipe does not match the behaviour's gradient");
  if (K <= 0) return fail("dxm_synth_gradients: K must be positive");
  if (set_device(h)) return -1;
  const double kfrac = (double)k / (double)K;
  synth_kernel<<<h->num_sms * 8, 256, 0, h->stream>>>(h->gen[1 - h->i0], h->ld, h->n, h->ngrad,
                                                       recipe, seed, amp, kfrac, start);
  LAUNCH_CHECK();
  return 0;
}

int dxm_host_mirror_sym6(const double* packed, double* full, int64_t n, int threads) {
  if (n < 0 || (n > 0 && (!packed || !full))) return fail("dxm_host_mirror_sym6: bad argument");
  dxm_host::mirror_sym6(packed, full, n, threads);
  return 0;
}

int dxm_host_gather_rows(const double* src, const int64_t* rows, int64_t n, int64_t row_len, double* dst, int threads) {
  if (n < 0 || row_len < 0 || (n > 0 && row_len > 0 && (!src || !rows || !dst))) return fail("dxm_host_gather_rows: bad argument");
  dxm_host::gather_rows(src, rows, n, row_len, dst, threads);
  return 0;
}

int dxm_host_scatter_rows(double* dst, const int64_t* rows, int64_t n, int64_t row_len, const double* src, int threads) {
  if (n < 0 || row_len < 0 || (n > 0 && row_len > 0 && (!src || !rows || !dst))) return fail("dxm_host_scatter_rows: bad argument");
  dxm_host::scatter_rows(dst, rows, n, row_len, src, threads);
  return 0;
}

int dxm_host_alloc(void** ptr, int64_t bytes) {
  if (!ptr || bytes < 0) return fail("dxm_host_alloc: bad argument");
  CK(cudaMallocHost(ptr, (size_t)(bytes ? bytes : 1)));
  return 0;
}

// registrations made through this library: address -> (bytes, reference count).  Registering the same range again only
// bumps the count; a range that overlaps someone else's registration is reported, not silently treated as valid.
static std::mutex g_reg_mu;
static std::map<void*, std::pair<int64_t, int>> g_reg;

int dxm_host_register(void* ptr, int64_t bytes) {
  if (!ptr || bytes <= 0) return fail("dxm_host_register: bad argument");
  std::lock_guard<std::mutex> lock(g_reg_mu);
  auto it = g_reg.find(ptr);
  if (it != g_reg.end()) {
    if (it->second.first != bytes)
      return fail("dxm_host_register: this address is already page-locked with a different size (" +
                  std::to_string(it->second.first) + " bytes); unregister it first");
    ++it->second.second;
    return 0;
  }
  cudaError_t e = cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable);
  if (e == cudaErrorHostMemoryAlreadyRegistered) {
    cudaGetLastError();
    return fail("dxm_host_register: the range overlaps memory page-locked by someone else (stale or partial "
                "registration); it cannot be trusted to cover the array");
  }
  CK(e);
  g_reg[ptr] = {bytes, 1};
  return 0;
}

int dxm_host_unregister(void* ptr) {
  if (!ptr) return 0;
  std::lock_guard<std::mutex> lock(g_reg_mu);
  auto it = g_reg.find(ptr);
  if (it == g_reg.end()) return 0;  // not ours (or already released)
  if (--it->second.second > 0) return 0;
  g_reg.erase(it);
  cudaError_t e = cudaHostUnregister(ptr);
  if (e == cudaErrorHostMemoryNotRegistered) {
    cudaGetLastError();
    return 0;
  }
  CK(e);
  return 0;
}

int dxm_host_free(void* ptr) {
  if (ptr) CK(cudaFreeHost(ptr));
  return 0;
}

}  // extern "C"
