// C ABI of libflimo_cuda.so (declared in include/flimo.h).  Host-side plumbing only: device
// buffers, streams, pose constants, launch sequencing, the MAX_NUM_MATCHES first-N rule, and the
// host IKFoM update (ekf_host.hpp).  No CPU fallback exists: without a working CUDA device every
// entry point fails with a negative status.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <immintrin.h>
#include <new>
#include <string>
#include <vector>

#include "../../include/flimo.h"
#include "ekf_host.hpp"
#include "ekf_step.hpp"
#include "imu_host.hpp"
#include "flimo_dev.cuh"
#include "pose_consts.hpp"
#include "scan_prep.cuh"

using namespace flimo;

// Index ladder defaults (tools/tune_knn.py sweeps on the 5 M-point headline map, B200): finest cell 0.25 m,
// cells grow 1.5x per level, a query starts on the finest level whose block holds >= 8 candidates.
constexpr float kDefaultRatio = 1.5f;
constexpr int kDefaultTau = 8;

struct flimo_ctx {
  flimo_cfg cfg{};
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::string err;

  MapIndex map;
  bool map_exists = false;
  double last_time = -1.0;       // Mapper::last_map_time (Mapper.cpp:23)

  float4* scan = nullptr;        // packed scan (w = original index); built on demand unless sort_scan
  const unsigned char* raw_scan = nullptr;   // the bound scan as uploaded (strided xyz, device memory)
  size_t raw_stride = 0;
  bool packed_valid = false;     // h->scan holds the bound scan
  unsigned int raw_inv = 1;      // inverse of the storage permutation's stride (mod scan_n)
  // host uploads: two staging buffers, so that the next scan can be copied while this one is registered
  void* scan_stage[2] = {nullptr, nullptr};
  size_t scan_stage_cap[2] = {0, 0};
  int scan_stage_bound = 0;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_prefetch = nullptr;
  const void* pref_src = nullptr;   // pending prefetch: host pointer, size, stride, staging buffer
  size_t pref_n = 0, pref_stride = 0;
  int pref_idx = -1;
  bool pref_issued = false;
  float4* scan_tmp = nullptr;
  size_t scan_cap = 0, scan_n = 0;
  size_t scan_n_full = 0;         // points of the bound cloud before the MAX_NUM_PC2MATCH cap (all of them go to the map)
  size_t shard_begin = 0, shard_end = 0;
  uint32_t* scan_keys = nullptr;
  size_t scan_keys_cap = 0;
  void* scan_cub = nullptr;
  size_t scan_cub_bytes = 0;

  void* stage = nullptr;         // raw strided uploads
  size_t stage_cap = 0;
  unsigned int* d_count = nullptr;
  float4* batch = nullptr;                // packed incoming batch
  size_t batch_cap = 0;
  unsigned long long* batch_keys = nullptr;
  uint8_t* batch_accept = nullptr;
  OctreeLattice lattice;                  // the reference octree's cell lattice (insert rule)
  CountTable counts;

  double* partials = nullptr;
  size_t partials_cap = 0;       // in doubles
  unsigned int* ticket = nullptr;
  size_t ticket_cap = 0;
  double* out96 = nullptr;       // device
  double* h_out96 = nullptr;     // pinned + mapped host copy: 96 records {double value, u64 seq} (kernel tail)
  double* d_h_out96 = nullptr;   // device alias of h_out96
  unsigned long long seq = 0;    // sequence number of the last blocking pass
  // multi-process exchange segment (flimo_exchange_attach)
  double* xch_host = nullptr;
  double* xch_dev = nullptr;
  int xch_rank = 0, xch_world = 0;
  unsigned long long xch_seq = 0;
  struct ScanGraph {
    const void* src;
    size_t n, stride;
    int sort;
    cudaGraphExec_t exec;
  };
  std::vector<ScanGraph> scan_graphs;   // captured upload pipelines (pack + sort + gather)
  float* dbg16 = nullptr;
  size_t dbg_cap = 0;
  uint8_t* valid_flags = nullptr;
  size_t valid_cap = 0;
  float* xyz_out = nullptr;
  unsigned long long* timing = nullptr;   // tools/warp_timing.py only
  size_t xyz_cap = 0;

  int knn_tau = 24;              // level choice threshold (MatchParams::tau)
  int probe_mode = 1, wide_loads = 1, pair_scan = 0;
  int interleave = 0, scan_perm = 1, l2_prefetch = 0, stage_runs = 0, fx_reduce = 1;
  int index_incremental = 1;     // FLIMO_INDEX_INCREMENTAL=0: every Mapper::add rebuilds the whole index
  uint64_t stats_index_builds = 0, stats_index_updates = 0, stats_update_stalls = 0;
  uint4* cta_trace = nullptr;     // FLIMO_DEBUG_CTA_TRACE=1: per-CTA progress records of the registration tiles kernel (diagnostics)
  bool pdl = true;                // FLIMO_PDL=0: filter and tiles kernels on two streams instead of a programmatic dependent launch
  int debug_stall_every = 0;      // FLIMO_DEBUG_STALL_EVERY: fault injection for the stall recovery of flimo_update (tests)
  int time_every = 8;            // every n-th flimo_update runs one launch per pass, each timed with CUDA events (0 = never)
  // persistent kernel (one launch per flimo_update)
  PassCtlWire* h_ctl = nullptr;  // mapped pinned host control block (tagged 16-byte records)
  PassCtlWire* d_h_ctl = nullptr;   // its device alias
  PassCtl* dev_ctl = nullptr;    // device copy
  int persistent = 1;            // 1 = flimo_update keeps one kernel resident over all passes
  unsigned long long ctl_seq = 0;   // tag of the last command posted to the persistent kernel
  int persist_capacity = 0;      // co-resident CTAs of the persistent kernel
  // registration kernel (whole update on the device, filter step included)
  double wait_timeout_s = 30.0;  // wall-clock limit of a wait for result records (FLIMO_WAIT_TIMEOUT_S)
  int device_ekf = 1;            // FLIMO_DEVICE_EKF=0: keep the filter step on the host (persistent kernel + handshake)
  int reg_capacity = 0;
  ekf::UpdState* upd_state = nullptr;   // device
  ekf::UpdInit* h_in = nullptr;         // mapped pinned input block of the filter kernel
  ekf::UpdInit* d_h_in = nullptr;       // its device alias
  ekf::UpdInit* dev_in = nullptr;       // device copy (made by the filter kernel)
  cudaStream_t filter_stream = nullptr;
  double* h_res = nullptr;       // mapped pinned result block: kResRecords records {value, seq}
  double* d_h_res = nullptr;
  unsigned long long res_seq = 0;
  uint32_t* flag_words = nullptr;
  size_t flag_words_cap = 0;
  uint64_t device_updates = 0, device_redone = 0;
  double last_x_dev[26] = {0};   // the device's own state after the last pass of the last update (tests)
  // peer exchange over NVLink (flimo_peer_*): every rank's inbox mapped through CUDA IPC
  double* inbox = nullptr;
  double* peer_inbox[kMaxPeers] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int peer_rank = 0, peer_world = 0;
  unsigned long long peer_xseq = 0;
  double persist_ns_total = 0;   // in-kernel device time of persistent passes
  double exchange_ns_total = 0;  // of it: own tiles complete -> sums of all ranks in hand (device-resident updates)
  uint64_t persist_passes = 0;
  uint64_t update_calls = 0;
  int test_stall_pass = -1;      // FLIMO_TEST_STALL_PASS=k: the host sleeps 60 ms before command k of every update (watchdog test)
  bool prof = false;             // FLIMO_PROFILE=1: host-side wall-clock breakdown printed by flimo_destroy
  double prof_launch = 0, prof_wait = 0, prof_step = 0, prof_other = 0;
  double prof_add_pack = 0, prof_add_insert = 0, prof_add_index = 0;
  uint64_t prof_adds = 0;
  uint64_t prof_passes = 0;

  // scan preparation (scan_prep.cu)
  PrepBuffers prep;
  flimo_prep_cfg prep_cfg{};
  const float4* prep_pc2match = nullptr;
  size_t prep_n_pc2match = 0;
  bool prep_deskewed = false;

  ekf::IteratedUpdate upd;
  ekf::CoopUpdate coop;            // FLIMO_EKF_COOP=1: flimo_ekf_* run the device code path's algebra (ekf_step.hpp) on the host
  bool use_coop = false;
  bool upd_active = false;
  ekf::PropagatedRing propagated;                 // Localizer::propagated_buffer
  std::vector<ekf::PropagatedState> frames_tmp;

  flimo_stats stats{};
  std::vector<cudaEvent_t> ev_pending, ev_pool;   // async match launches awaiting timing
};

namespace {

thread_local std::string g_err;   // for errors without a handle

int fail(flimo_handle h, int code, const std::string& msg) {
  if (h) h->err = msg;
  else g_err = msg;
  return code;
}

#define CU(h, call)                                                                           \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess)                                                                   \
      return fail(h, e__ == cudaErrorMemoryAllocation ? FLIMO_ERR_NOMEM : FLIMO_ERR_CUDA,     \
                  std::string(#call) + ": " + cudaGetErrorString(e__));                       \
  } while (0)

#define NEED_GPU(h)                                                                            \
  do {                                                                                        \
    if ((h)->device < 0) return fail(h, FLIMO_ERR_NO_DEVICE, "host-only handle: no GPU bound"); \
    CU(h, cudaSetDevice((h)->device));                                                        \
  } while (0)

template <typename T>
cudaError_t grow(T** p, size_t* cap, size_t need) {
  if (*cap >= need) return cudaSuccess;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  size_t want = need + need / 4 + 256;
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), want * sizeof(T));
  if (e != cudaSuccess) return e;
  *cap = want;
  return cudaSuccess;
}

// stride ~ 0.618 n that is co-prime with n: i -> (i * stride) mod n is a permutation of [0, n)
uint32_t coprime_stride(uint32_t n) {
  if (n < 3) return 1;
  uint32_t s = (uint32_t)(0.6180339887 * n) | 1u;
  auto gcd = [](uint32_t a, uint32_t b) {
    while (b) {
      const uint32_t t = a % b;
      a = b;
      b = t;
    }
    return a;
  };
  while (gcd(s, n) != 1) s += 2;
  return s % n ? s % n : 1;
}

// smallest float f with (double)f >= D:  (double)x < D  <=>  x < f  for every float x
float ceil_to_float(double D) {
  float f = (float)D;
  if ((double)f < D) f = std::nextafterf(f, INFINITY);
  return f;
}

int upload(flimo_handle h, const void* host, size_t bytes) {
  if (bytes > h->stage_cap) {
    if (h->stage) cudaFree(h->stage);
    h->stage = nullptr;
    h->stage_cap = 0;
    CU(h, cudaMalloc(&h->stage, bytes + bytes / 4 + 4096));
    h->stage_cap = bytes + bytes / 4 + 4096;
  }
  CU(h, cudaMemcpyAsync(h->stage, host, bytes, cudaMemcpyHostToDevice, h->stream));
  return FLIMO_OK;
}

int fill_params(flimo_handle h, const double state14[14], MatchParams& P, double* d_out, uint32_t orig_limit,
                float* dbg, uint8_t* valid) {
  const size_t n = h->shard_end - h->shard_begin;
  const int tiles = match_num_tiles((int)n);
  const int groups = (tiles + 31) / 32;
  const size_t need_part = (size_t)(tiles + groups) * kPartialStride;
  CU(h, grow(&h->partials, &h->partials_cap, need_part));
  if ((size_t)groups + 1 + kTicketWords > h->ticket_cap) {       // counters + the fixed-point pass accumulators (flimo_dev.cuh)
    CU(h, grow(&h->ticket, &h->ticket_cap, (size_t)groups + 1 + kTicketWords));
    CU(h, cudaMemsetAsync(h->ticket, 0, h->ticket_cap * sizeof(unsigned int), h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
  }
  P.scan = h->packed_valid ? h->scan : nullptr;
  P.raw_scan = h->raw_scan;
  P.raw_stride = (uint32_t)h->raw_stride;
  P.raw_n = (uint32_t)h->scan_n;
  P.raw_inv = h->raw_inv;
  P.raw_vec4 = (h->raw_stride % 16 == 0 && (reinterpret_cast<uintptr_t>(h->raw_scan) % 16) == 0) ? 1 : 0;
  P.raw_magic = h->scan_n ? (~0ull) / (unsigned long long)h->scan_n : 0ull;
  P.n_levels = h->map.n_levels;
  for (int l = 0; l < kMaxLevels; ++l) {
    P.lv[l].pts = h->map.lv[l].pts;
    P.lv[l].cell_start = h->map.lv[l].cell_start;
    P.lv[l].row_stride = h->map.lv[l].g.nx + 1;
    P.lv[l].g = h->map.lv[l].g;
  }
  make_pose(state14, P.pc);
  P.q_begin = (int)h->shard_begin;
  P.interleave = h->interleave;
  P.q_end = (int)h->shard_end;
  P.tau = h->knn_tau;
  P.probe_mode = h->probe_mode;
  P.wide_loads = h->wide_loads;
  P.pair_scan = h->pair_scan;
  P.l2_prefetch = h->l2_prefetch;
  P.stage_runs = h->stage_runs;
  P.fx_reduce = h->fx_reduce;
  P.max_dist_f = ceil_to_float(h->cfg.MAX_DIST_PLANE);
  P.plane_thr = (float)h->cfg.PLANE_THRESHOLD;
  P.estimate_extrinsics = h->cfg.estimate_extrinsics;
  P.orig_limit = orig_limit;
  P.partials = h->partials;
  P.ticket = h->ticket;
  P.out96 = d_out;
  P.dbg16 = dbg;
  P.valid_by_orig = valid;
  P.timing = h->timing;
  P.cta_trace = h->cta_trace;
  P.host_out96 = nullptr;
  P.host_out_alt = 0;
  P.seq = 0;
  P.host_ctl = h->d_h_ctl;
  P.dev_ctl = h->dev_ctl;
  P.watchdog_ns = 20ull * 1000ull * 1000ull;          // 20 ms of host silence ends the persistent kernel
  P.ctl_seq = 0;
  return FLIMO_OK;
}

// Waits until all 96 records {double value, u64 seq} of a result block in mapped host memory carry
// `seq` and copies the values out.  Each record is written by the kernel with ONE 16-byte store, so
// value and sequence word become visible together: reading the sequence word first (x86 keeps load
// order) makes the value that follows current.  Returns 0, 1 (own kernel finished without
// publishing; only when `own`), or a negative status.
int wait_records_n(flimo_handle h, const double* block, unsigned long long seq, double* out, int n, bool own);
int wait_records(flimo_handle h, const double* block, unsigned long long seq, double out[96], bool own) {
  return wait_records_n(h, block, seq, out, 96, own);
}
int wait_records_n(flimo_handle h, const double* block, unsigned long long seq, double* out, int n, bool own) {
  const volatile unsigned long long* rec = reinterpret_cast<const volatile unsigned long long*>(block);
  unsigned long long spins = 0;
  std::chrono::steady_clock::time_point t0;
  for (int i = n - 1; i >= 0; --i) {
    while (rec[2 * i + 1] != seq) {
      if ((++spins & 0xFFFFF) == 0) {                     // every ~1M polls (a few ms)
        if (own) {                                        // has the kernel died?
          const cudaError_t q = cudaStreamQuery(h->stream);
          if (q != cudaSuccess && q != cudaErrorNotReady) return fail(h, FLIMO_ERR_CUDA, std::string("match kernel: ") + cudaGetErrorString(q));
          if (q == cudaSuccess && rec[2 * i + 1] != seq) return 1;
        }
        // wall-clock limit: a peer rank that died (or never made the matching call) must not hang this one for ever
        const auto now = std::chrono::steady_clock::now();
        if (spins == 0x100000) t0 = now;
        else if (std::chrono::duration<double>(now - t0).count() > h->wait_timeout_s)
          return fail(h, FLIMO_ERR_STATE, own ? "timed out waiting for a measurement result" : "timed out waiting for a peer rank's measurement result");
      }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    unsigned long long bits = rec[2 * i];
    std::memcpy(&out[i], &bits, sizeof(double));
  }
  return 0;
}

// Issues a requested prefetch copy now (copy stream).  Called right after the first kernel launch of an
// update — the API call then overlaps the kernel the host would otherwise only wait for — or, at the
// latest, by the flimo_scan_set that consumes the prefetch.
int issue_prefetch(flimo_handle h) {
  if (h->pref_idx < 0 || h->pref_issued) return FLIMO_OK;
  CU(h, cudaMemcpyAsync(h->scan_stage[h->pref_idx], h->pref_src, h->pref_n * h->pref_stride, cudaMemcpyHostToDevice, h->copy_stream));
  CU(h, cudaEventRecord(h->ev_prefetch, h->copy_stream));
  h->pref_issued = true;
  return FLIMO_OK;
}

// Blocking pass.  The last CTA of the kernel also writes the 96 result doubles, each tagged with a
// sequence number, into MAPPED pinned host memory; the host spins on those records instead of issuing a
// D2H copy and a stream synchronise (saves ~10 us of launch/sync latency per pass).  Device time is taken from a
// pair of events that is resolved lazily in flimo_get_stats.
int run_pass_blocking(flimo_handle h, const double state14[14], uint32_t orig_limit, float* dbg, uint8_t* valid,
                      double packed[96]) {
  const auto tp0 = std::chrono::steady_clock::now();
  MatchParams P;
  int rc = fill_params(h, state14, P, h->out96, orig_limit, dbg, valid);
  if (rc) return rc;
  P.host_out96 = h->d_h_out96;
  P.seq = ++h->seq;
  // device time of the kernel: a pair of events around every `time_every`-th launch
  const bool timed = h->time_every > 0 && (h->persistent || (h->stats.match_launches % (uint64_t)h->time_every) == 0);
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (timed) {
    if (h->ev_pool.size() >= 2) {
      e0 = h->ev_pool.back(); h->ev_pool.pop_back();
      e1 = h->ev_pool.back(); h->ev_pool.pop_back();
    } else {
      CU(h, cudaEventCreate(&e0));
      CU(h, cudaEventCreate(&e1));
    }
    CU(h, cudaEventRecord(e0, h->stream));
  }
  CU(h, launch_match(P, h->stream));
  if (timed) {
    CU(h, cudaEventRecord(e1, h->stream));
    h->ev_pending.push_back(e0);
    h->ev_pending.push_back(e1);
  }
  h->stats.kernel_launches++;
  h->stats.match_launches++;
  if (h->pref_idx >= 0 && !h->pref_issued) {               // the requested copy of the next scan overlaps this kernel
    rc = issue_prefetch(h);
    if (rc) return rc;
  }
  const auto tp1 = std::chrono::steady_clock::now();
  const int wr = wait_records(h, h->h_out96, P.seq, packed, true);
  if (h->prof) {
    const auto tp2 = std::chrono::steady_clock::now();
    h->prof_launch += std::chrono::duration<double, std::micro>(tp1 - tp0).count();
    h->prof_wait += std::chrono::duration<double, std::micro>(tp2 - tp1).count();
    h->prof_passes++;
  }
  if (wr == 1) {                                           // finished without publishing: read the device copy
    CU(h, cudaMemcpy(packed, h->out96, 96 * sizeof(double), cudaMemcpyDeviceToHost));
  } else if (wr < 0) {
    return wr;
  }
  if (h->ev_pending.size() > 64) {                         // keep the lazy list short
    flimo_stats tmp;
    flimo_get_stats(h, &tmp);
  }
  return FLIMO_OK;
}

}  // namespace

extern "C" {

const char* flimo_version(void) { return "fast_limo_b200 0.1 (sm_100a)"; }

void flimo_cfg_default(flimo_cfg* c) {
  std::memset(c, 0, sizeof(*c));
  c->NUM_MATCH_POINTS = 5;
  c->MAX_NUM_MATCHES = 2000;        // src/main.cpp:149
  c->MAX_NUM_PC2MATCH = 10000;      // Mapper.cpp:27
  c->estimate_extrinsics = 1;
  c->MAX_DIST_PLANE = 2.0;
  c->PLANE_THRESHOLD = 5.e-2;
  c->octree_bucket_size = 2;        // YAML value; ignored like the reference ignores it
  c->octree_downsampling = 1;
  c->octree_min_extent = 0.2f;
  c->knn_cell = 0.f;
  c->sort_scan = 0;
  c->knn_level_ratio = 0.f;
  c->knn_tau = 0;
}

const char* flimo_last_error(flimo_handle h) { return h ? h->err.c_str() : g_err.c_str(); }

int flimo_create(const flimo_cfg* cfg, int device, flimo_handle* out) {
  if (!cfg || !out) return fail(nullptr, FLIMO_ERR_INVALID, "null argument");
  *out = nullptr;
  if (cfg->NUM_MATCH_POINTS != 5)
    return fail(nullptr, FLIMO_ERR_INVALID, "NUM_MATCH_POINTS != 5 is not compiled in (all shipped configs use 5)");
  if (device == -1) {
    // Host-only handle: ONLY the flimo_ekf_* state machine works (pure host algebra, used by the
    // CPU unit tests and by drivers that run the measurement pass elsewhere).  Every entry point
    // that needs the GPU fails with FLIMO_ERR_NO_DEVICE; there is no CPU measurement path.
    flimo_ctx* hh = new (std::nothrow) flimo_ctx;
    if (!hh) return fail(nullptr, FLIMO_ERR_NOMEM, "host allocation failed");
    hh->cfg = *cfg;
    hh->device = -1;
    if (const char* e = std::getenv("FLIMO_EKF_REFERENCE_FORM")) hh->upd.reference_form_ = std::atoi(e) != 0;
    if (const char* e = std::getenv("FLIMO_EKF_COOP")) hh->use_coop = std::atoi(e) != 0;
    *out = hh;
    return FLIMO_OK;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(nullptr, FLIMO_ERR_NO_DEVICE, "no CUDA device: libflimo_cuda has no CPU path");
  if (device < 0 || device >= ndev) return fail(nullptr, FLIMO_ERR_INVALID, "device ordinal out of range");
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10)
    return fail(nullptr, FLIMO_ERR_NO_DEVICE, "device is not sm_100 class; this library ships sm_100a code only");
  flimo_ctx* h = new (std::nothrow) flimo_ctx;
  if (!h) return fail(nullptr, FLIMO_ERR_NOMEM, "host allocation failed");
  h->cfg = *cfg;
  h->device = device;
  if (const char* e = std::getenv("FLIMO_EKF_REFERENCE_FORM")) h->upd.reference_form_ = std::atoi(e) != 0;
  if (const char* e = std::getenv("FLIMO_EKF_COOP")) h->use_coop = std::atoi(e) != 0;
  // index ladder: finest cell, growth ratio, candidates-per-block threshold.  The FLIMO_KNN_* environment
  // variables override the configuration (tuning runs of tools/ only).
  if (const char* e = std::getenv("FLIMO_KNN_CELL")) h->cfg.knn_cell = (float)std::atof(e);
  if (const char* e = std::getenv("FLIMO_KNN_RATIO")) h->cfg.knn_level_ratio = (float)std::atof(e);
  if (const char* e = std::getenv("FLIMO_KNN_TAU")) h->cfg.knn_tau = std::atoi(e);
  if (!(h->cfg.knn_cell > 0.f)) h->cfg.knn_cell = 0.f;         // 0 = chosen from the map's density at every full index build (map_index_build)
  if (!(h->cfg.knn_level_ratio > 1.05f)) h->cfg.knn_level_ratio = kDefaultRatio;
  if (h->cfg.knn_tau <= 0) h->cfg.knn_tau = kDefaultTau;
  h->knn_tau = h->cfg.knn_tau;
  if (const char* e = std::getenv("FLIMO_KNN_PROBE")) h->probe_mode = std::atoi(e);
  if (const char* e = std::getenv("FLIMO_PROFILE")) h->prof = std::atoi(e) != 0;
  if (const char* e = std::getenv("FLIMO_TIME_EVERY")) h->time_every = std::atoi(e);
  if (const char* e = std::getenv("FLIMO_KNN_WIDE")) h->wide_loads = std::atoi(e);
  if (const char* e = std::getenv("FLIMO_KNN_PAIR")) h->pair_scan = std::atoi(e);
  if (const char* e = std::getenv("FLIMO_KNN_SORT")) h->cfg.sort_scan = std::atoi(e);
  h->interleave = h->cfg.sort_scan ? 1 : 0;
  if (const char* e = std::getenv("FLIMO_KNN_INTERLEAVE")) h->interleave = std::atoi(e);
  if (const char* e = std::getenv("FLIMO_KNN_PERM")) h->scan_perm = std::atoi(e);
  if (const char* e = std::getenv("FLIMO_KNN_PREFETCH")) h->l2_prefetch = std::atoi(e);
  if (const char* e = std::getenv("FLIMO_KNN_STAGE")) h->stage_runs = std::atoi(e);
  if (const char* e = std::getenv("FLIMO_KNN_FX")) h->fx_reduce = std::atoi(e);
  if (const char* e = std::getenv("FLIMO_INDEX_INCREMENTAL")) h->index_incremental = std::atoi(e);
  if (const char* e = std::getenv("FLIMO_DEBUG_STALL_EVERY")) h->debug_stall_every = std::atoi(e);
  if (const char* e = std::getenv("FLIMO_PDL")) h->pdl = std::atoi(e) != 0;
  const bool want_cta_trace = std::getenv("FLIMO_DEBUG_CTA_TRACE") != nullptr;
  CU(h, cudaSetDevice(device));
  CU(h, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CU(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  CU(h, cudaEventCreateWithFlags(&h->ev_prefetch, cudaEventDisableTiming));
  CU(h, cudaEventCreate(&h->ev0));
  CU(h, cudaEventCreate(&h->ev1));
  CU(h, cudaMalloc(&h->out96, 96 * sizeof(double)));
  CU(h, cudaMemset(h->out96, 0, 96 * sizeof(double)));
  CU(h, cudaHostAlloc(&h->h_out96, 2 * 96 * sizeof(double), cudaHostAllocMapped));
  std::memset(h->h_out96, 0, 2 * 96 * sizeof(double));
  CU(h, cudaHostGetDevicePointer(reinterpret_cast<void**>(&h->d_h_out96), h->h_out96, 0));
  CU(h, cudaMalloc(&h->d_count, sizeof(unsigned int)));
  CU(h, cudaHostAlloc(&h->h_ctl, sizeof(PassCtlWire), cudaHostAllocMapped));
  std::memset(h->h_ctl, 0, sizeof(PassCtlWire));
  CU(h, cudaHostGetDevicePointer(reinterpret_cast<void**>(&h->d_h_ctl), h->h_ctl, 0));
  if (want_cta_trace) {
    CU(h, cudaMalloc(&h->cta_trace, 4096 * sizeof(uint4)));
    CU(h, cudaMemset(h->cta_trace, 0, 4096 * sizeof(uint4)));
  }
  CU(h, cudaMalloc(&h->dev_ctl, sizeof(PassCtl)));
  CU(h, cudaMemset(h->dev_ctl, 0, sizeof(PassCtl)));
  h->persist_capacity = match_persistent_capacity();
  CU(h, preload_match_kernels());
  CU(h, preload_filter_kernel());
  h->reg_capacity = registration_capacity();
  CU(h, cudaMalloc(&h->upd_state, sizeof(ekf::UpdState)));
  CU(h, cudaMemset(h->upd_state, 0, sizeof(ekf::UpdState)));
  CU(h, cudaHostAlloc(&h->h_in, sizeof(ekf::UpdInit), cudaHostAllocMapped));
  CU(h, cudaHostGetDevicePointer(reinterpret_cast<void**>(&h->d_h_in), h->h_in, 0));
  CU(h, cudaMalloc(&h->dev_in, sizeof(ekf::UpdInit)));
  CU(h, cudaStreamCreateWithFlags(&h->filter_stream, cudaStreamNonBlocking));
  CU(h, cudaHostAlloc(&h->h_res, (size_t)kResRecords * 16, cudaHostAllocMapped));
  std::memset(h->h_res, 0, (size_t)kResRecords * 16);
  CU(h, cudaHostGetDevicePointer(reinterpret_cast<void**>(&h->d_h_res), h->h_res, 0));
  if (const char* e = std::getenv("FLIMO_DEVICE_EKF")) h->device_ekf = std::atoi(e);
  if (const char* e = std::getenv("FLIMO_WAIT_TIMEOUT_S")) h->wait_timeout_s = std::atof(e);
  if (const char* e = std::getenv("FLIMO_PERSISTENT")) h->persistent = std::atoi(e);
  if (const char* e = std::getenv("FLIMO_TEST_STALL_PASS")) h->test_stall_pass = std::atoi(e);
  *out = h;
  return FLIMO_OK;
}

void flimo_destroy(flimo_handle h) {
  if (!h) return;
  if (h->prof && h->prof_passes)
    std::fprintf(stderr, "[flimo profile] passes=%llu  per pass: launch %.2f us, wait %.2f us, filter step %.2f us\n",
                 (unsigned long long)h->prof_passes, h->prof_launch / h->prof_passes, h->prof_wait / h->prof_passes,
                 h->prof_step / h->prof_passes);
  if (h->prof && h->prof_adds)
    std::fprintf(stderr, "[flimo profile] index: %llu full builds, %llu incremental merges\n", (unsigned long long)h->stats_index_builds,
                 (unsigned long long)h->stats_index_updates);
  if (h->prof && h->prof_adds)
    std::fprintf(stderr, "[flimo profile] map adds=%llu  per add: pack+bbox %.1f us, insert rule %.1f us, index build %.1f us\n",
                 (unsigned long long)h->prof_adds, h->prof_add_pack / h->prof_adds, h->prof_add_insert / h->prof_adds,
                 h->prof_add_index / h->prof_adds);
  if (h->device < 0) {
    delete h;
    return;
  }
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  map_index_free(h->map);
  cudaFree(h->scan);
  cudaFree(h->scan_tmp);
  cudaFree(h->scan_keys);
  cudaFree(h->scan_cub);
  cudaFree(h->stage);
  cudaFree(h->d_count);
  cudaFree(h->batch);
  cudaFree(h->batch_keys);
  cudaFree(h->batch_accept);
  table_free(h->counts);
  cudaFree(h->partials);
  cudaFree(h->ticket);
  cudaFree(h->out96);
  for (auto& g : h->scan_graphs) cudaGraphExecDestroy(g.exec);
  if (h->xch_host) cudaHostUnregister(h->xch_host);
  cudaFreeHost(h->h_out96);
  prep_free(h->prep);
  cudaFreeHost(h->h_ctl);
  cudaFree(h->scan_stage[0]);
  cudaFree(h->scan_stage[1]);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->ev_prefetch) cudaEventDestroy(h->ev_prefetch);
  cudaFree(h->dev_ctl);
  cudaFree(h->cta_trace);
  cudaFree(h->upd_state);
  cudaFreeHost(h->h_in);
  cudaFree(h->dev_in);
  if (h->filter_stream) {
    cudaStreamSynchronize(h->filter_stream);
    cudaStreamDestroy(h->filter_stream);
  }
  cudaFreeHost(h->h_res);
  cudaFree(h->flag_words);
  for (int r = 0; r < h->peer_world; ++r)
    if (h->peer_inbox[r] && r != h->peer_rank) cudaIpcCloseMemHandle(h->peer_inbox[r]);
  cudaFree(h->inbox);
  cudaFree(h->dbg16);
  cudaFree(h->valid_flags);
  cudaFree(h->xyz_out);
  for (auto e : h->ev_pending) cudaEventDestroy(e);
  for (auto e : h->ev_pool) cudaEventDestroy(e);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

void* flimo_stream(flimo_handle h) { return h ? (void*)h->stream : nullptr; }

// Profiling aid (not part of flimo.h): per-warp timestamps of the next passes are written to a
// device buffer of 6 u64 per warp; returns it to the host.  Pass enable=0 to switch off.
int flimo_debug_timing(flimo_handle h, int enable, unsigned long long* host_out, size_t n_warps) {
  if (!h || h->device < 0) return FLIMO_ERR_INVALID;
  cudaSetDevice(h->device);
  if (enable && !h->timing) {
    if (cudaMalloc(&h->timing, (n_warps * 8 + 8) * sizeof(unsigned long long)) != cudaSuccess) return FLIMO_ERR_NOMEM;
    cudaMemset(h->timing, 0, (n_warps * 8 + 8) * sizeof(unsigned long long));
  }
  if (host_out && h->timing) {
    cudaStreamSynchronize(h->stream);
    cudaMemcpy(host_out, h->timing, (n_warps * 8 + 8) * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  }
  if (!enable && h->timing) {
    cudaFree(h->timing);
    h->timing = nullptr;
  }
  return FLIMO_OK;
}

int flimo_get_stats(flimo_handle h, flimo_stats* out) {
  if (!h || !out) return FLIMO_ERR_INVALID;
  for (size_t i = 0; i + 1 < h->ev_pending.size(); i += 2) {
    float ms = 0.f;
    if (cudaEventSynchronize(h->ev_pending[i + 1]) == cudaSuccess &&
        cudaEventElapsedTime(&ms, h->ev_pending[i], h->ev_pending[i + 1]) == cudaSuccess) {
      h->stats.match_ms_total += ms;
      h->stats.match_timed++;
      h->stats.last_match_ms = ms;
    }
    h->ev_pool.push_back(h->ev_pending[i]);
    h->ev_pool.push_back(h->ev_pending[i + 1]);
  }
  h->ev_pending.clear();
  h->stats.knn_cell = h->map.lv[0].g.cell;
  h->stats.grid_nx = h->map.lv[0].g.nx;
  h->stats.grid_ny = h->map.lv[0].g.ny;
  h->stats.grid_nz = h->map.lv[0].g.nz;
  h->stats.n_levels = h->map.n_levels;
  h->stats.table_bytes = 0;
  h->stats.map_bytes = h->map.n_pts * sizeof(float4);
  for (int l = 0; l < h->map.n_levels; ++l) {
    h->stats.table_bytes += (h->map.lv[l].cap_slots + 2 * h->map.lv[l].cap_rows) * sizeof(uint32_t);
    h->stats.map_bytes += h->map.lv[l].cap_entries * sizeof(float4);   // super-row entries (16 bytes each) incl. the rows' head-room
  }
  h->stats.persist_ms_total = h->persist_ns_total * 1e-6;
  h->stats.exchange_ms_total = h->exchange_ns_total * 1e-6;
  h->stats.persist_passes = h->persist_passes;
  h->stats.index_builds = h->stats_index_builds;
  h->stats.index_updates = h->stats_index_updates;
  h->stats.index_rows_moved = h->map.rows_moved;
  h->stats.update_stalls = h->stats_update_stalls;
  *out = h->stats;
  return FLIMO_OK;
}

// ------------------------------------------------------------------------------------------------
int flimo_map_add_device(flimo_handle h, const void* d_xyz, size_t n, size_t stride_bytes, double stamp) {
  if (!h || (!d_xyz && n)) return fail(h, FLIMO_ERR_INVALID, "null argument");
  if (stride_bytes < 12 || stride_bytes % 4) return fail(h, FLIMO_ERR_INVALID, "stride must be a multiple of 4 and >= 12");
  if (n < 1) return FLIMO_OK;                                   // Mapper::add: size < 1 -> return
  NEED_GPU(h);
  const size_t old_n = h->map_exists ? h->map.n_pts : 0;
  const auto ta0 = std::chrono::steady_clock::now();
  // 1. pack the batch (drops NaN points, Octree::processPoints :243) and take its bounding box
  if (n > h->batch_cap) {
    cudaFree(h->batch);
    cudaFree(h->batch_keys);
    cudaFree(h->batch_accept);
    h->batch = nullptr;
    h->batch_keys = nullptr;
    h->batch_accept = nullptr;
    h->batch_cap = 0;
    const size_t want = n + n / 4 + 1024;
    CU(h, cudaMalloc(&h->batch, want * sizeof(float4)));
    CU(h, cudaMalloc(&h->batch_keys, want * sizeof(unsigned long long)));
    CU(h, cudaMalloc(&h->batch_accept, want));
    h->batch_cap = want;
  }
  CU(h, cudaMemsetAsync(h->d_count, 0, sizeof(unsigned int), h->stream));
  CU(h, pack_points(d_xyz, n, stride_bytes, h->batch, h->d_count, h->stream));
  unsigned int kept = 0;
  CU(h, cudaMemcpyAsync(&kept, h->d_count, sizeof(kept), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  h->stats.kernel_launches += 1;
  if (kept == 0) return FLIMO_OK;                               // Octree::initialize: empty -> no root
  const auto tb0 = std::chrono::steady_clock::now();
  CU(h, map_index_reserve(h->map, old_n + kept));
  const auto tb1 = std::chrono::steady_clock::now();
  float lo[3], hi[3];
  CU(h, points_bbox(h->batch, kept, h->map.bbox, lo, hi, h->stream));
  if (h->prof) {
    const auto tb2 = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[add] n=%zu pack %.0f us reserve %.0f us bbox %.0f us\n", n,
                 std::chrono::duration<double, std::micro>(tb0 - ta0).count(), std::chrono::duration<double, std::micro>(tb1 - tb0).count(),
                 std::chrono::duration<double, std::micro>(tb2 - tb1).count());
  }
  h->stats.kernel_launches += 2;
  // 2. the octree lattice: first batch defines it (Octree::initialize), later ones may double the root
  const bool first = !h->map_exists;
  if (first) {
    lattice_init(h->lattice, lo, hi, h->cfg.octree_min_extent);
    table_free(h->counts);
  } else {
    lattice_grow(h->lattice, hi);                               // expandTree(max) then expandTree(min), Octree.hpp:373-374
    lattice_grow(h->lattice, lo);
  }
  // 3. accept / drop per point, append the accepted ones to the canonical list, update the counts
  const auto ta1 = std::chrono::steady_clock::now();
  unsigned int accepted = 0;
  CU(h, map_insert_batch(h->lattice, h->counts, h->batch, kept, old_n, h->cfg.octree_downsampling, first, h->map.pts + old_n,
                         h->d_count, h->batch_keys, h->batch_accept, &accepted, h->stream, &h->stats.kernel_launches));
  const size_t total = old_n + accepted;
  h->map_exists = true;                                         // the tree exists even if this batch was dropped entirely
  h->last_time = stamp;
  if (accepted == 0) return FLIMO_OK;
  // 4. rebuild the search index over all map points
  const auto ta2 = std::chrono::steady_clock::now();
  h->map.n_pts = total;
  const size_t max_cells = (size_t)1 << 30;
  // coarsest level: cell >= sqrt(MAX_DIST_PLANE) so its 3x3x3 block covers the close_enough radius
  const float coarsest = (float)(std::sqrt(std::max(h->cfg.MAX_DIST_PLANE, 1e-6)) * 1.0001 + 1e-4);
  if (h->index_incremental && !first && map_index_can_update(h->map, old_n, lo, hi)) {
    // merge the accepted points into every level (the arrays end up identical to a rebuild)
    bool full = false;
    CU(h, map_index_update(h->map, old_n, h->stream, &h->stats.kernel_launches, &full));
    for (int a = 0; a < 3; ++a) {
      h->map.lo[a] = std::min(h->map.lo[a], lo[a]);
      h->map.hi[a] = std::max(h->map.hi[a], hi[a]);
    }
    h->stats_index_updates++;
    if (full) {                                                  // no room left behind a level's segments: rebuild (compacts)
      CU(h, map_index_build(h->map, h->cfg.knn_cell, h->cfg.knn_level_ratio, coarsest, max_cells, h->stream,
                            &h->stats.kernel_launches));
      h->stats_index_builds++;
    }
  } else {
    CU(h, map_index_build(h->map, h->cfg.knn_cell, h->cfg.knn_level_ratio, coarsest, max_cells, h->stream,
                          &h->stats.kernel_launches));
    h->stats_index_builds++;
  }
  CU(h, cudaStreamSynchronize(h->stream));
  if (h->prof) {
    const auto ta3 = std::chrono::steady_clock::now();
    h->prof_add_pack += std::chrono::duration<double, std::micro>(ta1 - ta0).count();
    h->prof_add_insert += std::chrono::duration<double, std::micro>(ta2 - ta1).count();
    h->prof_add_index += std::chrono::duration<double, std::micro>(ta3 - ta2).count();
    h->prof_adds++;
    std::fprintf(stderr, "[add] total %zu: insert %.0f us, index %.0f us\n", total, std::chrono::duration<double, std::micro>(ta2 - ta1).count(),
                 std::chrono::duration<double, std::micro>(ta3 - ta2).count());
  }
  return FLIMO_OK;
}

int flimo_map_add(flimo_handle h, const float* xyz, size_t n, size_t stride_bytes, double stamp) {
  if (!h || (!xyz && n)) return fail(h, FLIMO_ERR_INVALID, "null argument");
  if (n < 1) return FLIMO_OK;
  NEED_GPU(h);
  int rc = upload(h, xyz, n * stride_bytes);
  if (rc) return rc;
  return flimo_map_add_device(h, h->stage, n, stride_bytes, stamp);
}

int flimo_map_size(flimo_handle h, size_t* n_points) {
  if (!h || !n_points) return FLIMO_ERR_INVALID;
  *n_points = h->map_exists ? h->map.n_pts : 0;
  return FLIMO_OK;
}
int flimo_map_exists(flimo_handle h) { return (h && h->map_exists && h->map.n_pts > 0) ? 1 : 0; }
double flimo_map_last_time(flimo_handle h) { return h ? h->last_time : -1.0; }

int flimo_map_get_points(flimo_handle h, float* out_xyz, size_t cap_points, size_t* n_points) {
  if (!h || !n_points) return fail(h, FLIMO_ERR_INVALID, "null argument");
  const size_t n = h->map_exists ? h->map.n_pts : 0;
  *n_points = n;
  if (!out_xyz || n == 0) return FLIMO_OK;
  NEED_GPU(h);
  const size_t m = n < cap_points ? n : cap_points;
  std::vector<float4> tmp(m);
  CU(h, cudaMemcpyAsync(tmp.data(), h->map.pts, m * sizeof(float4), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  for (size_t i = 0; i < m; ++i) {
    out_xyz[3 * i] = tmp[i].x;
    out_xyz[3 * i + 1] = tmp[i].y;
    out_xyz[3 * i + 2] = tmp[i].z;
  }
  return FLIMO_OK;
}

// ------------------------------------------------------------------------------------------------
// modular inverse of a (mod n), gcd(a, n) = 1
static uint32_t mod_inverse(uint32_t a, uint32_t n) {
  long long t = 0, nt = 1, r = n, nr = a % n;
  while (nr != 0) {
    const long long q = r / nr;
    const long long tt = t - q * nt; t = nt; nt = tt;
    const long long rr = r - q * nr; r = nr; nr = rr;
  }
  if (t < 0) t += n;
  return (uint32_t)t;
}

// Builds the packed copy h->scan of the bound scan (pack [+ Morton sort]); needed by sort_scan and by the
// helpers that walk the scan in storage order.  The pipeline is captured once per (source, size) into a
// CUDA graph and replayed with a single launch afterwards.
static int pack_bound_scan(flimo_handle h) {
  const void* d_xyz = h->raw_scan;
  const size_t nq = h->scan_n, stride_bytes = h->raw_stride;
  if (nq == 0 || h->packed_valid) return FLIMO_OK;
  if (nq > h->scan_cap) {
    cudaFree(h->scan);
    cudaFree(h->scan_tmp);
    h->scan = h->scan_tmp = nullptr;
    h->scan_cap = 0;
    const size_t want = nq + nq / 4 + 1024;
    CU(h, cudaMalloc(&h->scan, want * sizeof(float4)));
    CU(h, cudaMalloc(&h->scan_tmp, want * sizeof(float4)));
    h->scan_cap = want;
    for (auto& g : h->scan_graphs) cudaGraphExecDestroy(g.exec);   // captured pointers are stale
    h->scan_graphs.clear();
  }
  const int sort = (h->cfg.sort_scan && nq > 1) ? 1 : 0;
  const unsigned int perm = (sort || !h->scan_perm) ? 0u : coprime_stride((uint32_t)nq);
  h->packed_valid = true;
  if (!sort) {                                                   // one kernel: launch it directly
    CU(h, scan_prepare(d_xyz, nq, stride_bytes, false, perm, h->scan, h->scan_tmp, &h->scan_cub, &h->scan_cub_bytes, &h->scan_keys,
                       &h->scan_keys_cap, h->stream, &h->stats.kernel_launches));
    return FLIMO_OK;
  }
  for (auto& g : h->scan_graphs) {
    if (g.src == d_xyz && g.n == nq && g.stride == stride_bytes && g.sort == sort) {
      CU(h, cudaGraphLaunch(g.exec, h->stream));
      h->stats.kernel_launches += sort ? 7 : 1;
      return FLIMO_OK;
    }
  }
  const size_t old_keys_cap = h->scan_keys_cap;
  if (sort) CU(h, scan_prepare_reserve(nq, &h->scan_cub, &h->scan_cub_bytes, &h->scan_keys, &h->scan_keys_cap));
  if (h->scan_keys_cap != old_keys_cap) {                        // scratch moved: drop graphs that point at the old one
    for (auto& g : h->scan_graphs) cudaGraphExecDestroy(g.exec);
    h->scan_graphs.clear();
  }
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  uint64_t dummy = 0;
  CU(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
  const cudaError_t ce = scan_prepare(d_xyz, nq, stride_bytes, sort != 0, perm, h->scan, h->scan_tmp, &h->scan_cub, &h->scan_cub_bytes,
                                      &h->scan_keys, &h->scan_keys_cap, h->stream, &dummy);
  const cudaError_t ee = cudaStreamEndCapture(h->stream, &graph);
  if (ce != cudaSuccess || ee != cudaSuccess || graph == nullptr) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    // capture unavailable: run the pipeline directly
    CU(h, scan_prepare(d_xyz, nq, stride_bytes, sort != 0, perm, h->scan, h->scan_tmp, &h->scan_cub, &h->scan_cub_bytes, &h->scan_keys,
                       &h->scan_keys_cap, h->stream, &h->stats.kernel_launches));
    return FLIMO_OK;
  }
  CU(h, cudaGraphInstantiate(&exec, graph, 0));
  cudaGraphDestroy(graph);
  if (h->scan_graphs.size() >= 8) {
    cudaGraphExecDestroy(h->scan_graphs.front().exec);
    h->scan_graphs.erase(h->scan_graphs.begin());
  }
  h->scan_graphs.push_back({d_xyz, nq, stride_bytes, sort, exec});
  CU(h, cudaGraphLaunch(exec, h->stream));
  h->stats.kernel_launches += sort ? 7 : 1;
  return FLIMO_OK;
}

int flimo_scan_set_device(flimo_handle h, const void* d_xyz, size_t n, size_t stride_bytes) {
  if (!h || (!d_xyz && n)) return fail(h, FLIMO_ERR_INVALID, "null argument");
  if (stride_bytes < 12 || stride_bytes % 4) return fail(h, FLIMO_ERR_INVALID, "stride must be a multiple of 4 and >= 12");
  NEED_GPU(h);
  const size_t cap = h->cfg.MAX_NUM_PC2MATCH > 0 ? (size_t)h->cfg.MAX_NUM_PC2MATCH : 0;
  const size_t nq = n > cap ? cap : n;                           // first-N rule (Mapper.cpp:63-69)
  // The measurement kernel reads the caller's array IN PLACE (no copy, no launch): stored position j of
  // its pseudo-random query order is original point (j * raw_inv) mod n.  The array must stay valid and
  // unchanged until the next scan is bound.
  h->raw_scan = static_cast<const unsigned char*>(d_xyz);
  h->raw_stride = stride_bytes;
  h->scan_n = nq;
  h->scan_n_full = n;
  h->shard_begin = 0;
  h->shard_end = nq;
  h->packed_valid = false;
  h->raw_inv = 1;
  if (nq == 0) return FLIMO_OK;
  if (h->cfg.sort_scan && nq > 1) return pack_bound_scan(h);     // Morton order needs the packed copy
  if (h->scan_perm && nq > 2) h->raw_inv = mod_inverse(coprime_stride((uint32_t)nq), (uint32_t)nq);
  return FLIMO_OK;
}

static int stage_reserve(flimo_handle h, int idx, size_t bytes) {
  if (bytes <= h->scan_stage_cap[idx]) return FLIMO_OK;
  if (h->scan_stage[idx]) cudaFree(h->scan_stage[idx]);
  h->scan_stage[idx] = nullptr;
  h->scan_stage_cap[idx] = 0;
  CU(h, cudaMalloc(&h->scan_stage[idx], bytes + bytes / 4 + 4096));
  h->scan_stage_cap[idx] = bytes + bytes / 4 + 4096;
  return FLIMO_OK;
}

int flimo_scan_prefetch(flimo_handle h, const float* xyz_body, size_t n, size_t stride_bytes) {
  if (!h || (!xyz_body && n)) return fail(h, FLIMO_ERR_INVALID, "null argument");
  if (stride_bytes < 12 || stride_bytes % 4) return fail(h, FLIMO_ERR_INVALID, "stride must be a multiple of 4 and >= 12");
  NEED_GPU(h);
  const size_t nq = n;                                            // the whole cloud travels: Mapper::add takes all of it
  h->pref_idx = -1;
  h->pref_issued = false;
  if (nq == 0) return FLIMO_OK;
  const int idx = h->scan_stage_bound ^ 1;                        // the staging buffer the bound scan does not use
  int rc = stage_reserve(h, idx, nq * stride_bytes);
  if (rc) return rc;
  h->pref_src = xyz_body;
  h->pref_n = nq;
  h->pref_stride = stride_bytes;
  h->pref_idx = idx;
  return FLIMO_OK;
}

int flimo_scan_set(flimo_handle h, const float* xyz_body, size_t n, size_t stride_bytes) {
  if (!h || (!xyz_body && n)) return fail(h, FLIMO_ERR_INVALID, "null argument");
  if (stride_bytes < 12 || stride_bytes % 4) return fail(h, FLIMO_ERR_INVALID, "stride must be a multiple of 4 and >= 12");
  NEED_GPU(h);
  const size_t nq = n;           // the first-N rule (Mapper.cpp:63-69) is applied by flimo_scan_set_device; the map gets all points
  if (nq == 0) return flimo_scan_set_device(h, h->scan_stage[0], 0, stride_bytes);
  int idx;
  if (h->pref_idx >= 0 && h->pref_src == xyz_body && h->pref_n == nq && h->pref_stride == stride_bytes) {
    idx = h->pref_idx;                                            // already on its way: order the passes after the copy
    int rc = issue_prefetch(h);
    if (rc) return rc;
    CU(h, cudaStreamWaitEvent(h->stream, h->ev_prefetch, 0));
  } else {
    idx = h->scan_stage_bound ^ 1;
    if (h->pref_idx >= 0 && h->pref_issued)                      // an unrelated prefetch is filling this buffer: let it finish first
      CU(h, cudaStreamWaitEvent(h->stream, h->ev_prefetch, 0));
    int rc = stage_reserve(h, idx, nq * stride_bytes);
    if (rc) return rc;
    CU(h, cudaMemcpyAsync(h->scan_stage[idx], xyz_body, nq * stride_bytes, cudaMemcpyHostToDevice, h->stream));
  }
  h->pref_idx = -1;
  h->scan_stage_bound = idx;
  return flimo_scan_set_device(h, h->scan_stage[idx], nq, stride_bytes);
}

int flimo_scan_shard(flimo_handle h, size_t begin, size_t end) {
  if (!h) return FLIMO_ERR_INVALID;
  if (begin > end || end > h->scan_n) return fail(h, FLIMO_ERR_INVALID, "shard out of range");
  h->shard_begin = begin;
  h->shard_end = end;
  return FLIMO_OK;
}

int flimo_exchange_attach(flimo_handle h, void* shared_host_mem, size_t bytes, int rank, int world) {
  if (!h || !shared_host_mem || world < 1 || rank < 0 || rank >= world) return fail(h, FLIMO_ERR_INVALID, "bad exchange arguments");
  if (bytes < (size_t)world * FLIMO_EXCHANGE_BYTES_PER_RANK) return fail(h, FLIMO_ERR_INVALID, "exchange segment too small");
  NEED_GPU(h);
  if (h->xch_host) {
    cudaHostUnregister(h->xch_host);
    h->xch_host = nullptr;
  }
  CU(h, cudaHostRegister(shared_host_mem, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable));
  CU(h, cudaHostGetDevicePointer(reinterpret_cast<void**>(&h->xch_dev), shared_host_mem, 0));
  h->xch_host = static_cast<double*>(shared_host_mem);
  h->xch_rank = rank;
  h->xch_world = world;
  h->xch_seq = 0;
  return FLIMO_OK;
}

int flimo_match_reduce_exchange(flimo_handle h, const double state14[14], double HTH[144], double HTh[12], int64_t* n_valid,
                                int64_t* n_rows, double* sum_sq_res) {
  if (!h || !state14 || !HTH || !HTh) return fail(h, FLIMO_ERR_INVALID, "null argument");
  NEED_GPU(h);
  if (!h->xch_host) return fail(h, FLIMO_ERR_STATE, "flimo_exchange_attach not called");
  const unsigned long long seq = ++h->xch_seq;
  // two buffers per rank, alternating by sequence parity: a rank can be at most one pass ahead of a peer
  // that is still reading (it needs that peer's data of the current pass to advance)
  const size_t slot_doubles = FLIMO_EXCHANGE_BYTES_PER_RANK / sizeof(double);      // 512
  const size_t buf_doubles = slot_doubles / 2;                                      // 256 (96 records of 16 bytes + pad)
  const size_t my_off = (size_t)h->xch_rank * slot_doubles + (seq & 1) * buf_doubles;
  if (flimo_map_exists(h) && h->shard_end > h->shard_begin) {
    MatchParams P;
    int rc = fill_params(h, state14, P, h->out96, 0xFFFFFFFFu, nullptr, nullptr);
    if (rc) return rc;
    P.host_out96 = h->xch_dev + my_off;
    P.seq = seq;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (h->ev_pool.size() >= 2) {
      e0 = h->ev_pool.back(); h->ev_pool.pop_back();
      e1 = h->ev_pool.back(); h->ev_pool.pop_back();
    } else {
      CU(h, cudaEventCreate(&e0));
      CU(h, cudaEventCreate(&e1));
    }
    CU(h, cudaEventRecord(e0, h->stream));
    CU(h, launch_match(P, h->stream));
    CU(h, cudaEventRecord(e1, h->stream));
    h->ev_pending.push_back(e0);
    h->ev_pending.push_back(e1);
    h->stats.kernel_launches++;
    h->stats.match_launches++;
  } else {                                         // nothing to match on this rank: publish zeros
    volatile unsigned long long* mine = reinterpret_cast<volatile unsigned long long*>(h->xch_host + my_off);
    for (int i = 0; i < 96; ++i) mine[2 * i] = 0ull;
    std::atomic_thread_fence(std::memory_order_release);
    for (int i = 0; i < 96; ++i) mine[2 * i + 1] = seq;
  }
  double packed[96];
  for (int i = 0; i < 96; ++i) packed[i] = 0.0;
  for (int r = 0; r < h->xch_world; ++r) {          // fixed rank order => identical sums on every rank
    const double* slot = h->xch_host + (size_t)r * slot_doubles + (seq & 1) * buf_doubles;
    double vals[96];
    const int wr = wait_records(h, slot, seq, vals, r == h->xch_rank);
    if (wr < 0) return wr;
    if (wr == 1) return fail(h, FLIMO_ERR_CUDA, "match kernel finished without publishing its result");
    for (int i = 0; i < 96; ++i) packed[i] += vals[i];
  }
  if (h->ev_pending.size() > 64) {
    flimo_stats tmp;
    flimo_get_stats(h, &tmp);
  }
  if ((int64_t)std::llround(packed[92]) > (int64_t)h->cfg.MAX_NUM_MATCHES)
    return fail(h, FLIMO_ERR_STATE, "more accepted matches than MAX_NUM_MATCHES: the host-segment exchange cannot apply the first-N rule across shards (use flimo_update_peer)");
  flimo_unpack96(packed, HTH, HTh, n_valid, n_rows, sum_sq_res);
  return FLIMO_OK;
}

void flimo_unpack96(const double p[96], double HTH[144], double HTh[12], int64_t* n_valid, int64_t* n_rows,
                    double* sum_sq_res) {
  int e = 0;
  for (int i = 0; i < 12; ++i)
    for (int j = i; j < 12; ++j) {
      HTH[i * 12 + j] = p[e];
      HTH[j * 12 + i] = p[e];
      ++e;
    }
  for (int i = 0; i < 12; ++i) HTh[i] = p[78 + i];
  if (n_rows) *n_rows = (int64_t)std::llround(p[90]);
  if (sum_sq_res) *sum_sq_res = p[91];
  if (n_valid) *n_valid = (int64_t)std::llround(p[92]);
}

int flimo_match_reduce_async(flimo_handle h, const double state14[14], double* d_out96, void* cuda_stream) {
  if (!h || !state14 || !d_out96) return fail(h, FLIMO_ERR_INVALID, "null argument");
  NEED_GPU(h);
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
  if (!flimo_map_exists(h) || h->shard_end == h->shard_begin) {   // Mapper::match on an empty map -> no matches
    CU(h, cudaMemsetAsync(d_out96, 0, 96 * sizeof(double), st));
    return FLIMO_OK;
  }
  MatchParams P;
  int rc = fill_params(h, state14, P, d_out96, 0xFFFFFFFFu, nullptr, nullptr);
  if (rc) return rc;
  // device-time bookkeeping for bench.py: event pairs are resolved lazily in flimo_get_stats
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (h->ev_pool.size() >= 2) {
    e0 = h->ev_pool.back(); h->ev_pool.pop_back();
    e1 = h->ev_pool.back(); h->ev_pool.pop_back();
  } else {
    CU(h, cudaEventCreate(&e0));
    CU(h, cudaEventCreate(&e1));
  }
  CU(h, cudaEventRecord(e0, st));
  CU(h, launch_match(P, st));
  CU(h, cudaEventRecord(e1, st));
  h->ev_pending.push_back(e0);
  h->ev_pending.push_back(e1);
  h->stats.kernel_launches++;
  h->stats.match_launches++;
  return FLIMO_OK;
}

int flimo_match_reduce(flimo_handle h, const double state14[14], double HTH[144], double HTh[12], int64_t* n_valid,
                       int64_t* n_rows, double* sum_sq_res) {
  if (!h || !state14 || !HTH || !HTh) return fail(h, FLIMO_ERR_INVALID, "null argument");
  NEED_GPU(h);
  double packed[96];
  std::memset(packed, 0, sizeof(packed));
  if (flimo_map_exists(h) && h->shard_end > h->shard_begin) {
    int rc = run_pass_blocking(h, state14, 0xFFFFFFFFu, nullptr, nullptr, packed);
    if (rc) return rc;
    const int64_t nv = (int64_t)std::llround(packed[92]);
    const int64_t cap = h->cfg.MAX_NUM_MATCHES;
    if (nv > cap) {
      // Localizer::calculate_H keeps the FIRST MAX_NUM_MATCHES accepted matches in scan order
      // (Localizer.cpp:539,547-548).  Find the original index after which rows stop counting.
      const size_t nq = h->scan_n;
      CU(h, grow(&h->valid_flags, &h->valid_cap, nq));
      CU(h, cudaMemsetAsync(h->valid_flags, 0, nq, h->stream));
      rc = run_pass_blocking(h, state14, 0xFFFFFFFFu, nullptr, h->valid_flags, packed);
      if (rc) return rc;
      std::vector<uint8_t> flags(nq);
      CU(h, cudaMemcpyAsync(flags.data(), h->valid_flags, nq, cudaMemcpyDeviceToHost, h->stream));
      CU(h, cudaStreamSynchronize(h->stream));
      int64_t seen = 0;
      uint32_t limit = (uint32_t)nq;
      for (size_t i = 0; i < nq; ++i) {
        if (flags[i] && ++seen == cap) {
          limit = (uint32_t)(i + 1);
          break;
        }
      }
      if (cap <= 0) limit = 0;
      rc = run_pass_blocking(h, state14, limit, nullptr, nullptr, packed);
      if (rc) return rc;
    }
  }
  flimo_unpack96(packed, HTH, HTh, n_valid, n_rows, sum_sq_res);
  return FLIMO_OK;
}

int flimo_match_debug(flimo_handle h, const double state14[14], float* out16, size_t cap_points, size_t* n_points) {
  if (!h || !state14 || !n_points) return fail(h, FLIMO_ERR_INVALID, "null argument");
  NEED_GPU(h);
  const size_t nq = h->scan_n;
  *n_points = nq;
  if (!out16 || nq == 0) return FLIMO_OK;
  CU(h, grow(&h->dbg16, &h->dbg_cap, nq * 16));
  CU(h, cudaMemsetAsync(h->dbg16, 0, nq * 16 * sizeof(float), h->stream));
  if (flimo_map_exists(h) && h->shard_end > h->shard_begin) {
    double packed[96];
    int rc = run_pass_blocking(h, state14, 0xFFFFFFFFu, h->dbg16, nullptr, packed);
    if (rc) return rc;
  }
  const size_t m = nq < cap_points ? nq : cap_points;
  CU(h, cudaMemcpyAsync(out16, h->dbg16, m * 16 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return FLIMO_OK;
}

int flimo_scan_to_world(flimo_handle h, const double state14[14], float* out_xyz, size_t cap_points, size_t* n_points) {
  if (!h || !state14 || !n_points) return fail(h, FLIMO_ERR_INVALID, "null argument");
  NEED_GPU(h);
  const size_t nq = h->scan_n_full;                              // the whole pc2match, not only the matched first N (Localizer.cpp:361)
  *n_points = nq;
  if (!out_xyz || nq == 0) return FLIMO_OK;
  CU(h, grow(&h->xyz_out, &h->xyz_cap, nq * 3));
  PoseConsts pc;
  make_pose(state14, pc);
  CU(h, transform_raw(h->raw_scan, nq, h->raw_stride, pc, h->xyz_out, h->stream));
  h->stats.kernel_launches++;
  const size_t m = nq < cap_points ? nq : cap_points;
  CU(h, cudaMemcpyAsync(out_xyz, h->xyz_out, m * 3 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return FLIMO_OK;
}

int flimo_map_add_scan(flimo_handle h, const double state14[14], double stamp) {
  if (!h || !state14) return fail(h, FLIMO_ERR_INVALID, "null argument");
  NEED_GPU(h);
  const size_t nq = h->scan_n_full;
  if (nq < 1) return FLIMO_OK;
  CU(h, grow(&h->xyz_out, &h->xyz_cap, nq * 3));
  PoseConsts pc;
  make_pose(state14, pc);
  CU(h, transform_raw(h->raw_scan, nq, h->raw_stride, pc, h->xyz_out, h->stream));
  h->stats.kernel_launches++;
  return flimo_map_add_device(h, h->xyz_out, nq, 12, stamp);
}

// ------------------------------------------------------------------------------------------------
int flimo_ekf_begin(flimo_handle h, const double state26[26], const double P529[529], int max_iter,
                    const double limit23[23], double R_noise, double D_degeneracy) {
  if (!h || !state26 || !P529 || !limit23) return fail(h, FLIMO_ERR_INVALID, "null argument");
  if (h->use_coop) h->coop.begin(state26, P529, max_iter, limit23, R_noise, D_degeneracy);
  else h->upd.begin(state26, P529, max_iter, limit23, R_noise, D_degeneracy);
  h->upd_active = true;
  return FLIMO_OK;
}
int flimo_ekf_state(flimo_handle h, double state26[26]) {
  if (!h || !state26 || !h->upd_active) return fail(h, FLIMO_ERR_STATE, "flimo_ekf_begin not called");
  if (h->use_coop) h->coop.state(state26);
  else h->upd.state(state26);
  return FLIMO_OK;
}
int flimo_ekf_step(flimo_handle h, const double HTH[144], const double HTh[12], int64_t n_rows, int* done) {
  if (!h || !HTH || !HTh || !h->upd_active) return fail(h, FLIMO_ERR_STATE, "flimo_ekf_begin not called");
  const bool d = h->use_coop ? h->coop.step(HTH, HTh, n_rows) : h->upd.step(HTH, HTh, n_rows);
  if (done) *done = d ? 1 : 0;
  if (h->use_coop ? h->coop.failed() : h->upd.failed())
    return fail(h, FLIMO_ERR_STATE, "singular or non-finite normal equations: state left at the prediction");
  return FLIMO_OK;
}
int flimo_ekf_end(flimo_handle h, double state26[26], double P529[529]) {
  if (!h || !state26 || !P529 || !h->upd_active) return fail(h, FLIMO_ERR_STATE, "flimo_ekf_begin not called");
  if (h->use_coop) h->coop.end(state26, P529);
  else h->upd.end(state26, P529);
  h->upd_active = false;
  return FLIMO_OK;
}

// ------------------------------------------------------------------------------------------------
// IMU rate (imu_host.hpp): host algebra only, usable on a handle without a device.
static_assert(sizeof(flimo_frame) == sizeof(ekf::PropagatedState), "flimo_frame mirrors ekf::PropagatedState");

int flimo_ekf_predict(flimo_handle h, double state26[26], double P529[529], const flimo_imu* imu,
                      const double cov4[4]) {
  if (!h || !state26 || !P529 || !imu || !cov4) return fail(h, FLIMO_ERR_INVALID, "null argument");
  ekf::State x;
  x.load(state26);
  ekf::Mat<ekf::N, ekf::N> P;
  std::memcpy(P.a, P529, sizeof(P.a));
  const ekf::V3 acc = {imu->lin_accel[0], imu->lin_accel[1], imu->lin_accel[2]};   // .cast<double>() (Localizer.cpp:585-586)
  const ekf::V3 gyro = {imu->ang_vel[0], imu->ang_vel[1], imu->ang_vel[2]};
  ekf::predict(x, P, acc, gyro, imu->dt, ekf::ProcessNoise{cov4[0], cov4[1], cov4[2], cov4[3]});
  x.store(state26);
  std::memcpy(P529, P.a, sizeof(P.a));
  h->propagated.push_front(ekf::make_propagated(x, imu->stamp, imu->lin_accel, imu->ang_vel));
  return FLIMO_OK;
}

int flimo_propagated_frames(flimo_handle h, double start_time, double end_time, flimo_frame* out, size_t cap,
                            size_t* n_frames) {
  if (!h || !n_frames || (cap && !out)) return fail(h, FLIMO_ERR_INVALID, "null argument");
  *n_frames = 0;
  const long n = h->propagated.frames(start_time, end_time, h->frames_tmp);
  if (n < 0) return fail(h, FLIMO_ERR_STATE, "propagated states end before end_time (IMU behind the scan)");
  *n_frames = static_cast<size_t>(n);
  if (cap) {
    if (cap < static_cast<size_t>(n)) return fail(h, FLIMO_ERR_INVALID, "frame buffer too small");
    std::memcpy(out, h->frames_tmp.data(), static_cast<size_t>(n) * sizeof(flimo_frame));
  }
  return FLIMO_OK;
}

int flimo_propagated_clear(flimo_handle h) {
  if (!h) return fail(h, FLIMO_ERR_INVALID, "null handle");
  h->propagated.clear();
  return FLIMO_OK;
}

// Hands the persistent kernel its next command through the mapped control block.
static void post_ctl(flimo_handle h, uint32_t cmd, uint32_t orig_limit, const double* state14) {
  const unsigned long long seq = ++h->ctl_seq;         // commands carry their own numbering
  uint32_t w[3 * kCtlRecords];
  std::memset(w, 0, sizeof(w));
  w[0] = cmd;
  w[1] = orig_limit;
  if (state14) {
    PoseConsts pc;
    make_pose(state14, pc);
    std::memcpy(&w[2], &pc, sizeof(pc));
  }
  for (int r = 0; r < kCtlRecords; ++r) {                // one 16-byte store per record
    const __m128i v = _mm_set_epi32((int)(uint32_t)seq, (int)w[3 * r + 2], (int)w[3 * r + 1], (int)w[3 * r]);
    _mm_store_si128(reinterpret_cast<__m128i*>(&h->h_ctl->rec[r][0]), v);
  }
}

// flimo_update with ONE kernel launch for all passes (match_persistent_kernel).  Returns 1 when the caller
// has to finish the update with per-pass launches (kernel ended early, or the MAX_NUM_MATCHES truncation
// needs the general path); `u` is then positioned at the pass that has to be (re)done.
static int update_persistent(flimo_handle h, ekf::IteratedUpdate& u, bool exchange) {
  const size_t n = h->shard_end - h->shard_begin;
  const int tiles = match_num_tiles((int)n);
  const size_t slot_doubles = FLIMO_EXCHANGE_BYTES_PER_RANK / sizeof(double), buf_doubles = slot_doubles / 2;
  unsigned long long& counter = exchange ? h->xch_seq : h->seq;
  double x[26], HTH[144], HTh[12], packed[96];
  u.state(x);
  MatchParams P;
  int rc = fill_params(h, x, P, h->out96, 0xFFFFFFFFu, nullptr, nullptr);
  if (rc) return rc;
  if (exchange) {       // results go straight into this rank's slot of the shared segment, two blocks by sequence parity
    P.host_out96 = h->xch_dev + (size_t)h->xch_rank * slot_doubles;
    P.host_out_alt = (unsigned int)buf_doubles;
  } else {
    P.host_out96 = h->d_h_out96;
  }
  P.seq = counter + 1;                                  // sequence number of the first pass
  P.ctl_seq = h->ctl_seq + 1;                           // tag of the first command
  const int grid = std::min(tiles, h->persist_capacity);
  CU(h, launch_match_persistent(P, grid, h->stream));
  h->stats.kernel_launches++;
  int fallback = 0;
  while (!u.done()) {
    u.state(x);
    const unsigned long long seq = ++counter;
    if (h->test_stall_pass >= 0 && u.passes() == h->test_stall_pass) {      // test hook: outlast the kernel's watchdog
      const auto until = std::chrono::steady_clock::now() + std::chrono::milliseconds(60);
      while (std::chrono::steady_clock::now() < until) {}
    }
    const auto tp0 = std::chrono::steady_clock::now();
    post_ctl(h, 0u, 0xFFFFFFFFu, x);
    if (h->pref_idx >= 0 && !h->pref_issued) {             // the requested copy of the next scan overlaps this pass
      rc = issue_prefetch(h);
      if (rc) return rc;
    }
    double own_ns = 0.0;
    if (exchange) {
      for (int i = 0; i < 96; ++i) packed[i] = 0.0;
      for (int r = 0; r < h->xch_world && !fallback; ++r) {   // fixed rank order => identical sums on every rank
        const double* slot = h->xch_host + (size_t)r * slot_doubles + (seq & 1) * buf_doubles;
        double vals[96];
        const int wr = wait_records(h, slot, seq, vals, r == h->xch_rank);
        if (wr < 0) return wr;
        if (wr == 1) {
          fallback = 1;
          break;
        }
        if (r == h->xch_rank) own_ns = vals[93];
        vals[93] = 0.0;
        for (int i = 0; i < 96; ++i) packed[i] += vals[i];
      }
      if (fallback) {
        --counter;                                          // the classic exchange pass re-publishes under the same number
        break;
      }
    } else {
      const int wr = wait_records(h, h->h_out96, seq, packed, true);
      if (wr < 0) return wr;
      if (wr == 1) {                                        // watchdog fired / kernel gone: redo this pass classically
        fallback = 1;
        break;
      }
      own_ns = packed[93];
      packed[93] = 0.0;
    }
    h->stats.match_launches++;
    h->persist_ns_total += own_ns;
    h->persist_passes++;
    if (h->prof) {
      h->prof_wait += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - tp0).count();
      h->prof_passes++;
    }
    if ((int64_t)std::llround(packed[92]) > (int64_t)h->cfg.MAX_NUM_MATCHES) {   // first-N truncation: general path
      if (exchange) {
        post_ctl(h, 1u, 0u, nullptr);
        return fail(h, FLIMO_ERR_STATE, "more accepted matches than MAX_NUM_MATCHES: the host-segment exchange cannot apply the first-N rule across shards (use flimo_update_peer)");
      }
      fallback = 1;
      break;
    }
    int64_t nv = 0, nr = 0;
    double ss = 0;
    flimo_unpack96(packed, HTH, HTh, &nv, &nr, &ss);
    const auto ts0 = std::chrono::steady_clock::now();
    u.step(HTH, HTh, nr);
    if (h->prof) h->prof_step += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - ts0).count();
  }
  post_ctl(h, 1u, 0u, nullptr);                          // stop: the kernel exits
  if (fallback && h->ticket) {
    // the kernel may have ended in the middle of a pass: clear the last-CTA-done counters before the
    // per-pass launches reuse them
    CU(h, cudaMemsetAsync(h->ticket, 0, h->ticket_cap * sizeof(unsigned int), h->stream));
  }
  return fallback;
}

// flimo_update entirely on the device (registration_kernel): one launch, the filter step runs in the CTA that completes a
// pass, the host only waits for the final state + covariance in the mapped result block.  With `exchange` the scan is
// sharded over the ranks attached with flimo_peer_attach and the pass sums travel over NVLink peer memory.
// Returns 0, or 1 when the caller has to run the update the classic way (nothing was changed), or a negative status.
static int update_device(flimo_handle h, double state26[26], double P529[529], int max_iter, const double limit23[23], double R_noise,
                         double D_degeneracy, int* passes_out, bool exchange) {
  const size_t n = h->shard_end - h->shard_begin;
  if (n == 0 || max_iter < 0 || h->reg_capacity <= 0) return 1;
  if (exchange && h->peer_world < 1) return fail(h, FLIMO_ERR_STATE, "flimo_peer_attach not called");
  RegParams RP;
  std::memset(&RP, 0, sizeof(RP));
  int rc = fill_params(h, state26, RP.m, h->out96, 0xFFFFFFFFu, nullptr, nullptr);
  if (rc) return rc;
  ekf::UpdInit& in = *h->h_in;                           // read by the filter kernel over PCIe at its start
  std::memcpy(in.x, state26, sizeof(in.x));
  std::memcpy(in.P, P529, sizeof(in.P));
  std::memcpy(in.limit, limit23, sizeof(in.limit));
  in.R = R_noise;
  in.D = D_degeneracy;
  in.max_iter = max_iter;
  in.max_matches = h->cfg.MAX_NUM_MATCHES;
  RP.host_in = h->d_h_in;
  RP.dev_in = h->dev_in;
  RP.st = h->upd_state;
  RP.host_res = h->d_h_res;
  RP.res_seq = ++h->res_seq;
  RP.world = exchange ? h->peer_world : 1;
  RP.rank = exchange ? h->peer_rank : 0;
  for (int r = 0; r < kMaxPeers; ++r) RP.inbox[r] = exchange ? h->peer_inbox[r] : nullptr;
  RP.peer_timeout_ns = 2000ull * 1000ull * 1000ull;
  // hang protection only: tiles without a filter kernel (or vice versa) give up.  One GPU: a pass takes well under a
  // millisecond, and a stalled update is redone with one launch per pass (flimo_update); several ranks may enter late.
  RP.m.watchdog_ns = (exchange ? 4000ull : 250ull) * 1000ull * 1000ull;
  const int max_cmds = 2 * (max_iter + 2) + 2;          // every pass may be repeated once (first-N rule) + stop
  RP.xseq = h->peer_xseq + 1;
  h->peer_xseq += (unsigned long long)max_cmds;
  RP.m.ctl_seq = h->ctl_seq + 1;
  h->ctl_seq += (unsigned long long)max_cmds;
  // first-N rule: only possible when the scan holds more points than the cap (summed over all shards)
  if ((long long)h->scan_n > (long long)h->cfg.MAX_NUM_MATCHES) {
    const size_t nq = h->scan_n, nw = (nq + 31) / 32;
    if (nw > (size_t)kFlagWordsCap && exchange) return 1;
    CU(h, grow(&h->valid_flags, &h->valid_cap, nq + 64));
    CU(h, grow(&h->flag_words, &h->flag_words_cap, nw + 64));
    CU(h, cudaMemsetAsync(h->valid_flags, 0, nq, h->stream));
    RP.m.valid_by_orig = h->valid_flags;
    RP.flag_words = h->flag_words;
  }
  const int tiles = match_num_tiles((int)n);
  const int grid = std::min(tiles, h->reg_capacity);
  // the filter CTA first (its own stream: it has to run beside the tiles), then the tiles
  // (h->pdl: both on the handle's stream, the tiles as a programmatic dependent of the filter kernel — see
  //  launch_registration_tiles; FLIMO_PDL=0: the round-2 arrangement on two streams, ordered by launch time only)
  CU(h, launch_filter(RP, h->pdl ? h->stream : h->filter_stream));
  if (h->debug_stall_every > 0 && (h->device_updates + 1) % (uint64_t)h->debug_stall_every == 0) {
    // fault injection (tests): these tiles listen for the wrong command numbers, i.e. they go silent after the first pass
    MatchParams deaf = RP.m;
    deaf.ctl_seq += 1000000ull;
    CU(h, launch_registration_tiles(deaf, grid, h->stream, h->pdl));
  } else {
    CU(h, launch_registration_tiles(RP.m, grid, h->stream, h->pdl));
  }
  h->stats.kernel_launches += 2;
  if (h->pref_idx >= 0 && !h->pref_issued) {               // the requested copy of the next scan overlaps the update
    rc = issue_prefetch(h);
    if (rc) return rc;
  }
  double res[kResRecords];
  const int wr = wait_records_n(h, h->h_res, RP.res_seq, res, kResRecords, true);
  if (wr < 0) return wr;
  if (wr == 1) return fail(h, FLIMO_ERR_CUDA, "registration kernel finished without publishing its result");
  const int passes = (int)std::llround(res[kResPasses]);
  int failed = (int)std::llround(res[kResFailed]);
  if (passes_out) *passes_out = passes;
  const uint64_t redone = (uint64_t)std::llround(res[kResRedone]);
  h->stats.match_launches += (uint64_t)passes + redone;
  h->persist_ns_total += res[kResDevNs];
  h->exchange_ns_total += res[kResXchNs];
  h->persist_passes += (uint64_t)passes + redone;
  h->device_updates++;
  h->device_redone += redone;
  if (!failed) {
    // the last pass: final state and covariance (esekfom.hpp:1764-1819) from the state it was evaluated at and its sums
    double HTH[144], HTh[12];
    int64_t nv = 0, nr = 0;
    double ss = 0;
    flimo_unpack96(&res[kResSums], HTH, HTh, &nv, &nr, &ss);
    ekf::IteratedUpdate& u = h->upd;
    u.begin(state26, P529, max_iter, limit23, R_noise, D_degeneracy);
    const auto ts0 = std::chrono::steady_clock::now();
    u.finish(&res[kResXEval], HTH, HTh, nr);
    if (h->prof) h->prof_step += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - ts0).count();
    u.end(state26, P529);
    std::memcpy(h->last_x_dev, state26, sizeof(h->last_x_dev));   // the state after the last pass (formed here)
    if (u.failed()) failed = 1;
  }
  if (failed) {
    // the kernels may have stopped in the middle of a pass: drain them and clear the last-CTA-done counters
    const cudaError_t qf = cudaStreamQuery(h->filter_stream), qt = cudaStreamQuery(h->stream);
    cudaStreamSynchronize(h->filter_stream);
    cudaStreamSynchronize(h->stream);
    if (failed == 2) {                                           // diagnostics of a pass that never completed
      unsigned int tk[4] = {0, 0, 0, 0};
      PassCtl dc{};
      if (h->ticket) cudaMemcpy(tk, h->ticket, sizeof(tk), cudaMemcpyDeviceToHost);
      if (h->dev_ctl) cudaMemcpy(&dc, h->dev_ctl, sizeof(dc), cudaMemcpyDeviceToHost);
      std::fprintf(stderr, "[flimo] at the time-out: %u tiles had delivered, %u CTAs had begun the pass\n", dc.pad[0], tk[2]);
      if (h->cta_trace) {
        std::vector<uint4> tr((size_t)grid);
        cudaMemcpy(tr.data(), h->cta_trace, tr.size() * sizeof(uint4), cudaMemcpyDeviceToHost);
        int by_state[16][12] = {};
        for (const uint4& r : tr) by_state[r.x & 15u][((r.x >> 8) & 255u) < 12u ? ((r.x >> 8) & 255u) : 11u]++;
        for (int pss = 0; pss < 16; ++pss)
          for (int ph = 0; ph < 12; ++ph)
            if (by_state[pss][ph]) std::fprintf(stderr, "[flimo]   %d CTAs: pass %d phase %d (1 waiting, 2 running, 3 delivered, 9 gave up)\n", by_state[pss][ph], pss, ph);
        std::fprintf(stderr, "[flimo]   command posted at %u us, filter time-out at %u us\n", dc.pad[1], dc.pad[2]);
        int shown = 0;
        for (size_t b = 0; b < tr.size() && shown < 10; ++b)
          if (((tr[b].x >> 8) & 255u) == 9u && (tr[b].x & 15u) == 1u) {
            std::fprintf(stderr, "[flimo]   CTA %zu on SM %u gave up after %u polls and %u k cycles, last saw %u\n", b, tr[b].x >> 16, tr[b].z, tr[b].w, tr[b].y);
            ++shown;
          }
        shown = 0;
        for (size_t b = 0; b < tr.size() && shown < 4; ++b)
          if ((tr[b].x & 15u) == 2u) {
            std::fprintf(stderr, "[flimo]   CTA %zu on SM %u (ran the pass): phase %u, then waited for command %u from %u us, last saw %u\n", b, tr[b].x >> 16,
                         (tr[b].x >> 8) & 255u, tr[b].z, tr[b].w, tr[b].y);
            ++shown;
          }
      }
      std::fprintf(stderr, "[flimo] update %llu: pass sums never completed: ticket[0]=%u of %d tiles (grid %d), device command seq=%llu cmd=%u, "
                           "first command of this update %llu, passes done %d, streams at failure: filter %s, tiles %s\n",
                   (unsigned long long)h->device_updates, tk[0], tiles, grid, (unsigned long long)dc.seq, dc.cmd, (unsigned long long)RP.m.ctl_seq, passes,
                   cudaGetErrorName(qf), cudaGetErrorName(qt));
    }
    if (h->ticket) cudaMemset(h->ticket, 0, h->ticket_cap * sizeof(unsigned int));
  }
  if (failed == 2 && !exchange) {                         // own kernels only: the caller redoes the update with one launch per pass
    h->stats_update_stalls++;
    return 2;
  }
  if (failed == 2) return fail(h, FLIMO_ERR_STATE, "a peer rank (or this rank's tile kernel) did not deliver its pass sums in time");
  if (failed) return fail(h, FLIMO_ERR_STATE, "singular or non-finite normal equations: state left at the prediction");
  return FLIMO_OK;
}

// ---- scan preparation --------------------------------------------------------------------------------
// Stages after the 32-byte records are in h->prep.raw.
static int prep_filter_sort_common(flimo_handle h, size_t n, double sweep_ref_time, const flimo_prep_cfg* cfg, size_t* n_kept,
                                   double* t_last) {
  PrepDev c{};
  c.crop_active = cfg->crop_active; c.dist_active = cfg->dist_active; c.rate_active = cfg->rate_active; c.fov_active = cfg->fov_active;
  for (int i = 0; i < 3; ++i) { c.crop_min[i] = cfg->cropBoxMin[i]; c.crop_max[i] = cfg->cropBoxMax[i]; }
  c.min_dist = static_cast<float>(cfg->min_dist);
  c.rate_value = cfg->rate_value > 0 ? cfg->rate_value : 1;
  c.fov_angle = cfg->fov_angle;
  c.sensor_type = cfg->sensor_type;
  c.end_of_sweep = cfg->end_of_sweep;
  c.sweep_ref_time = sweep_ref_time;
  uint32_t m = 0;
  CU(h, prep_filter_sort(h->prep, n, c, h->stream, &m, t_last, &h->stats.kernel_launches));
  *n_kept = m;
  return FLIMO_OK;
}

static int prep_check_args(flimo_handle h, const void* data, size_t n, const flimo_prep_cfg* cfg, size_t* n_kept, double* t_last) {
  if (!h || !cfg || !n_kept || !t_last || (!data && n)) return fail(h, FLIMO_ERR_INVALID, "null argument");
  if (cfg->sensor_type < 0 || cfg->sensor_type > 3) return fail(h, FLIMO_ERR_INVALID, "unknown LiDAR sensor type");   // Localizer.cpp:778-783
  if (cfg->rate_active && cfg->rate_value < 1) return fail(h, FLIMO_ERR_INVALID, "rate_value must be >= 1");
  if (n >= 0x7FFFFFF0ull) return fail(h, FLIMO_ERR_INVALID, "cloud too large");
  return FLIMO_OK;
}

int flimo_prep_filter_sort(flimo_handle h, const void* raw_points, size_t n, double sweep_ref_time, const flimo_prep_cfg* cfg,
                           size_t* n_kept, double* t_last) {
  int rc = prep_check_args(h, raw_points, n, cfg, n_kept, t_last);
  if (rc) return rc;
  NEED_GPU(h);
  h->prep_cfg = *cfg;
  h->prep_deskewed = false;
  h->prep_pc2match = nullptr;
  h->prep_n_pc2match = 0;
  *n_kept = 0;
  *t_last = 0.0;
  if (n == 0) return FLIMO_OK;
  CU(h, prep_reserve(h->prep, n));
  CU(h, cudaMemcpyAsync(h->prep.raw, raw_points, n * 32, cudaMemcpyHostToDevice, h->stream));
  return prep_filter_sort_common(h, n, sweep_ref_time, cfg, n_kept, t_last);
}

int flimo_prep_filter_sort_msg(flimo_handle h, const void* data, size_t n, size_t point_step, const flimo_msg_layout* layout,
                               double sweep_ref_time, const flimo_prep_cfg* cfg, size_t* n_kept, double* t_last) {
  int rc = prep_check_args(h, data, n, cfg, n_kept, t_last);
  if (rc) return rc;
  if (!layout) return fail(h, FLIMO_ERR_INVALID, "null argument");
  // debug_limo::checkPointcloudStructure (src/main.cpp:17-20): the message must carry xyz and the sensor's time field
  const int tsz = layout->time_datatype == 8 ? 8 : 4;
  if (layout->off_x < 0 || layout->off_y < 0 || layout->off_z < 0) return fail(h, FLIMO_ERR_INVALID, "invalid pointcloud structure: xyz missing");
  if (layout->off_time >= 0 && layout->time_datatype != 6 && layout->time_datatype != 7 && layout->time_datatype != 8)
    return fail(h, FLIMO_ERR_INVALID, "invalid pointcloud structure: time field must be UINT32, FLOAT32 or FLOAT64");
  const long long ends[5] = {layout->off_x + 4LL, layout->off_y + 4LL, layout->off_z + 4LL,
                             layout->off_intensity >= 0 ? layout->off_intensity + 4LL : 0LL,
                             layout->off_time >= 0 ? layout->off_time + (long long)tsz : 0LL};
  for (long long e : ends)
    if (e > (long long)point_step) return fail(h, FLIMO_ERR_INVALID, "invalid pointcloud structure: field beyond point_step");
  NEED_GPU(h);
  h->prep_cfg = *cfg;
  h->prep_deskewed = false;
  h->prep_pc2match = nullptr;
  h->prep_n_pc2match = 0;
  *n_kept = 0;
  *t_last = 0.0;
  if (n == 0) return FLIMO_OK;
  CU(h, prep_reserve(h->prep, n));
  rc = upload(h, data, n * point_step);
  if (rc) return rc;
  CU(h, prep_decode_msg(h->prep, static_cast<const unsigned char*>(h->stage), n, point_step, *layout, h->stream, &h->stats.kernel_launches));
  return prep_filter_sort_common(h, n, sweep_ref_time, cfg, n_kept, t_last);
}

int flimo_prep_deskew(flimo_handle h, const flimo_frame* frames, int n_frames, const float last_q[4], const float last_p[3],
                      const float T_lidar2baselink[16], double offset, size_t* n_pc2match) {
  if (!h || !frames || !last_q || !last_p || !T_lidar2baselink || !n_pc2match) return fail(h, FLIMO_ERR_INVALID, "null argument");
  NEED_GPU(h);
  *n_pc2match = 0;
  if (n_frames < 1) return fail(h, FLIMO_ERR_STATE, "no frames obtained from IMU propagation");   // Localizer.cpp:807-811
  const size_t m = h->prep.n_sorted;
  if (m == 0) return flimo_scan_set_device(h, h->prep.xt2, 0, 16);
  DeskewDev d{};
  d.offset = offset;
  d.n_frames = n_frames;
  std::memcpy(d.T_l2b, T_lidar2baselink, sizeof(d.T_l2b));
  {   // last_state.get_RT_inv() (State.cpp:145-153), float, Eigen order
    float R[9];
    quat_matrix<float>(last_q, R);
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) d.Tinv[4 * i + j] = R[3 * j + i];
      d.Tinv[4 * i + 3] = (-R[i] * last_p[0] + -R[3 + i] * last_p[1]) + -R[6 + i] * last_p[2];
    }
    d.Tinv[12] = d.Tinv[13] = d.Tinv[14] = 0.f;
    d.Tinv[15] = 1.f;
  }
  CU(h, prep_deskew(h->prep, d, frames, true, h->stream, &h->stats.kernel_launches));
  h->prep_deskewed = true;
  const float4* out = h->prep.xt2;
  size_t n_out = m;
  if (h->prep_cfg.voxel_active) {
    uint32_t nv = 0;
    bool pass = false;
    CU(h, prep_voxel(h->prep, h->prep.xt2, (uint32_t)m, reinterpret_cast<const uint32_t*>(h->prep.small), h->prep_cfg.leafSize,
                     h->stream, &nv, &pass, &h->stats.kernel_launches));
    if (!pass) {
      out = h->prep.vox_out;
      n_out = nv;
    }
  }
  h->prep_pc2match = out;
  h->prep_n_pc2match = n_out;
  *n_pc2match = n_out;
  return flimo_scan_set_device(h, out, n_out, 16);
}

int flimo_prep_get(flimo_handle h, int what, void* out, size_t cap_items, size_t* n_items) {
  if (!h || !n_items) return fail(h, FLIMO_ERR_INVALID, "null argument");
  NEED_GPU(h);
  const void* src = nullptr;
  size_t n = 0, item = 16;
  switch (what) {
    case 0: src = h->prep.order; n = h->prep.n_sorted; item = 4; break;
    case 1: src = h->prep.world; n = h->prep_deskewed ? h->prep.n_sorted : 0; break;
    case 2: src = h->prep.xt2; n = h->prep_deskewed ? h->prep.n_sorted : 0; break;
    case 3: src = h->prep_pc2match; n = h->prep_n_pc2match; break;
    default: return fail(h, FLIMO_ERR_INVALID, "unknown cloud selector");
  }
  *n_items = n;
  if (!out || n == 0) return FLIMO_OK;
  const size_t take = n < cap_items ? n : cap_items;
  CU(h, cudaMemcpyAsync(out, src, take * item, cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return FLIMO_OK;
}

int flimo_voxel_grid(flimo_handle h, const float* xyz4, size_t n, float leaf, float* out_xyz4, size_t cap_points, size_t* n_out) {
  if (!h || !n_out || (!xyz4 && n) || !(leaf > 0.f)) return fail(h, FLIMO_ERR_INVALID, "bad argument");
  NEED_GPU(h);
  *n_out = 0;
  if (n == 0) return FLIMO_OK;
  if (n >= 0x7FFFFFF0ull) return fail(h, FLIMO_ERR_INVALID, "cloud too large");
  CU(h, prep_reserve(h->prep, n));
  const uint32_t n32 = (uint32_t)n;
  CU(h, cudaMemcpyAsync(h->prep.xt2, xyz4, n * sizeof(float4), cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaMemcpyAsync(h->prep.small, &n32, sizeof(n32), cudaMemcpyHostToDevice, h->stream));
  uint32_t nv = 0;
  bool pass = false;
  CU(h, prep_voxel(h->prep, h->prep.xt2, n32, reinterpret_cast<const uint32_t*>(h->prep.small), leaf, h->stream, &nv, &pass,
                   &h->stats.kernel_launches));
  const float4* src = pass ? h->prep.xt2 : h->prep.vox_out;
  const size_t m = pass ? n : nv;
  *n_out = m;
  if (out_xyz4 && m) {
    CU(h, cudaMemcpyAsync(out_xyz4, src, std::min(m, cap_points) * sizeof(float4), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
  }
  return FLIMO_OK;
}

int flimo_update(flimo_handle h, double state26[26], double P529[529], int max_iter, const double limit23[23],
                 double R_noise, double D_degeneracy, int* passes_out) {
  if (!h || !state26 || !P529 || !limit23) return fail(h, FLIMO_ERR_INVALID, "null argument");
  ekf::IteratedUpdate& u = h->upd;
  ++h->update_calls;
  // Every time_every-th update runs with one launch per pass so that the kernel can be timed with CUDA
  // events (flimo_stats::match_ms_total); the others keep one persistent kernel resident over all passes.
  const bool classic = !h->persistent || h->device < 0 || h->persist_capacity <= 0 || !flimo_map_exists(h) ||
                       h->shard_end <= h->shard_begin || h->timing != nullptr ||
                       (h->time_every > 0 && (h->update_calls % (uint64_t)h->time_every) == 0);
  bool stalled = false;
  if (!classic && h->device_ekf) {                        // the whole update on the device
    NEED_GPU(h);
    const int rc = update_device(h, state26, P529, max_iter, limit23, R_noise, D_degeneracy, passes_out, false);
    if (rc <= 0) return rc;
    stalled = rc == 2;                                     // the resident kernels stopped answering: one launch per pass below
  }
  u.begin(state26, P529, max_iter, limit23, R_noise, D_degeneracy);
  double x[26], HTH[144], HTh[12];
  if (!classic && !stalled && !u.done()) {
    const int rc = update_persistent(h, u, false);
    if (rc < 0) return rc;
  }
  while (!u.done()) {
    u.state(x);
    int64_t nv = 0, nr = 0;
    double ss = 0;
    int rc = flimo_match_reduce(h, x, HTH, HTh, &nv, &nr, &ss);
    if (rc) return rc;
    const auto ts0 = std::chrono::steady_clock::now();
    u.step(HTH, HTh, nr);
    if (h->prof) h->prof_step += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - ts0).count();
  }
  u.end(state26, P529);
  if (passes_out) *passes_out = u.passes();
  if (u.failed()) return fail(h, FLIMO_ERR_STATE, "singular or non-finite normal equations: state left at the prediction");
  return FLIMO_OK;
}

int flimo_update_trace(flimo_handle h, double* out32, size_t cap_passes, size_t* n_passes) {
  if (!h || !n_passes) return fail(h, FLIMO_ERR_INVALID, "null argument");
  NEED_GPU(h);
  *n_passes = 0;
  static ekf::UpdState tmp;                                // 10 KB: not on the stack of a callback thread
  CU(h, cudaStreamSynchronize(h->stream));
  CU(h, cudaMemcpy(&tmp, h->upd_state, sizeof(tmp), cudaMemcpyDeviceToHost));
  const size_t n = (size_t)std::min(std::max(tmp.passes, 0), ekf::kMaxTrace);
  *n_passes = n;
  if (out32)
    for (size_t i = 0; i < n && i < cap_passes; ++i) {
      std::memcpy(out32 + 32 * i, tmp.trace[i], 32 * sizeof(double));
      if (i + 1 == n) std::memcpy(out32 + 32 * i, h->last_x_dev, 26 * sizeof(double));   // the last pass is completed on the host
    }
  if (h->prof)
    for (size_t i = 0; i < n; ++i) {
      std::fprintf(stderr, "[step %zu] ns after tree:", i);
      for (int k = 0; k < 10; ++k) std::fprintf(stderr, " %c=%.0f", "ABCDEFGHIJ"[k], tmp.phase_ns[i][k]);
      std::fprintf(stderr, "  pass total %.0f\n", tmp.trace[i][28]);
    }
  return FLIMO_OK;
}

// ---- peer exchange over NVLink -------------------------------------------------------------------------
int flimo_peer_export(flimo_handle h, void* ipc_handle_64) {
  if (!h || !ipc_handle_64) return fail(h, FLIMO_ERR_INVALID, "null argument");
  NEED_GPU(h);
  static_assert(sizeof(cudaIpcMemHandle_t) == FLIMO_PEER_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");
  if (!h->inbox) {
    CU(h, cudaMalloc(&h->inbox, kInboxBytes));
    CU(h, cudaMemset(h->inbox, 0, kInboxBytes));
    CU(h, cudaDeviceSynchronize());
  }
  cudaIpcMemHandle_t hd;
  CU(h, cudaIpcGetMemHandle(&hd, h->inbox));
  std::memcpy(ipc_handle_64, &hd, sizeof(hd));
  return FLIMO_OK;
}

int flimo_peer_attach(flimo_handle h, int rank, int world, const void* ipc_handles) {
  if (!h || !ipc_handles || world < 1 || world > kMaxPeers || rank < 0 || rank >= world)
    return fail(h, FLIMO_ERR_INVALID, "bad peer arguments (at most 8 ranks)");
  NEED_GPU(h);
  if (!h->inbox) return fail(h, FLIMO_ERR_STATE, "flimo_peer_export not called");
  for (int r = 0; r < h->peer_world; ++r)
    if (h->peer_inbox[r] && r != h->peer_rank) cudaIpcCloseMemHandle(h->peer_inbox[r]);
  for (int r = 0; r < kMaxPeers; ++r) h->peer_inbox[r] = nullptr;
  h->peer_world = 0;
  for (int r = 0; r < world; ++r) {
    if (r == rank) {
      h->peer_inbox[r] = h->inbox;
      continue;
    }
    cudaIpcMemHandle_t hd;
    std::memcpy(&hd, static_cast<const unsigned char*>(ipc_handles) + (size_t)r * sizeof(hd), sizeof(hd));
    void* p = nullptr;
    CU(h, cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
    h->peer_inbox[r] = static_cast<double*>(p);
  }
  h->peer_rank = rank;
  h->peer_world = world;
  h->peer_xseq = 0;
  return FLIMO_OK;
}

int flimo_update_peer(flimo_handle h, double state26[26], double P529[529], int max_iter, const double limit23[23],
                      double R_noise, double D_degeneracy, int* passes_out) {
  if (!h || !state26 || !P529 || !limit23) return fail(h, FLIMO_ERR_INVALID, "null argument");
  NEED_GPU(h);
  if (h->peer_world < 1) return fail(h, FLIMO_ERR_STATE, "flimo_peer_attach not called");
  if (!flimo_map_exists(h)) return fail(h, FLIMO_ERR_STATE, "flimo_update_peer needs a map on every rank");
  if (max_iter < 0) {
    if (passes_out) *passes_out = 0;
    return FLIMO_OK;
  }
  ++h->update_calls;
  const int rc = update_device(h, state26, P529, max_iter, limit23, R_noise, D_degeneracy, passes_out, true);
  if (rc == 1) return fail(h, FLIMO_ERR_STATE, "this rank has no scan points to match (every rank needs a non-empty shard)");
  return rc;
}

int flimo_update_exchange(flimo_handle h, double state26[26], double P529[529], int max_iter, const double limit23[23],
                          double R_noise, double D_degeneracy, int* passes_out) {
  if (!h || !state26 || !P529 || !limit23) return fail(h, FLIMO_ERR_INVALID, "null argument");
  if (!h->xch_host) return fail(h, FLIMO_ERR_STATE, "flimo_exchange_attach not called");
  ekf::IteratedUpdate& u = h->upd;
  u.begin(state26, P529, max_iter, limit23, R_noise, D_degeneracy);
  double x[26], HTH[144], HTh[12];
  ++h->update_calls;
  const bool classic = !h->persistent || h->persist_capacity <= 0 || !flimo_map_exists(h) || h->shard_end <= h->shard_begin ||
                       h->timing != nullptr || (h->time_every > 0 && (h->update_calls % (uint64_t)h->time_every) == 0);
  if (!classic && !u.done()) {
    const int rc = update_persistent(h, u, true);
    if (rc < 0) return rc;
  }
  while (!u.done()) {
    u.state(x);
    int64_t nv = 0, nr = 0;
    double ss = 0;
    int rc = flimo_match_reduce_exchange(h, x, HTH, HTh, &nv, &nr, &ss);
    if (rc) return rc;
    u.step(HTH, HTh, nr);
  }
  u.end(state26, P529);
  if (passes_out) *passes_out = u.passes();
  if (u.failed()) return fail(h, FLIMO_ERR_STATE, "singular or non-finite normal equations: state left at the prediction");
  return FLIMO_OK;
}

}  // extern "C"
