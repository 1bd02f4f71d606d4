// Registration, filter side (B200 / sm_100a): ONE resident CTA per update that turns the partial sums of every
// measurement pass into the next pose — the part of esekf::update_iterated_dyn_share_modified
// (esekfom.hpp:1620-1823) the reference runs on the host between two h_share_model evaluations.
//
//   tiles (match_kernel.cu, registration_tiles_kernel, stream A)      filter (this file, stream B)
//   ------------------------------------------------------------      ------------------------------------------
//   pass k: kNN + plane + row + HTH partials                            pre_step: dx, Jacobians, P = J P_prop J^T
//   two-level tree up to the GROUP partials;                            (while the tiles are still matching)
//   every finished group: ticket[0] += 1            ───────────▶       waits for ticket[0] == n_groups, sums the group
//                                                                       partials (fixed order), packs the 96 doubles
//                                                                       [N > 1: stores them into every rank's inbox over
//                                                                        NVLink peer memory, sums all ranks' records]
//                                                                       [n_valid > MAX_NUM_MATCHES: first-N limit by a prefix
//                                                                        count over the accepted-match bits, pass repeated]
//   wait for dev_ctl.seq                             ◀───────────       post_step: 12x25 elimination, dx, (+) of the pose,
//   pass k+1 ...                                                        pose constants -> dev_ctl, release
//                                                                       then (off the critical path) the rest of the state;
//                                                                       last pass: its sums + the state it was evaluated at
//                                                                       go to the host, which forms the final covariance
//
// No host between the passes, no collective launch; with several ranks every rank runs the identical step on
// bit-identical sums, so no pose exchange is needed either.
#include "flimo_dev.cuh"
#include "pose_consts.hpp"

namespace flimo {

namespace {

__device__ __forceinline__ unsigned long long gtime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_record(double* slot, double v, unsigned long long seq) {
  *reinterpret_cast<ulonglong2*>(slot) = make_ulonglong2((unsigned long long)__double_as_longlong(v), seq);
}
__device__ __forceinline__ ulonglong2 ld_record(const double* slot) {
  ulonglong2 r;
  asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(slot) : "memory");
  return r;
}
__device__ __forceinline__ double* inbox_slot(double* inbox, unsigned long long xs, int src, int i) {
  return inbox + ((((size_t)(xs & 1ull) * kMaxPeers + (size_t)src) * kInboxSlot + (size_t)i) * 2);
}
__device__ __forceinline__ uint32_t* inbox_flags(double* inbox, unsigned long long xs, int src) {
  return reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(inbox) + kInboxRecordBytes) +
         ((size_t)(xs & 1ull) * kMaxPeers + (size_t)src) * kFlagWordsCap;
}

constexpr int kFilterThreads = 128;

struct FilterShared {
  ekf::StepShared step;
  double sums[2][kPartialStride];       // packed pass sums: [0] = the ones the step uses, [1] = scratch of the exchange
  int scratch[kFilterThreads + 8];
  unsigned long long stamps[16];
  double xch_ns;                        // sum over the passes: own tiles complete -> sums of ALL ranks in hand (collect + peer exchange)
  int flag;
};

// (i,j) of the e-th entry of the row-major upper triangle of a 13x13 matrix -> slot of the packed layout (flimo.h)
__device__ __forceinline__ int packed_slot(int e) {
  if (e >= kTriEntries) return e == 91 ? 92 : (e == 92 ? 90 : e);     // n_valid, n_rows, 93..95 reserved
  int r = 0, base = 0;
  for (int q = 0; q < 12; ++q) {
    const int len = 13 - q;
    if (e >= base + len) {
      base += len;
      r = q + 1;
    } else {
      break;
    }
  }
  const int i = r, j = r + (e - base);
  if (j < 12) return i * 12 - (i * (i - 1)) / 2 + (j - i);
  if (i < 12) return 78 + i;
  return 91;
}

// Exchange of the pass sums (fs.sums[0]) over all ranks.  Returns false on a peer time-out.
__device__ __forceinline__ bool peer_exchange_sums(const RegParams& RP, FilterShared& fs, const unsigned long long xs) {
  const int t = (int)threadIdx.x;
  if (t == 0) fs.flag = 1;
  __syncthreads();
  if (t < kPartialStride) {
    const double mine = fs.sums[0][t];
    for (int r = 0; r < RP.world; ++r) st_record(inbox_slot(RP.inbox[r], xs, RP.rank, t), mine, xs);
    const long long w0 = watch_start();
    // the records of all ranks are polled TOGETHER (eight independent loads in flight per round): polling them one rank after
    // the other cost one memory latency per rank — 5.5 us per pass at eight ranks against 3.0 at two
    double* const my_inbox = RP.inbox[RP.rank];
    ulonglong2 rec[kMaxPeers];
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r) rec[r] = make_ulonglong2(0ull, xs);
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r)
      if (r < RP.world) rec[r] = ld_record(inbox_slot(my_inbox, xs, r, t));
    for (;;) {
      bool all = true;
#pragma unroll
      for (int r = 0; r < kMaxPeers; ++r)
        if (r < RP.world && rec[r].y != xs) {
          rec[r] = ld_record(inbox_slot(my_inbox, xs, r, t));
          all = false;                                     // (checked again in the next round)
        }
      if (all) break;
      if (watch_expired(w0, RP.peer_timeout_ns)) {
        fs.flag = 0;
        break;
      }
    }
    double sum = 0.0;
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r)                    // fixed rank order: identical sums on every rank
      if (r < RP.world) {
        double v = __longlong_as_double((long long)rec[r].x);
        if (t >= 93) v = 0.0;                              // slots 93.. are per-rank bookkeeping
        sum += v;
      }
    fs.sums[1][t] = sum;
  }
  __syncthreads();
  if (t < kPartialStride) fs.sums[0][t] = fs.sums[1][t];
  __syncthreads();
  return fs.flag != 0;
}

// First-N rule (Localizer.cpp:539,547-548): the original scan index after which accepted rows stop counting
// (position of the cap-th accepted match in scan order, over all ranks); 0xFFFFFFFF on a peer time-out.
__device__ __forceinline__ uint32_t first_n_limit(const RegParams& RP, FilterShared& fs, const unsigned long long xs, const long long cap) {
  const MatchParams& P = RP.m;
  const int t = (int)threadIdx.x;
  const uint32_t n = P.raw_n, n_words = (n + 31u) >> 5;
  uint32_t* words = RP.flag_words;
  int* s_i = fs.scratch;                                   // [0..127] counts, [128] chunk, [129] before, [130] limit, [131] ok
  if (cap <= 0) return 0u;
  // 1. pack this rank's byte flags into bits (flags of points outside the shard are zero)
  for (uint32_t w = (uint32_t)t; w < n_words; w += kFilterThreads) {
    uint32_t bits = 0;
    const uint32_t base = w << 5;
    if (base + 32u <= n) {
      const uint4 a = __ldcg(reinterpret_cast<const uint4*>(P.valid_by_orig + base)), b = __ldcg(reinterpret_cast<const uint4*>(P.valid_by_orig + base + 16));
      const uint32_t v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
      for (int k = 0; k < 8; ++k)
        bits |= (((v[k] & 1u) | ((v[k] >> 7) & 2u) | ((v[k] >> 14) & 4u) | ((v[k] >> 21) & 8u)) << (4 * k));
    } else {
      for (uint32_t i = base; i < n; ++i) bits |= (uint32_t)(__ldcg(P.valid_by_orig + i) & 1u) << (i - base);
    }
    words[w] = bits;
    if (RP.world > 1)
      for (int r = 0; r < RP.world; ++r) inbox_flags(RP.inbox[r], xs, RP.rank)[w] = bits;
  }
  if (t == 0) s_i[131] = 1;
  __syncthreads();
  if (RP.world > 1) {
    __threadfence_system();                               // the bit words are visible before the "stored" record
    __syncthreads();
    if (t < RP.world) st_record(inbox_slot(RP.inbox[t], xs, RP.rank, 96), (double)n_words, xs);
    if (t < RP.world) {
      const long long w0 = watch_start();
      const double* slot = inbox_slot(RP.inbox[RP.rank], xs, t, 96);
      while (ld_record(slot).y != xs)
        if (watch_expired(w0, RP.peer_timeout_ns)) {
          s_i[131] = 0;
          break;
        }
    }
    __syncthreads();
    if (!s_i[131]) return 0xFFFFFFFFu;
    for (uint32_t w = (uint32_t)t; w < n_words; w += kFilterThreads) {   // shards are disjoint: OR = union
      uint32_t bits = 0;
      for (int r = 0; r < RP.world; ++r) {
        uint32_t v;
        asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(inbox_flags(RP.inbox[RP.rank], xs, r) + w) : "memory");
        bits |= v;
      }
      words[w] = bits;
    }
    __syncthreads();
  }
  // 2. prefix count over contiguous chunks of words
  const uint32_t chunk = (n_words + kFilterThreads - 1) / kFilterThreads;
  const uint32_t w0 = min((uint32_t)t * chunk, n_words), w1 = min(w0 + chunk, n_words);
  int cnt = 0;
  for (uint32_t w = w0; w < w1; ++w) cnt += __popc(words[w]);
  s_i[t] = cnt;
  __syncthreads();
  if (t == 0) {
    long long before = 0;
    int sel = -1;
    for (int k = 0; k < kFilterThreads; ++k) {
      if (before + s_i[k] >= cap) {
        sel = k;
        break;
      }
      before += s_i[k];
    }
    s_i[128] = sel;
    s_i[129] = (int)before;
    s_i[130] = (int)n;                                    // fewer than cap accepted: everything counts
  }
  __syncthreads();
  if (s_i[128] == t) {
    long long need = cap - (long long)s_i[129];           // the need-th set bit of this chunk
    for (uint32_t w = w0; w < w1; ++w) {
      const uint32_t bits = words[w];
      const int c = __popc(bits);
      if (need <= c) {
        const uint32_t pos = __fns(bits, 0, (int)need);   // position of the need-th set bit
        s_i[130] = (int)((w << 5) + pos + 1u);
        break;
      }
      need -= c;
    }
  }
  __syncthreads();
  return (uint32_t)s_i[130];
}

__global__ void __launch_bounds__(kFilterThreads, 1) filter_kernel(const __grid_constant__ RegParams RP) {
  __shared__ FilterShared fs;
  const MatchParams& P = RP.m;
  const int tid = (int)threadIdx.x;
  const int n_tiles = (P.q_end - P.q_begin + kTileQueries - 1) / kTileQueries;
  const int n_groups = (n_tiles + 31) / 32;
  const double* group_part = P.partials + (size_t)n_tiles * kPartialStride;
  PassCtl* dctl = P.dev_ctl;
  ekf::UpdState& st = *RP.st;
  ekf::StepShared& ss = fs.step;
  // the dependent launch (the tiles kernel, next in this stream) may be dispatched from here on: this CTA is resident
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const unsigned long long t_kernel = gtime_ns();

  // the inputs of the update: mapped host block -> device memory (one PCIe round trip, all loads in flight at once)
  {
    const volatile double* hin = reinterpret_cast<const volatile double*>(RP.host_in);
    double* din = reinterpret_cast<double*>(RP.dev_in);
    constexpr int kWords = (int)(sizeof(ekf::UpdInit) / sizeof(double));
    static_assert(sizeof(ekf::UpdInit) % sizeof(double) == 0, "UpdInit is copied as doubles");
    double v[(kWords + kFilterThreads - 1) / kFilterThreads];
#pragma unroll
    for (int k = 0; k < (kWords + kFilterThreads - 1) / kFilterThreads; ++k) v[k] = (tid + k * kFilterThreads < kWords) ? hin[tid + k * kFilterThreads] : 0.0;
#pragma unroll
    for (int k = 0; k < (kWords + kFilterThreads - 1) / kFilterThreads; ++k)
      if (tid + k * kFilterThreads < kWords) din[tid + k * kFilterThreads] = v[k];
  }
  __syncthreads();
  const ekf::UpdInit& in = *RP.dev_in;
  if (tid == 0) {
    ekf::step_begin(ss, in);
    st.passes = 0;
    st.failed = 0;
    st.redone = 0;
  }
  __syncthreads();
  ekf::CtaExec ex{tid, kFilterThreads, fs.stamps};
  bool need_pre = true;
  int redone = 0;
  if (tid == 0) {
    fs.stamps[14] = t_kernel;                               // when the current command was posted
    fs.xch_ns = 0.0;
  }
  for (unsigned long long cmd_no = 0;; ++cmd_no) {
    // The pass that exhausts MAX_NUM_ITERS is known to be the last one before it starts: its sums go straight to the host,
    // which forms the final state and covariance (IteratedUpdate::finish) — nothing to prepare or to solve here.
    const bool known_last = ss.iter == in.max_iter - 1;
    if (need_pre && !known_last) ekf::pre_step(ex, ss, in);   // overlaps the tiles' work on this pass
    if (need_pre && known_last) {
      if (tid < 26) ss.x_eval[tid] = ss.x[tid];
      if (tid == 0) ss.singular = 0;
      __syncthreads();
    }
    need_pre = false;
    // ---- wait for the last group of this pass, sum the group partials in a fixed order -------------------
    if (tid == 0) {
      const long long w0 = watch_start();
      unsigned int naps = 0;
      fs.flag = 1;
      const unsigned int n_arrivals = P.fx_reduce ? (unsigned int)n_tiles : (unsigned int)n_groups;   // tiles (order-free sums) or groups (tree)
      while (ld_acquire_u32(&P.ticket[0]) != n_arrivals) {
        __nanosleep(20);
        if ((++naps & 1023u) == 0u && watch_expired(w0, P.watchdog_ns)) {
          fs.flag = 0;
          break;
        }
      }
      if (!fs.flag) {                                      // diagnostics of a pass that never completed (flimo_update prints them)
        dctl->pad[0] = ld_acquire_u32(&P.ticket[0]);
        dctl->pad[1] = (uint32_t)(fs.stamps[14] / 1000ull);      // when the command of this pass was posted
        dctl->pad[2] = (uint32_t)(gtime_ns() / 1000ull);
        P.ticket[2] = ld_acquire_u32(&P.ticket[1]);
      }
      P.ticket[1] = 0u;
      P.ticket[0] = 0u;                                    // for the next pass (its groups start after the next command)
      fs.stamps[15] = gtime_ns();
    }
    __syncthreads();
    bool ok = fs.flag != 0;
    const double pass_ns = timer_span_ns(fs.stamps[15], fs.stamps[14]);  // command posted -> pass sums complete
    const int pass_idx = ss.passes;                                      // (only post_step_rest changes it, barriers away)
    if (ok && P.fx_reduce) {
      __threadfence();
      if (tid < kPartialStride) fs.sums[0][packed_slot(tid)] = fx_collect(P.ticket, (int)(cmd_no & 1ull), tid);
    } else if (ok && tid < kPartialStride) {
      double s = 0.0;
      for (int g0 = 0; g0 < n_groups; g0 += 32) {
        const double* base = group_part + (size_t)g0 * kPartialStride + tid;
        const int cnt = min(32, n_groups - g0);
        double v[32];
#pragma unroll
        for (int u = 0; u < 32; ++u) v[u] = (u < cnt) ? __ldcg(base + (size_t)u * kPartialStride) : 0.0;
#pragma unroll
        for (int u = 0; u < 32; ++u) s += v[u];
      }
      fs.sums[0][packed_slot(tid)] = s;
    }
    __syncthreads();
    const unsigned long long xs = RP.xseq + cmd_no;
    if (ok && RP.world > 1) ok = peer_exchange_sums(RP, fs, xs);
    ex.stamp(0);
    if (tid == 0) fs.xch_ns += timer_span_ns(fs.stamps[0], fs.stamps[15]);

    uint32_t next_cmd = 0u, next_limit = 0xFFFFFFFFu;
    bool stepped = false;
    const long long n_valid = (long long)(fs.sums[0][92] + 0.5), cap = (long long)in.max_matches;
    const uint32_t cur_limit = (cmd_no == 0) ? P.orig_limit : dctl->orig_limit;   // written by this CTA
    if (!ok) {
      if (tid == 0) ss.failed = 2;                         // the tiles or a peer did not answer in time
      next_cmd = 1u;
    } else if (P.valid_by_orig != nullptr && cur_limit == 0xFFFFFFFFu && n_valid > cap) {
      next_limit = first_n_limit(RP, fs, xs, cap);         // repeat the pass: same pose, rows limited to the first `cap` matches
      if (next_limit == 0xFFFFFFFFu) {
        if (tid == 0) ss.failed = 2;
        next_cmd = 1u;
      } else {
        ++redone;
      }
    } else if (known_last) {
      next_cmd = 1u;
      if (pass_idx < ekf::kMaxTrace && tid < 32) {         // trace: sums of the pass; the state after it is the host's
        double* tr = st.trace[pass_idx];
        tr[tid] = tid < 26 ? ss.x_eval[tid] : (tid == 26 ? fs.sums[0][92] : (tid == 27 ? fs.sums[0][90] : (tid == 28 ? pass_ns : (tid == 29 ? (double)cur_limit : 0.0))));
      }
      __syncthreads();
      if (tid == 0) {
        ss.passes = ss.passes + 1;
        ss.done = 1;
      }
    } else {
      ekf::unpack_measurement(ex, ss, fs.sums[0]);
      ekf::post_step_pose(ex, ss, in, (long long)(fs.sums[0][90] + 0.5));
      stepped = true;
      if (ss.singular || ss.final_pass) next_cmd = 1u;
    }
    ex.stamp(1);
    __syncthreads();
    // ---- post the next command -------------------------------------------------------------------------
    if (next_cmd == 0u) {
      if (stepped && tid < 4 * 32 && (tid & 31) == 0) make_pose_part(ss.x, dctl->pc, tid >> 5);
      else if (!stepped && cmd_no == 0) {                 // repeated first pass: the pose is the one of the parameter block
        const uint32_t* src = reinterpret_cast<const uint32_t*>(&P.pc);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&dctl->pc);
        for (int w = tid; w < (int)(sizeof(PoseConsts) / 4); w += kFilterThreads) dst[w] = src[w];
      }                                                    // (a repeated later pass finds its pose still in dev_ctl)
      if (tid == 0) {
        dctl->cmd = stepped ? 0u : 3u;                       // 3 = same pose again, rows limited (the tiles may reuse their rows)
        dctl->orig_limit = next_limit;
        const unsigned long long now = gtime_ns();
        dctl->t_begin = now;
        fs.stamps[14] = now;
      }
      __syncthreads();                                     // the release below is cumulative over the CTA's writes ordered by this barrier
      if (tid == 0) {
        const unsigned long long pub = P.ctl_seq + cmd_no;
        asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(&dctl->seq), "l"(pub) : "memory");
      }
    }
    ex.stamp(2);
    // ---- off the critical path: the rest of the state, counters, the covariance of the last pass, the trace ----
    if (stepped) {
      ekf::post_step_rest(ex, ss, in);
      need_pre = true;
      if (pass_idx < ekf::kMaxTrace) {
        double* tr = st.trace[pass_idx];
        if (tid < 26) tr[tid] = ss.x[tid];
        if (tid == 26) tr[26] = fs.sums[0][92];
        if (tid == 27) tr[27] = fs.sums[0][90];
        if (tid == 28) tr[28] = pass_ns;
        if (tid == 29) tr[29] = (double)cur_limit;
        if (tid >= 32 && tid < 40) st.phase_ns[pass_idx][tid - 32] = timer_span_ns(fs.stamps[tid - 32], fs.stamps[15]);
      }
    }
    __syncthreads();
    if (next_cmd != 0u) {
      // the update is complete: state, covariance and counters go to the mapped host block as tagged records
      if (tid == 0) {
        st.passes = ss.passes;
        st.failed = ss.failed;
        st.redone = redone;
      }
      if (RP.host_res != nullptr)
        for (int i = tid; i < kResRecords; i += kFilterThreads) {
          double v = 0.0;
          if (i < kResSums) v = ss.x_eval[i];
          else if (i < kResSums + kPartialStride) v = fs.sums[0][i - kResSums];
          else if (i == kResPasses) v = (double)ss.passes;
          else if (i == kResFailed) v = (double)ss.failed;
          else if (i == kResDevNs) v = timer_span_ns(gtime_ns(), t_kernel);
          else if (i == kResRedone) v = (double)redone;
          else if (i == kResXchNs) v = fs.xch_ns;
          else if (i >= kResXDev && i < kResXDev + 26) v = ss.x[i - kResXDev];
          st_record(RP.host_res + 2 * (size_t)i, v, RP.res_seq);
        }
      // release the tiles: stop
      if (tid == 0) {
        dctl->cmd = 1u;
        dctl->orig_limit = 0u;
        dctl->t_begin = gtime_ns();
      }
      __threadfence();
      __syncthreads();
      if (tid == 0) {
        const unsigned long long pub = P.ctl_seq + cmd_no;
        asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(&dctl->seq), "l"(pub) : "memory");
      }
      return;
    }
  }
}

}  // namespace

// The filter CTA asks for (almost) all shared memory of an SM, so that no tile CTA shares the SM with it: its
// float64 code is one latency-bound warp most of the time and should not queue behind 28 tile warps for issue slots.
// (On c2 the step time was the same with and without neighbours; the reservation makes that independent of the load.)
static int filter_dyn_bytes();

cudaError_t preload_filter_kernel() {                     // see preload_match_kernels
  return filter_dyn_bytes() >= 0 ? cudaSuccess : cudaErrorUnknown;
}

cudaError_t launch_filter(const RegParams& p, cudaStream_t st) {
  const int dyn = filter_dyn_bytes();
  if (dyn < 0) return cudaErrorUnknown;
  filter_kernel<<<1, kFilterThreads, (size_t)dyn, st>>>(p);
  return cudaGetLastError();
}

static int filter_dyn_bytes() {
  static int dyn_bytes[64];                                // per device: function attributes belong to a device's context
  static bool init = false;
  if (!init) {
    for (int& v : dyn_bytes) v = -1;
    init = true;
  }
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
  if (dyn_bytes[dev] < 0) {
    int optin = 0;
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, filter_kernel) != cudaSuccess) return -1;
    int want = optin - (int)fa.sharedSizeBytes - 1024;
    if (want < 0) want = 0;
    if (cudaFuncSetAttribute(filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, want) != cudaSuccess) {
      cudaGetLastError();
      want = 0;
    }
    dyn_bytes[dev] = want;
  }
  return dyn_bytes[dev];
}

}  // namespace flimo
