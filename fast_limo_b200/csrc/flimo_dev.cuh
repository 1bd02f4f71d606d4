// Shared host/device declarations of libflimo_cuda (B200 / sm_100a registration hot path).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "ekf_step.hpp"

namespace flimo {

// One level of the device map index.
//
// A level is a uniform grid (cell side `cell`) over the map's bounding box.  Cell (ix,iy,iz) has
// linear index (iz*ny + iy)*nx + ix.  For every grid row (iy,iz) the level stores a SUPER-ROW: all
// map points whose cell lies in rows (iy-1..iy+1, iz-1..iz+1), sorted by ix.  Each point therefore
// appears in (up to) nine super-rows — HBM capacity is traded for access shape: the 3x3x3 cell
// neighbourhood of a query is ONE contiguous run
//     [cell_start[row + ix-1], cell_start[row + ix+2])        of float4 points,
// found with two table reads and streamed with perfectly predictable addresses.
// Levels grow geometrically in cell size; a query whose 5th neighbour is not provably inside the
// 3x3x3 block of level l simply rescans the (larger) block of level l+1 from scratch.
struct GridDesc {
  float ox, oy, oz;   // lower corner
  float inv_cell;     // 1 / cell (float)
  float cell;         // cell side in metres
  int nx, ny, nz;
};

constexpr int kMaxLevels = 12;

struct LevelView {
  const float4* pts;           // super-row storage (9x duplicated): one gapped segment per row, sorted by ix inside it
  const uint32_t* cell_start;  // [ny*nz rows][nx + 1]: slot ix = absolute start of cell ix of the row, slot nx = end of the row's entries
  GridDesc g;
  int row_stride;              // nx + 1
};

// Constants of one measurement pass.  Built on the host from the double filter state exactly the
// way the reference builds them (State.cpp:38-55,136-172; Localizer.cpp:554-555).
struct PoseConsts {
  float R_wb[9], t_wb[3];         // State::get_RT()             (float-cast quaternion)
  float Rinv_wb[9], tinv_wb[3];   // State::get_RT_inv()
  float Rinv_LI[9], tinv_LI[3];   // State::get_extr_RT_inv()
  float Rd_wb_inv[9];             // s.rot.conjugate().toRotationMatrix().cast<float>()
  float Rd_LI_inv[9];             // s.offset_R_L_I.conjugate().toRotationMatrix().cast<float>()
};

// Control block of the persistent kernel (one per handle in mapped pinned host memory, one in device
// memory).  The host fills everything after `t_begin`, then stores `seq` (release); CTA 0 copies the
// payload to the device copy, stamps t_begin and publishes seq there.
struct PassCtl {
  unsigned long long seq;      // sequence number of the pass this block describes (written last)
  unsigned long long t_begin;  // device copy only: %globaltimer when the pass was released
  uint32_t cmd;                // 0 = run the pass, 1 = stop (kernel exits), 3 = run it again with the same pose and a row limit
  uint32_t orig_limit;         // first-N cap on contributing rows
  PoseConsts pc;
  uint32_t pad[3];             // diagnostics of a stalled pass: tiles delivered at the time-out, time the command was posted, time of the time-out (us, low 32 bits)
};
static_assert(sizeof(PassCtl) % 8 == 0, "PassCtl must be a whole number of 8-byte words");
// Wire format of the HOST control block: the 56 payload words of a PassCtl (cmd, orig_limit, pose) in 19
// records of 16 bytes, each {3 payload words, tag = low 32 bits of the pass sequence number}.  The host
// stores every record with ONE 16-byte store, the device reads every record with ONE 16-byte load and all
// 19 in parallel: a command is complete when all tags match — a single PCIe read round trip per poll.
constexpr int kCtlPayloadWords = 2 + (int)(sizeof(PoseConsts) / 4);          // 56
constexpr int kCtlRecords = (kCtlPayloadWords + 2) / 3;                       // 19
struct PassCtlWire {
  uint32_t rec[kCtlRecords][4];
};
static_assert(kCtlRecords <= 32, "one warp reads the control block");
constexpr unsigned long long kPassAbort = 1ull << 63;   // flag bit: the watchdog ended the kernel at this command

constexpr int kTileQueries = 128;     // queries per CTA tile (== threads per CTA)
constexpr int kTriEntries = 91;       // upper triangle of the 13x13 outer product of [row(12), z]
constexpr int kPartialStride = 96;    // doubles per partial record

// Pass accumulators of the order-free reduction (MatchParams::fx_reduce): they live behind the ticket words,
//   ticket[0..3]                                   counters
//   u64 [parity 2][replica kFxReplicas][96][2]     fixed-point sums {units of 2^-18, units of 2^-66}
// (replicas spread the same-address REDs of tiles that finish together; parity alternates between consecutive passes so
// that the finisher's zeroing of one set is two passes ahead of its next use).
#ifdef __CUDACC__
// Watchdogs count SM cycles (clock64), not %globaltimer: the cycle counter of an SM is monotonic whatever happens to the
// global timer (an unsigned `globaltimer - t0 > limit` turns true at once if the timer is ever set back), and %globaltimer
// values of different SMs were found not to be comparable at the 100 us level.  2.25 cycles per nanosecond is above any
// B200 clock, so a limit never fires early (at 1.965 GHz it fires 15 % late).
__device__ __forceinline__ long long watch_start() { return clock64(); }
__device__ __forceinline__ bool watch_expired(long long c0, unsigned long long limit_ns) {
  return (unsigned long long)(clock64() - c0) > 2ull * limit_ns + (limit_ns >> 2);
}
// elapsed %globaltimer time for statistics: a negative difference (timer stepped back in between) counts as zero
__device__ __forceinline__ double timer_span_ns(unsigned long long t1, unsigned long long t0) {
  const long long d = (long long)(t1 - t0);
  return d > 0 ? (double)d : 0.0;
}
#endif

constexpr int kFxReplicas = 4;
constexpr int kTicketWords = 4 + 2 * kFxReplicas * kPartialStride * 2 * 2;
#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long* fx_slot(unsigned int* ticket, int sel, int rep, int e) {
  return reinterpret_cast<unsigned long long*>(ticket + 4) + ((size_t)((sel & 1) * kFxReplicas + rep) * kPartialStride + (size_t)e) * 2;
}
// Sum of entry e over the replicas as float64 (one rounding); the slots are cleared for their next use.
__device__ __forceinline__ double fx_collect(unsigned int* ticket, int sel, int e) {
  long long hi = 0, lo = 0;
  unsigned long long v[2 * kFxReplicas];
#pragma unroll
  for (int r = 0; r < kFxReplicas; ++r) {
    const unsigned long long* a = fx_slot(ticket, sel, r, e);
    v[2 * r] = __ldcg(a);
    v[2 * r + 1] = __ldcg(a + 1);
  }
#pragma unroll
  for (int r = 0; r < kFxReplicas; ++r) {
    hi += (long long)v[2 * r];
    lo += (long long)v[2 * r + 1];
    unsigned long long* a = fx_slot(ticket, sel, r, e);
    __stcg(reinterpret_cast<ulonglong2*>(a), make_ulonglong2(0ull, 0ull));
  }
  return (double)hi * 0x1p-18 + (double)lo * 0x1p-66;
}
#endif

struct MatchParams {
  const float4* scan;          // packed body-frame points; .w carries the original scan index (bit pattern)
  // Unpacked scan (scan == nullptr): the caller's strided xyz array is read in place.  Stored position j of
  // the pseudo-random order is original point (j * raw_inv) mod raw_n (raw_inv = inverse of the upload
  // permutation's stride), evaluated with a 64-bit Barrett reduction (raw_magic = floor(2^64 / raw_n)).
  const unsigned char* raw_scan;
  uint32_t raw_stride, raw_n, raw_inv;
  int raw_vec4;                // stride and base are multiples of 16 bytes: points are read with one 16-byte load
  unsigned long long raw_magic;
  LevelView lv[kMaxLevels];    // finest first
  int n_levels;
  PoseConsts pc;
  int q_begin, q_end;          // slice of the scan handled by this launch
  int interleave;              // 1 = slot s of tile t takes stored point s*n_tiles+t (Morton-sorted scans); 0 = tiles are contiguous
  int tau;                     // level choice: finest level whose 3x3x3 block holds >= tau candidates
  int probe_mode;              // 0 = probe all levels at once, 1 = climb one level at a time
  int wide_loads;              // 1 = 256-bit candidate loads
  int pair_scan;               // 1 = two lanes per query in the first scan (adjacent 32-byte loads share a wavefront)
  int l2_prefetch;             // 1 = request the whole run with prefetch.global.L2 right after the probe
  int stage_runs;              // 1 = one-launch-per-pass kernel stages the runs in shared memory with cp.async.bulk (TMA) + mbarrier
  int fx_reduce;               // 1 = tile sums are accumulated with 64-bit fixed-point REDs (one ticket round), 0 = two-level float64 tree
  float max_dist_f;            // smallest float >= MAX_DIST_PLANE  (d2_5 < MAX_DIST_PLANE test)
  float plane_thr;             // (float)PLANE_THRESHOLD
  int estimate_extrinsics;
  uint32_t orig_limit;         // rows contribute only if original index < orig_limit
  double* partials;            // [n_tiles][96] scratch
  unsigned int* ticket;        // last-CTA-done counter (self resetting)
  double* out96;               // packed result (see flimo.h)
  float* dbg16;                // optional per-point record [n][16], indexed by original index
  uint8_t* valid_by_orig;      // optional accepted flag per original index
  double* host_out96;          // optional mapped pinned copy of the result: 96 records {double value, u64 seq}
  unsigned int host_out_alt;   // doubles between the two alternating host result blocks (block = seq & 1); 0 = one block
  unsigned long long seq;      // sequence number stored with every record
  unsigned long long* timing;  // optional per-warp timestamps (profiling builds of the tools only)
  uint4* cta_trace;            // optional: per-CTA progress record of the registration tiles kernel {pass, phase, last seq seen (lo), smid}
  // persistent kernel only
  const PassCtlWire* host_ctl; // device alias of the host control block
  PassCtl* dev_ctl;            // device copy
  unsigned long long watchdog_ns;
  unsigned long long ctl_seq;  // tag of the first command (commands are numbered independently of the results)
};

// Parameter block of the filter kernel (filter_kernel.cu; with registration_tiles_kernel: one whole iterated update,
// no host between passes).
constexpr int kResRecords = 160;      // host result block: records {double value, u64 seq}
// x_eval = state the LAST pass was evaluated at, sums = its 96 packed sums (the host forms the final state and
// covariance from them, IteratedUpdate::finish), x_dev = the device's own state after the last pass (tests)
constexpr int kResXEval = 0, kResSums = 26, kResPasses = 122, kResFailed = 123, kResDevNs = 124, kResRedone = 125, kResXchNs = 126, kResXDev = 128;
constexpr int kMaxPeers = 8;
// Peer inbox (device memory of every rank, mapped by all ranks of the node through CUDA IPC):
//   records : [parity 2][source rank 8][128] x {double value, u64 seq}   (0..95 pass sums, 96 = "flag words stored")
//   flagbits: [parity 2][source rank 8][kFlagWordsCap] u32               accepted-match bits by original scan index
constexpr int kInboxSlot = 128;                         // records per (parity, source)
constexpr int kFlagWordsCap = 1 << 15;                  // 2^20 scan points
constexpr size_t kInboxRecordBytes = (size_t)2 * kMaxPeers * kInboxSlot * 16;
constexpr size_t kInboxBytes = kInboxRecordBytes + (size_t)2 * kMaxPeers * kFlagWordsCap * 4;
struct RegParams {
  MatchParams m;
  const ekf::UpdInit* host_in;   // device alias of the mapped pinned input block (read once, over PCIe, at kernel start)
  ekf::UpdInit* dev_in;          // its copy in device memory
  ekf::UpdState* st;             // results / trace (device memory)
  double* host_res;              // device alias of the mapped pinned result block
  unsigned long long res_seq;    // tag of this update's result records
  // scan sharded over several GPUs: every rank stores its 96 partial sums into every rank's inbox (peer memory
  // over NVLink), sums the inbox in rank order and runs the identical step
  int world, rank;
  double* inbox[kMaxPeers];      // inbox[r] = rank r's inbox as mapped in this process; layout [parity][source rank][96] records of 16 bytes
  unsigned long long xseq;       // sequence number of the first exchange of this update
  unsigned long long peer_timeout_ns;
  uint32_t* flag_words;          // scratch: this rank's accepted-match bits (ceil(raw_n / 32) words)
};

// map_index.cu
struct LevelIndex {
  float4* pts = nullptr;        // super-row entries (cell key kept in .w), one segment with head-room per row (map_index.cu)
  uint32_t* cell_start = nullptr;   // (nx + 1) slots per row
  uint32_t* row_base = nullptr;     // [n_rows + 1] segment start of every row
  uint32_t* row_cap = nullptr;      // [n_rows + 1] segment capacity; row_cap[n_rows + 1] = the tail pointer (first free entry)
  size_t n_entries = 0, cap_entries = 0;
  size_t n_cells = 0, n_rows = 0, cap_slots = 0, cap_rows = 0;
  GridDesc g{};
};

struct MapIndex {
  float4* pts = nullptr;        // canonical map points (each once, insertion order)
  size_t n_pts = 0, cap_pts = 0;
  LevelIndex lv[kMaxLevels];
  int n_levels = 0;
  float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};   // bounding box of the map
  float glo[3] = {0, 0, 0}, ghi[3] = {0, 0, 0}; // box the grids were laid out for (bounding box + margin)
  // incremental update scratch (new entries of all levels): sorted 64-bit (level, cell) keys, point ids, entries, row jobs
  void* upd_buf = nullptr;
  size_t upd_bytes = 0;
  uint32_t* upd_counters = nullptr;   // [0] jobs, [1] rows moved to the tail, [2] a level's array is full
  uint32_t* row_first = nullptr;      // build scratch: first sorted entry of every row
  size_t row_first_cap = 0;
  uint64_t rows_moved = 0;            // statistics
  // scratch
  uint32_t *keys = nullptr, *keys_alt = nullptr, *vals = nullptr, *vals_alt = nullptr;
  size_t cap_scratch = 0;       // entries (9 per point)
  void* cub_tmp = nullptr;
  size_t cub_tmp_bytes = 0;
  float* bbox = nullptr;        // 6 floats on device
};

// Rebuilds every level from the canonical points idx.pts[0..idx.n_pts).  cell0 <= 0 picks the
// finest cell automatically; levels grow by `ratio` until one cell >= coarsest_min (the level whose
// 3x3x3 block is guaranteed to cover the reference's MAX_DIST_PLANE ball).
cudaError_t map_index_build(MapIndex& idx, float cell0, float ratio, float coarsest_min, size_t max_cells,
                            cudaStream_t st, uint64_t* launches);
cudaError_t map_index_reserve(MapIndex& idx, size_t n_pts);
// Incremental form for Mapper::add: the points idx.pts[old_n..idx.n_pts) are merged into the rows they touch, all levels
// in one pass (sort of the 9*m*levels new entries, one CTA per touched row).  Answers are identical to a full rebuild's.
// Precondition (map_index_can_update): same grids, capacity in place, batch inside the grid box.
// *full = true: a level ran out of room behind its segments — the caller rebuilds (which compacts) before the next search.
bool map_index_can_update(const MapIndex& idx, size_t old_n, const float batch_lo[3], const float batch_hi[3]);
cudaError_t map_index_update(MapIndex& idx, size_t old_n, cudaStream_t st, uint64_t* launches, bool* full);
void map_index_free(MapIndex& idx);

// Packs strided xyz (device) into float4 with w = 0, dropping NaN points; writes the count.
cudaError_t pack_points(const void* d_src, size_t n, size_t stride_bytes, float4* dst, unsigned int* d_count,
                        cudaStream_t st);
// Packs a scan: float4 with w = original index bits (no NaN filtering: the reference matches them
// and they simply fail the kNN gate).
cudaError_t pack_scan(const void* d_src, size_t n, size_t stride_bytes, float4* dst, cudaStream_t st);
cudaError_t scan_prepare(const void* d_src, size_t n, size_t stride_bytes, bool sort, unsigned int perm_stride, float4* scan, float4* tmp, void** cub_tmp,
                         size_t* cub_tmp_bytes, uint32_t** keys, size_t* keys_cap, cudaStream_t st, uint64_t* launches);
cudaError_t scan_prepare_reserve(size_t n, void** cub_tmp, size_t* cub_tmp_bytes, uint32_t** keys, size_t* keys_cap);
cudaError_t transform_raw(const unsigned char* d_src, size_t n, size_t stride_bytes, const PoseConsts& pc, float* d_out_xyz, cudaStream_t st);
cudaError_t transform_scan(const float4* scan, size_t n, const PoseConsts& pc, float* d_out_xyz, cudaStream_t st);

// map_insert.cu — the reference's incremental insert rule (Octree::update, Octree.hpp:341-432)
constexpr int kMaxChain = 24;          // root doublings tracked (2^23 x the first scan's extent)

struct LatticeDesc {                   // kernel-parameter view of the octree lattice
  int n_chain;                         // chain[0] = original root ... chain[n_chain-1] = current root
  int min_depth;                       // depth of the min-level cells below the original root
  int off[3];                          // integer offset of the current root's lower corner (min-level cells)
  float chain_c[kMaxChain][3];
  float chain_ext[kMaxChain];
  int chain_slot[kMaxChain];           // child slot of chain[k-1] inside chain[k]
};

struct OctreeLattice {                 // host-side state, built with the reference's float arithmetic
  struct Root {
    float c[3];
    float ext;
    int slot;
  };
  std::vector<Root> chain;
  int min_depth = 0;
  int off[3] = {0, 0, 0};
  float min_extent = 0.2f;
  bool initialised = false;
};

struct CountTable {                    // open-addressing hash: cell key -> number of map points
  unsigned long long* keys = nullptr;
  uint32_t* vals = nullptr;
  size_t cap = 0;
  size_t used_bound = 0;               // upper bound on occupied slots
};

void lattice_init(OctreeLattice& L, const float lo[3], const float hi[3], float min_extent);
void lattice_grow(OctreeLattice& L, const float boundary[3]);
void table_free(CountTable& T);
cudaError_t map_insert_batch(OctreeLattice& L, CountTable& T, const float4* d_batch, size_t n, size_t map_points, int downsample, bool first_batch,
                             float4* d_dst, unsigned int* d_counter, unsigned long long* d_cell_keys, uint8_t* d_accept,
                             unsigned int* n_accepted, cudaStream_t st, uint64_t* launches);
// bounding box (lo[3], hi[3]) of n float4 points on the device -> host (synchronises the stream)
cudaError_t points_bbox(const float4* d_pts, size_t n, float* d_scratch8, float lo[3], float hi[3], cudaStream_t st);

// match_kernel.cu
cudaError_t launch_match(const MatchParams& p, cudaStream_t st);
cudaError_t launch_match_persistent(const MatchParams& p, int grid, cudaStream_t st);
cudaError_t launch_registration_tiles(const MatchParams& p, int grid, cudaStream_t st, bool after_primary);
cudaError_t launch_filter(const RegParams& p, cudaStream_t st);
int match_persistent_capacity();
int registration_capacity();
cudaError_t preload_match_kernels();
cudaError_t preload_filter_kernel();
int match_num_tiles(int n_queries);

}  // namespace flimo
