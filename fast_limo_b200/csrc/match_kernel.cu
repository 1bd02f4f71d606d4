// K1 — the fused registration measurement pass on B200 (sm_100a).
//
// One launch = one IKFoM::h_share_model evaluation (use-ikfom.cpp:10-31) fused with the two
// products that consume it (esekfom.hpp:1723 HTH = H^T H, :1727 H^T h):
//
//   per scan point (one thread each, reference loop Mapper.cpp:68-76):
//     g      = T_wb * p                              Mapper.cpp:71-72, State.cpp:136-143
//     5-NN   = exact nearest map points to g         Mapper.cpp:106-109, Octree.hpp:526-599
//     gates  : 5 found, d2_5 < MAX_DIST_PLANE        Plane.cpp:41-48
//     plane  : column-pivoted Householder QR of the 5x3 system A x = -1 (float32),
//              n = x/|x|, d = 1/|x|                  Plane.cpp:80-105 (Eigen ColPivHouseholderQR)
//     gate   : all |n.q_j + d| <= PLANE_THRESHOLD    Plane.cpp:107-114
//     dist   = n.g + d                               Plane.cpp:50-52, Match.cpp:27
//     row    = [n, p_imu x C, p_lid x (R_LI^T C), C], z = -dist      Localizer.cpp:549-572
//   reduction (float64): sum over accepted points of [row,z]^T [row,z]  (13x13 upper triangle =
//     78 HTH + 12 HTh + sum z^2), warp-cooperative, then a deterministic two-level
//     last-CTA-done tree over tiles.  H is never written to memory.
//
// This translation unit is compiled with --fmad=false: every float32 operation is an individually
// rounded IEEE add/mul/div/sqrt in the order the reference's x86-64 (no-FMA) build evaluates them,
// so per-point results are bit-identical to the CPU restatement in oracle/ (tests/ check that).
//
// Data layout: see map_index.cu.  The search reads map points as 16-byte float4 through the
// read-only path; neighbouring threads handle neighbouring scan points (ring order or Morton
// order), so the cell rows a warp touches overlap and are served from L1/L2.
#include <cfloat>

#include "flimo_dev.cuh"
#include "grid_math.cuh"

namespace flimo {

namespace {

constexpr int kGroupTiles = 32;   // tiles per first-level reduction group

struct Top5 {
  float d[5];
  uint32_t i[5];
  uint32_t sl;     // 3 bits per neighbour: its slot in the thread's neighbour stash (see kStashStride)
};
// Neighbour stash: the exact re-ranking loads the coordinates of (up to) six candidates anyway; they are parked in
// shared memory (slot j of lane l of a warp at [j * 32 + l], conflict-free for any per-lane slot) so that the plane fit
// reads the five winners from there instead of fetching them from the map a second time.
constexpr int kStashStride = 32;
constexpr uint32_t kStashIdentity = 0u | (1u << 3) | (2u << 6) | (3u << 9) | (4u << 12);

// Sorted insertion; on equal distance the earlier candidate stays in front and a candidate equal
// to the current worst is rejected (Octree.hpp:72-87).  Branch free; NaN is never inserted.
__device__ __forceinline__ void top5_offer(Top5& t, float d, uint32_t idx) {
  const bool c0 = d < t.d[0], c1 = d < t.d[1], c2 = d < t.d[2], c3 = d < t.d[3], c4 = d < t.d[4];
  t.i[4] = c3 ? t.i[3] : (c4 ? idx : t.i[4]);
  t.d[4] = c3 ? t.d[3] : (c4 ? d : t.d[4]);
  t.i[3] = c2 ? t.i[2] : (c3 ? idx : t.i[3]);
  t.d[3] = c2 ? t.d[2] : (c3 ? d : t.d[3]);
  t.i[2] = c1 ? t.i[1] : (c2 ? idx : t.i[2]);
  t.d[2] = c1 ? t.d[1] : (c2 ? d : t.d[2]);
  t.i[1] = c0 ? t.i[0] : (c1 ? idx : t.i[1]);
  t.d[1] = c0 ? t.d[0] : (c1 ? d : t.d[1]);
  t.i[0] = c0 ? idx : t.i[0];
  t.d[0] = c0 ? d : t.d[0];
}

// ---------------------------------------------------------------------------------------------------
// Block scan: exact top-5 of the 3x3x3 cell block around the query's home cell at one level.
// The block is one contiguous run of the level's super-row (flimo_dev.cuh), streamed four points at
// a time.  Selection is branch-free on PACKED keys:
//
//   key = (float_bits(d2) & ~low) | n            n = position in the run (unique per candidate)
//
// Squared distances are non-negative floats, so their bit patterns order like unsigned integers; the
// low B mantissa bits are replaced by the candidate's ordinal, which makes every key unique and lets a
// 6-slot sorted list be maintained with 11 integer min/max per candidate — no branch, no index
// bookkeeping.  Truncation can only reorder candidates whose distances agree in all but the low B
// bits, so the six smallest KEYS are re-ranked with exact arithmetic afterwards, and the result is
// accepted only if provably identical to the exact search:
//   every candidate not kept has key > k5  =>  bucket(d2) >= bucket(k5);
//   if bucket(k5) > bucket(exact 5th d2) then all of them have d2 > that 5th d2.     (else: exact path)
// Ties in d2 are ordered by position in the run (x cell, then map index).
// ---------------------------------------------------------------------------------------------------
// Keys are bit patterns of non-negative, non-NaN floats, so they are ranked with FLOAT min/max
// (FMNMX, full-rate ALU pipe) — the ordering equals the unsigned-integer ordering of the bits.
// Empty slot = +inf.  A NaN key (NaN query) is never inserted because fminf/fmaxf drop NaN operands.
__device__ __forceinline__ void keys6_insert(float (&k)[6], float x) {
  k[5] = fminf(k[5], fmaxf(x, k[4]));
  k[4] = fminf(k[4], fmaxf(x, k[3]));
  k[3] = fminf(k[3], fmaxf(x, k[2]));
  k[2] = fminf(k[2], fmaxf(x, k[1]));
  k[1] = fminf(k[1], fmaxf(x, k[0]));
  k[0] = fminf(k[0], x);
}
__device__ __forceinline__ float pack_key(float d, uint32_t low, uint32_t ord) {
  return __uint_as_float((__float_as_uint(d) & ~low) | ord);
}

// Insertion of a SORTED pair x <= y into the ascending 6-list (lowest six survive).  Merging two sorted
// sequences: c_j = min(a_j, max(a_{j-1}, x), max(a_{j-2}, y)); with 3-input FMNMX3 that is 15 min/max per
// pair (+2 to sort the pair) instead of 22 for two single insertions.
__device__ __forceinline__ float fmin3(float a, float b, float c) { return fminf(fminf(a, b), c); }
__device__ __forceinline__ void keys6_insert2(float (&k)[6], float a, float b) {
  const float x = fminf(a, b), y = fmaxf(a, b);
  const float c5 = fmin3(k[5], fmaxf(k[4], x), fmaxf(k[3], y));
  const float c4 = fmin3(k[4], fmaxf(k[3], x), fmaxf(k[2], y));
  const float c3 = fmin3(k[3], fmaxf(k[2], x), fmaxf(k[1], y));
  const float c2 = fmin3(k[2], fmaxf(k[1], x), fmaxf(k[0], y));
  const float c1 = fmin3(k[1], fmaxf(k[0], x), y);
  const float c0 = fminf(k[0], x);
  k[0] = c0; k[1] = c1; k[2] = c2; k[3] = c3; k[4] = c4; k[5] = c5;
}
// Ranking distance of the scan loop: fused multiply-adds (2 instructions fewer than the reference's
// rounding order).  It may differ from the exact value by a few ulp, which block_scan_private's
// acceptance test allows for; every distance that leaves the scan is recomputed exactly.
__device__ __forceinline__ float sqdist_rank(float qx, float qy, float qz, float px, float py, float pz) {
  const float dx = qx - px, dy = qy - py, dz = qz - pz;
  return __fmaf_rn(dx, dx, __fmaf_rn(dy, dy, dz * dz));
}

__device__ __forceinline__ void cmpswap64(unsigned long long& a, unsigned long long& b) {
  const bool sw = a > b;
  const unsigned long long ta = sw ? b : a, tb = sw ? a : b;
  a = ta;
  b = tb;
}

__device__ __forceinline__ float sqdist(float qx, float qy, float qz, const float4& p) {
  const float dx = qx - p.x, dy = qy - p.y, dz = qz - p.z;
  return dx * dx + (dy * dy + dz * dz);               // Eigen Vector3f::squaredNorm order
}

constexpr uint32_t kPrivateCap = 96;   // longest run a single thread scans by itself

// Two consecutive super-row entries with ONE 256-bit load (LDG.E.256, sm_100): each lane streams
// its own run, so every load instruction costs one L1 wavefront per lane whatever its width —
// 32 bytes per wavefront instead of 16 halves the wavefront count of the scan.
struct Pair {
  float ax, ay, az, aw, bx, by, bz, bw;
};
// kSrc: 0 = two 128-bit global loads, 1 = one 256-bit global load, 2 = the run has been staged in shared memory
template <int kSrc>
__device__ __forceinline__ Pair ldg_pair(const float4* p) {
  Pair r;
  if (kSrc == 2) {
    const float4 a = p[0], b = p[1];
    r.ax = a.x; r.ay = a.y; r.az = a.z; r.aw = a.w;
    r.bx = b.x; r.by = b.y; r.bz = b.z; r.bw = b.w;
  } else if (kSrc == 1) {
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r.ax), "=f"(r.ay), "=f"(r.az), "=f"(r.aw), "=f"(r.bx), "=f"(r.by), "=f"(r.bz), "=f"(r.bw)
        : "l"(p));
  } else {
    const float4 a = __ldg(p), b = __ldg(p + 1);
    r.ax = a.x; r.ay = a.y; r.az = a.z; r.aw = a.w;
    r.bx = b.x; r.by = b.y; r.bz = b.z; r.bw = b.w;
  }
  return r;
}
constexpr uint32_t kKeyLow = 127u;     // ordinal bits of a packed key (kPrivateCap + 1 < 128)

// Ranks four consecutive run entries (two 32-byte pairs) whose first ordinal is n.
__device__ __forceinline__ void rank4(float (&k)[6], const Pair& p0, const Pair& p1, float qx, float qy, float qz, uint32_t n,
                                      float lead) {
  const float d0 = fmaxf(sqdist_rank(qx, qy, qz, p0.ax, p0.ay, p0.az), lead), d1 = sqdist_rank(qx, qy, qz, p0.bx, p0.by, p0.bz),
              d2 = sqdist_rank(qx, qy, qz, p1.ax, p1.ay, p1.az), d3 = sqdist_rank(qx, qy, qz, p1.bx, p1.by, p1.bz);
  keys6_insert2(k, pack_key(d0, kKeyLow, n), pack_key(d1, kKeyLow, n + 1));
  keys6_insert2(k, pack_key(d2, kKeyLow, n + 2), pack_key(d3, kKeyLow, n + 3));
}

// Exact re-ranking + acceptance of a packed 6-list (see block_scan_private).  pts = L.pts + a (ordinals are
// relative to it), a = absolute position of ordinal 0.
template <int kSrc = 1>
__device__ __forceinline__ bool finish_private(const float (&k)[6], const float4* __restrict__ pts, uint32_t a, float qx, float qy,
                                               float qz, Top5& t, float4* stash) {
  const uint32_t low = kKeyLow;
  const float inf = __int_as_float(0x7f800000);
  // exact re-ranking of the (up to) six kept candidates: sort (exact d2 bits, ordinal) pairs
  unsigned long long ek[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const uint32_t kb = __float_as_uint(k[j]);
    const bool have = kb < 0x7f800000u;
    const uint32_t ord = kb & low;
    float d = inf;
    if (have) {
      const float4 p = kSrc == 2 ? pts[ord] : __ldg(&pts[ord]);
      stash[j * kStashStride] = p;
      d = sqdist(qx, qy, qz, p);
    }
    ek[j] = ((unsigned long long)__float_as_uint(d) << 32) | (have ? ((ord << 3) | (uint32_t)j) : 0xFFFFFFFFu);   // ties: by ordinal
  }
  cmpswap64(ek[0], ek[5]); cmpswap64(ek[1], ek[3]); cmpswap64(ek[2], ek[4]);
  cmpswap64(ek[1], ek[2]); cmpswap64(ek[3], ek[4]);
  cmpswap64(ek[0], ek[3]); cmpswap64(ek[2], ek[5]);
  cmpswap64(ek[0], ek[1]); cmpswap64(ek[2], ek[3]); cmpswap64(ek[4], ek[5]);
  cmpswap64(ek[1], ek[2]); cmpswap64(ek[3], ek[4]);
  t.sl = 0u;
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    t.d[j] = __uint_as_float((uint32_t)(ek[j] >> 32));
    t.i[j] = a + ((uint32_t)(ek[j] & 0xFFFFFFFFu) >> 3);
    t.sl |= ((uint32_t)ek[j] & 7u) << (3 * j);
  }
  // Acceptance.  Every candidate that was not kept has a ranking key >= k5 (the sixth smallest), hence a
  // ranking distance >= bucket_floor(k5) and an exact distance within 4 ulp of that.  If k5's bucket
  // (128 ulp wide) lies at least TWO buckets above the bucket of the exact fifth distance, all of them
  // are strictly farther than the fifth neighbour, so the kept set contains the exact answer.
  const uint32_t k5 = __float_as_uint(k[5]);
  return (k5 >= 0x7f800000u) || ((k5 & ~low) >= ((uint32_t)(ek[4] >> 32) & ~low) + 2u * (low + 1u));
}

// Thread-private scan of a short run [s, e) of level L (<= kPrivateCap candidates).  Returns false
// when the packed selection cannot be proven exact (or the run is too long): the caller then hands
// the query to the warp-cooperative exact scan below.  On success t.d[] holds the exact ascending
// squared distances of the block's five nearest points and t.i[] their positions in L.pts (+inf / 0
// if missing).  The run is read from its 32-byte aligned start a = s & ~1; ordinals are relative to a,
// the (at most one) leading entry before s is masked, trailing entries are never ranked.
template <int kWide>
__device__ __forceinline__ bool block_scan_private(const LevelView& L, uint32_t s, uint32_t e, float qx, float qy, float qz,
                                                   Top5& t, float4* stash, const float4* staged = nullptr) {
  const uint32_t total = e - s;
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    t.d[j] = __int_as_float(0x7f800000);
    t.i[j] = 0;
  }
  if (total == 0) return true;
  if (total > kPrivateCap) return false;
  const uint32_t a = s & ~1u;
  const uint32_t tot = (s - a) + total;          // ordinals [s - a, tot) are candidates
  const float4* __restrict__ pts = (kWide == 2) ? staged : L.pts + a;   // staged: the copy of [a, a + tot) in shared memory
  const uint32_t low = kKeyLow;
  const float inf = __int_as_float(0x7f800000);
  float lead = (s != a) ? inf : 0.f;             // raises the masked leading entry's distance to +inf
  float k[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) k[j] = inf;

  // Software-pipelined, two register sets in ping-pong (no register moves): while four entries are
  // ranked the next four are in flight.
  const uint32_t last = (tot - 1) & ~1u;         // first entry of the last pair
  Pair a0 = ldg_pair<kWide>(pts), a1 = ldg_pair<kWide>(pts + min(2u, last));
  uint32_t n = 0;
#pragma unroll 1
  for (; n + 8 <= tot; n += 8) {
    const Pair b0 = ldg_pair<kWide>(pts + min(n + 4, last)), b1 = ldg_pair<kWide>(pts + min(n + 6, last));
    rank4(k, a0, a1, qx, qy, qz, n, lead);
    lead = 0.f;
    if (n + 8 < tot) {                            // nothing left to rank: do not fetch
      a0 = ldg_pair<kWide>(pts + min(n + 8, last));
      a1 = ldg_pair<kWide>(pts + min(n + 10, last));
    }
    rank4(k, b0, b1, qx, qy, qz, n + 4, 0.f);
  }
  if (n + 4 <= tot) {
    const Pair b0 = ldg_pair<kWide>(pts + min(n + 4, last)), b1 = ldg_pair<kWide>(pts + min(n + 6, last));
    rank4(k, a0, a1, qx, qy, qz, n, lead);
    lead = 0.f;
    a0 = b0;
    a1 = b1;
    n += 4;
  }
  // tail (< 4 candidates): a0 / a1 hold entries n..n+3 (clamped)
  if (n < tot) keys6_insert(k, pack_key(fmaxf(sqdist_rank(qx, qy, qz, a0.ax, a0.ay, a0.az), lead), low, n));
  if (n + 1 < tot) keys6_insert(k, pack_key(sqdist_rank(qx, qy, qz, a0.bx, a0.by, a0.bz), low, n + 1));
  if (n + 2 < tot) keys6_insert(k, pack_key(sqdist_rank(qx, qy, qz, a1.ax, a1.ay, a1.az), low, n + 2));

  return finish_private<kWide>(k, pts, a, qx, qy, qz, t, stash);
}

// ---- two lanes per query (P.pair_scan) ------------------------------------------------------------------
// The lane pair (2i, 2i+1) scans ONE query's run together, the warp's 32 queries in two rounds: lane parity b
// ranks every second 32-byte pair of the run, so the pair's two loads are adjacent (one 64-byte segment, one
// L1 wavefront for both lanes instead of one each), then the two packed 6-lists are merged (6 shuffles +
// bitonic min + sorting network).  Ordinals are positions in the run counted from its 64-byte aligned start,
// unique across both lanes, so the merged list is exactly what one lane scanning alone would hold.
__device__ __forceinline__ void sort6f(float (&m)[6]) {
#define FL_CS(i, j) { const float lo_ = fminf(m[i], m[j]), hi_ = fmaxf(m[i], m[j]); m[i] = lo_; m[j] = hi_; }
  FL_CS(0, 5) FL_CS(1, 3) FL_CS(2, 4)
  FL_CS(1, 2) FL_CS(3, 4)
  FL_CS(0, 3) FL_CS(2, 5)
  FL_CS(0, 1) FL_CS(2, 3) FL_CS(4, 5)
  FL_CS(1, 2) FL_CS(3, 4)
#undef FL_CS
}
template <bool kWide>
__device__ __forceinline__ void pair_scan_round(const float4* __restrict__ pts, uint32_t lead_n, uint32_t total, int b, bool valid,
                                                float qx, float qy, float qz, float (&k)[6]) {
  const float inf = __int_as_float(0x7f800000);
#pragma unroll
  for (int j = 0; j < 6; ++j) k[j] = inf;
  if (valid) {
    const uint32_t tot = lead_n + total;
    const uint32_t lastp = (tot - 1) & ~1u;
    uint32_t n = 2u * (uint32_t)b;                       // this lane's pairs start at entries 2b, 2b+4, 2b+8, ...
    Pair c0 = ldg_pair<kWide>(pts + min(n, lastp)), c1 = ldg_pair<kWide>(pts + min(n + 4, lastp));
#pragma unroll 1
    for (; n < tot; n += 8) {
      const Pair p0 = c0, p1 = c1;
      if (n + 8 < tot) {
        c0 = ldg_pair<kWide>(pts + min(n + 8, lastp));
        c1 = ldg_pair<kWide>(pts + min(n + 12, lastp));
      }
      const float d0 = sqdist_rank(qx, qy, qz, p0.ax, p0.ay, p0.az), d1 = sqdist_rank(qx, qy, qz, p0.bx, p0.by, p0.bz),
                  d2 = sqdist_rank(qx, qy, qz, p1.ax, p1.ay, p1.az), d3 = sqdist_rank(qx, qy, qz, p1.bx, p1.by, p1.bz);
      // entries before the run (ordinal < lead_n) and past its end rank as +inf (exactly: a NaN key would be dropped
      // by fmin/fmax and duplicate its partner)
      const float k0 = (n - lead_n < total) ? pack_key(d0, kKeyLow, n) : inf;
      const float k1 = (n + 1 - lead_n < total) ? pack_key(d1, kKeyLow, n + 1) : inf;
      const float k2 = (n + 4 - lead_n < total) ? pack_key(d2, kKeyLow, n + 4) : inf;
      const float k3 = (n + 5 - lead_n < total) ? pack_key(d3, kKeyLow, n + 5) : inf;
      keys6_insert2(k, k0, k1);
      keys6_insert2(k, k2, k3);
    }
  }
  // merge with the partner lane: min(k[j], partner[5-j]) are the six smallest of the union (bitonic), then sort
  float m[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) m[j] = __shfl_xor_sync(0xffffffffu, k[5 - j], 1);
#pragma unroll
  for (int j = 0; j < 6; ++j) m[j] = fminf(k[j], m[j]);
  sort6f(m);
#pragma unroll
  for (int j = 0; j < 6; ++j) k[j] = m[j];
}

// TEAM scan: exact top-5 of one query's block at level L by a team of T consecutive lanes
// (T a power of two, 1..32; every lane of the warp calls this, teams work on different queries).
// Team lanes stride through the contiguous run (coalesced reads), keep private exact top-5 lists,
// and the lists are merged inside the team by five rounds of (d2 bits, position) arg-min built from
// xor-shuffles of width T.  Results are identical in all lanes of a team.  A team with nothing to do
// passes valid = false.
// `bound`: an upper bound of the query's true 5th squared distance (the exact 5th distance inside a smaller block that was
// scanned before, +inf if none): candidates beyond it cannot be among the five nearest and skip the insertion.
__device__ __forceinline__ void team_offer(Top5& mine, float d, uint32_t i, float bound) {
  if (d <= bound && d < mine.d[4]) top5_offer(mine, d, i);
}
__device__ __forceinline__ void block_scan_team(const LevelView& L, int T, int tl, bool valid, float qx, float qy, float qz, float bound,
                                                float (&rd)[5], uint32_t (&ri)[5]) {
  const unsigned int full = 0xffffffffu;
  const GridDesc& G = L.g;
  uint32_t s = 0, e = 0;
  if (valid) {
    const int hx = cell_coord(qx, G.ox, G.inv_cell, G.nx), hy = cell_coord(qy, G.oy, G.inv_cell, G.ny),
              hz = cell_coord(qz, G.oz, G.inv_cell, G.nz);
    const size_t row = (size_t)(hz * G.ny + hy) * (size_t)L.row_stride;
    s = __ldg(&L.cell_start[row + max(hx - 1, 0)]);
    e = __ldg(&L.cell_start[row + min(hx + 1, G.nx - 1) + 1]);
  }
  Top5 mine;
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    mine.d[j] = __int_as_float(0x7f800000);
    mine.i[j] = 0xFFFFFFFFu;
  }
  const float4* __restrict__ pts = L.pts;
  // four loads in flight per lane (the offers form one dependent chain; the loads do not depend on it)
  uint32_t i = s + tl;
  const uint32_t T4 = 4u * (uint32_t)T;
#pragma unroll 1
  for (; i + 3u * (uint32_t)T < e; i += T4) {
    const float4 p0 = __ldg(&pts[i]), p1 = __ldg(&pts[i + T]), p2 = __ldg(&pts[i + 2 * T]), p3 = __ldg(&pts[i + 3 * T]);
    team_offer(mine, sqdist(qx, qy, qz, p0), i, bound);
    team_offer(mine, sqdist(qx, qy, qz, p1), i + T, bound);
    team_offer(mine, sqdist(qx, qy, qz, p2), i + 2 * T, bound);
    team_offer(mine, sqdist(qx, qy, qz, p3), i + 3 * T, bound);
  }
#pragma unroll 1
  for (; i < e; i += T) team_offer(mine, sqdist(qx, qy, qz, __ldg(&pts[i])), i, bound);
#pragma unroll
  for (int round = 0; round < 5; ++round) {
    const uint32_t db = __float_as_uint(mine.d[0]);
    uint32_t mb = db;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
      if (o < T) mb = min(mb, __shfl_xor_sync(full, mb, o));
    const uint32_t cand = (db == mb) ? mine.i[0] : 0xFFFFFFFFu;
    uint32_t mi = cand;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
      if (o < T) mi = min(mi, __shfl_xor_sync(full, mi, o));
    if (db == mb && mine.i[0] == mi) {                        // the winner pops its head
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        mine.d[j] = mine.d[j + 1];
        mine.i[j] = mine.i[j + 1];
      }
      mine.d[4] = __int_as_float(0x7f800000);
      mine.i[4] = 0xFFFFFFFFu;
    }
    rd[round] = __uint_as_float(mb);
    ri[round] = (mi == 0xFFFFFFFFu) ? 0u : mi;
  }
}

// Is the block's answer provably the answer of an unbounded search (or irrelevant for the reference's
// outcome)?  m = distance, in cells, from q to the nearest face of the 3x3x3 block that is not on the
// grid boundary (nothing lies beyond boundary faces); `slack` covers the ~1 ulp binning error of the
// point and of the query.
__device__ __forceinline__ bool block_is_final(const GridDesc& G, float max_dist_f, float qx, float qy, float qz, int hx,
                                               int hy, int hz, float d5) {
  const float ux = cell_units(qx, G.ox, G.inv_cell), uy = cell_units(qy, G.oy, G.inv_cell),
              uz = cell_units(qz, G.oz, G.inv_cell);
  const float fx = ux - (float)hx, fy = uy - (float)hy, fz = uz - (float)hz;
  const float slack = 6.0e-7f * fmaxf(fmaxf(fabsf(ux), fabsf(uy)), fabsf(uz)) + 1.0e-6f;
  float m = FLT_MAX;
  if (hx - 1 > 0) m = fminf(m, fx + 1.0f);
  if (hx + 1 < G.nx - 1) m = fminf(m, (1.0f - fx) + 1.0f);
  if (hy - 1 > 0) m = fminf(m, fy + 1.0f);
  if (hy + 1 < G.ny - 1) m = fminf(m, (1.0f - fy) + 1.0f);
  if (hz - 1 > 0) m = fminf(m, fz + 1.0f);
  if (hz + 1 < G.nz - 1) m = fminf(m, (1.0f - fz) + 1.0f);
  if (m == FLT_MAX) return true;                              // the block is the whole grid
  const float me = (m - slack) * G.cell * 0.999999f;
  const float gr2 = me > 0.f ? me * me : 0.f;
  if (d5 <= gr2) return true;                                 // exact: nothing closer can be outside
  return gr2 >= max_dist_f;                                   // outside points fail Plane::close_enough
}

// One probe of level L: the query's home cell and the run [s, e) of its 3x3x3 block.
struct Probe {
  int hx, hy, hz;
  uint32_t s, e;
};
__device__ __forceinline__ void probe_level(const LevelView& L, float qx, float qy, float qz, Probe& p) {
  const GridDesc& G = L.g;
  p.hx = cell_coord(qx, G.ox, G.inv_cell, G.nx);
  p.hy = cell_coord(qy, G.oy, G.inv_cell, G.ny);
  p.hz = cell_coord(qz, G.oz, G.inv_cell, G.nz);
  const size_t row = (size_t)(p.hz * G.ny + p.hy) * (size_t)L.row_stride;
  p.s = __ldg(&L.cell_start[row + max(p.hx - 1, 0)]);
  p.e = __ldg(&L.cell_start[row + min(p.hx + 1, G.nx - 1) + 1]);
}

// Level to rescan after the block of level `lvl` was not final.  d5 is the exact 5th squared
// distance inside that block, i.e. an upper bound of the true one: the first level whose CELL is at
// least sqrt(d5) long has a block that provably contains the answer (block_is_final: m >= 1), so one
// more scan settles the query.  Fewer than five points found (d5 = +inf): climb two levels.
__device__ __forceinline__ int next_level(const MatchParams& P, int lvl, float d5) {
  const int top = P.n_levels - 1;
  if (!(d5 < __int_as_float(0x7f800000))) return min(lvl + 2, top);
  const float need = sqrtf(d5) * 1.002f + 1.0e-6f;       // covers block_is_final's slack
  int l = lvl + 1;
  while (l < top && P.lv[l].g.cell < need) ++l;
  return min(l, top);
}

// Exact 5-NN of q, outcome-equivalent to the reference's unbounded octree search
// (Octree.hpp:526-599) for every query the reference would accept: a level whose 3x3x3 block
// provably contains the 5th neighbour answers; the coarsest level's cell is >= sqrt(MAX_DIST_PLANE),
// so its block always covers the radius beyond which Plane::close_enough rejects the match anyway.
//   level choice : the grids form a geometric ladder of cell sizes and every query starts on the FINEST
//                  level whose block holds at least tau candidates (two table reads per probe, the
//                  next probe is predicted from the count) — the octree's "a leaf holds ~a bucket of
//                  points" adapted to a flat layout: scan length is ~tau..2 tau whatever the local
//                  density, lanes of a warp do equal work and few queries need a second scan;
//   first scan   : one thread per query (block_scan_private);
//   the rest     : the queries that are not settled are shared out among TEAMS of lanes
//                  (block_scan_team) at the level next_level() names, until every query is settled.
// Correctness never depends on the level choice: a block's answer is used only if block_is_final.
// Must be called by all 32 lanes (inactive lanes pass active = false).  `lvl` returns the level whose
// storage t.i[] indexes into.
// Level choice of a query (see knn_search): finest level whose 3x3x3 block holds at least tau candidates.
__device__ __forceinline__ void knn_probe(const MatchParams& P, float qx, float qy, float qz, int& lvl, Probe& pr) {
  const int top = P.n_levels - 1;
  lvl = 0;
  if (P.probe_mode == 0) {
    // all levels probed at once (independent table reads, one memory latency): finest level with >= tau
    uint32_t cs[kMaxLevels], ce[kMaxLevels];
#pragma unroll
    for (int l = 0; l < kMaxLevels; ++l) {
      cs[l] = ce[l] = 0;
      if (l < P.n_levels) {
        Probe q;
        probe_level(P.lv[l], qx, qy, qz, q);
        cs[l] = q.s;
        ce[l] = q.e;
      }
    }
    int sel = top;
#pragma unroll
    for (int l = kMaxLevels - 1; l >= 0; --l)
      if (l < P.n_levels && (int)(ce[l] - cs[l]) >= P.tau) sel = l;
    lvl = sel;
    probe_level(P.lv[lvl], qx, qy, qz, pr);
  } else {
    probe_level(P.lv[0], qx, qy, qz, pr);
    while (lvl < top && (int)(pr.e - pr.s) < P.tau) {
      lvl = min(lvl + 1, top);
      probe_level(P.lv[lvl], qx, qy, qz, pr);
    }
  }
}

// ---- staging of the runs with bulk asynchronous copies (TMA, 1-D) -----------------------------------------------
// Variant kStage (FLIMO_KNN_STAGE=1, one-launch-per-pass kernel only): after the probe every lane asks the TMA unit for ITS
// run with one cp.async.bulk (global -> its slice of the warp's shared-memory stage, completion on the warp's mbarrier);
// the 32 copies of a warp are in flight together, the scan and the exact re-ranking then read shared memory.  Runs longer
// than kStageCap entries keep the streaming path.  Measured slower than streaming on c2 (profiles/README.md, round 2): the
// stage costs 24.5 KB per warp, i.e. 2 CTAs per SM instead of 7, and the run loads were not the critical path.
constexpr uint32_t kStageCap = 48;      // entries per lane
constexpr uint32_t kStageStride = 49;   // entries between lane slices (784 bytes: conflict-free 128-bit reads across a quarter warp)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
struct StageCtx {
  float4* lane_slice;           // this lane's slice of the warp's stage (nullptr: no staging)
  unsigned long long* bar;      // the warp's mbarrier
  uint32_t* phase;              // its phase parity (per thread copy, all lanes of the warp agree)
};

template <bool kWide, bool kPair, bool kStage = false>
__device__ __forceinline__ void knn_search(const MatchParams& P, int lane, bool active, float qx, float qy, float qz, Top5& t,
                                           int& lvl, int& first_lvl, uint32_t& first_cnt, unsigned long long& t_priv, unsigned long long& t_probe,
                                           float4* stash, const StageCtx stage = StageCtx{nullptr, nullptr, nullptr}) {
  const unsigned int full = 0xffffffffu;
  lvl = 0;
  first_lvl = 0;
  first_cnt = 0;
  bool pending = false;
  float ub = __int_as_float(0x7f800000);       // upper bound of the true 5th squared distance known so far (see block_scan_team)
  uint32_t scan_s = 0, scan_e = 0;
  int hx0 = 0, hy0 = 0, hz0 = 0;
  if (active) {
    const int top = P.n_levels - 1;
    Probe pr;
    knn_probe(P, qx, qy, qz, lvl, pr);
    first_lvl = lvl;
    first_cnt = pr.e - pr.s;
    if (P.timing) {
      if (pr.e == 0xFFFFFFFFu) first_cnt = 0;   // keep the dependence on the probe
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_probe) :: "memory");
    }
    if (P.l2_prefetch) {
      // Pull the whole run towards L2 now (fire and forget, no registers): every lane streams a different
      // run, so a warp-level load waits for the slowest of 32 independent lines — with the run requested
      // up front the scan's loads find their lines in L2 (or on their way) instead of paying DRAM latency
      // once per loop iteration.
      const char* pf = reinterpret_cast<const char*>(P.lv[lvl].pts + (pr.s & ~7u));
      const char* pe = reinterpret_cast<const char*>(P.lv[lvl].pts + min(pr.e, pr.s + kPrivateCap));
      for (; pf < pe; pf += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
    }
    if (kPair) {
      scan_s = pr.s;
      scan_e = pr.e;
      hx0 = pr.hx; hy0 = pr.hy; hz0 = pr.hz;
    }
    if (!kPair && !kStage) {
      const bool exact_here = block_scan_private<kWide ? 1 : 0>(P.lv[lvl], pr.s, pr.e, qx, qy, qz, t, stash);
      if (!exact_here) {
        pending = true;                                          // redo this level cooperatively
      } else if (!block_is_final(P.lv[lvl].g, P.max_dist_f, qx, qy, qz, pr.hx, pr.hy, pr.hz, t.d[4])) {
        pending = lvl < top;
        if (pending) lvl = next_level(P, lvl, t.d[4]);
        ub = t.d[4];
      }
    }
    if (kStage) {
      scan_s = pr.s;
      scan_e = pr.e;
      hx0 = pr.hx; hy0 = pr.hy; hz0 = pr.hz;
    }
  }
  if (kStage) {                                                  // warp-converged: stage the runs, then scan them
    const uint32_t total = scan_e - scan_s, a = scan_s & ~1u, tot = (scan_s - a) + total;
    const bool staged = active && total > 0u && tot <= kStageCap;
    const uint32_t bytes = staged ? ((tot + 1u) & ~1u) * (uint32_t)sizeof(float4) : 0u;   // whole 32-byte pairs (the arrays carry slack)
    const uint32_t warp_bytes = __reduce_add_sync(full, bytes);
    if (warp_bytes != 0u) {
      if (lane == 0) mbar_expect_tx(stage.bar, warp_bytes);
      __syncwarp();
      if (staged) bulk_copy_g2s(stage.lane_slice, P.lv[lvl].pts + a, bytes, stage.bar);
      unsigned int spins = 0;
      while (!mbar_try_wait(stage.bar, *stage.phase)) {
        if (++spins > (1u << 22)) break;                         // never observed; a lost copy must not hang the GPU
      }
      *stage.phase ^= 1u;
    }
    if (active) {
      const int top = P.n_levels - 1;
      const bool exact_here = staged ? block_scan_private<2>(P.lv[lvl], scan_s, scan_e, qx, qy, qz, t, stash, stage.lane_slice)
                                     : block_scan_private<kWide ? 1 : 0>(P.lv[lvl], scan_s, scan_e, qx, qy, qz, t, stash);
      if (!exact_here) {
        pending = true;
      } else if (!block_is_final(P.lv[lvl].g, P.max_dist_f, qx, qy, qz, hx0, hy0, hz0, t.d[4])) {
        pending = lvl < top;
        if (pending) lvl = next_level(P, lvl, t.d[4]);
        ub = t.d[4];
      }
    }
    __syncwarp();                                                // the slices are free again (next tile of a resident kernel)
  }
  if (kPair) {                                                   // warp-converged: two lanes per query, two rounds
    const uint32_t total_own = scan_e - scan_s;
    const bool do_scan = active && total_own > 0u && total_own <= kPrivateCap;
    float kk[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) kk[j] = __int_as_float(0x7f800000);
#pragma unroll 1
    for (int r = 0; r < 2; ++r) {
      const int src = (lane & ~1) + r;
      const float bx = __shfl_sync(full, qx, src), by = __shfl_sync(full, qy, src), bz = __shfl_sync(full, qz, src);
      const uint32_t bs = __shfl_sync(full, scan_s, src), be = __shfl_sync(full, scan_e, src);
      const int bl = __shfl_sync(full, lvl, src);
      const bool bdo = __shfl_sync(full, do_scan ? 1 : 0, src) != 0;
      const uint32_t ba = bs & ~3u;                              // 64-byte aligned start of the run
      float k[6];
      pair_scan_round<kWide>(P.lv[bl].pts + ba, bs - ba, be - bs, lane & 1, bdo, bx, by, bz, k);
      if (lane == src) {
#pragma unroll
        for (int j = 0; j < 6; ++j) kk[j] = k[j];
      }
    }
    if (active) {
      const int top = P.n_levels - 1;
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        t.d[j] = __int_as_float(0x7f800000);
        t.i[j] = 0;
      }
      bool exact_here = total_own == 0u;                         // empty block: exact (nothing there)
      if (do_scan) {
        const uint32_t a = scan_s & ~3u;
        exact_here = finish_private(kk, P.lv[lvl].pts + a, a, qx, qy, qz, t, stash);
      }
      if (!exact_here) {
        pending = true;
      } else if (!block_is_final(P.lv[lvl].g, P.max_dist_f, qx, qy, qz, hx0, hy0, hz0, t.d[4])) {
        pending = lvl < top;
        if (pending) lvl = next_level(P, lvl, t.d[4]);
        ub = t.d[4];
      }
    }
  }
  if (P.timing) {
    __syncwarp();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_priv));
  }
  bool teamed = false;
#pragma unroll 1
  for (;;) {
    const unsigned int todo = __ballot_sync(full, pending);
    if (todo == 0) break;
    teamed = teamed || pending;
    // Split the warp into G = pow2ceil(#pending) teams of T = 32/G lanes; team g serves the pending
    // lane of rank g.  Few stragglers => wide teams (short, parallel scans); many => T = 1, which is
    // the thread-per-query regime with every lane busy.
    const int n_pend = __popc(todo);
    const int G = (n_pend <= 1) ? 1 : (1 << (32 - __clz(n_pend - 1)));
    const int T = 32 / G;
    const int team = lane / T, tl = lane - team * T;
    const bool valid = team < n_pend;
    const int src_lane = valid ? (int)__fns(todo, 0, team + 1) : 0;
    const float bx = __shfl_sync(full, qx, src_lane), by = __shfl_sync(full, qy, src_lane), bz = __shfl_sync(full, qz, src_lane);
    const int bl = __shfl_sync(full, lvl, src_lane);
    float rd[5];
    uint32_t ri[5];
    const float bb = __shfl_sync(full, ub, src_lane);
    block_scan_team(P.lv[bl], T, tl, valid, bx, by, bz, bb, rd, ri);
    // hand the results back: pending lane of rank r reads from the first lane of team r
    const int my_rank = __popc(todo & ((1u << lane) - 1u));
    const int from = pending ? my_rank * T : lane;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const float dj = __shfl_sync(full, rd[j], from);
      const uint32_t ij = __shfl_sync(full, ri[j], from);
      if (pending) {
        t.d[j] = dj;
        t.i[j] = ij;
      }
    }
    if (pending) {
      const LevelView& L = P.lv[lvl];
      const int hx = cell_coord(qx, L.g.ox, L.g.inv_cell, L.g.nx), hy = cell_coord(qy, L.g.oy, L.g.inv_cell, L.g.ny),
                hz = cell_coord(qz, L.g.oz, L.g.inv_cell, L.g.nz);
      if (block_is_final(L.g, P.max_dist_f, qx, qy, qz, hx, hy, hz, t.d[4]) || lvl + 1 >= P.n_levels) pending = false;
      else lvl = next_level(P, lvl, t.d[4]);
      ub = t.d[4];
    }
  }
  if (teamed) {                                                  // a team's answer: its five points go to the stash now
    const float4* __restrict__ src = P.lv[lvl].pts;
#pragma unroll
    for (int j = 0; j < 5; ++j) stash[j * kStashStride] = __ldg(&src[t.i[j]]);
    t.sl = kStashIdentity;
  }
}

// Column swap helper with static register indexing.
__device__ __forceinline__ void cswap(bool c, float& a, float& b) {
  const float ta = c ? b : a, tb = c ? a : b;
  a = ta;
  b = tb;
}

// Least-squares solve of the 5x3 system A x = -1 by column-pivoted Householder QR, float32, same
// operation order as Eigen 3.3's ColPivHouseholderQR::computeInPlace + _solve_impl restated in
// oracle/plane_match.hpp (reference call site Plane.cpp:95).  A[r][c] is destroyed.
__device__ __forceinline__ void plane_qr_solve(float (&A)[5][3], float (&x)[3]) {
  constexpr int R = 5;
  const float eps = 1.1920929e-07f;
  float nu[3], nd[3], hc[3];
  int perm[3] = {0, 1, 2};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < R; ++r) s = s + A[r][c] * A[r][c];
    nd[c] = sqrtf(s);
    nu[c] = nd[c];
  }
  float maxn = nu[0];
  if (nu[1] > maxn) maxn = nu[1];
  if (nu[2] > maxn) maxn = nu[2];
  const float th = maxn * eps;
  const float threshold_helper = (th * th) / 5.0f;
  const float downdate_thr = sqrtf(eps);
  int nonzero = 3;

#pragma unroll
  for (int k = 0; k < 3; ++k) {
    // pivot: first column of maximal updated norm among k..2
    int big = k;
    float bign = nu[k];
#pragma unroll
    for (int c = k + 1; c < 3; ++c)
      if (nu[c] > bign) {
        bign = nu[c];
        big = c;
      }
    const float big_sq = bign * bign;
    if (nonzero == 3 && big_sq < threshold_helper * (float)(R - k)) nonzero = k;
#pragma unroll
    for (int c = k + 1; c < 3; ++c) {
      const bool sw = (big == c);
#pragma unroll
      for (int r = 0; r < R; ++r) cswap(sw, A[r][k], A[r][c]);
      cswap(sw, nu[k], nu[c]);
      cswap(sw, nd[k], nd[c]);
      const int pk = sw ? perm[c] : perm[k], pc = sw ? perm[k] : perm[c];
      perm[k] = pk;
      perm[c] = pc;
    }
    // Householder vector of column k
    float tail = 0.f;
#pragma unroll
    for (int r = k + 1; r < R; ++r) tail = tail + A[r][k] * A[r][k];
    const float c0 = A[k][k];
    float tau, beta;
    if (tail <= FLT_MIN) {
      tau = 0.f;
      beta = c0;
#pragma unroll
      for (int r = k + 1; r < R; ++r) A[r][k] = 0.f;
    } else {
      beta = sqrtf(c0 * c0 + tail);
      if (c0 >= 0.f) beta = -beta;
      const float den = c0 - beta;
#pragma unroll
      for (int r = k + 1; r < R; ++r) A[r][k] = A[r][k] / den;
      tau = (beta - c0) / beta;
    }
    A[k][k] = beta;
    hc[k] = tau;
    if (tau != 0.f) {
#pragma unroll
      for (int c = k + 1; c < 3; ++c) {
        float tmp = 0.f;
#pragma unroll
        for (int r = k + 1; r < R; ++r) tmp = tmp + A[r][k] * A[r][c];
        tmp = tmp + A[k][c];
        A[k][c] = A[k][c] - tau * tmp;
#pragma unroll
        for (int r = k + 1; r < R; ++r) A[r][c] = A[r][c] - tmp * (tau * A[r][k]);
      }
    }
    // column-norm downdate (LAWN 176)
#pragma unroll
    for (int c = k + 1; c < 3; ++c) {
      if (nu[c] != 0.f) {
        float temp = fabsf(A[k][c]) / nu[c];
        temp = (1.f + temp) * (1.f - temp);
        temp = temp < 0.f ? 0.f : temp;
        const float ratio = nu[c] / nd[c];
        const float temp2 = temp * (ratio * ratio);
        if (temp2 <= downdate_thr) {
          float s = 0.f;
#pragma unroll
          for (int r = k + 1; r < R; ++r) s = s + A[r][c] * A[r][c];
          nd[c] = sqrtf(s);
          nu[c] = nd[c];
        } else {
          nu[c] = nu[c] * sqrtf(temp);
        }
      }
    }
  }

  float cv[R];
#pragma unroll
  for (int r = 0; r < R; ++r) cv[r] = -1.0f;
  x[0] = x[1] = x[2] = 0.f;
  if (nonzero == 0) return;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (k < nonzero) {
      const float tau = hc[k];
      if (tau != 0.f) {
        float tmp = 0.f;
#pragma unroll
        for (int r = k + 1; r < R; ++r) tmp = tmp + A[r][k] * cv[r];
        tmp = tmp + cv[k];
        cv[k] = cv[k] - tau * tmp;
#pragma unroll
        for (int r = k + 1; r < R; ++r) cv[r] = cv[r] - tmp * (tau * A[r][k]);
      }
    }
  }
#pragma unroll
  for (int i = 2; i >= 0; --i) {
    if (i < nonzero) {
      cv[i] = cv[i] / A[i][i];
#pragma unroll
      for (int j = 0; j < i; ++j) cv[j] = cv[j] - cv[i] * A[j][i];
    }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    if (i < nonzero) {
#pragma unroll
      for (int j = 0; j < 3; ++j)
        if (perm[i] == j) x[j] = cv[i];
    }
  }
}

__device__ __forceinline__ void affine_apply(const float* __restrict__ Rm, const float* __restrict__ tv, float px,
                                             float py, float pz, float (&o)[3]) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    float acc = Rm[3 * r] * px;
    acc = Rm[3 * r + 1] * py + acc;
    acc = Rm[3 * r + 2] * pz + acc;
    o[r] = tv[r] * 1.0f + acc;
  }
}
__device__ __forceinline__ void mat3_vec(const float* __restrict__ M, const float (&v)[3], float (&o)[3]) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const float a = M[3 * r] * v[0], b = M[3 * r + 1] * v[1], c = M[3 * r + 2] * v[2];
    o[r] = a + (b + c);
  }
}

// (i,j) of the e-th entry of the row-major upper triangle of a 13x13 matrix.
__device__ __forceinline__ void tri13(int e, int& i, int& j) {
  int r = 0, base = 0;
  bool go = true;
#pragma unroll
  for (int q = 0; q < 12; ++q) {
    const int len = 13 - q;
    if (go && e >= base + len) {
      base += len;
      r = q + 1;
    } else {
      go = false;
    }
  }
  i = r;
  j = r + (e - base);
}

// pack: 13x13 triangle entry e=(i,j) -> [0..77] HTH tri (12x12), [78..89] HTh, [91] sum z^2 (flimo.h layout)
__device__ __forceinline__ int packed_slot13(int e) {
  if (e < kTriEntries) {
    int i, j;
    tri13(e, i, j);
    if (j < 12) return i * 12 - (i * (i - 1)) / 2 + (j - i);
    if (i < 12) return 78 + i;
    return 91;
  }
  if (e == 91) return 92;     // n_valid
  if (e == 92) return 90;     // n_rows
  return e;                   // 93,94,95 reserved (zero)
}

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ int match_num_tiles_dev(int n_queries) { return (n_queries + kTileQueries - 1) / kTileQueries; }

// Shared memory of one CTA (one 128-query tile at a time).
struct TileShared {
  __align__(16) float tile[kTileQueries / 32][32][24];   // rows of [row(12), z] as floats; stride 24: conflict-free fragment reads
  double wsum[kTileQueries / 32][kPartialStride];
  float4 held[kTileQueries];   // resident kernel, one tile per CTA: the tile's scan points (x, y, z, original index) stay here over the passes
  int s_last;
};

// One tile of one measurement pass: 128 queries -> partial normal equations -> deterministic tree over
// the tiles; the CTA that completes the tree publishes the pass result.  `pc` is the pose (kernel
// parameter bank in the one-launch-per-pass kernel, shared memory in the persistent kernel), `seq` the
// sequence number of the pass, `orig_limit` the first-N cap on contributing rows.
// Returns true (in every thread of the CTA) when this CTA completed the tree: the packed result of the pass
// (flimo.h layout, 96 doubles) is then staged in sh.wsum[0] and the caller decides what happens with it
// (publish_result for the host-driven kernels, the filter step in the registration kernel).
template <bool kWide, bool kPair, bool kExternalFinal = false, bool kStage = false>
__device__ __forceinline__ bool match_tile(const MatchParams& P, const PoseConsts& pc, TileShared& sh, const int tile_idx,
                                           const int n_tiles, const uint32_t orig_limit, const bool reuse_rows = false,
                                           const StageCtx stage = StageCtx{nullptr, nullptr, nullptr}, const int acc_sel = 0,
                                           const int held_mode = 0 /* 1: keep the scan points in sh.held, 2: take them from there */) {
  auto& tile = sh.tile;
  auto& wsum = sh.wsum;
  int& s_last = sh.s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // Query assignment.  The scan is stored in a pseudo-random order (pack_scan_kernel scatters the points
  // with a multiplicative permutation on upload): neighbouring scan points share sparse / dense regions
  // of the map and therefore search cost, and must not pile up in one warp.  A tile simply takes 128
  // consecutive stored points (coalesced read).  Morton-sorted scans (sort_scan) are interleaved over the
  // tiles instead.  Results do not depend on the assignment.
  const int i_lin = P.interleave ? (int)threadIdx.x * n_tiles + tile_idx : tile_idx * kTileQueries + (int)threadIdx.x;
  const int q = min(P.q_begin + i_lin, P.q_end);
  const bool in_range = q < P.q_end;

  float v13[13];
#pragma unroll
  for (int c = 0; c < 13; ++c) v13[c] = 0.f;
  bool accepted = false;   // Match::lisanAlGaib()
  uint32_t orig = 0;

  unsigned long long tm0 = 0, tm1 = 0, tm2 = 0;
  int lvl = 0;
  int first_lvl = 0;
  unsigned long long t_priv = 0;
  unsigned long long t_probe = 0;
  uint32_t first_cnt = 0;
  if (reuse_rows) {
    // Repeated pass with the SAME pose and a row limit (first-N rule): the rows of this tile are still in shared memory
    // from the first evaluation (this CTA owns exactly one tile); only their selection changes.
    orig = __float_as_uint(tile[warp][lane][16]);
    accepted = tile[warp][lane][17] != 0.f;
  } else {
    float g[3] = {0.f, 0.f, 0.f};
    if (in_range) {
      float sx, sy, sz;
      if (held_mode == 2) {                                  // the same tile as in the previous pass of this registration
        const float4 sp = sh.held[threadIdx.x];
        orig = __float_as_uint(sp.w);
        sx = sp.x; sy = sp.y; sz = sp.z;
      } else if (P.scan != nullptr) {
        const float4 sp = __ldg(&P.scan[q]);
        orig = __float_as_uint(sp.w);
        sx = sp.x; sy = sp.y; sz = sp.z;
      } else {
        // in-place read of the caller's array: original index = (q * raw_inv) mod raw_n
        const unsigned long long prod = (unsigned long long)(uint32_t)q * P.raw_inv;
        unsigned long long r = prod - __umul64hi(prod, P.raw_magic) * P.raw_n;
        if (r >= P.raw_n) r -= P.raw_n;
        orig = (uint32_t)r;
        const unsigned char* sp = P.raw_scan + (size_t)orig * P.raw_stride;
        if (P.raw_vec4) {                                    // 16-byte aligned records: one load
          const float4 v = __ldg(reinterpret_cast<const float4*>(sp));
          sx = v.x; sy = v.y; sz = v.z;
        } else {
          const float* sf = reinterpret_cast<const float*>(sp);
          sx = __ldg(sf); sy = __ldg(sf + 1); sz = __ldg(sf + 2);
        }
      }
      if (held_mode == 1) sh.held[threadIdx.x] = make_float4(sx, sy, sz, __uint_as_float(orig));
      affine_apply(pc.R_wb, pc.t_wb, sx, sy, sz, g);
    }
    Top5 t;
  #pragma unroll
    for (int j = 0; j < 5; ++j) {
      t.d[j] = __int_as_float(0x7f800000);
      t.i[j] = 0;
    }
    t.sl = kStashIdentity;
    if (P.timing) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tm0));
    float4* stash = reinterpret_cast<float4*>(&tile[warp][0][0]) + lane;   // the warp's row block is free until its rows are written below
    knn_search<kWide, kPair, kStage>(P, lane, in_range, g[0], g[1], g[2], t, lvl, first_lvl, first_cnt, t_priv, t_probe, stash, stage);      // warp-converged call
    if (P.timing) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tm1));

    if (in_range) {
      float n4[4] = {0.f, 0.f, 0.f, 0.f};
      float dist = 0.f;
      // Plane::enough_points + close_enough: an empty slot is +inf and fails the strict '<'.
      if (t.d[4] < P.max_dist_f) {
        float A[5][3];
  #pragma unroll
        for (int j = 0; j < 5; ++j) {
          const float4 nbj = stash[((t.sl >> (3 * j)) & 7u) * kStashStride];
          A[j][0] = nbj.x;
          A[j][1] = nbj.y;
          A[j][2] = nbj.z;
        }
        float x[3];
        plane_qr_solve(A, x);
        const float nrm = sqrtf(x[0] * x[0] + (x[1] * x[1] + x[2] * x[2]));
        n4[0] = x[0] / nrm;
        n4[1] = x[1] / nrm;
        n4[2] = x[2] / nrm;
        n4[3] = 1.0f / nrm;   // == float(1.0 / double(nrm)): double rounding is innocuous for division
        bool ok = true;
  #pragma unroll
        for (int j = 0; j < 5; ++j) {
          const float4 nbj = stash[((t.sl >> (3 * j)) & 7u) * kStashStride];   // (re-read: cheaper than 15 registers across the QR)
          const float res = ((n4[0] * nbj.x + n4[1] * nbj.y) + n4[2] * nbj.z) + n4[3];
          if (fabsf(res) > P.plane_thr) ok = false;
        }
        accepted = ok;
        dist = ((n4[0] * g[0] + n4[1] * g[1]) + n4[2] * g[2]) + n4[3];
      }

      if (accepted) {
        float p_imu[3], p_lid[3], C[3], RC[3];
        affine_apply(pc.Rinv_wb, pc.tinv_wb, g[0], g[1], g[2], p_imu);
        affine_apply(pc.Rinv_LI, pc.tinv_LI, p_imu[0], p_imu[1], p_imu[2], p_lid);
        const float nv[3] = {n4[0], n4[1], n4[2]};
        mat3_vec(pc.Rd_wb_inv, nv, C);
        mat3_vec(pc.Rd_LI_inv, C, RC);
        v13[0] = n4[0];
        v13[1] = n4[1];
        v13[2] = n4[2];
        v13[3] = p_imu[1] * C[2] - p_imu[2] * C[1];
        v13[4] = p_imu[2] * C[0] - p_imu[0] * C[2];
        v13[5] = p_imu[0] * C[1] - p_imu[1] * C[0];
        if (P.estimate_extrinsics) {
          v13[6] = p_lid[1] * RC[2] - p_lid[2] * RC[1];
          v13[7] = p_lid[2] * RC[0] - p_lid[0] * RC[2];
          v13[8] = p_lid[0] * RC[1] - p_lid[1] * RC[0];
          v13[9] = C[0];
          v13[10] = C[1];
          v13[11] = C[2];
        }
        v13[12] = -dist;
      }

      if (P.dbg16 != nullptr) {
        float4* o = reinterpret_cast<float4*>(P.dbg16 + (size_t)orig * 16);
        o[0] = make_float4(g[0], g[1], g[2], n4[0]);
        o[1] = make_float4(n4[1], n4[2], n4[3], dist);
        o[2] = make_float4(accepted ? 1.f : 0.f, t.d[0], t.d[1], t.d[2]);
        o[3] = make_float4(t.d[3], t.d[4], (float)(first_lvl + 16 * lvl), (float)first_cnt);
      }
      if (P.valid_by_orig != nullptr) P.valid_by_orig[orig] = accepted ? 1 : 0;
    }

  }
  if (P.timing) {
    __syncwarp();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tm2));
  }
  // ---- warp-cooperative float64 accumulation of [row,z]^T [row,z] on the FP64 tensor pipe -----------
  // The warp's 32 rows v (13 floats, zero for lanes that do not contribute) form V[32][16]; the upper
  // triangle of V^T V is three 8x8 tiles (00, 01, 11), each accumulated over the 32 rows with eight
  // mma.m8n8k4.f64 (products of two floats are exact in float64).  Fragment of step k: lane l supplies
  // V[4k + l%4][l/4] (columns 0-7) and V[4k + l%4][8 + l/4] (columns 8-15) — A and B fragments of a
  // diagonal tile coincide.  H is never written.
  const bool contributes = accepted && (orig < orig_limit);
  const unsigned int m_all = __ballot_sync(0xffffffffu, accepted);
  const unsigned int m_rows = __ballot_sync(0xffffffffu, contributes);
  {
    float4* rowp = reinterpret_cast<float4*>(&tile[warp][lane][0]);
    if (reuse_rows) {
      if (!contributes) rowp[0] = rowp[1] = rowp[2] = rowp[3] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      const float zf = contributes ? 1.f : 0.f;     // v13 is already zero unless accepted; orig_limit may veto
      rowp[0] = make_float4(v13[0] * zf, v13[1] * zf, v13[2] * zf, v13[3] * zf);
      rowp[1] = make_float4(v13[4] * zf, v13[5] * zf, v13[6] * zf, v13[7] * zf);
      rowp[2] = make_float4(v13[8] * zf, v13[9] * zf, v13[10] * zf, v13[11] * zf);
      rowp[3] = make_float4(v13[12] * zf, 0.f, 0.f, 0.f);
      rowp[4] = make_float4(__uint_as_float(orig), accepted ? 1.f : 0.f, 0.f, 0.f);   // columns 16.. are padding of the fragment layout: kept for a repeated pass
    }
  }
  __syncwarp();
  double c00[2] = {0.0, 0.0}, c01[2] = {0.0, 0.0}, c11[2] = {0.0, 0.0};
  if (m_rows != 0u) {
    const int fr = lane & 3, fc = lane >> 2;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if ((m_rows >> (4 * k)) & 0xFu) {           // warp-uniform: skip row groups without contributions
        const double lo = (double)tile[warp][4 * k + fr][fc], hi = (double)tile[warp][4 * k + fr][8 + fc];
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c00[0]), "+d"(c00[1]) : "d"(lo), "d"(lo));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c01[0]), "+d"(c01[1]) : "d"(lo), "d"(hi));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c11[0]), "+d"(c11[1]) : "d"(hi), "d"(hi));
      }
    }
  }
  // accumulator element u of a tile sits at (i, j) = (lane/4, 2*(lane%4) + u); scatter the upper triangle
  // of the 13x13 result to its row-major triangle slot e = i*13 - i(i-1)/2 + (j - i)
  {
    const int ti = lane >> 2, tj = 2 * (lane & 3);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int j = tj + u;
      if (ti <= j) wsum[warp][ti * 13 - (ti * (ti - 1)) / 2 + (j - ti)] = c00[u];                       // tile 00
      if (8 + j < 13) wsum[warp][ti * 13 - (ti * (ti - 1)) / 2 + (8 + j - ti)] = c01[u];              // tile 01
      const int gi = 8 + ti, gj = 8 + j;
      if (gi <= gj && gj < 13) wsum[warp][gi * 13 - (gi * (gi - 1)) / 2 + (gj - gi)] = c11[u];        // tile 11
    }
    if (lane < kPartialStride - kTriEntries) wsum[warp][kTriEntries + lane] = 0.0;
  }
  __syncwarp();
  if (lane == 0) {
    wsum[warp][91] = (double)__popc(m_all);    // n_valid
    wsum[warp][92] = (double)__popc(m_rows);   // n_rows
  }
  __syncthreads();

  if (P.timing && lane == 0) {
    unsigned long long tm3;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tm3));
    unsigned int smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    unsigned long long* o = P.timing + ((size_t)tile_idx * (kTileQueries / 32) + warp) * 8;
    o[0] = smid; o[1] = tm0; o[2] = tm1; o[3] = tm2; o[4] = tm3;
    o[5] = (unsigned long long)__popc(__ballot_sync(0xffffffffu, lvl != first_lvl));
    o[6] = t_priv;
    o[7] = (unsigned long long)__reduce_max_sync(0xffffffffu, first_cnt) | (t_probe << 16);
  } else if (P.timing) {
    (void)__ballot_sync(0xffffffffu, lvl != first_lvl);
    (void)__reduce_max_sync(0xffffffffu, first_cnt);
  }
  // ---- CTA partial -> order-free exact accumulation (P.fx_reduce) ------------------------------------
  // The tile's 96 sums are split into two 64-bit fixed-point words (units 2^-18 and 2^-66) and added to the pass
  // accumulators with integer REDs: integer addition is associative, so the pass sum is the same whatever order the
  // tiles finish in — deterministic like the tree below, but ONE ticket round instead of two (the finisher converts
  // the 96 sums back to float64 with a single rounding each).  Range: |tile sum| < 2^44, up to 2^15 tiles.
  if (P.fx_reduce) {
    if (threadIdx.x < kPartialStride) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < kTileQueries / 32; ++w) s += wsum[w][threadIdx.x];
      const double hi_d = rint(s * 0x1p18);
      const long long hi = (long long)hi_d;
      const long long lo = __double2ll_rn((s - hi_d * 0x1p-18) * 0x1p66);
      unsigned long long* a = fx_slot(P.ticket, acc_sel, tile_idx & (kFxReplicas - 1), threadIdx.x);
      if (hi != 0) atomicAdd(a, (unsigned long long)hi);
      if (lo != 0) atomicAdd(a + 1, (unsigned long long)lo);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned int prev = atomicAdd(&P.ticket[0], 1u);
      s_last = (prev == (unsigned int)(n_tiles - 1)) ? 1 : 0;
    }
    if (kExternalFinal) return false;   // the filter kernel waits for ticket[0] == n_tiles and reads the accumulators itself
    __syncthreads();
    if (!s_last) return false;
    __threadfence();
    if (threadIdx.x < kPartialStride) {
      const double s = fx_collect(P.ticket, acc_sel, threadIdx.x);
      wsum[0][packed_slot13(threadIdx.x)] = s;                 // stage in packed order
    }
    if (threadIdx.x == 0) P.ticket[0] = 0u;                  // self reset for the next pass / launch
    __threadfence();
    __syncthreads();
    if (P.timing && threadIdx.x == 0) {
      unsigned long long te;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(te));
      P.timing[(size_t)n_tiles * (kTileQueries / 32) * 8] = te;
    }
    return true;
  }
  // ---- CTA partial, then a deterministic two-level tree over tiles ----------------------------------
  const int n_groups = (n_tiles + kGroupTiles - 1) / kGroupTiles;
  const int group = tile_idx / kGroupTiles;
  double* tile_part = P.partials;                                   // [n_tiles][96]
  double* group_part = P.partials + (size_t)n_tiles * kPartialStride;   // [n_groups][96]
  if (threadIdx.x < kPartialStride) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kTileQueries / 32; ++w) s += wsum[w][threadIdx.x];
    __stcg(&tile_part[(size_t)tile_idx * kPartialStride + threadIdx.x], s);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int g_first = group * kGroupTiles;
    const int g_count = min(kGroupTiles, n_tiles - g_first);
    const unsigned int prev = atomicAdd(&P.ticket[kTicketWords + group], 1u);
    s_last = (prev == (unsigned int)(g_count - 1)) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  {
    const int g_first = group * kGroupTiles;
    const int g_count = min(kGroupTiles, n_tiles - g_first);
    if (threadIdx.x < kPartialStride) {
      // fixed summation order (tile index ascending); all loads of the group are in flight at once so
      // the L2 latency is paid once
      const double* base = tile_part + (size_t)g_first * kPartialStride + threadIdx.x;
      double v[kGroupTiles];
#pragma unroll
      for (int u = 0; u < kGroupTiles; ++u) v[u] = (u < g_count) ? __ldcg(base + (size_t)u * kPartialStride) : 0.0;
      double s = 0.0;
#pragma unroll
      for (int u = 0; u < kGroupTiles; ++u) s += v[u];
      __stcg(&group_part[(size_t)group * kPartialStride + threadIdx.x], s);
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    P.ticket[kTicketWords + group] = 0u;   // self reset for the next launch
    const unsigned int prev = atomicAdd(&P.ticket[0], 1u);
    s_last = (prev == (unsigned int)(n_groups - 1)) ? 1 : 0;
  }
  if (kExternalFinal) return false;   // the filter kernel waits for ticket[0] == n_groups and sums the group partials itself
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  if (threadIdx.x < kPartialStride) {
    double s = 0.0;
    for (int g0 = 0; g0 < n_groups; g0 += 32) {
      const double* base = group_part + (size_t)g0 * kPartialStride + threadIdx.x;
      const int cnt = min(32, n_groups - g0);
      double v[32];
#pragma unroll
      for (int u = 0; u < 32; ++u) v[u] = (u < cnt) ? __ldcg(base + (size_t)u * kPartialStride) : 0.0;
#pragma unroll
      for (int u = 0; u < 32; ++u) s += v[u];
    }
    wsum[0][packed_slot13((int)threadIdx.x)] = s;           // stage in packed order
  }
  if (threadIdx.x == 0) P.ticket[0] = 0u;                  // self reset for the next pass / launch
  __syncthreads();
  if (P.timing && threadIdx.x == 0) {
    unsigned long long te;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(te));
    P.timing[(size_t)n_tiles * (kTileQueries / 32) * 8] = te;
  }
  return true;
}

// Publication of a pass result by the CTA that completed the tree (host-driven kernels): device copy + the
// mapped pinned host copy.
__device__ __forceinline__ void publish_result(const MatchParams& P, TileShared& sh, const unsigned long long seq,
                                               const unsigned long long t_begin) {
  if (threadIdx.x == 93 && t_begin != 0ull) sh.wsum[0][93] = timer_span_ns(globaltimer_ns(), t_begin);   // persistent kernel: device time of the pass (ns)
  if (threadIdx.x < kPartialStride) {
    const double v = sh.wsum[0][threadIdx.x];
    P.out96[threadIdx.x] = v;
    // Mapped pinned host copy: 96 records of {value, sequence number}, each ONE 16-byte store, so a
    // record is complete as soon as its sequence word is visible — no system-scope fence, no flag.
    // Consecutive threads write consecutive records: the 1.5 KB leave the GPU as twelve full 128-byte
    // PCIe writes instead of 96 partial ones.
    if (P.host_out96 != nullptr)
      reinterpret_cast<ulonglong2*>(P.host_out96 + (size_t)(seq & 1ull) * P.host_out_alt)[threadIdx.x] =
          make_ulonglong2((unsigned long long)__double_as_longlong(v), seq);
  }
}

// One launch = one pass (the pose travels in the kernel parameters).
template <bool kWide, bool kPair>
__global__ void __launch_bounds__(kTileQueries, 7) match_reduce_kernel(const __grid_constant__ MatchParams P) {
  __shared__ TileShared sh;
  if (match_tile<kWide, kPair>(P, P.pc, sh, (int)blockIdx.x, (int)gridDim.x, P.orig_limit, false, StageCtx{nullptr, nullptr, nullptr}, (int)(P.seq & 1ull)))
    publish_result(P, sh, P.seq, 0ull);
}

// One launch = one pass, runs staged in shared memory by bulk asynchronous copies (see StageCtx above).
__global__ void __launch_bounds__(kTileQueries, 2) match_reduce_staged_kernel(const __grid_constant__ MatchParams P) {
  __shared__ TileShared sh;
  __shared__ __align__(8) unsigned long long s_bar[kTileQueries / 32];
  extern __shared__ __align__(16) float4 s_stage[];          // [warps][32 lanes][kStageStride]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) mbar_init(&s_bar[warp], 1u);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  uint32_t phase = 0u;
  const StageCtx stage{s_stage + ((size_t)warp * 32 + lane) * kStageStride, &s_bar[warp], &phase};
  if (match_tile<true, false, false, true>(P, P.pc, sh, (int)blockIdx.x, (int)gridDim.x, P.orig_limit, false, stage, (int)(P.seq & 1ull))) publish_result(P, sh, P.seq, 0ull);
}

// ---------------------------------------------------------------------------------------------------
// Persistent form: ONE launch per scan registration.  The CTAs stay resident over all passes of
// esekf::update_iterated_dyn_share_modified (esekfom.hpp:1634-1764); between passes the host runs the
// 23x23 filter algebra and hands the next pose over through MAPPED PINNED HOST MEMORY instead of a new
// launch: CTA 0 polls the host control block (one PCIe read per poll), copies the pose to device memory
// and raises a device flag; every CTA waits for that flag, takes the pose into shared memory and runs
// its tile(s).  The pass result goes back the same way it does in the per-pass kernel (tagged 16-byte
// records).  This removes the launch latency (~10 us of host + device time) from every pass.
// A watchdog ends the kernel if the host stays silent (the host then continues with per-pass launches).
// ---------------------------------------------------------------------------------------------------
template <bool kWide, bool kPair>
__global__ void __launch_bounds__(kTileQueries, 7) match_persistent_kernel(const __grid_constant__ MatchParams P) {
  __shared__ TileShared sh;
  __shared__ PassCtl s_ctl;
  const int n_tiles = match_num_tiles_dev(P.q_end - P.q_begin);
  PassCtl* dctl = P.dev_ctl;
  unsigned long long ctl = P.ctl_seq;
  for (unsigned long long want = P.seq;; ++want, ++ctl) {
    if (blockIdx.x == 0 && threadIdx.x < 32) {
      // the host's control block: 19 tagged 16-byte records, read by 19 lanes in one PCIe round trip
      const PassCtlWire* hctl = P.host_ctl;
      const int lane = (int)threadIdx.x;
      const uint32_t tag = (uint32_t)ctl;
      uint32_t r0 = 0, r1 = 0, r2 = 0, r3 = tag;
      bool ok = false;
      const long long w0 = watch_start();
      for (;;) {
        if (lane < kCtlRecords)
          asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "l"(&hctl->rec[lane][0]) : "memory");
        ok = __all_sync(0xffffffffu, r3 == tag);
        if (ok) break;
        if (__any_sync(0xffffffffu, watch_expired(w0, P.watchdog_ns))) break;
      }
      if (ok) {
        uint32_t* dst = reinterpret_cast<uint32_t*>(&dctl->cmd);          // cmd, orig_limit, pose words follow each other
        if (lane < kCtlRecords) {
          if (3 * lane < kCtlPayloadWords) dst[3 * lane] = r0;
          if (3 * lane + 1 < kCtlPayloadWords) dst[3 * lane + 1] = r1;
          if (3 * lane + 2 < kCtlPayloadWords) dst[3 * lane + 2] = r2;
        }
        if (lane == 0) dctl->t_begin = globaltimer_ns();
      } else if (lane == 0) {
        dctl->cmd = 1u;
      }
      __threadfence();
      __syncwarp();
      if (lane == 0) {
        const unsigned long long pub = ok ? ctl : (ctl | kPassAbort);   // the device flag carries the command number (never stale)
        asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(&dctl->seq), "l"(pub) : "memory");
      }
    }
    if (threadIdx.x == 0) {
      unsigned long long f;
      for (;;) {
        f = ld_acquire_u64(&dctl->seq);
        if (f == ctl || f == (ctl | kPassAbort)) break;
        __nanosleep(64);
      }
    }
    __syncthreads();
    {
      const uint32_t* src = reinterpret_cast<const uint32_t*>(dctl);
      uint32_t* dst = reinterpret_cast<uint32_t*>(&s_ctl);
      for (int w = (int)threadIdx.x; w < (int)(sizeof(PassCtl) / 4); w += kTileQueries) dst[w] = __ldcg(src + w);
    }
    __syncthreads();
    if (s_ctl.cmd != 0u) return;
    for (int t = (int)blockIdx.x; t < n_tiles; t += (int)gridDim.x)
      if (match_tile<kWide, kPair>(P, s_ctl.pc, sh, t, n_tiles, s_ctl.orig_limit, false, StageCtx{nullptr, nullptr, nullptr}, (int)(want & 1ull)))
        publish_result(P, sh, want, s_ctl.t_begin);
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------
// Registration, tile side: ONE launch = all measurement passes of one iterated update, with NO host between the
// passes.  The CTAs stay resident; the pose of the first pass travels in the parameter block, the poses of the
// following passes are posted to `dev_ctl` by the FILTER kernel (filter_kernel.cu: one CTA on a second stream that
// sums the group partials of a pass as soon as the last group is complete, runs the filter step and derives the
// next pose — or ends the update).  A CTA leaves when it reads the stop command, or after 0.5 s of silence.
// ---------------------------------------------------------------------------------------------------
template <bool kWide, bool kPair>
__global__ void __launch_bounds__(kTileQueries, 7) registration_tiles_kernel(const __grid_constant__ MatchParams P) {
  __shared__ TileShared sh;
  __shared__ PassCtl s_ctl;
  const int n_tiles = match_num_tiles_dev(P.q_end - P.q_begin);
  PassCtl* dctl = P.dev_ctl;
  const int tid = (int)threadIdx.x;
  for (unsigned long long cmd_no = 0;; ++cmd_no) {
    if (cmd_no == 0) {                                     // first pass: the pose travels in the parameter block
      const uint32_t* src = reinterpret_cast<const uint32_t*>(&P.pc);
      uint32_t* dst = reinterpret_cast<uint32_t*>(&s_ctl.pc);
      for (int w = tid; w < (int)(sizeof(PoseConsts) / 4); w += kTileQueries) dst[w] = src[w];
      if (tid == 0) {
        s_ctl.cmd = 0u;
        s_ctl.orig_limit = P.orig_limit;
      }
    } else {
      const unsigned long long want = P.ctl_seq + cmd_no - 1ull;
      if (tid == 0) {
        const unsigned long long t0 = globaltimer_ns();   // (diagnostics only)
        const long long w0 = watch_start();
        unsigned int naps = 0;
        unsigned int smid = 0;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        const uint32_t tag = (uint32_t)cmd_no | (smid << 16);
        if (P.cta_trace) P.cta_trace[blockIdx.x] = make_uint4(tag | (1u << 8), 0u, (uint32_t)want, (uint32_t)(t0 / 1000ull));
        unsigned long long seen;
        while ((seen = ld_acquire_u64(&dctl->seq)) != want) {
          __nanosleep(32);
          if ((++naps & 1023u) == 0u) {
            if (P.cta_trace) P.cta_trace[blockIdx.x] = make_uint4(tag | (1u << 8), (uint32_t)seen, (uint32_t)want, (uint32_t)(t0 / 1000ull));
            if (watch_expired(w0, P.watchdog_ns)) {        // the filter kernel is gone
              s_ctl.cmd = 2u;
              break;
            }
          }
        }
        if (s_ctl.cmd != 2u) s_ctl.cmd = 0u;
        if (P.cta_trace)
          P.cta_trace[blockIdx.x] = s_ctl.cmd == 2u ? make_uint4(tag | (9u << 8), (uint32_t)seen, naps, (uint32_t)((clock64() - w0) >> 10))
                                                    : make_uint4(tag | (2u << 8), (uint32_t)seen, (uint32_t)want, (uint32_t)(t0 / 1000ull));
      }
      __syncthreads();
      if (s_ctl.cmd == 2u) return;
      const uint32_t* src = reinterpret_cast<const uint32_t*>(dctl);
      uint32_t* dst = reinterpret_cast<uint32_t*>(&s_ctl);
      for (int w = tid; w < (int)(sizeof(PassCtl) / 4); w += kTileQueries) dst[w] = __ldcg(src + w);
    }
    __syncthreads();
    if (s_ctl.cmd != 0u && s_ctl.cmd != 3u) return;
    if (tid == 0 && P.cta_trace) atomicAdd(&P.ticket[1], 1u);   // diagnostics (FLIMO_DEBUG_CTA_TRACE): CTAs that began this pass
    // cmd 3 = the pass is repeated with the same pose and a row limit: a CTA that owns exactly one tile re-selects the rows
    // it still holds in shared memory instead of matching the tile again
    const bool one_tile = (int)gridDim.x >= n_tiles;          // this CTA owns exactly one tile: its scan points are loaded once
    const bool reuse = s_ctl.cmd == 3u && one_tile;
    const int held_mode = one_tile ? (cmd_no == 0 ? 1 : 2) : 0;
    for (int t = (int)blockIdx.x; t < n_tiles; t += (int)gridDim.x)
      match_tile<kWide, kPair, true>(P, s_ctl.pc, sh, t, n_tiles, s_ctl.orig_limit, reuse, StageCtx{nullptr, nullptr, nullptr}, (int)(cmd_no & 1ull),
                                     held_mode);
    __syncthreads();
    if (tid == 0 && P.cta_trace) P.cta_trace[blockIdx.x].x = (P.cta_trace[blockIdx.x].x & ~0xFF00u) | (3u << 8);   // delivered
  }
}

}  // namespace

int match_num_tiles(int n_queries) { return (n_queries + kTileQueries - 1) / kTileQueries; }

cudaError_t launch_match_persistent(const MatchParams& p, int grid, cudaStream_t st) {
  const int n = p.q_end - p.q_begin;
  if (n <= 0 || grid <= 0) return cudaErrorInvalidValue;
  if (p.pair_scan) match_persistent_kernel<true, true><<<grid, kTileQueries, 0, st>>>(p);
  else if (p.wide_loads) match_persistent_kernel<true, false><<<grid, kTileQueries, 0, st>>>(p);
  else match_persistent_kernel<false, false><<<grid, kTileQueries, 0, st>>>(p);
  return cudaGetLastError();
}

// CTAs of the persistent kernel that can be resident at once on the current device.
int match_persistent_capacity() {
  int dev = 0, sms = 0, per_sm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, match_persistent_kernel<true, false>, kTileQueries, 0) != cudaSuccess) return 0;
  return sms * per_sm;
}

// `after_primary`: programmatic dependent launch — the tiles follow the filter kernel IN THE SAME STREAM and may start as soon
// as the filter CTA has executed griddepcontrol.launch_dependents (its first instruction), not when it completes.  That is
// the residency guarantee of the pair: the filter CTA wants an SM of its own, and tile CTAs dispatched before it would fill
// every SM and never leave (they wait for the filter's commands) — measured: about one update in 20 000 when the two
// kernels were launched on two streams in the right order.
cudaError_t launch_registration_tiles(const MatchParams& p, int grid, cudaStream_t st, bool after_primary) {
  const int n = p.q_end - p.q_begin;
  if (n <= 0 || grid <= 0) return cudaErrorInvalidValue;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned int)grid);
  cfg.blockDim = dim3(kTileQueries);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = after_primary ? 1 : 0;
  if (p.pair_scan) return cudaLaunchKernelEx(&cfg, registration_tiles_kernel<true, true>, p);
  if (p.wide_loads) return cudaLaunchKernelEx(&cfg, registration_tiles_kernel<true, false>, p);
  return cudaLaunchKernelEx(&cfg, registration_tiles_kernel<false, false>, p);
}

// CUDA loads kernels lazily at their first launch, and that load can wait for the device to drain — which never
// happens while the resident filter CTA spins for the tiles of the very kernel being loaded.  Load every variant up front.
cudaError_t preload_match_kernels() {
  cudaFuncAttributes fa;
  cudaError_t e;
#define FL_LOAD(k) if ((e = cudaFuncGetAttributes(&fa, k)) != cudaSuccess) return e;
  FL_LOAD((registration_tiles_kernel<true, true>))
  FL_LOAD((registration_tiles_kernel<true, false>))
  FL_LOAD((registration_tiles_kernel<false, false>))
  FL_LOAD((match_persistent_kernel<true, true>))
  FL_LOAD((match_persistent_kernel<true, false>))
  FL_LOAD((match_persistent_kernel<false, false>))
  FL_LOAD((match_reduce_kernel<true, true>))
  FL_LOAD((match_reduce_kernel<true, false>))
  FL_LOAD((match_reduce_kernel<false, false>))
  FL_LOAD(match_reduce_staged_kernel)
#undef FL_LOAD
  return cudaSuccess;
}

int registration_capacity() {
  int dev = 0, sms = 0, per_sm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, registration_tiles_kernel<true, false>, kTileQueries, 0) != cudaSuccess) return 0;
  return sms * per_sm - per_sm;        // one SM's worth of slots is left to the filter kernel's CTA (it needs many registers)
}

cudaError_t launch_match(const MatchParams& p, cudaStream_t st) {
  const int n = p.q_end - p.q_begin;
  if (n <= 0) return cudaErrorInvalidValue;
  if (p.stage_runs && !p.pair_scan) {
    const size_t dyn = (size_t)(kTileQueries / 32) * 32 * kStageStride * sizeof(float4);
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
      const cudaError_t e = cudaFuncSetAttribute(match_reduce_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
      if (e != cudaSuccess) return e;
      attr_set[dev] = true;
    }
    match_reduce_staged_kernel<<<match_num_tiles(n), kTileQueries, dyn, st>>>(p);
    return cudaGetLastError();
  }
  if (p.pair_scan) match_reduce_kernel<true, true><<<match_num_tiles(n), kTileQueries, 0, st>>>(p);
  else if (p.wide_loads) match_reduce_kernel<true, false><<<match_num_tiles(n), kTileQueries, 0, st>>>(p);
  else match_reduce_kernel<false, false><<<match_num_tiles(n), kTileQueries, 0, st>>>(p);
  return cudaGetLastError();
}

}  // namespace flimo
