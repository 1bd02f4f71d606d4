// K3 — Mapper::add after the first scan: the reference's incremental-octree insert rule on the device.
//
// Reference: Octree::update / updateOctant / createOctant (fast_limo/Objects/Octree.hpp:301-432).
// Only the OUTCOME of that recursion matters for the registration path — which incoming points end
// up in the map.  SURVEY H3 derives it in closed form and the CPU oracle (the pointer octree) is the
// authority the tests compare against, point for point:
//
//   * the first batch (Octree::initialize, :282-298) fixes the lattice (root centre/extent from its
//     bounding box) and is never down-sampled;
//   * afterwards the root only doubles around the old root (expandTree, :354-371), so the lattice of
//     MIN-LEVEL cells (first depth whose half-extent <= 2*min_extent) is preserved;
//   * a node above min level is a leaf iff its cube holds <= 32 points (bucket; the YAML value is
//     ignored, :178-180), and a leaf there accepts everything (:385-405);
//   * hence an incoming point is DROPPED iff down-sampling is on AND, before this batch, its min-level
//     cell already holds > 4 points (bucket/8) AND that cell's parent cube holds > 32 (the cell exists
//     as a min-level leaf).  Decisions use the counts BEFORE the batch (the recursion partitions the
//     batch first), and every accepted point is counted afterwards.
//
// Cell membership must follow the reference's float arithmetic at cell boundaries: children are
// chosen with `p > centre` against centres computed as centre + (+-0.5f * extent) level by level
// (:269-275,:322-325), and an expanded root keeps the OLD root's stored centre for that child
// (:362-369).  descend_kernel replays exactly that walk (root chain in constant-size parameter
// memory), so membership is bit-faithful; the integer cell coordinates it produces only serve as hash keys.
//
// Counts live in one open-addressing hash table (64-bit keys: level tag + 3 x 21-bit signed cell
// coordinates), sized for 180 GB-class memory: never shrinks, rebuilt at 2x when half full.
#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>

#include "flimo_dev.cuh"

namespace flimo {

#define FL_TRY(x)                     \
  do {                                \
    cudaError_t e_ = (x);             \
    if (e_ != cudaSuccess) return e_; \
  } while (0)

namespace {

constexpr unsigned long long kEmpty = 0xFFFFFFFFFFFFFFFFull;

__host__ __device__ inline unsigned long long cell_key(int ix, int iy, int iz, int parent) {
  const unsigned long long m = (1ull << 21) - 1ull;
  return ((unsigned long long)parent << 63) | (((unsigned long long)(ix + (1 << 20)) & m) << 42) |
         (((unsigned long long)(iy + (1 << 20)) & m) << 21) | ((unsigned long long)(iz + (1 << 20)) & m);
}

__device__ inline unsigned long long mix64(unsigned long long x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
  return x;
}

__device__ inline uint32_t table_get(const unsigned long long* keys, const uint32_t* vals, uint32_t mask, unsigned long long k) {
  uint32_t slot = (uint32_t)mix64(k) & mask;
  for (;;) {
    const unsigned long long cur = keys[slot];
    if (cur == k) return vals[slot];
    if (cur == kEmpty) return 0u;
    slot = (slot + 1) & mask;
  }
}

__device__ inline void table_add(unsigned long long* keys, uint32_t* vals, uint32_t mask, unsigned long long k, uint32_t inc) {
  uint32_t slot = (uint32_t)mix64(k) & mask;
  for (;;) {
    const unsigned long long cur = atomicCAS(&keys[slot], kEmpty, k);
    if (cur == kEmpty || cur == k) {
      atomicAdd(&vals[slot], inc);
      return;
    }
    slot = (slot + 1) & mask;
  }
}

// Walk from the current root to the min-level cell of p exactly like Octree::updateOctant does.
__device__ inline void descend(const LatticeDesc& lat, float px, float py, float pz, int& ix, int& iy, int& iz) {
  float cx = lat.chain_c[lat.n_chain - 1][0], cy = lat.chain_c[lat.n_chain - 1][1], cz = lat.chain_c[lat.n_chain - 1][2];
  float ext = lat.chain_ext[lat.n_chain - 1];
  int k = lat.n_chain - 1;          // index in the chain while still on it, -1 once off
  int x = 0, y = 0, z = 0;
  const int depth = (lat.n_chain - 1) + lat.min_depth;
  for (int d = 0; d < depth; ++d) {
    const int bx = px > cx ? 1 : 0, by = py > cy ? 1 : 0, bz = pz > cz ? 1 : 0;
    const int m = bx | (by << 1) | (bz << 2);
    x = 2 * x + bx;
    y = 2 * y + by;
    z = 2 * z + bz;
    if (k > 0 && m == lat.chain_slot[k]) {          // the child that IS the older root: stored centre
      --k;
      cx = lat.chain_c[k][0];
      cy = lat.chain_c[k][1];
      cz = lat.chain_c[k][2];
      ext = lat.chain_ext[k];
    } else {
      k = -1;
      cx = __fadd_rn(cx, __fmul_rn(bx ? 0.5f : -0.5f, ext));
      cy = __fadd_rn(cy, __fmul_rn(by ? 0.5f : -0.5f, ext));
      cz = __fadd_rn(cz, __fmul_rn(bz ? 0.5f : -0.5f, ext));
      ext = __fmul_rn(ext, 0.5f);
    }
  }
  ix = x + lat.off[0];
  iy = y + lat.off[1];
  iz = z + lat.off[2];
}

__global__ void __launch_bounds__(256) decide_kernel(const float4* __restrict__ pts, size_t n, LatticeDesc lat,
                                                     const unsigned long long* __restrict__ keys,
                                                     const uint32_t* __restrict__ vals, uint32_t mask, int downsample,
                                                     int first_batch, unsigned long long* __restrict__ cell_keys,
                                                     uint8_t* __restrict__ accept) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = pts[i];
  int ix, iy, iz;
  descend(lat, p.x, p.y, p.z, ix, iy, iz);
  const unsigned long long kc = cell_key(ix, iy, iz, 0);
  cell_keys[i] = kc;
  bool ok = true;
  if (downsample && !first_batch) {
    const uint32_t c_cell = table_get(keys, vals, mask, kc);
    if (c_cell > 4u) {                                              // bucket/8 with the effective bucket of 32
      // the parent cube; with min_depth == 0 the min-level node is the original root, whose parent
      // (if any) is an expansion node and therefore always interior
      const uint32_t c_par = lat.min_depth > 0 ? table_get(keys, vals, mask, cell_key(ix >> 1, iy >> 1, iz >> 1, 1)) : 0xFFFFFFFFu;
      if (c_par > 32u) ok = false;
    }
  }
  accept[i] = ok ? 1 : 0;
}

__global__ void __launch_bounds__(256) commit_kernel(const float4* __restrict__ pts, const unsigned long long* __restrict__ cell_keys,
                                                     const uint8_t* __restrict__ accept, size_t n, unsigned long long* keys,
                                                     uint32_t* vals, uint32_t mask, float4* __restrict__ dst,
                                                     unsigned int* __restrict__ counter) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const bool keep = i < n && accept[i];
  const unsigned int m = __ballot_sync(0xffffffffu, keep);
  if (m == 0) return;
  const int lane = threadIdx.x & 31;
  unsigned int base = 0;
  if (lane == __ffs(m) - 1) base = atomicAdd(counter, __popc(m));
  base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
  if (!keep) return;
  dst[base + __popc(m & ((1u << lane) - 1u))] = pts[i];
  const unsigned long long kc = cell_keys[i];
  table_add(keys, vals, mask, kc, 1u);
  // parent key: arithmetic shift of the signed coordinates
  const unsigned long long mm = (1ull << 21) - 1ull;
  const int ix = (int)((kc >> 42) & mm) - (1 << 20), iy = (int)((kc >> 21) & mm) - (1 << 20), iz = (int)(kc & mm) - (1 << 20);
  table_add(keys, vals, mask, cell_key(ix >> 1, iy >> 1, iz >> 1, 1), 1u);
}

__global__ void __launch_bounds__(256) rehash_kernel(const unsigned long long* __restrict__ old_keys, const uint32_t* __restrict__ old_vals,
                                                     size_t old_cap, unsigned long long* keys, uint32_t* vals, uint32_t mask) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= old_cap) return;
  const unsigned long long k = old_keys[i];
  if (k != kEmpty) table_add(keys, vals, mask, k, old_vals[i]);
}

}  // namespace

// ---- host side: the lattice (root chain) exactly as Octree::initialize / expandTree build it ----------
void lattice_init(OctreeLattice& L, const float lo[3], const float hi[3], float min_extent) {
  L.min_extent = min_extent;
  L.chain.clear();
  float half[3], c[3];
  for (int a = 0; a < 3; ++a) {
    half[a] = 0.5f * (hi[a] - lo[a]);        // Octree.hpp:294
    c[a] = lo[a] + half[a];                  // :295
  }
  float e = half[0];
  if (half[1] > e) e = half[1];
  if (half[2] > e) e = half[2];
  OctreeLattice::Root r;
  std::memcpy(r.c, c, sizeof(c));
  r.ext = e;
  r.slot = 0;
  L.chain.push_back(r);
  // depth of the min level below the ORIGINAL root: first d with ext*0.5^d <= 2*min_extent (:310)
  int d = 0;
  float x = e;
  while (x > 2 * min_extent && d < 40) {
    x *= 0.5f;
    ++d;
  }
  L.min_depth = d;
  L.off[0] = L.off[1] = L.off[2] = 0;
  L.initialised = true;
}

// expandTree(boundary) (Octree.hpp:354-371)
void lattice_grow(OctreeLattice& L, const float b[3]) {
  static const float f[2] = {-0.5f, 0.5f};
  for (;;) {
    const OctreeLattice::Root& r = L.chain.back();
    float m = std::fabs(b[0] - r.c[0]);
    const float my = std::fabs(b[1] - r.c[1]), mz = std::fabs(b[2] - r.c[2]);
    if (my > m) m = my;
    if (mz > m) m = mz;
    if (!(m > r.ext)) break;
    OctreeLattice::Root up;
    up.ext = 2 * r.ext;
    int up_bits[3];
    for (int a = 0; a < 3; ++a) {
      up_bits[a] = b[a] > r.c[a] ? 1 : 0;
      up.c[a] = r.c[a] + f[up_bits[a]] * up.ext;
    }
    // child slot of the OLD root inside the new one = mortonCode(old centre, new centre) (:367)
    int slot = 0;
    for (int a = 0; a < 3; ++a)
      if (r.c[a] > up.c[a]) slot |= 1 << a;
    up.slot = slot;
    // integer offset of the new root's lower corner, in min-level cells: the old root spans
    // 2^(levels so far + min_depth) cells; if it is the upper child on an axis the corner moves down
    const int span = 1 << ((int)L.chain.size() - 1 + L.min_depth);
    for (int a = 0; a < 3; ++a)
      if ((slot >> a) & 1) L.off[a] -= span;
    L.chain.push_back(up);
    if ((int)L.chain.size() >= kMaxChain) break;     // > 2^20 x the first scan's extent: give up growing
  }
}

static void lattice_desc(const OctreeLattice& L, LatticeDesc& d) {
  d.n_chain = (int)L.chain.size();
  d.min_depth = L.min_depth;
  for (int a = 0; a < 3; ++a) d.off[a] = L.off[a];
  for (int k = 0; k < d.n_chain; ++k) {
    for (int a = 0; a < 3; ++a) d.chain_c[k][a] = L.chain[k].c[a];
    d.chain_ext[k] = L.chain[k].ext;
    d.chain_slot[k] = L.chain[k].slot;
  }
}

static cudaError_t table_reserve(CountTable& T, size_t need_entries, cudaStream_t st) {
  if (T.cap && T.cap >= 2 * need_entries) return cudaSuccess;   // load <= 1/2: keep the table
  size_t want = T.cap ? T.cap : (1u << 22);      // start large: growing means a rehash and two large (de)allocations
  while (want < 4 * need_entries) want <<= 1;    // (re)size to load <= 1/4, so that the next growth is a doubling of the map away
  unsigned long long* nk = nullptr;
  uint32_t* nv = nullptr;
  FL_TRY(cudaMalloc(&nk, want * sizeof(unsigned long long)));
  FL_TRY(cudaMalloc(&nv, want * sizeof(uint32_t)));
  FL_TRY(cudaMemsetAsync(nk, 0xFF, want * sizeof(unsigned long long), st));
  FL_TRY(cudaMemsetAsync(nv, 0, want * sizeof(uint32_t), st));
  if (T.cap) {
    rehash_kernel<<<(unsigned int)((T.cap + 255) / 256), 256, 0, st>>>(T.keys, T.vals, T.cap, nk, nv, (uint32_t)(want - 1));
    FL_TRY(cudaStreamSynchronize(st));
    cudaFree(T.keys);
    cudaFree(T.vals);
  }
  T.keys = nk;
  T.vals = nv;
  T.cap = want;
  return cudaGetLastError();
}

void table_free(CountTable& T) {
  cudaFree(T.keys);
  cudaFree(T.vals);
  T = CountTable{};
}

// Applies the insert rule to `n` packed batch points (float4, NaN already removed) and appends the
// accepted ones to dst[0..); *n_accepted receives their number.  Every call adds at most 2n table
// entries, so the table is grown up front.
cudaError_t map_insert_batch(OctreeLattice& L, CountTable& T, const float4* d_batch, size_t n, size_t map_points, int downsample, bool first_batch,
                             float4* d_dst, unsigned int* d_counter, unsigned long long* d_cell_keys, uint8_t* d_accept,
                             unsigned int* n_accepted, cudaStream_t st, uint64_t* launches) {
  *n_accepted = 0;
  if (n == 0) return cudaSuccess;
  // occupied slots <= distinct min-level cells + distinct parents <= 2 x map points (only accepted points add entries)
  T.used_bound = 2 * (map_points + n);
  FL_TRY(table_reserve(T, T.used_bound, st));
  LatticeDesc d;
  lattice_desc(L, d);
  const unsigned int blocks = (unsigned int)((n + 255) / 256);
  decide_kernel<<<blocks, 256, 0, st>>>(d_batch, n, d, T.keys, T.vals, (uint32_t)(T.cap - 1), downsample, first_batch ? 1 : 0,
                                        d_cell_keys, d_accept);
  FL_TRY(cudaMemsetAsync(d_counter, 0, sizeof(unsigned int), st));
  commit_kernel<<<blocks, 256, 0, st>>>(d_batch, d_cell_keys, d_accept, n, T.keys, T.vals, (uint32_t)(T.cap - 1), d_dst, d_counter);
  FL_TRY(cudaMemcpyAsync(n_accepted, d_counter, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
  FL_TRY(cudaStreamSynchronize(st));
  if (launches) *launches += 2;
  return cudaGetLastError();
}

}  // namespace flimo
