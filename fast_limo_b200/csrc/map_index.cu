// Device-resident map index of the registration hot path (B200 / sm_100a).
//
// Replaces the pointer-based i-Octree container of the reference
// (fast_limo/Objects/Octree.hpp:103-132, built by createOctant :301-338) with a layout made for
// HBM3e capacity and streaming access (see LevelView in flimo_dev.cuh): per level, every grid row
// owns a SUPER-ROW holding the points of its 3x3 neighbouring rows sorted by x cell, so that a
// query's 3x3x3 neighbourhood is one contiguous float4 run.  kNN stays exact for any cell size
// (match_kernel.cu escalates to the next, coarser level until the 5th distance is provably inside
// the scanned block), so the grids are a performance choice only and need not be the octree lattice.
//
// Build (per level) = 9 (row key, point id) pairs per point -> radix sort (CUB, 32-bit keys, only
// the bits needed) -> gather -> boundary scatter + suffix-min scan for the prefix table.
// All on the device; one small D2H (6 floats) sizes the grids.  Off the per-pass hot path
// (once per Mapper::add).
#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cstdlib>
#include <cmath>
#include <cstdio>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <thrust/iterator/reverse_iterator.h>

#include "flimo_dev.cuh"
#include "grid_math.cuh"

namespace flimo {

#define FL_TRY(x)                     \
  do {                                \
    cudaError_t e_ = (x);             \
    if (e_ != cudaSuccess) return e_; \
  } while (0)

// ---------------------------------------------------------------------------------------------
// order-preserving float <-> uint encoding for atomicMin/Max
__device__ __forceinline__ unsigned int f2ord(float f) {
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord2f(unsigned int u) {
  unsigned int v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
  return __uint_as_float(v);
#else
  float f;
  memcpy(&f, &v, 4);
  return f;
#endif
}

__global__ void bbox_init_kernel(unsigned int* bb) {
  if (threadIdx.x < 3) bb[threadIdx.x] = 0xFFFFFFFFu;       // mins
  else if (threadIdx.x < 6) bb[threadIdx.x] = 0u;            // maxs
}

__global__ void __launch_bounds__(256) bbox_kernel(const float4* __restrict__ p, size_t n, unsigned int* bb) {
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = p[i];
    lo[0] = fminf(lo[0], v.x); hi[0] = fmaxf(hi[0], v.x);
    lo[1] = fminf(lo[1], v.y); hi[1] = fmaxf(hi[1], v.y);
    lo[2] = fminf(lo[2], v.z); hi[2] = fmaxf(hi[2], v.z);
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      atomicMin(&bb[a], f2ord(lo[a]));
      atomicMax(&bb[3 + a], f2ord(hi[a]));
    }
  }
}

// Nine (super-row key, point id) pairs per point: the point's cell column ix in each of the 3x3
// rows around its own row.  Rows outside the grid get the sentinel key n_cells (sorted to the end).
__global__ void __launch_bounds__(256) keys9_kernel(const float4* __restrict__ p, size_t n, GridDesc g, uint32_t n_cells,
                                                    uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, uint32_t id_base = 0) {
  const size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (e >= 9 * n) return;
  const size_t i = e / 9;
  const int c = (int)(e - 9 * i);
  const float4 v = p[i];
  const int ix = cell_coord(v.x, g.ox, g.inv_cell, g.nx);
  const int iy = cell_coord(v.y, g.oy, g.inv_cell, g.ny) + (c % 3 - 1);
  const int iz = cell_coord(v.z, g.oz, g.inv_cell, g.nz) + (c / 3 - 1);
  const bool ok = iy >= 0 && iy < g.ny && iz >= 0 && iz < g.nz;
  keys[e] = ok ? (uint32_t)((iz * g.ny + iy) * g.nx + ix) : n_cells;
  vals[e] = id_base + (uint32_t)i;
}

__global__ void __launch_bounds__(256) gather_kernel(const float4* __restrict__ src, const uint32_t* __restrict__ order,
                                                     size_t n, float4* __restrict__ dst) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[order[i]];
}

// Super-row entry = point coordinates + the key of its super-row cell in .w (bit pattern): the array is sorted by that
// key, which is all the incremental merge needs (the search never looks at .w).
__global__ void __launch_bounds__(256) gather_tag_kernel(const float4* __restrict__ src, const uint32_t* __restrict__ order,
                                                         const uint32_t* __restrict__ sorted_keys, size_t n, float4* __restrict__ dst) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 v = src[order[i]];
  dst[i] = make_float4(v.x, v.y, v.z, __uint_as_float(sorted_keys[i]));
}

// Incremental update of the prefix table: every cell start moves up by the number of NEW entries whose
// cell key is smaller.  new_keys: the sorted 32-bit cell keys of the new entries.  One block shifts 1024
// consecutive cells; two binary searches bound the window of new keys that fall inside it.
__global__ void __launch_bounds__(256) table_shift_kernel(uint32_t* __restrict__ cell_start, size_t n_slots,
                                                          const uint32_t* __restrict__ new_keys, uint32_t m) {
  __shared__ uint32_t s_lo, s_hi;
  const size_t first = (size_t)blockIdx.x * 1024;
  if (threadIdx.x < 2) {
    const unsigned long long target = threadIdx.x == 0 ? first : first + 1024;   // #keys < target
    uint32_t lo = 0, hi = m;
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if ((unsigned long long)new_keys[mid] < target) lo = mid + 1; else hi = mid;
    }
    if (threadIdx.x == 0) s_lo = lo; else s_hi = lo;
  }
  __syncthreads();
  const uint32_t wlo = s_lo, whi = s_hi;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const size_t c = first + (size_t)j * 256 + threadIdx.x;
    if (c >= n_slots) continue;
    uint32_t lo = wlo, hi = whi;                      // #keys < c lies in [wlo, whi]
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if ((unsigned long long)new_keys[mid] < (unsigned long long)c) lo = mid + 1; else hi = mid;
    }
    if (lo) cell_start[c] += lo;
  }
}

// cell_start[c] = first sorted position whose key is >= c.  Occupied cells get their start here,
// empty cells are filled by the suffix-min scan that follows.
__global__ void __launch_bounds__(256) boundaries_kernel(const uint32_t* __restrict__ keys, size_t n, size_t n_cells,
                                                         uint32_t* __restrict__ cell_start) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i == 0) cell_start[n_cells + 1] = (uint32_t)n;   // slot n_cells = start of the sentinel entries
  if (i >= n) return;
  const uint32_t k = keys[i];
  if (i == 0 || keys[i - 1] != k) cell_start[k] = (uint32_t)i;
}

// ---- gapped super-rows ---------------------------------------------------------------------------------------
// Every super-row (iy, iz) of a level owns a segment [row_base[r], row_base[r] + row_cap[r]) of the entry array with
// head-room behind its entries, and nx + 1 slots of the prefix table (slot j = absolute start of cell j, slot nx = end of
// the row's entries).  A query's run [slot[ix - 1], slot[ix + 2]) never leaves its row, so the gaps are invisible to the
// search — and Mapper::add only rewrites the rows a batch touches instead of merging the whole level.

// row_first[r] = first sorted position whose key is >= r * nx  (r == n_rows: first out-of-grid sentinel)
__global__ void __launch_bounds__(256) row_first_kernel(const uint32_t* __restrict__ keys, uint32_t n, int nx, uint32_t n_rows,
                                                        uint32_t* __restrict__ row_first) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > n_rows) return;
  const unsigned long long target = (unsigned long long)r * (unsigned long long)nx;
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if ((unsigned long long)keys[mid] < target) lo = mid + 1; else hi = mid;
  }
  row_first[r] = lo;
}

// capacity of row r: its entries + 1/8 + 8, even (the search reads 32-byte pairs)
__device__ __host__ __forceinline__ uint32_t row_capacity(uint32_t cnt) { return (cnt + (cnt >> 3) + 9u) & ~1u; }

__global__ void __launch_bounds__(256) row_caps_kernel(const uint32_t* __restrict__ row_first, uint32_t n_rows, uint32_t* __restrict__ row_cap) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > n_rows) return;
  row_cap[r] = (r == n_rows) ? 0u : row_capacity(row_first[r + 1] - row_first[r]);
}

// slot nx of every row = end of its entries; the tail pointer = end of the last segment
__global__ void __launch_bounds__(256) row_ends_kernel(const uint32_t* __restrict__ row_first, const uint32_t* __restrict__ row_base, int nx,
                                                       uint32_t n_rows, uint32_t* __restrict__ cell_start, uint32_t* __restrict__ tail) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > n_rows) return;
  if (r == n_rows) {
    *tail = row_base[n_rows];
    return;
  }
  cell_start[(size_t)r * (nx + 1) + nx] = row_base[r] + (row_first[r + 1] - row_first[r]);
}

// first entry of every occupied cell -> its slot; entries -> their place inside their row's segment
__global__ void __launch_bounds__(256) row_scatter_kernel(const float4* __restrict__ src, const uint32_t* __restrict__ order,
                                                          const uint32_t* __restrict__ sorted_keys, size_t n, const uint32_t* __restrict__ row_first,
                                                          const uint32_t* __restrict__ row_base, int nx, uint32_t n_cells,
                                                          uint32_t* __restrict__ cell_start, float4* __restrict__ dst) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t k = sorted_keys[i];
  if (k >= n_cells) return;                               // rows outside the grid are not stored
  const uint32_t r = k / (uint32_t)nx;
  const uint32_t at = row_base[r] + ((uint32_t)i - row_first[r]);
  if (i == 0 || sorted_keys[i - 1] != k) cell_start[(size_t)r * (nx + 1) + (k - r * (uint32_t)nx)] = at;
  const float4 v = src[order[i]];
  dst[at] = make_float4(v.x, v.y, v.z, __uint_as_float(k));
}

// Sum over every `step`-th map point of the size of the 3x3x3 block around its own cell (= the run a query at that point
// would scan): the density measure behind the automatic choice of the finest cell.
__global__ void __launch_bounds__(256) block_size_kernel(const float4* __restrict__ p, size_t n, size_t step, GridDesc g,
                                                         const uint32_t* __restrict__ cell_start, unsigned long long* __restrict__ sum) {
  const size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t i = j * step;
  unsigned int len = 0;
  if (i < n) {
    const float4 v = p[i];
    const int ix = cell_coord(v.x, g.ox, g.inv_cell, g.nx), iy = cell_coord(v.y, g.oy, g.inv_cell, g.ny), iz = cell_coord(v.z, g.oz, g.inv_cell, g.nz);
    const size_t row = (size_t)(iz * g.ny + iy) * (size_t)(g.nx + 1);
    len = cell_start[row + min(ix + 1, g.nx - 1) + 1] - cell_start[row + max(ix - 1, 0)];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) len += __shfl_xor_sync(0xffffffffu, len, o);
  if ((threadIdx.x & 31) == 0 && len) atomicAdd(sum, (unsigned long long)len);
}

// ---- incremental update: all levels in one pass over 9 entries per new point and level -------------------------
struct LevelDev {
  float4* pts;
  uint32_t* cell_start;
  uint32_t* row_base;      // segment of every row ...
  uint32_t* row_cap;       // ... and its capacity (a row that outgrows it moves to the free tail of the array)
  uint32_t* tail;          // first free entry behind all segments
  uint32_t cap_entries;
  GridDesc g;
  uint32_t n_cells;
};
struct UpdLevels {
  LevelDev lv[kMaxLevels];
  int n_levels;
};
struct RowJob {          // one touched row of one level
  uint32_t level, row, begin, end;   // [begin, end) = its new entries in the sorted update arrays
  uint32_t cnt;                      // entries the row holds now
  uint32_t new_base, new_cap;        // new_cap != 0: the row moves to [new_base, new_base + new_cap)
  uint32_t skip;                     // the array is full: the host rebuilds the index
};

__global__ void __launch_bounds__(256) upd_keys_kernel(const float4* __restrict__ p, size_t m, UpdLevels U, uint32_t id_base,
                                                       unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals) {
  const size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t per_level = 9 * m;
  if (e >= per_level * (size_t)U.n_levels) return;
  const int l = (int)(e / per_level);
  const size_t q = e - (size_t)l * per_level;
  const size_t i = q / 9;
  const int c = (int)(q - 9 * i);
  const GridDesc& g = U.lv[l].g;
  const float4 v = p[i];
  const int ix = cell_coord(v.x, g.ox, g.inv_cell, g.nx);
  const int iy = cell_coord(v.y, g.oy, g.inv_cell, g.ny) + (c % 3 - 1);
  const int iz = cell_coord(v.z, g.oz, g.inv_cell, g.nz) + (c / 3 - 1);
  const bool ok = iy >= 0 && iy < g.ny && iz >= 0 && iz < g.nz;
  const uint32_t key = ok ? (uint32_t)((iz * g.ny + iy) * g.nx + ix) : 0xFFFFFFFFu;      // outside the grid: sorted behind the level's rows
  keys[e] = ((unsigned long long)l << 32) | key;
  vals[e] = id_base + (uint32_t)i;
}

__global__ void __launch_bounds__(256) upd_gather_kernel(const float4* __restrict__ src, const uint32_t* __restrict__ order,
                                                         const unsigned long long* __restrict__ keys, size_t n, float4* __restrict__ dst) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 v = src[order[i]];
  dst[i] = make_float4(v.x, v.y, v.z, __uint_as_float((uint32_t)keys[i]));
}

// first new entry of every (level, row) group -> a job; checks the row's head-room
__global__ void __launch_bounds__(256) upd_jobs_kernel(const unsigned long long* __restrict__ keys, uint32_t n, UpdLevels U, RowJob* __restrict__ jobs,
                                                       uint32_t* __restrict__ counters /* [0] jobs, [1] rows moved, [2] overflow */) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long k = keys[i];
  const uint32_t l = (uint32_t)(k >> 32), ck = (uint32_t)k;
  if (ck == 0xFFFFFFFFu) return;
  const uint32_t nx = (uint32_t)U.lv[l].g.nx, r = ck / nx;
  if (i > 0) {
    const unsigned long long kp = keys[i - 1];
    if ((uint32_t)(kp >> 32) == l && (uint32_t)kp != 0xFFFFFFFFu && (uint32_t)kp / nx == r) return;   // not the head of its group
  }
  const unsigned long long limit = ((unsigned long long)l << 32) | ((unsigned long long)(r + 1u) * nx);   // first key of the next row (<= 2^32 - 1 for the last one)
  uint32_t lo = i + 1, hi = n;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    const unsigned long long km = keys[mid];
    if (km < limit && (uint32_t)km != 0xFFFFFFFFu) lo = mid + 1; else hi = mid;
  }
  const uint32_t base = U.lv[l].row_base[r], cap = U.lv[l].row_cap[r];
  const uint32_t cnt = U.lv[l].cell_start[(size_t)r * (nx + 1) + nx] - base;
  const uint32_t need = cnt + (lo - i);
  RowJob j;
  j.level = l; j.row = r; j.begin = i; j.end = lo; j.cnt = cnt;
  j.new_base = 0u; j.new_cap = 0u; j.skip = 0u;
  if (need > cap) {                                        // the row outgrows its segment: it moves to the tail with room to double
    j.new_cap = (2u * need + 33u) & ~1u;
    j.new_base = atomicAdd(U.lv[l].tail, j.new_cap);
    atomicAdd(&counters[1], 1u);
    if ((unsigned long long)j.new_base + j.new_cap > (unsigned long long)U.lv[l].cap_entries) {   // array full: rebuild (compacts)
      counters[2] = 1u;
      j.skip = 1u;
    }
  }
  jobs[atomicAdd(&counters[0], 1u)] = j;
}

__device__ __forceinline__ uint32_t lower_bound_w(const float4* __restrict__ a, uint32_t n, uint32_t key) {   // #entries with cell key < key
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (__float_as_uint(a[mid].w) < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// One CTA per touched row (grid-stride over the job list): the row's entries move up by the number of new entries with a
// smaller cell key — in place, back to front in CTA-wide chunks (a chunk's destinations lie at or behind its own sources,
// never in front of a chunk that has not been read yet) — the new entries drop into the holes (old entries stay in front
// among equal keys), and the row's table slots move up by the new entries in front of them.
__global__ void __launch_bounds__(256) upd_merge_kernel(const RowJob* __restrict__ jobs, const uint32_t* __restrict__ counters, UpdLevels U,
                                                        const float4* __restrict__ new_pts) {
  const uint32_t n_jobs = counters[0];
  const uint32_t B = blockDim.x, tid = threadIdx.x;
  for (uint32_t job = blockIdx.x; job < n_jobs; job += gridDim.x) {
    const RowJob j = jobs[job];
    if (j.skip) continue;
    const LevelDev& L = U.lv[j.level];
    const uint32_t nx = (uint32_t)L.g.nx, base = L.row_base[j.row];
    const bool moves = j.new_cap != 0u;
    const float4* __restrict__ nw = new_pts + j.begin;
    const uint32_t n_new = j.end - j.begin, cnt = j.cnt;
    uint32_t* slots = L.cell_start + (size_t)j.row * (nx + 1);
    const uint32_t first_key = j.row * nx;
    if (moves) {
      const float4* src = L.pts + base;
      float4* dst = L.pts + j.new_base;
      for (uint32_t i = tid; i < cnt; i += B) {
        const float4 v = src[i];
        dst[i + lower_bound_w(nw, n_new, __float_as_uint(v.w))] = v;
      }
    } else {
      float4* row = L.pts + base;
      // entries in front of the first new entry's cell stay where they are
      const uint32_t keep = slots[__float_as_uint(nw[0].w) - first_key] - base;
      const uint32_t span = cnt - keep;
      for (uint32_t c = (span + B - 1) / B; c-- > 0;) {
        const uint32_t i = keep + c * B + tid;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t lo = 0;
        if (i < cnt) {
          v = row[i];
          lo = lower_bound_w(nw, n_new, __float_as_uint(v.w));
        }
        __syncthreads();
        if (i < cnt && lo) row[i + lo] = v;
        __syncthreads();
      }
    }
    __syncthreads();
    {
      float4* dst = L.pts + (moves ? j.new_base : base);
      for (uint32_t i = tid; i < n_new; i += B) {              // new entry i: behind the old entries of its own and all earlier cells
        const float4 v = nw[i];
        dst[i + (slots[__float_as_uint(v.w) - first_key + 1u] - base)] = v;
      }
    }
    __syncthreads();
    const uint32_t shift = moves ? j.new_base - base : 0u;   // (modular arithmetic: new_base may lie below base)
    for (uint32_t c = tid; c <= nx; c += B) {                  // slot c moves up by the new entries of cells < c
      const uint32_t lo = lower_bound_w(nw, n_new, first_key + c);
      if (lo | shift) slots[c] += lo + shift;
    }
    if (moves && tid == 0) {
      L.row_base[j.row] = j.new_base;
      L.row_cap[j.row] = j.new_cap;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) pack_points_kernel(const unsigned char* __restrict__ src, size_t n, size_t stride,
                                                          float4* __restrict__ dst, unsigned int* __restrict__ count) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  bool keep = false;
  float x = 0, y = 0, z = 0;
  if (i < n) {
    const float* p = reinterpret_cast<const float*>(src + i * stride);
    x = p[0]; y = p[1]; z = p[2];
    keep = !(isnan(x) || isnan(y) || isnan(z));      // Octree::processPoints (Octree.hpp:243)
  }
  // warp-aggregated append (order inside the map array is irrelevant: it is re-sorted by cell)
  const unsigned int m = __ballot_sync(0xffffffffu, keep);
  if (m == 0) return;
  const int lane = threadIdx.x & 31;
  unsigned int base = 0;
  if (lane == __ffs(m) - 1) base = atomicAdd(count, __popc(m));
  base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
  if (keep) dst[base + __popc(m & ((1u << lane) - 1))] = make_float4(x, y, z, 0.f);
}

// Scan upload: float4 with w = original index.  perm_stride != 0 stores point i at position
// (i * perm_stride) mod n (stride co-prime with n, ~0.618 n): a fixed pseudo-random order, so that
// the tiles of the match kernel mix points of every region of the scan.
__global__ void __launch_bounds__(256) pack_scan_kernel(const unsigned char* __restrict__ src, size_t n, size_t stride,
                                                        unsigned int perm_stride, float4* __restrict__ dst) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = reinterpret_cast<const float*>(src + i * stride);
  const size_t at = perm_stride ? (size_t)(((unsigned long long)i * perm_stride) % (unsigned long long)n) : i;
  dst[at] = make_float4(p[0], p[1], p[2], __uint_as_float((unsigned int)i));
}

__device__ __forceinline__ uint32_t spread10(uint32_t v) {
  v &= 0x3ffu;
  v = (v | (v << 16)) & 0x030000FFu;
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

__global__ void __launch_bounds__(256) transform_kernel(const float4* __restrict__ scan, size_t n, PoseConsts pc,
                                                        float* __restrict__ out) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 v = scan[i];
  const unsigned int orig = __float_as_uint(v.w);
  float g[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    float acc = __fmul_rn(pc.R_wb[3 * r], v.x);
    acc = __fadd_rn(__fmul_rn(pc.R_wb[3 * r + 1], v.y), acc);
    acc = __fadd_rn(__fmul_rn(pc.R_wb[3 * r + 2], v.z), acc);
    g[r] = __fadd_rn(pc.t_wb[r], acc);
  }
  out[3 * (size_t)orig] = g[0];
  out[3 * (size_t)orig + 1] = g[1];
  out[3 * (size_t)orig + 2] = g[2];
}

// ---------------------------------------------------------------------------------------------
static inline unsigned int nblk(size_t n, int t = 256) { return (unsigned int)((n + t - 1) / t); }

cudaError_t pack_points(const void* d_src, size_t n, size_t stride_bytes, float4* dst, unsigned int* d_count,
                        cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  pack_points_kernel<<<nblk(n), 256, 0, st>>>(static_cast<const unsigned char*>(d_src), n, stride_bytes, dst, d_count);
  return cudaGetLastError();
}

cudaError_t pack_scan(const void* d_src, size_t n, size_t stride_bytes, float4* dst, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  pack_scan_kernel<<<nblk(n), 256, 0, st>>>(static_cast<const unsigned char*>(d_src), n, stride_bytes, 0u, dst);
  return cudaGetLastError();
}

// pcl::transformPointCloud of a strided xyz cloud, original order (same float operation order as the measurement pass).
__global__ void __launch_bounds__(256) transform_raw_kernel(const unsigned char* __restrict__ src, size_t n, size_t stride, const PoseConsts pc,
                                                            float* __restrict__ out) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = reinterpret_cast<const float*>(src + i * stride);
  const float x = p[0], y = p[1], z = p[2];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    float acc = __fmul_rn(pc.R_wb[3 * r], x);
    acc = __fadd_rn(__fmul_rn(pc.R_wb[3 * r + 1], y), acc);
    acc = __fadd_rn(__fmul_rn(pc.R_wb[3 * r + 2], z), acc);
    out[3 * i + r] = __fadd_rn(pc.t_wb[r], acc);
  }
}

cudaError_t transform_raw(const unsigned char* d_src, size_t n, size_t stride_bytes, const PoseConsts& pc, float* d_out_xyz, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  transform_raw_kernel<<<(unsigned int)((n + 255) / 256), 256, 0, st>>>(d_src, n, stride_bytes, pc, d_out_xyz);
  return cudaGetLastError();
}

cudaError_t transform_scan(const float4* scan, size_t n, const PoseConsts& pc, float* d_out_xyz, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  transform_kernel<<<nblk(n), 256, 0, st>>>(scan, n, pc, d_out_xyz);
  return cudaGetLastError();
}

static cudaError_t ensure(void** p, size_t* cap, size_t need) {
  if (*cap >= need) return cudaSuccess;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  const size_t want = need + need / 2 + (1u << 20);      // head-room: the requirement creeps up with every Mapper::add
  FL_TRY(cudaMalloc(p, want));
  *cap = want;
  return cudaSuccess;
}

// Scan upload pipeline (all launches on `st`, capturable into a CUDA graph):
//   pack (+ Morton key of the body-frame position) -> radix sort of (key, index) -> gather.
// The key has 9 bits per axis at 0.5 m resolution in a +-128 m cube: neighbouring scan points end up in
// neighbouring tiles, which is all the ordering is for (results do not depend on it).
__global__ void __launch_bounds__(256) pack_scan_keys_kernel(const unsigned char* __restrict__ src, size_t n, size_t stride,
                                                             float4* __restrict__ dst, uint32_t* __restrict__ keys,
                                                             uint32_t* __restrict__ vals) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = reinterpret_cast<const float*>(src + i * stride);
  const float x = p[0], y = p[1], z = p[2];
  dst[i] = make_float4(x, y, z, __uint_as_float((unsigned int)i));
  auto q = [](float c) {
    float u = (c + 128.0f) * 2.0f;
    u = isnan(u) ? 0.f : fminf(fmaxf(u, 0.f), 511.f);
    return (uint32_t)u;
  };
  keys[i] = spread10(q(x)) | (spread10(q(y)) << 1) | (spread10(q(z)) << 2);
  vals[i] = (uint32_t)i;
}

cudaError_t scan_prepare(const void* d_src, size_t n, size_t stride_bytes, bool sort, unsigned int perm_stride, float4* scan, float4* tmp, void** cub_tmp,
                         size_t* cub_tmp_bytes, uint32_t** keys, size_t* keys_cap, cudaStream_t st, uint64_t* launches) {
  if (n == 0) return cudaSuccess;
  if (!sort || n < 2) {
    pack_scan_kernel<<<nblk(n), 256, 0, st>>>(static_cast<const unsigned char*>(d_src), n, stride_bytes, perm_stride, scan);
    if (launches) *launches += 1;
    return cudaGetLastError();
  }
  uint32_t *k0 = *keys, *k1 = k0 + n, *v0 = k1 + n, *v1 = v0 + n;
  pack_scan_keys_kernel<<<nblk(n), 256, 0, st>>>(static_cast<const unsigned char*>(d_src), n, stride_bytes, tmp, k0, v0);
  cub::DoubleBuffer<uint32_t> dk(k0, k1), dv(v0, v1);
  size_t bytes = *cub_tmp_bytes;
  FL_TRY(cub::DeviceRadixSort::SortPairs(*cub_tmp, bytes, dk, dv, (int)n, 0, 27, st));
  gather_kernel<<<nblk(n), 256, 0, st>>>(tmp, dv.Current(), n, scan);
  if (launches) *launches += 7;
  return cudaGetLastError();
}

// Sizes the scratch buffers of scan_prepare for n points (allocation must not happen during capture).
cudaError_t scan_prepare_reserve(size_t n, void** cub_tmp, size_t* cub_tmp_bytes, uint32_t** keys, size_t* keys_cap) {
  if (n < 2) return cudaSuccess;
  FL_TRY(ensure(reinterpret_cast<void**>(keys), keys_cap, 4 * n * sizeof(uint32_t)));
  cub::DoubleBuffer<uint32_t> dk(*keys, *keys + n), dv(*keys + 2 * n, *keys + 3 * n);
  size_t bytes = 0;
  FL_TRY(cub::DeviceRadixSort::SortPairs(nullptr, bytes, dk, dv, (int)n, 0, 27, (cudaStream_t)0));
  FL_TRY(ensure(cub_tmp, cub_tmp_bytes, bytes));
  return cudaSuccess;
}

cudaError_t map_index_reserve(MapIndex& idx, size_t n) {
  if (n > idx.cap_pts) {
    // Capacity: 1 M points to start with, then the requested size plus 25 % head-room (every growth re-allocates the 54
    // super-row copies of the map and forces a full rebuild, so it should be rare — but doubling made a 20 M-point map pay for
    // 32 M: 43.8 GB of index instead of 34).
    size_t cap = idx.cap_pts ? idx.cap_pts : ((size_t)1 << 20);
    if (cap < n) cap = n + n / 4;
    float4* np = nullptr;
    FL_TRY(cudaMalloc(&np, cap * sizeof(float4)));
    if (idx.pts) {
      if (idx.n_pts) FL_TRY(cudaMemcpy(np, idx.pts, idx.n_pts * sizeof(float4), cudaMemcpyDeviceToDevice));
      cudaFree(idx.pts);
    }
    idx.pts = np;
    idx.cap_pts = cap;
  }
  if (9 * idx.cap_pts > idx.cap_scratch) {
    if (idx.keys) cudaFree(idx.keys);
    idx.keys = nullptr;
    idx.cap_scratch = 0;
    const size_t cap = 9 * idx.cap_pts;
    FL_TRY(cudaMalloc(&idx.keys, 4 * cap * sizeof(uint32_t)));
    idx.keys_alt = idx.keys + cap;
    idx.vals = idx.keys_alt + cap;
    idx.vals_alt = idx.vals + cap;
    idx.cap_scratch = cap;
  }
  if (!idx.bbox) FL_TRY(cudaMalloc(&idx.bbox, 8 * sizeof(float)));
  return cudaSuccess;
}

void map_index_free(MapIndex& idx) {
  cudaFree(idx.pts);
  for (auto& l : idx.lv) {
    cudaFree(l.pts);
    cudaFree(l.cell_start);
    cudaFree(l.row_base);
    cudaFree(l.row_cap);
  }
  cudaFree(idx.upd_buf);
  cudaFree(idx.upd_counters);
  cudaFree(idx.row_first);
  cudaFree(idx.keys);
  cudaFree(idx.cub_tmp);
  cudaFree(idx.bbox);
  idx = MapIndex{};
}

cudaError_t points_bbox(const float4* d_pts, size_t n, float* d_scratch8, float lo[3], float hi[3], cudaStream_t st) {
  unsigned int* bb = reinterpret_cast<unsigned int*>(d_scratch8);
  bbox_init_kernel<<<1, 32, 0, st>>>(bb);
  bbox_kernel<<<min(nblk(n), 148u * 8u), 256, 0, st>>>(d_pts, n, bb);
  unsigned int hb[6];
  FL_TRY(cudaMemcpyAsync(hb, bb, sizeof(hb), cudaMemcpyDeviceToHost, st));
  FL_TRY(cudaStreamSynchronize(st));
  for (int a = 0; a < 3; ++a) {
    lo[a] = ord2f(hb[a]);
    hi[a] = ord2f(hb[3 + a]);
  }
  return cudaGetLastError();
}

static GridDesc make_grid(const float lo[3], const float hi[3], float cell) {
  GridDesc g{};
  g.cell = cell;
  g.inv_cell = 1.0f / cell;
  g.ox = lo[0] - 0.5f * cell;
  g.oy = lo[1] - 0.5f * cell;
  g.oz = lo[2] - 0.5f * cell;
  g.nx = (int)std::min(2.0e9, std::floor(((double)hi[0] - g.ox) / cell) + 2);
  g.ny = (int)std::min(2.0e9, std::floor(((double)hi[1] - g.oy) / cell) + 2);
  g.nz = (int)std::min(2.0e9, std::floor(((double)hi[2] - g.oz) / cell) + 2);
  return g;
}

static cudaError_t build_level(MapIndex& idx, LevelIndex& L, const GridDesc& g, cudaStream_t st, uint64_t* launches, double* alloc_us = nullptr,
                               unsigned long long* d_block_sum = nullptr, size_t sample_step = 1) {
  const auto t_alloc0 = std::chrono::steady_clock::now();
  const size_t n = idx.n_pts, n9 = 9 * n;
  const size_t n_cells = (size_t)g.nx * g.ny * g.nz;
  const size_t n_rows = (size_t)g.ny * g.nz, n_slots = n_rows * (size_t)(g.nx + 1);
  if (n_slots + 2 > L.cap_slots) {
    const bool had = L.cell_start != nullptr;
    if (L.cell_start) cudaFree(L.cell_start);
    L.cell_start = nullptr;
    L.cap_slots = 0;
    // the grid grows with the map's bounding box: exact the first time (floor 16 M slots), doubled when it has to grow
    const size_t cap = std::max((had ? 2 * n_slots : n_slots) + 1024, (size_t)1 << 24);
    FL_TRY(cudaMalloc(&L.cell_start, cap * sizeof(uint32_t)));
    L.cap_slots = cap;
  }
  if (n_rows + 2 > L.cap_rows) {
    cudaFree(L.row_base);
    cudaFree(L.row_cap);
    L.row_base = L.row_cap = nullptr;
    L.cap_rows = 0;
    const size_t cap = n_rows + n_rows / 2 + 1024;
    FL_TRY(cudaMalloc(&L.row_base, cap * sizeof(uint32_t)));
    FL_TRY(cudaMalloc(&L.row_cap, cap * sizeof(uint32_t)));
    L.cap_rows = cap;
  }
  if (n_rows + 2 > idx.row_first_cap) {
    cudaFree(idx.row_first);
    idx.row_first = nullptr;
    idx.row_first_cap = 0;
    const size_t cap = n_rows + n_rows / 2 + 1024;
    FL_TRY(cudaMalloc(&idx.row_first, cap * sizeof(uint32_t)));
    idx.row_first_cap = cap;
  }
  // Room for the entries with every row's head-room (entries + 1/8 + 10 per row), plus a free tail of 1/8 for rows that
  // outgrow their segment — all of it below 2^32 entries.  The first allocation is exact (a static map wastes nothing); a
  // map that has outgrown its arrays once will grow again, so a re-allocation takes 2.5 x: the free tail then absorbs ~40 % of
  // growth (moved rows leave holes behind) before the next rebuild compacts the level.
  // (Maps below a million points are given the room of a million: allocation is the expensive part of a rebuild.)
  // A rebuild after growth keeps the arrays as long as the segments plus a small tail (1/32) still fit: the rebuild compacts
  // (rows that moved left holes behind), so a map that was allocated exactly and then grows by a few per cent is not
  // re-allocated — (de)allocating gigabytes took hundreds of milliseconds in the stream replays.
  const size_t segs = n9 + n9 / 8 + 10 * n_rows;
  const size_t need = segs + std::max(segs / 8, (size_t)1 << 16);
  const size_t min_need = segs + std::max(segs / 32, (size_t)1 << 14);
  const size_t floor_entries = 9 * std::min(idx.cap_pts, (size_t)1 << 20) * 5 / 4 + 10 * n_rows;
  if (need >= 0xFFFFFFF0ull / 3) return cudaErrorInvalidValue;
  if (min_need > L.cap_entries) {
    const size_t want = std::max(L.pts ? 2 * need + need / 2 : need, floor_entries);
    if (L.pts) cudaFree(L.pts);
    L.pts = nullptr;
    L.cap_entries = 0;
    FL_TRY(cudaMalloc(&L.pts, (want + 8) * sizeof(float4)));     // + slack: the search reads whole 32-byte pairs
    L.cap_entries = want;
  }
  L.g = g;
  L.n_cells = n_cells;
  L.n_rows = n_rows;
  if (alloc_us) *alloc_us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_alloc0).count();
  keys9_kernel<<<nblk(n9), 256, 0, st>>>(idx.pts, n, g, (uint32_t)n_cells, idx.keys, idx.vals);
  int bits = 1;
  while (bits < 32 && ((size_t)1 << bits) <= n_cells) ++bits;      // keys go up to n_cells inclusive
  cub::DoubleBuffer<uint32_t> dk(idx.keys, idx.keys_alt), dv(idx.vals, idx.vals_alt);
  size_t bytes = 0, bytes2 = 0, bytes3 = 0;
  FL_TRY(cub::DeviceRadixSort::SortPairs(nullptr, bytes, dk, dv, (int)n9, 0, bits, st));
  auto rb = thrust::make_reverse_iterator(L.cell_start + n_slots);
  FL_TRY(cub::DeviceScan::InclusiveScan(nullptr, bytes2, rb, rb, cub::Min(), (long long)n_slots, st));
  FL_TRY(cub::DeviceScan::ExclusiveSum(nullptr, bytes3, L.row_cap, L.row_base, (int)(n_rows + 1), st));
  bytes = std::max(bytes, std::max(bytes2, bytes3));
  FL_TRY(ensure(&idx.cub_tmp, &idx.cub_tmp_bytes, bytes));
  FL_TRY(cub::DeviceRadixSort::SortPairs(idx.cub_tmp, bytes, dk, dv, (int)n9, 0, bits, st));
  // rows: first sorted entry, capacity, segment start (exclusive sum of the capacities)
  const unsigned int rblk = nblk(n_rows + 1);
  row_first_kernel<<<rblk, 256, 0, st>>>(dk.Current(), (uint32_t)n9, g.nx, (uint32_t)n_rows, idx.row_first);
  row_caps_kernel<<<rblk, 256, 0, st>>>(idx.row_first, (uint32_t)n_rows, L.row_cap);
  FL_TRY(cub::DeviceScan::ExclusiveSum(idx.cub_tmp, bytes, L.row_cap, L.row_base, (int)(n_rows + 1), st));
  // table: occupied cells and row ends are written, empty cells take the next start (suffix minimum: the segments of a
  // fresh build lie in row order)
  FL_TRY(cudaMemsetAsync(L.cell_start, 0xFF, n_slots * sizeof(uint32_t), st));
  row_ends_kernel<<<rblk, 256, 0, st>>>(idx.row_first, L.row_base, g.nx, (uint32_t)n_rows, L.cell_start, L.row_cap + n_rows + 1);
  row_scatter_kernel<<<nblk(n9), 256, 0, st>>>(idx.pts, dv.Current(), dk.Current(), n9, idx.row_first, L.row_base, g.nx, (uint32_t)n_cells,
                                                L.cell_start, L.pts);
  FL_TRY(cub::DeviceScan::InclusiveScan(idx.cub_tmp, bytes, rb, rb, cub::Min(), (long long)n_slots, st));
  if (d_block_sum) {
    FL_TRY(cudaMemsetAsync(d_block_sum, 0, sizeof(unsigned long long), st));
    block_size_kernel<<<nblk((n + sample_step - 1) / sample_step), 256, 0, st>>>(idx.pts, n, sample_step, g, L.cell_start, d_block_sum);
  }
  L.n_entries = n9;
  if (launches) *launches += 12;
  return cudaGetLastError();
}

cudaError_t map_index_build(MapIndex& idx, float cell0, float ratio, float coarsest_min, size_t max_cells, cudaStream_t st,
                            uint64_t* launches) {
  const size_t n = idx.n_pts;
  if (n == 0 || 9 * n > 0x7FFFFFF0ull) return cudaErrorInvalidValue;
  if (n > idx.cap_pts || 9 * n > idx.cap_scratch) return cudaErrorInvalidValue;   // caller reserves
  unsigned int* bb = reinterpret_cast<unsigned int*>(idx.bbox);
  bbox_init_kernel<<<1, 32, 0, st>>>(bb);
  bbox_kernel<<<min(nblk(n), 148u * 8u), 256, 0, st>>>(idx.pts, n, bb);
  unsigned int hb[6];
  FL_TRY(cudaMemcpyAsync(hb, bb, sizeof(hb), cudaMemcpyDeviceToHost, st));
  FL_TRY(cudaStreamSynchronize(st));
  if (launches) *launches += 2;
  for (int a = 0; a < 3; ++a) {
    idx.lo[a] = ord2f(hb[a]);
    idx.hi[a] = ord2f(hb[3 + a]);
  }
  // The grids are laid out for the bounding box plus a margin, so that a map that grows at its rim keeps
  // its grids (and the incremental update applies) for many scans.
  for (int a = 0; a < 3; ++a) {
    const float ext = idx.hi[a] - idx.lo[a];
    const float margin = std::min(32.0f, std::max(4.0f, 0.10f * ext));
    idx.glo[a] = idx.lo[a] - margin;
    idx.ghi[a] = idx.hi[a] + margin;
  }
  // cell0 <= 0: the finest cell is chosen from the map's density.  Level 0 is built with 0.25 m; if the 3x3x3 block around a
  // map point then holds more than 45 candidates on average (every query scans long runs) it is rebuilt once with a cell
  // that brings the average back to ~28 (surfaces: candidates scale with the cell's area).  Results never depend on the
  // cells (the search is exact for any grid); this is a speed choice only.
  bool auto_cell = !(cell0 > 0.f);
  if (auto_cell) cell0 = 0.25f;
  if (!(ratio > 1.05f)) ratio = 1.5f;
  auto fit_budget = [&](float c) {
    for (;;) {   // finest level must fit the table budget
      const GridDesc g = make_grid(idx.glo, idx.ghi, c);
      if ((double)g.nx * g.ny * (g.nz + 1.0) <= (double)max_cells) return c;
      c *= 1.25992105f;
    }
  };
  cell0 = fit_budget(cell0);
  if (!idx.upd_counters) FL_TRY(cudaMalloc(&idx.upd_counters, 8 * sizeof(uint32_t)));   // [0..3] update counters, [4..5] density sum (u64)
  int nl = 0;
  float cell = cell0;
  static const bool prof = std::getenv("FLIMO_PROFILE_INDEX") != nullptr;
  for (;;) {
    if (nl == kMaxLevels - 1 && cell < coarsest_min) cell = coarsest_min;   // force termination
    const auto t0 = std::chrono::steady_clock::now();
    double alloc_us = 0.0;
    const bool probe_density = auto_cell && nl == 0;
    const size_t sample_step = std::max<size_t>(1, n >> 18);   // ~260 k samples
    FL_TRY(build_level(idx, idx.lv[nl], make_grid(idx.glo, idx.ghi, cell), st, launches, &alloc_us,
                       probe_density ? reinterpret_cast<unsigned long long*>(idx.upd_counters + 4) : nullptr, sample_step));
    if (probe_density) {
      auto_cell = false;                                     // one retry at most
      unsigned long long block_sum = 0;
      FL_TRY(cudaMemcpyAsync(&block_sum, idx.upd_counters + 4, sizeof(block_sum), cudaMemcpyDeviceToHost, st));
      FL_TRY(cudaStreamSynchronize(st));
      const double mean = (double)block_sum / (double)((n + sample_step - 1) / sample_step);
      if (prof) std::fprintf(stderr, "[index] %.1f candidates per 3x3x3 block around a map point at %.3f m\n", mean, cell0);
      if (mean > 45.0) {
        const float finer = fit_budget(std::max(0.125f, cell0 * (float)std::sqrt(28.0 / mean)));
        if (finer < 0.9f * cell0) {
          if (prof) std::fprintf(stderr, "[index] %.1f candidates per 3x3x3 block at %.3f m: finest cell -> %.3f m\n", mean, cell0, finer);
          cell0 = cell = finer;
          continue;                                          // rebuild level 0
        }
      }
    }
    if (prof) {
      const auto t1 = std::chrono::steady_clock::now();
      cudaStreamSynchronize(st);
      const auto t2 = std::chrono::steady_clock::now();
      std::fprintf(stderr, "[index] level %d cell %.3f n %zu cells %zu entries cap %zu: host %.0f us (allocation %.0f us), +sync %.0f us\n", nl, cell, n,
                   idx.lv[nl].n_cells, idx.lv[nl].cap_entries, std::chrono::duration<double, std::micro>(t1 - t0).count(), alloc_us,
                   std::chrono::duration<double, std::micro>(t2 - t1).count());
    }
    ++nl;
    if (cell >= coarsest_min || nl == kMaxLevels) break;
    cell *= ratio;
  }
  idx.n_levels = nl;
  return cudaSuccess;
}


bool map_index_can_update(const MapIndex& idx, size_t old_n, const float batch_lo[3], const float batch_hi[3]) {
  if (idx.n_levels <= 0 || old_n == 0 || idx.n_pts <= old_n) return false;
  const size_t m = idx.n_pts - old_n;
  if (4 * m > old_n) return false;                                  // a large batch: the full rebuild is as cheap
  if (idx.n_pts > idx.cap_pts) return false;
  if (9 * m * (size_t)idx.n_levels >= 0x7FFFFFF0ull) return false;
  for (int a = 0; a < 3; ++a)
    if (!(batch_lo[a] >= idx.glo[a] && batch_hi[a] <= idx.ghi[a])) return false;   // also rejects NaN boxes
  for (int l = 0; l < idx.n_levels; ++l) {
    const LevelIndex& L = idx.lv[l];
    if (!L.pts || !L.row_base || L.n_entries != 9 * old_n) return false;
  }
  return true;
}

cudaError_t map_index_update(MapIndex& idx, size_t old_n, cudaStream_t st, uint64_t* launches, bool* full) {
  const size_t n = idx.n_pts, m = n - old_n, E = 9 * m * (size_t)idx.n_levels;
  *full = false;
  // scratch: keys (2 x u64), ids (2 x u32), entries (float4), jobs
  const size_t off_keys = 0, off_vals = off_keys + 2 * E * sizeof(unsigned long long), off_pts = (off_vals + 2 * E * sizeof(uint32_t) + 15) & ~(size_t)15,
               off_jobs = off_pts + E * sizeof(float4), total = off_jobs + E * sizeof(RowJob);
  FL_TRY(ensure(&idx.upd_buf, &idx.upd_bytes, total));
  if (!idx.upd_counters) FL_TRY(cudaMalloc(&idx.upd_counters, 8 * sizeof(uint32_t)));
  unsigned char* buf = static_cast<unsigned char*>(idx.upd_buf);
  unsigned long long* k0 = reinterpret_cast<unsigned long long*>(buf + off_keys);
  uint32_t* v0 = reinterpret_cast<uint32_t*>(buf + off_vals);
  float4* upd_pts = reinterpret_cast<float4*>(buf + off_pts);
  RowJob* jobs = reinterpret_cast<RowJob*>(buf + off_jobs);
  UpdLevels U{};
  U.n_levels = idx.n_levels;
  for (int l = 0; l < idx.n_levels; ++l) {
    LevelIndex& L = idx.lv[l];
    U.lv[l].pts = L.pts;
    U.lv[l].cell_start = L.cell_start;
    U.lv[l].row_base = L.row_base;
    U.lv[l].row_cap = L.row_cap;
    U.lv[l].tail = L.row_cap + L.n_rows + 1;
    U.lv[l].cap_entries = (uint32_t)L.cap_entries;
    U.lv[l].g = L.g;
    U.lv[l].n_cells = (uint32_t)L.n_cells;
  }
  FL_TRY(cudaMemsetAsync(idx.upd_counters, 0, 4 * sizeof(uint32_t), st));
  // 1. the nine (level, super-row cell) keys of every new point on every level, sorted (stable: (key, id) order)
  upd_keys_kernel<<<nblk(E), 256, 0, st>>>(idx.pts + old_n, m, U, (uint32_t)old_n, k0, v0);
  int lbits = 1;
  while ((1 << lbits) < idx.n_levels) ++lbits;
  cub::DoubleBuffer<unsigned long long> dk(k0, k0 + E);
  cub::DoubleBuffer<uint32_t> dv(v0, v0 + E);
  size_t bytes = 0;
  FL_TRY(cub::DeviceRadixSort::SortPairs(nullptr, bytes, dk, dv, (int)E, 0, 32 + lbits, st));
  FL_TRY(ensure(&idx.cub_tmp, &idx.cub_tmp_bytes, bytes));
  FL_TRY(cub::DeviceRadixSort::SortPairs(idx.cub_tmp, bytes, dk, dv, (int)E, 0, 32 + lbits, st));
  upd_gather_kernel<<<nblk(E), 256, 0, st>>>(idx.pts, dv.Current(), dk.Current(), E, upd_pts);
  // 2. one job per touched row, 3. one CTA per job
  upd_jobs_kernel<<<nblk(E), 256, 0, st>>>(dk.Current(), (uint32_t)E, U, jobs, idx.upd_counters);
  upd_merge_kernel<<<148 * 8, 256, 0, st>>>(jobs, idx.upd_counters, U, upd_pts);
  uint32_t hc[4] = {0, 0, 0, 0};
  FL_TRY(cudaMemcpyAsync(hc, idx.upd_counters, sizeof(hc), cudaMemcpyDeviceToHost, st));
  FL_TRY(cudaStreamSynchronize(st));
  for (int l = 0; l < idx.n_levels; ++l) idx.lv[l].n_entries = 9 * n;
  idx.rows_moved += hc[1];
  *full = hc[2] != 0u;
  if (launches) *launches += 9;
  return cudaGetLastError();
}

}  // namespace flimo
