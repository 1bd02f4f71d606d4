// K4 — scan preparation on the device: everything Localizer::updatePointCloud does to a raw LiDAR
// message before the registration hot path sees it, so that pc2match is born in
// HBM and never crosses PCIe again:
//   point filters     fast_limo/Modules/Localizer.cpp:262-302   NaN removal, negative crop box, min
//                     distance, rate sampling on the index of the cropped cloud, field of view
//   time sort         Localizer.cpp:744-789                     (stable; the reference's partial_sort_copy is not)
//   deskew            Localizer.cpp:822-843 + State::update (fast_limo/Objects/State.cpp:76-119),
//                     binary_search_tailored (fast_limo/Utils/Algorithms.hpp:25-38)
//   voxel grid        Localizer.cpp:313-321 = pcl::VoxelGrid (PCL 1.10 voxel_grid.hpp:211-400): centroid per
//                     leaf-sized voxel, output in ascending voxel index
// All stages are order preserving (prefix sums, stable radix sorts, one thread walks one voxel), so the
// output sequence equals the CPU restatement's (oracle/prep.hpp) element by element.
// This TU is compiled with --fmad=false: float32 operations round individually in the order Eigen's
// fixed-size expressions evaluate them on the reference's x86-64 build; sinf/cosf/atan2f are CUDA's.
#include <cfloat>
#include <cmath>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "scan_prep.cuh"

namespace flimo {

#define FP_TRY(x)                     \
  do {                                \
    cudaError_t e_ = (x);             \
    if (e_ != cudaSuccess) return e_; \
  } while (0)

namespace {

struct RawPt {
  float x, y, z;
  uint32_t lo, hi;   // the time union (fast_limo::Point, Common.hpp:100-113): bytes 24..31
};
__device__ __forceinline__ RawPt load_raw(const unsigned char* raw, size_t i) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(raw + i * 32));
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(raw + i * 32 + 24));
  return RawPt{a.x, a.y, a.z, u.x, u.y};
}

// stage 1: finite + crop box (Localizer.cpp:262-271)
__global__ void __launch_bounds__(256) flag1_kernel(const unsigned char* __restrict__ raw, uint32_t n, PrepDev c,
                                                    uint32_t* __restrict__ flag1) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const RawPt p = load_raw(raw, i);
  bool ok = isfinite(p.x) && isfinite(p.y) && isfinite(p.z);
  if (ok && c.crop_active) {
    const bool outside = (p.x < c.crop_min[0] || p.y < c.crop_min[1] || p.z < c.crop_min[2]) ||
                         (p.x > c.crop_max[0] || p.y > c.crop_max[1] || p.z > c.crop_max[2]);
    ok = outside;
  }
  flag1[i] = ok ? 1u : 0u;
}

// stage 2: distance / rate / field of view on the index of the cropped cloud (Localizer.cpp:273-302)
__global__ void __launch_bounds__(256) flag2_kernel(const unsigned char* __restrict__ raw, uint32_t n, PrepDev c,
                                                    const uint32_t* __restrict__ flag1, const uint32_t* __restrict__ idx1,
                                                    uint32_t* __restrict__ flag2) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool ok = flag1[i] != 0u;
  if (ok) {
    const RawPt p = load_raw(raw, i);
    if (c.dist_active) ok = ok && (sqrtf(p.x * p.x + (p.y * p.y + p.z * p.z)) > c.min_dist);
    if (c.rate_active) ok = ok && (idx1[i] % (uint32_t)c.rate_value == 0u);
    if (c.fov_active) ok = ok && (fabsf(atan2f(p.y, p.x)) < c.fov_angle);
  }
  flag2[i] = ok ? 1u : 0u;
}

__device__ __forceinline__ unsigned long long sort_key(const RawPt& p, const PrepDev& c) {
  if (c.sensor_type == 0) return c.end_of_sweep ? (unsigned long long)(~p.lo) : (unsigned long long)p.lo;
  if (c.sensor_type == 1) {
    uint32_t u = p.lo;                                       // float bits -> order preserving unsigned
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return c.end_of_sweep ? (unsigned long long)(~u) : (unsigned long long)u;
  }
  unsigned long long u = ((unsigned long long)p.hi << 32) | p.lo;
  return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}

// extract_point_time (Localizer.cpp:747-777)
__device__ __forceinline__ double point_time(const RawPt& p, const PrepDev& c) {
  if (c.sensor_type == 0) {
    const float f = __fmul_rn((float)p.lo, 1e-9f);
    return c.end_of_sweep ? c.sweep_ref_time - (double)f : c.sweep_ref_time + (double)f;
  }
  if (c.sensor_type == 1) {
    const float f = __uint_as_float(p.lo);
    return c.end_of_sweep ? c.sweep_ref_time - (double)f : c.sweep_ref_time + (double)f;
  }
  const double ts = __longlong_as_double((long long)(((unsigned long long)p.hi << 32) | p.lo));
  return c.sensor_type == 2 ? ts : ts * (double)1e-9f;
}

__global__ void __launch_bounds__(256) compact_kernel(const unsigned char* __restrict__ raw, uint32_t n, PrepDev c,
                                                      const uint32_t* __restrict__ flag2, const uint32_t* __restrict__ pos,
                                                      unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals,
                                                      uint32_t* __restrict__ count) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (flag2[i]) {
    const RawPt p = load_raw(raw, i);
    keys[pos[i]] = sort_key(p, c);
    vals[pos[i]] = i;
  }
  if (i == n - 1) *count = pos[i] + flag2[i];
}

__global__ void __launch_bounds__(256) times_kernel(const unsigned char* __restrict__ raw, const uint32_t* __restrict__ order,
                                                    uint32_t m, PrepDev c, double* __restrict__ t_out) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= m) return;
  t_out[k] = point_time(load_raw(raw, order[k]), c);
}

// ---- float algebra in Eigen's evaluation order (restated in oracle/prep.hpp) ----------------------------
__device__ __forceinline__ void quat_to_R(const float* q, float* R) {
  const float tx = 2.f * q[0], ty = 2.f * q[1], tz = 2.f * q[2];
  const float twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  const float txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  const float tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  R[0] = 1.f - (tyy + tzz); R[1] = txy - twz;         R[2] = txz + twy;
  R[3] = txy + twz;         R[4] = 1.f - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;         R[7] = tyz + twx;         R[8] = 1.f - (txx + tyy);
}
__device__ __forceinline__ void R_to_quat(const float* m, float* q) {
  float t = m[0] + (m[4] + m[8]);
  if (t > 0.f) {
    t = sqrtf(t + 1.0f);
    q[3] = 0.5f * t;
    t = 0.5f / t;
    q[0] = (m[7] - m[5]) * t;
    q[1] = (m[2] - m[6]) * t;
    q[2] = (m[3] - m[1]) * t;
  } else {
    int i = 0;
    if (m[4] > m[0]) i = 1;
    if (m[8] > m[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrtf(m[4 * i] - m[4 * j] - m[4 * k] + 1.0f);
    q[i] = 0.5f * t;
    t = 0.5f / t;
    q[3] = (m[3 * k + j] - m[3 * j + k]) * t;
    q[j] = (m[3 * j + i] + m[3 * i + j]) * t;
    q[k] = (m[3 * k + i] + m[3 * i + k]) * t;
  }
}
__device__ __forceinline__ void quat_mul(const float* a, const float* b, float* o) {   // Eigen's SSE quat_product order
  o[0] = (a[0] * b[3] - a[2] * b[1]) + (a[1] * b[2] + a[3] * b[0]);
  o[1] = (a[1] * b[3] - a[0] * b[2]) + (a[2] * b[0] + a[3] * b[1]);
  o[2] = (a[2] * b[3] - a[1] * b[0]) + (a[0] * b[1] + a[3] * b[2]);
  o[3] = (a[3] * b[3] - a[0] * b[0]) + -(a[2] * b[2] + a[1] * b[1]);
}
__device__ __forceinline__ void quat_rotate(const float* q, const float* v, float* o) {
  float uv[3] = {q[1] * v[2] - q[2] * v[1], q[2] * v[0] - q[0] * v[2], q[0] * v[1] - q[1] * v[0]};
  uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
  const float c[3] = {q[1] * uv[2] - q[2] * uv[1], q[2] * uv[0] - q[0] * uv[2], q[0] * uv[1] - q[1] * uv[0]};
#pragma unroll
  for (int i = 0; i < 3; ++i) o[i] = (v[i] + q[3] * uv[i]) + c[i];
}

// One point of deskewPointCloud's loop (Localizer.cpp:822-843): frame lookup, State::update to the point's
// time, lidar -> world with the integrated pose, world -> body frame of the last propagated state.
__global__ void __launch_bounds__(128) deskew_kernel(const unsigned char* __restrict__ raw, const uint32_t* __restrict__ order,
                                                     const double* __restrict__ t_sorted, uint32_t m, DeskewDev d,
                                                     const flimo_frame* __restrict__ frames, float4* __restrict__ out_world,
                                                     float4* __restrict__ out_xt2) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= m) return;
  const RawPt p = load_raw(raw, order[k]);
  const double t = t_sorted[k] + d.offset;
  int low = 0, high = d.n_frames - 1;                        // binary_search_tailored
  while (high >= low) {
    const int mid = (low + high) / 2;
    if (frames[mid].time > t) high = mid - 1; else low = mid + 1;
  }
  const flimo_frame f = frames[high < 0 ? 0 : high];
  // State::update(t)
  const double dt = t - f.time;
  const float w[3] = {f.w[0] - f.bg[0], f.w[1] - f.bg[1], f.w[2] - f.bg[2]};
  const float w_norm = sqrtf(w[0] * w[0] + (w[1] * w[1] + w[2] * w[2]));
  float R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  if ((double)w_norm > 1.e-7) {
    const float r[3] = {w[0] / w_norm, w[1] / w_norm, w[2] / w_norm};
    const float K[9] = {0.f, -r[2], r[1], r[2], 0.f, -r[0], -r[1], r[0], 0.f};
    const float r_ang = (float)((double)w_norm * dt);
    const float sn = sinf(r_ang);
    const float c1 = (float)(1.0 - (double)cosf(r_ang));
    float cK[9], KK[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) cK[i] = c1 * K[i];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) KK[3 * i + j] = (cK[3 * i] * K[j] + cK[3 * i + 1] * K[3 + j]) + cK[3 * i + 2] * K[6 + j];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = R[i] + (sn * K[i] + KK[i]);
  }
  const float ab[3] = {f.a[0] - f.ba[0], f.a[1] - f.ba[1], f.a[2] - f.ba[2]};
  float a0[3];
  quat_rotate(f.q, ab, a0);
#pragma unroll
  for (int i = 0; i < 3; ++i) a0[i] += f.g[i];
  float qu[4], q[4], pos[3];
  R_to_quat(R, qu);
  quat_mul(f.q, qu, q);
  const float dtf = (float)dt;
#pragma unroll
  for (int i = 0; i < 3; ++i) pos[i] = f.p[i] + (f.v[i] * dtf + ((0.5f * a0[i]) * dtf) * dtf);
  // T = X0.get_RT() * lidar2baselink_T ; pt = T * pt ; pt2 = last_state.get_RT_inv() * pt
  float Rq[9];
  quat_to_R(q, Rq);
  const float RT[16] = {Rq[0], Rq[1], Rq[2], pos[0], Rq[3], Rq[4], Rq[5], pos[1], Rq[6], Rq[7], Rq[8], pos[2], 0.f, 0.f, 0.f, 1.f};
  float T[16];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      T[4 * i + j] = ((RT[4 * i] * d.T_l2b[j] + RT[4 * i + 1] * d.T_l2b[4 + j]) + RT[4 * i + 2] * d.T_l2b[8 + j]) + RT[4 * i + 3] * d.T_l2b[12 + j];
  const float v[4] = {p.x, p.y, p.z, 1.f};
  float wv[4], bv[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) wv[i] = ((T[4 * i] * v[0] + T[4 * i + 1] * v[1]) + T[4 * i + 2] * v[2]) + T[4 * i + 3] * v[3];
#pragma unroll
  for (int i = 0; i < 4; ++i) bv[i] = ((d.Tinv[4 * i] * wv[0] + d.Tinv[4 * i + 1] * wv[1]) + d.Tinv[4 * i + 2] * wv[2]) + d.Tinv[4 * i + 3] * wv[3];
  if (out_world) out_world[k] = make_float4(wv[0], wv[1], wv[2], wv[3]);
  out_xt2[k] = make_float4(bv[0], bv[1], bv[2], bv[3]);
}

// ---- voxel grid (pcl::VoxelGrid::applyFilter, dense cloud, xyz centroid) --------------------------------
__device__ __forceinline__ unsigned int f2ord(float f) {
  const unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned int u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

__global__ void vox_init_kernel(VoxelDev* v) {
  if (threadIdx.x < 3) v->mn[threadIdx.x] = 0xFFFFFFFFu;
  else if (threadIdx.x < 6) v->mx[threadIdx.x - 3] = 0u;
}
__global__ void __launch_bounds__(256) vox_bbox_kernel(const float4* __restrict__ p, const uint32_t* __restrict__ n_ptr, VoxelDev* v) {
  const uint32_t n = *n_ptr;
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 q = p[i];
    lo[0] = fminf(lo[0], q.x); hi[0] = fmaxf(hi[0], q.x);
    lo[1] = fminf(lo[1], q.y); hi[1] = fmaxf(hi[1], q.y);
    lo[2] = fminf(lo[2], q.z); hi[2] = fmaxf(hi[2], q.z);
  }
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
  if ((threadIdx.x & 31) == 0)
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      atomicMin(&v->mn[a], f2ord(lo[a]));
      atomicMax(&v->mx[a], f2ord(hi[a]));
    }
}
// min_b / div_b / divb_mul and the "leaf size too small" test (voxel_grid.hpp:232-262)
__global__ void vox_setup_kernel(VoxelDev* v, float leaf) {
  const float inv = 1.0f / leaf;
  float mn[3], mx[3];
  for (int a = 0; a < 3; ++a) {
    mn[a] = ord2f(v->mn[a]);
    mx[a] = ord2f(v->mx[a]);
  }
  const long long dx = (long long)((mx[0] - mn[0]) * inv) + 1, dy = (long long)((mx[1] - mn[1]) * inv) + 1,
                  dz = (long long)((mx[2] - mn[2]) * inv) + 1;
  v->overflow = (dx * dy * dz > 2147483647LL) ? 1 : 0;
  int div_b[3];
  for (int a = 0; a < 3; ++a) {
    v->min_b[a] = (int)floorf(mn[a] * inv);
    div_b[a] = (int)floorf(mx[a] * inv) - v->min_b[a] + 1;
  }
  v->mul[0] = 1;
  v->mul[1] = div_b[0];
  v->mul[2] = div_b[0] * div_b[1];
  v->inv_leaf = inv;
}
__global__ void __launch_bounds__(256) vox_keys_kernel(const float4* __restrict__ p, const uint32_t* __restrict__ n_ptr,
                                                       const VoxelDev* __restrict__ v, uint32_t* __restrict__ keys,
                                                       uint32_t* __restrict__ vals, uint32_t cap) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cap) return;
  if (i >= *n_ptr) {                 // padding sorts to the end
    keys[i] = 0xFFFFFFFFu;
    vals[i] = i;
    return;
  }
  const float4 q = p[i];
  const float inv = v->inv_leaf;
  const int i0 = (int)(floorf(q.x * inv) - (float)v->min_b[0]), i1 = (int)(floorf(q.y * inv) - (float)v->min_b[1]),
            i2 = (int)(floorf(q.z * inv) - (float)v->min_b[2]);
  keys[i] = (uint32_t)(i0 * v->mul[0] + i1 * v->mul[1] + i2 * v->mul[2]);
  vals[i] = i;
}
__global__ void __launch_bounds__(256) vox_heads_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ n_ptr,
                                                        uint32_t* __restrict__ head, uint32_t cap) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cap) return;
  head[i] = (i < *n_ptr && (i == 0 || keys[i] != keys[i - 1])) ? 1u : 0u;
}
// one thread per voxel walks its points in sorted (= input) order: the float sums equal the CPU's
__global__ void __launch_bounds__(256) vox_centroid_kernel(const float4* __restrict__ p, const uint32_t* __restrict__ keys,
                                                           const uint32_t* __restrict__ vals, const uint32_t* __restrict__ head,
                                                           const uint32_t* __restrict__ pos, const uint32_t* __restrict__ n_ptr,
                                                           float4* __restrict__ out, uint32_t* __restrict__ n_out, uint32_t cap) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cap) return;
  const uint32_t n = *n_ptr;
  if (i == cap - 1) *n_out = pos[i] + head[i];
  if (i >= n || !head[i]) return;
  const uint32_t key = keys[i];
  float s[3] = {0.f, 0.f, 0.f};
  uint32_t j = i;
  for (; j < n && keys[j] == key; ++j) {
    const float4 q = p[vals[j]];
    s[0] += q.x;
    s[1] += q.y;
    s[2] += q.z;
  }
  const float cnt = (float)(j - i);
  out[pos[i]] = make_float4(s[0] / cnt, s[1] / cnt, s[2] / cnt, 1.f);
}

// pcl::fromROSMsg (src/main.cpp:23) for the fields fast_limo::Point registers: byte-wise field reads (any
// alignment), written as the canonical 32-byte record {x, y, z, 1, intensity, 0, time union}.
__device__ __forceinline__ uint32_t load_u32_any(const unsigned char* p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
__global__ void __launch_bounds__(256) decode_msg_kernel(const unsigned char* __restrict__ msg, uint32_t n, uint32_t step,
                                                         flimo_msg_layout L, unsigned char* __restrict__ raw) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned char* p = msg + (size_t)i * step;
  uint4 a, b;
  a.x = L.off_x >= 0 ? load_u32_any(p + L.off_x) : 0u;
  a.y = L.off_y >= 0 ? load_u32_any(p + L.off_y) : 0u;
  a.z = L.off_z >= 0 ? load_u32_any(p + L.off_z) : 0u;
  a.w = __float_as_uint(1.0f);
  b.x = L.off_intensity >= 0 ? load_u32_any(p + L.off_intensity) : 0u;
  b.y = 0u;
  b.z = b.w = 0u;
  if (L.off_time >= 0) {
    b.z = load_u32_any(p + L.off_time);
    if (L.time_datatype == 8) b.w = load_u32_any(p + L.off_time + 4);
  }
  uint4* o = reinterpret_cast<uint4*>(raw + (size_t)i * 32);
  o[0] = a;
  o[1] = b;
}

inline unsigned int nblk(size_t n, int t = 256) { return (unsigned int)((n + t - 1) / t); }

cudaError_t ensure(void** p, size_t* cap, size_t need) {
  if (*cap >= need) return cudaSuccess;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  FP_TRY(cudaMalloc(p, need));
  *cap = need;
  return cudaSuccess;
}

}  // namespace

void prep_free(PrepBuffers& b) {
  cudaFree(b.raw);
  cudaFree(b.u32);
  cudaFree(b.keys64);
  cudaFree(b.t_sorted);
  cudaFree(b.world);
  cudaFree(b.xt2);
  cudaFree(b.vox_out);
  cudaFree(b.frames);
  cudaFree(b.cub_tmp);
  cudaFree(b.small);
  b = PrepBuffers{};
}

cudaError_t prep_reserve(PrepBuffers& b, size_t n) {
  if (n <= b.cap && b.small) return cudaSuccess;
  const size_t cap = n + n / 4 + 1024;
  cudaFree(b.raw); cudaFree(b.u32); cudaFree(b.keys64); cudaFree(b.t_sorted); cudaFree(b.world); cudaFree(b.xt2); cudaFree(b.vox_out);
  b.raw = nullptr; b.u32 = nullptr; b.keys64 = nullptr; b.t_sorted = nullptr; b.world = b.xt2 = b.vox_out = nullptr;
  b.cap = 0;
  FP_TRY(cudaMalloc(&b.raw, cap * 32));
  FP_TRY(cudaMalloc(&b.u32, 10 * cap * sizeof(uint32_t)));
  FP_TRY(cudaMalloc(&b.keys64, 2 * cap * sizeof(unsigned long long)));
  FP_TRY(cudaMalloc(&b.t_sorted, cap * sizeof(double)));
  FP_TRY(cudaMalloc(&b.world, cap * sizeof(float4)));
  FP_TRY(cudaMalloc(&b.xt2, cap * sizeof(float4)));
  FP_TRY(cudaMalloc(&b.vox_out, cap * sizeof(float4)));
  if (!b.small) FP_TRY(cudaMalloc(&b.small, 256));
  b.cap = cap;
  // scratch for the CUB calls at this capacity
  size_t need = 0, t = 0;
  cub::DoubleBuffer<unsigned long long> dk(b.keys64, b.keys64 + cap);
  cub::DoubleBuffer<uint32_t> dv(b.u32, b.u32 + cap);
  FP_TRY(cub::DeviceRadixSort::SortPairs(nullptr, t, dk, dv, (int)cap, 0, 64, (cudaStream_t)0));
  need = t;
  cub::DoubleBuffer<uint32_t> dk32(b.u32, b.u32 + cap);
  FP_TRY(cub::DeviceRadixSort::SortPairs(nullptr, t, dk32, dv, (int)cap, 0, 32, (cudaStream_t)0));
  need = t > need ? t : need;
  FP_TRY(cub::DeviceScan::ExclusiveSum(nullptr, t, b.u32, b.u32, (int)cap, (cudaStream_t)0));
  need = t > need ? t : need;
  FP_TRY(ensure(&b.cub_tmp, &b.cub_tmp_bytes, need));
  return cudaSuccess;
}

cudaError_t prep_decode_msg(PrepBuffers& b, const unsigned char* d_msg, size_t n, size_t point_step, const flimo_msg_layout& L,
                            cudaStream_t st, uint64_t* launches) {
  if (n == 0) return cudaSuccess;
  decode_msg_kernel<<<nblk(n), 256, 0, st>>>(d_msg, (uint32_t)n, (uint32_t)point_step, L, b.raw);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

// u32 scratch layout (10 arrays of cap): 0 flag1/head | 1 idx1/pos | 2 flag2 | 3 pos | 4 order | 5 order_alt | 6 keys32 | 7 keys32_alt | 8 vox vals | 9 vox vals_alt
cudaError_t prep_filter_sort(PrepBuffers& b, size_t n, const PrepDev& c, cudaStream_t st, uint32_t* n_kept, double* t_last,
                             uint64_t* launches) {
  const size_t cap = b.cap;
  uint32_t *flag1 = b.u32, *idx1 = b.u32 + cap, *flag2 = b.u32 + 2 * cap, *pos = b.u32 + 3 * cap, *vals = b.u32 + 4 * cap,
           *vals_alt = b.u32 + 5 * cap;
  uint32_t* d_count = reinterpret_cast<uint32_t*>(b.small);
  *n_kept = 0;
  *t_last = 0.0;
  b.n_sorted = 0;
  if (n == 0) return cudaSuccess;
  size_t tmp = b.cub_tmp_bytes;
  flag1_kernel<<<nblk(n), 256, 0, st>>>(b.raw, (uint32_t)n, c, flag1);
  FP_TRY(cub::DeviceScan::ExclusiveSum(b.cub_tmp, tmp, flag1, idx1, (int)n, st));
  flag2_kernel<<<nblk(n), 256, 0, st>>>(b.raw, (uint32_t)n, c, flag1, idx1, flag2);
  tmp = b.cub_tmp_bytes;
  FP_TRY(cub::DeviceScan::ExclusiveSum(b.cub_tmp, tmp, flag2, pos, (int)n, st));
  compact_kernel<<<nblk(n), 256, 0, st>>>(b.raw, (uint32_t)n, c, flag2, pos, b.keys64, vals, d_count);
  uint32_t m = 0;
  FP_TRY(cudaMemcpyAsync(&m, d_count, sizeof(m), cudaMemcpyDeviceToHost, st));
  FP_TRY(cudaStreamSynchronize(st));
  if (launches) *launches += 5;
  if (m == 0) return cudaSuccess;
  cub::DoubleBuffer<unsigned long long> dk(b.keys64, b.keys64 + cap);
  cub::DoubleBuffer<uint32_t> dv(vals, vals_alt);
  tmp = b.cub_tmp_bytes;
  FP_TRY(cub::DeviceRadixSort::SortPairs(b.cub_tmp, tmp, dk, dv, (int)m, 0, c.sensor_type <= 1 ? 32 : 64, st));
  b.order = dv.Current();
  times_kernel<<<nblk(m), 256, 0, st>>>(b.raw, b.order, m, c, b.t_sorted);
  FP_TRY(cudaMemcpyAsync(t_last, b.t_sorted + (m - 1), sizeof(double), cudaMemcpyDeviceToHost, st));
  FP_TRY(cudaStreamSynchronize(st));
  if (launches) *launches += 6;
  b.n_sorted = m;
  *n_kept = m;
  return cudaGetLastError();
}

cudaError_t prep_deskew(PrepBuffers& b, const DeskewDev& d, const flimo_frame* h_frames, bool keep_world, cudaStream_t st,
                        uint64_t* launches) {
  const uint32_t m = (uint32_t)b.n_sorted;
  if (m == 0 || d.n_frames <= 0) return cudaSuccess;
  if ((size_t)d.n_frames > b.frames_cap) {
    cudaFree(b.frames);
    b.frames = nullptr;
    b.frames_cap = 0;
    FP_TRY(cudaMalloc(&b.frames, ((size_t)d.n_frames + 64) * sizeof(flimo_frame)));
    b.frames_cap = (size_t)d.n_frames + 64;
  }
  FP_TRY(cudaMemcpyAsync(b.frames, h_frames, (size_t)d.n_frames * sizeof(flimo_frame), cudaMemcpyHostToDevice, st));
  deskew_kernel<<<nblk(m, 128), 128, 0, st>>>(b.raw, b.order, b.t_sorted, m, d, b.frames, keep_world ? b.world : nullptr, b.xt2);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

// in: n_ptr = device count of points in `in`; out: centroids + device/host count.  cap = upper bound of *n_ptr.
cudaError_t prep_voxel(PrepBuffers& b, const float4* in, uint32_t cap_n, const uint32_t* d_n, float leaf, cudaStream_t st,
                       uint32_t* n_out_host, bool* passthrough, uint64_t* launches) {
  const size_t cap = b.cap;
  uint32_t *head = b.u32, *pos = b.u32 + cap, *vals = b.u32 + 8 * cap, *vals_alt = b.u32 + 9 * cap, *keys = b.u32 + 6 * cap,
           *keys_alt = b.u32 + 7 * cap;
  VoxelDev* v = reinterpret_cast<VoxelDev*>(b.small + 64);
  uint32_t* d_n_out = reinterpret_cast<uint32_t*>(b.small + 8);
  *n_out_host = 0;
  *passthrough = false;
  if (cap_n == 0) return cudaSuccess;
  vox_init_kernel<<<1, 32, 0, st>>>(v);
  vox_bbox_kernel<<<min(nblk(cap_n), 148u * 4u), 256, 0, st>>>(in, d_n, v);
  vox_setup_kernel<<<1, 1, 0, st>>>(v, leaf);
  vox_keys_kernel<<<nblk(cap_n), 256, 0, st>>>(in, d_n, v, keys, vals, cap_n);
  cub::DoubleBuffer<uint32_t> dk(keys, keys_alt), dv(vals, vals_alt);
  size_t tmp = b.cub_tmp_bytes;
  FP_TRY(cub::DeviceRadixSort::SortPairs(b.cub_tmp, tmp, dk, dv, (int)cap_n, 0, 32, st));
  vox_heads_kernel<<<nblk(cap_n), 256, 0, st>>>(dk.Current(), d_n, head, cap_n);
  tmp = b.cub_tmp_bytes;
  FP_TRY(cub::DeviceScan::ExclusiveSum(b.cub_tmp, tmp, head, pos, (int)cap_n, st));
  vox_centroid_kernel<<<nblk(cap_n), 256, 0, st>>>(in, dk.Current(), dv.Current(), head, pos, d_n, b.vox_out, d_n_out, cap_n);
  struct { uint32_t n; int overflow; } res = {0, 0};
  FP_TRY(cudaMemcpyAsync(&res.n, d_n_out, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  FP_TRY(cudaMemcpyAsync(&res.overflow, &v->overflow, sizeof(int), cudaMemcpyDeviceToHost, st));
  FP_TRY(cudaStreamSynchronize(st));
  if (launches) *launches += 11;
  *n_out_host = res.n;
  *passthrough = res.overflow != 0;     // "leaf size is too small": pcl returns the input cloud unchanged
  return cudaGetLastError();
}

}  // namespace flimo
