// The iterated error-state Kalman step (esekfom.hpp:1652-1819) as a COOPERATIVE routine: the work of one
// pass is cut into phases of independent items separated by barriers, so that the same source runs
//   * on the device, by the 128 threads of the CTA that completes the reduction tree of a measurement pass
//     (match_kernel.cu, registration kernel: no host between the passes of an update), and
//   * on the host, one item after the other (SerialExec) — that is how tests/ check it against the plain
//     host implementation (ekf_host.hpp, IteratedUpdate::step) and against the oracle without a GPU.
//
// Algebra.  The reference forms P_inv = ((P/R)^-1 + U HTH U^T)^-1 (esekfom.hpp:1722-1726, U = [I12; 0]) and uses its
// first 12 columns.  By the block-inverse identities these are  [I12; G] S^-1  with
//     S = (P11/R)^-1 + HTH   (symmetric positive definite),     G = P21 P11^-1,
// so everything that involves only the covariance is prepared BEFORE the measurement arrives:
//   pre  : dx = x (-) x_prop, dn = J dx, P = J P_prop J^T, Pi = (P11/R)^-1, G            (pivoted elimination, hidden)
//   post : S = Pi + HTH, v = HTh + HTH dn(0:12), z = S^-1 v  (13-column elimination, no pivot search: S is SPD)
//          dxk = [z; G z] - dn   (= K_h + (K_x - I) dn of esekfom.hpp:1733),  x = x (+) filter(dxk)
// The covariance of the LAST pass (esekfom.hpp:1764-1819) is formed by the caller on the host from the state the
// last pass was evaluated at and its sums (IteratedUpdate::finish) — once per scan, after the device is done.
#pragma once
#include "ekf_host.hpp"

namespace flimo {
namespace ekf {

// Inputs of one update that stay fixed over its passes (travel in the kernel parameter block).
struct UpdInit {
  double x[26];        // propagated state
  double P[N * N];     // propagated covariance
  double limit[N];     // convergence limits
  double R, D;         // measurement noise, degeneracy threshold (Localizer.cpp:333)
  int max_iter;        // MAX_NUM_ITERS
  int max_matches;     // MAX_NUM_MATCHES (first-N cap on contributing rows)
};

constexpr int kMaxTrace = 16;   // passes whose state is recorded for tests / debugging

// State carried from pass to pass in device memory (the CTA that runs the step differs from pass to pass).
struct UpdState {
  double x[26];
  double P[N * N];                // final covariance (written by the last pass)
  double last_dx[N];
  double trace[kMaxTrace][32];    // per pass: x after the pass (26), n_valid, n_rows, device ns of the pass, orig_limit
  double phase_ns[kMaxTrace][16]; // per pass: ns from the moment the pass sums were complete to the milestones of the step
  int passes, failed, redone, pad_;
};

// Scratch of one step; overlays the tile buffers of the CTA (they are idle while the step runs).
struct StepShared {
  double P[N * N];        // projected covariance of this pass
  double aug[12][26];     // pre: [P11/R | I] -> [I | Pi];  post: [S | v] -> [I | z] (columns 0..12)
  double Pi[144];         // (P11 / R)^-1
  double G[11][12];       // P21 P11^-1
  double HTH[144];
  double HTh[12];
  double x[26], xp[26], x_eval[26];   // current state, propagated state, state the last pass was evaluated at
  double dx[N], dn[N], dxk[N], dxf[N];
  double J3[2][9];        // A_matrix^T of the two SO3 blocks
  double J2[4];           // S2 block
  double colc[12];        // pivot column of the current elimination step
  double invR;
  int clear, converge, final_pass, singular;
  int iter, conv_count, passes, done, failed, pad_;   // IteratedUpdate's counters (carried from pass to pass)
};

// Host executor: items one after the other.
struct SerialExec {
  static constexpr bool kDevice = false;
  template <class F> FLIMO_HD void par(int n, F f) { for (int i = 0; i < n; ++i) f(i); }
  template <class F> FLIMO_HD void spread(int n, F f) { for (int i = 0; i < n; ++i) f(i); }
  template <class F> FLIMO_HD void warp0(int n, F f) { for (int i = 0; i < n; ++i) f(i); }
  FLIMO_HD void sync() {}
  FLIMO_HD void warp0_sync() {}
  FLIMO_HD void stamp(int) {}
  FLIMO_HD void stamp_here(int) {}
};

#if defined(__CUDACC__)
// Device executor: one CTA of `nt` threads (a multiple of 32, at least 128 for spread()).
struct CtaExec {
  static constexpr bool kDevice = true;
  int tid, nt;
  template <class F> __device__ void par(int n, F f) { for (int i = tid; i < n; i += nt) f(i); }
  template <class F> __device__ void spread(int n, F f) { if ((tid & 31) == 0 && (tid >> 5) < n) f(tid >> 5); }   // item j on warp j
  template <class F> __device__ void warp0(int n, F f) { if (tid < 32) for (int i = tid; i < n; i += 32) f(i); }
  __device__ void sync() { __syncthreads(); }
  __device__ void warp0_sync() { if (tid < 32) __syncwarp(); }
  unsigned long long* stamps;   // optional: [16] %globaltimer values at the phase boundaries (profiling)
  __device__ void stamp_here(int k) {               // by the calling thread, whoever it is
    if (stamps != nullptr) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      stamps[k] = t;
    }
  }
  __device__ void stamp(int k) {
    if (stamps != nullptr && tid == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      stamps[k] = t;
    }
  }
};
#endif

// block Jacobian helpers: index range of the block a row/column belongs to, and the entry of J
FLIMO_HD inline int blk_begin(int r) { return (r >= 3 && r < 9) ? (r < 6 ? 3 : 6) : (r >= 21 ? 21 : r); }
FLIMO_HD inline int blk_size(int r) { return (r >= 3 && r < 9) ? 3 : (r >= 21 ? 2 : 1); }
FLIMO_HD inline double blk_J(const double (*J3)[9], const double* J2, int r, int k) {   // J(r, k), k inside the block of r
  if (r >= 3 && r < 6) return J3[0][(r - 3) * 3 + (k - 3)];
  if (r >= 6 && r < 9) return J3[1][(r - 6) * 3 + (k - 6)];
  if (r >= 21) return J2[(r - 21) * 2 + (k - 21)];
  return 1.0;
}
// (J S J^T)(r, c): rows first, then columns
FLIMO_HD inline double congruence(const double* S, const double (*J3)[9], const double* J2, int r, int c) {
  const int rb = blk_begin(r), rn = blk_size(r), cb = blk_begin(c), cn = blk_size(c);
  double acc = 0.0;
  for (int l = 0; l < cn; ++l) {
    double t = 0.0;
    for (int k = 0; k < rn; ++k) t += blk_J(J3, J2, r, rb + k) * S[(rb + k) * N + (cb + l)];
    acc += t * blk_J(J3, J2, c, cb + l);
  }
  return acc;
}

FLIMO_HD inline void store9T(const Mat<3, 3>& A, double* out) {   // out = A^T, row-major
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) out[i * 3 + j] = A(j, i);
}

// ---------------------------------------------------------------------------------------------------
// The step is split where the measurement enters:
//   pre_step  — everything that depends only on the current state and the propagated covariance
//               (dx, the block Jacobians, dn, P = J P_prop J^T).  On the device it runs WHILE the
//               measurement pass is still matching points, off the critical path;
//   post_step — from the summed normal equations to the new state (D .. I above), ordered so that the
//               pose of the next pass is available as early as possible (the covariance of the last pass,
//               gravity, biases come after it).
// The carried state (x, iteration counters) lives in StepShared between the two and from pass to pass.
// ---------------------------------------------------------------------------------------------------
FLIMO_HD inline void step_begin(StepShared& s, const UpdInit& in) {   // IteratedUpdate::begin; single caller
  for (int i = 0; i < 26; ++i) {
    s.x[i] = in.x[i];
    s.xp[i] = in.x[i];
  }
  s.iter = -1;
  s.conv_count = 0;
  s.passes = 0;
  s.done = (s.iter >= in.max_iter) ? 1 : 0;     // max_iter < 0: the reference's loop body never runs
  s.failed = 0;
  s.invR = 1.0 / in.R;
}

template <class Ex>
FLIMO_HD inline void gauss_jordan(Ex& ex, StepShared& s, int n_cols, bool pivot);

template <class Ex>
FLIMO_HD inline void pre_step(Ex& ex, StepShared& s, const UpdInit& in) {
  // B. dx = x (-) x_prop and the block Jacobians (four independent scalar jobs)
  ex.spread(4, [&](int j) {
    State x, xp;
    x.load(s.x);
    xp.load(s.xp);
    if (j == 0) {
      const V3 d = so3_minus(x.rot, xp.rot);
      for (int i = 0; i < 3; ++i) s.dx[3 + i] = d[i];
      store9T(A_matrix(d), s.J3[0]);
    } else if (j == 1) {
      const V3 d = so3_minus(x.offR, xp.offR);
      for (int i = 0; i < 3; ++i) s.dx[6 + i] = d[i];
      store9T(A_matrix(d), s.J3[1]);
    } else if (j == 2) {
      double d0, d1;
      s2_minus(x.grav, xp.grav, d0, d1);
      s.dx[21] = d0;
      s.dx[22] = d1;
      const Mat<2, 2> J = s2_Nx_yy(x.grav) * s2_Mx(xp.grav, d0, d1);
      for (int i = 0; i < 4; ++i) s.J2[i] = J.a[i];
    } else {
      for (int i = 0; i < 3; ++i) {
        s.dx[i] = x.pos[i] - xp.pos[i];
        s.dx[9 + i] = x.offT[i] - xp.offT[i];
        s.dx[12 + i] = x.vel[i] - xp.vel[i];
        s.dx[15 + i] = x.bg[i] - xp.bg[i];
        s.dx[18 + i] = x.ba[i] - xp.ba[i];
      }
    }
  });
  ex.sync();
  // C. dn = J dx, P = J P_prop J^T
  ex.par(N, [&](int r) {
    const int rb = blk_begin(r), rn = blk_size(r);
    double t = 0.0;
    for (int k = 0; k < rn; ++k) t += blk_J(s.J3, s.J2, r, rb + k) * s.dx[rb + k];
    s.dn[r] = t;
  });
  ex.par(N * N, [&](int e) { s.P[e] = congruence(in.P, s.J3, s.J2, e / N, e % N); });
  ex.par(26, [&](int i) { s.x_eval[i] = s.x[i]; });
  ex.par(1, [&](int) { s.singular = 0; });
  ex.sync();
  // Pi = (P11 / R)^-1 by pivoted elimination of [P11/R | I]
  ex.par(12 * 26, [&](int e) {
    const int r = e / 26, c = e % 26;
    s.aug[r][c] = c < 12 ? s.P[r * N + c] * s.invR : ((c < 24 && c - 12 == r) ? 1.0 : 0.0);
  });
  ex.sync();
  gauss_jordan(ex, s, 24, true);
  ex.sync();
  ex.par(144, [&](int e) { s.Pi[e] = s.aug[e / 12][12 + e % 12]; });
  ex.sync();
  // G = P21 P11^-1 = (P21 / R) Pi
  ex.par(11 * 12, [&](int e) {
    const int r = e / 12, c = e % 12;
    double t = 0.0;
    for (int k = 0; k < 12; ++k) t += (s.P[(12 + r) * N + k] * s.invR) * s.Pi[k * 12 + c];
    s.G[r][c] = t;
  });
  ex.sync();
}

// Gauss-Jordan elimination of the 12 x n_cols system in s.aug ([A | B] -> [I | A^-1 B]), one column per item.
// pivot = true: row pivoting (first maximum of the column, as ekf_host.hpp's invert); pivot = false: none (SPD systems).
template <class Ex>
FLIMO_HD inline void gauss_jordan_phases(Ex& ex, StepShared& s, int n_cols, bool pivot) {
  for (int c = 0; c < 12; ++c) {
    ex.warp0(12, [&](int r) { s.colc[r] = s.aug[r][c]; });
    ex.warp0_sync();
    ex.warp0(n_cols, [&](int l) {
      int p = c;
      double best = fabs(s.colc[c]);
      if (pivot)
        for (int r = c + 1; r < 12; ++r) {
          const double v = fabs(s.colc[r]);
          if (v > best) {
            best = v;
            p = r;
          }
        }
      if (!(best > 0.0) || best > 1.7e308) {
        if (l == 0) s.singular = 1;
        return;
      }
      const double pv = s.colc[p];
      const double top = s.aug[p][l];            // row p moves to row c ...
      s.aug[p][l] = s.aug[c][l];                 // ... and row c to row p
      const double scaled = top * (1.0 / pv);
      s.aug[c][l] = scaled;
      for (int r = 0; r < 12; ++r) {
        if (r == c) continue;
        const double f = (r == p) ? s.colc[c] : s.colc[r];   // column c after the swap
        if (f != 0.0) s.aug[r][l] -= f * scaled;
      }
    });
    ex.warp0_sync();
  }
}

#if defined(__CUDACC__)
// The same elimination for the device: lane l of the first warp owns column l (n_cols <= 26).  The loop over the
// pivot columns is ROLLED; inside a step everything works on registers with static indices: the pivot column and
// the lane's own column are fetched with independent shared-memory loads, swapped / scaled / eliminated in registers
// and stored back.  (A fully unrolled variant with the matrix in registers and shuffles was not faster: 7.7 us
// against 7.0 us for the pivoted 25-column system — the chain of dependent float64 operations of a step sets the time.)
template <bool kPivot>
__device__ __forceinline__ void gauss_jordan_warp(StepShared& s, int tid, int n_cols) {
  if (tid >= 32) return;
  const int l = tid < n_cols ? tid : n_cols - 1;           // surplus lanes shadow the last column (they do not store)
  bool singular = false;
#pragma unroll 1
  for (int c = 0; c < 12; ++c) {
    double col[12], a[12];
#pragma unroll
    for (int r = 0; r < 12; ++r) col[r] = s.aug[r][c];     // broadcast reads
#pragma unroll
    for (int r = 0; r < 12; ++r) a[r] = s.aug[r][l];
    int p = c;
    double pv = 0.0, top = 0.0, colc = 0.0, ac = 0.0;
#pragma unroll
    for (int r = 0; r < 12; ++r)
      if (r == c) {
        colc = col[r];
        ac = a[r];
      }
    if (kPivot) {
      double best = -1.0;
#pragma unroll
      for (int r = 0; r < 12; ++r) {
        const double v = fabs(col[r]);
        if (r >= c && v > best) {                          // first maximum among rows c..11
          best = v;
          p = r;
        }
      }
#pragma unroll
      for (int r = 0; r < 12; ++r)
        if (r == p) {
          pv = col[r];
          top = a[r];
        }
    } else {
      pv = colc;
      top = ac;
    }
    const double apv = fabs(pv);
    if (!(apv > 0.0) || apv > 1.7e308) {                   // warp-uniform
      singular = true;
      break;
    }
    const double scaled = top * __drcp_rn(pv);             // == top * (1.0 / pv), correctly rounded
    __syncwarp();                                          // every lane has read column c
#pragma unroll
    for (int r = 0; r < 12; ++r) {
      double ar = (kPivot && r == p) ? ac : a[r];          // row c moves to row p ...
      const double f = (kPivot && r == p) ? colc : col[r]; // ... in the pivot column too
      if (r == c) ar = scaled;
      else if (f != 0.0) ar -= f * scaled;
      if (tid < n_cols) s.aug[r][l] = ar;
    }
    __syncwarp();
  }
  if (singular && tid == 0) s.singular = 1;
}
#endif

template <class Ex>
FLIMO_HD inline void gauss_jordan(Ex& ex, StepShared& s, int n_cols, bool pivot) {
#if defined(__CUDACC__)
  if constexpr (Ex::kDevice) {
    if (pivot) gauss_jordan_warp<true>(s, ex.tid, n_cols);
    else gauss_jordan_warp<false>(s, ex.tid, n_cols);
  } else
#endif
  {
    gauss_jordan_phases(ex, s, n_cols, pivot);
  }
}

// Pose constants of the next pass are derived by the caller between `post_step_pose` and `post_step_rest`.
//   post_step_pose: S, v, the elimination, dxk, the filter, the convergence test and the (+) of the pose part
//   post_step_rest: (+) of the remaining components, counters
// packed sums (flimo.h layout) -> s.HTH (full symmetric 12x12), s.HTh; resets the per-pass flags
template <class Ex>
FLIMO_HD inline void unpack_measurement(Ex& ex, StepShared& s, const double* packed96) {
  ex.par(144, [&](int e) {
    const int i = e / 12, j = e % 12, a = i < j ? i : j, b = i < j ? j : i;
    s.HTH[e] = packed96[a * 12 - (a * (a - 1)) / 2 + (b - a)];
  });
  ex.par(12, [&](int i) { s.HTh[i] = packed96[78 + i]; });
  ex.par(1, [&](int) { s.converge = 1; });
  ex.sync();
}

template <class Ex>
FLIMO_HD inline void post_step_pose(Ex& ex, StepShared& s, const UpdInit& in, const long long n_rows) {
  if (s.singular) return;                                  // P11 could not be inverted (pre_step)
  // D. [S | v]
  ex.par(12 * 13, [&](int e) {
    const int r = e / 13, c = e % 13;
    double v;
    if (c < 12) {
      v = s.Pi[r * 12 + c] + s.HTH[r * 12 + c];
    } else {
      v = s.HTh[r];
      for (int k = 0; k < 12; ++k) v += s.HTH[r * 12 + k] * s.dn[k];
    }
    s.aug[r][c] = v;
  });
  ex.sync();
  ex.stamp(3);
  // E. elimination (first warp) | fast test of the degeneracy filter (second warp)
  ex.spread(2, [&](int j) {
    if (j == 1) {
      s.clear = (n_rows >= N && all_eigs_above6(s.HTH, fmax(in.D, 1e-3))) ? 1 : 0;
      ex.stamp_here(6);
    }
  });
  gauss_jordan(ex, s, 13, false);
  ex.stamp_here(7);
  ex.sync();
  ex.stamp(4);
  if (s.singular) return;
  // F. dxk = [z; G z] - dn, convergence test
  ex.par(N, [&](int r) {
    double t;
    if (r < 12) {
      t = s.aug[r][12];
    } else {
      t = 0.0;
      for (int k = 0; k < 12; ++k) t += s.G[r - 12][k] * s.aug[k][12];
    }
    const double d = t - s.dn[r];
    s.dxk[r] = d;
    s.dxf[r] = d;
    if (fabs(d) > in.limit[r]) s.converge = 0;       // benign race: every writer stores 0
  });
  ex.sync();
  // G. degeneracy filter (rare: only when the pose block of HTH has a small eigenvalue)
  if (!s.clear) {
    ex.par(1, [&](int) { degeneracy_filter(s.HTH, n_rows >= N, in.D, s.dxk, s.dxf); });
    ex.sync();
  }
  ex.stamp(5);
  // H1. (+) of the components the measurement model reads: pos, rot, offset_R_L_I, offset_T_L_I
  ex.spread(3, [&](int j) {
    if (j == 0) {
      Q4 q = {s.x[3], s.x[4], s.x[5], s.x[6]};
      so3_plus(q, {s.dxf[3], s.dxf[4], s.dxf[5]});
      for (int i = 0; i < 4; ++i) s.x[3 + i] = q[i];
    } else if (j == 1) {
      Q4 q = {s.x[7], s.x[8], s.x[9], s.x[10]};
      so3_plus(q, {s.dxf[6], s.dxf[7], s.dxf[8]});
      for (int i = 0; i < 4; ++i) s.x[7 + i] = q[i];
    } else {
      for (int i = 0; i < 3; ++i) {
        s.x[i] += s.dxf[i];
        s.x[11 + i] += s.dxf[9 + i];
      }
      const int cc = s.conv_count + s.converge;
      s.final_pass = (cc > 1 || s.iter == in.max_iter - 1) ? 1 : 0;
    }
  });
  ex.sync();
}

template <class Ex>
FLIMO_HD inline void post_step_rest(Ex& ex, StepShared& s, const UpdInit& in) {
  if (s.singular) {     // singular / non-finite system: abandon the update at the propagated state
    ex.par(26, [&](int i) { s.x[i] = in.x[i]; });
    ex.par(1, [&](int) {
      s.failed = 1;
      s.done = 1;
      s.passes = s.passes + 1;
    });
    ex.sync();
    return;
  }
  // H2. (+) of the remaining components
  ex.spread(2, [&](int j) {
    if (j == 0) {
      V3 g = {s.x[23], s.x[24], s.x[25]};
      s2_plus(g, s.dxf[21], s.dxf[22]);
      for (int i = 0; i < 3; ++i) s.x[23 + i] = g[i];
    } else {
      for (int i = 0; i < 3; ++i) {
        s.x[14 + i] += s.dxf[12 + i];
        s.x[17 + i] += s.dxf[15 + i];
        s.x[20 + i] += s.dxf[18 + i];
      }
      s.conv_count = s.conv_count + s.converge;
      s.passes = s.passes + 1;
      if (s.final_pass) {
        s.done = 1;
      } else {
        s.iter = s.iter + 1;
        if (s.iter >= in.max_iter) s.done = 1;   // unreachable (final_pass fires at max_iter - 1); defensive
      }
    }
  });
  ex.sync();
}

// The cooperative step driven from the host, one item at a time: same interface as IteratedUpdate.  Used by the
// CPU tests (flimo_ekf_* with FLIMO_EKF_COOP=1) to pin the device code path's algebra without a GPU; the last pass is
// completed with IteratedUpdate::finish exactly as flimo_update does after the device-resident passes.
class CoopUpdate {
 public:
  void begin(const double* x26, const double* P529, int max_iter, const double* limit23, double R, double D) {
    for (int i = 0; i < 26; ++i) in_.x[i] = x26[i];
    for (int i = 0; i < N * N; ++i) in_.P[i] = P529[i];
    for (int i = 0; i < N; ++i) in_.limit[i] = limit23[i];
    in_.R = R;
    in_.D = D;
    in_.max_iter = max_iter;
    in_.max_matches = 0;
    step_begin(sh_, in_);
    fin_.begin(x26, P529, max_iter, limit23, R, D);
    for (int i = 0; i < 26; ++i) x_out_[i] = x26[i];
    for (int i = 0; i < N * N; ++i) P_out_[i] = P529[i];
  }
  bool done() const { return sh_.done != 0; }
  bool failed() const { return sh_.failed != 0; }
  int passes() const { return sh_.passes; }
  void state(double* x26) const { for (int i = 0; i < 26; ++i) x26[i] = sh_.done ? x_out_[i] : sh_.x[i]; }
  void end(double* x26, double* P529) const {
    state(x26);
    for (int i = 0; i < N * N; ++i) P529[i] = P_out_[i];
  }
  bool step(const double* HTH144, const double* HTh12, long long n_rows) {
    if (done()) return true;
    double packed[96] = {0};
    int e = 0;
    for (int i = 0; i < 12; ++i)
      for (int j = i; j < 12; ++j) packed[e++] = HTH144[i * 12 + j];
    for (int i = 0; i < 12; ++i) packed[78 + i] = HTh12[i];
    packed[90] = (double)n_rows;
    SerialExec ex;
    pre_step(ex, sh_, in_);
    unpack_measurement(ex, sh_, packed);
    post_step_pose(ex, sh_, in_, n_rows);
    const bool last = sh_.final_pass != 0 && !sh_.singular;
    post_step_rest(ex, sh_, in_);
    if (sh_.failed) {
      for (int i = 0; i < 26; ++i) x_out_[i] = in_.x[i];
    } else if (last) {
      fin_.finish(sh_.x_eval, HTH144, HTh12, n_rows);      // state + covariance of the last pass (host algebra)
      fin_.end(x_out_, P_out_);
      if (fin_.failed()) sh_.failed = 1;
    }
    return done();
  }

 private:
  UpdInit in_;
  StepShared sh_;
  IteratedUpdate fin_;
  double x_out_[26], P_out_[N * N];
};

}  // namespace ekf
}  // namespace flimo
