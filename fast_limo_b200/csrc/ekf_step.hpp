// The iterated error-state Kalman step (esekfom.hpp:1652-1819) as a COOPERATIVE routine: the work of one
// pass is cut into phases of independent items separated by barriers, so that the same source runs
//   * on the device, by the 128 threads of the CTA that completes the reduction tree of a measurement pass
//     (match_kernel.cu, registration kernel: no host between the passes of an update), and
//   * on the host, one item after the other (SerialExec) — that is how tests/ check it against the plain
//     host implementation (ekf_host.hpp, IteratedUpdate::step) and against the oracle without a GPU.
//
// Algebra (same as ekf_host.hpp, regrouped so that a pass needs ONE 12x12 elimination and no 23x23 inverse):
//   dx     = x (-) x_prop,  dn = J dx,  P = J P_prop J^T          J = blockdiag(I, A(dth)^T, A(dthLI)^T, I, .., Nx Mx)
//   M      = I12 + HTH (P11 / R)
//   [y Z]  = M^-1 [HTh + HTH dn(0:12) | HTH]                       Gauss-Jordan with row pivoting, 25 columns
//   dxk    = (P(:,0:12) / R) y - dn                                = K_h + (K_x - I) dn   of esekfom.hpp:1733
//   x      = x (+) filter(dxk)
//   last pass:  K_x = (P(:,0:12) / R) Z,  P <- Jf P Jf^T - (Jf K_x) (P Jf^T)(0:12, :)
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
  int iter, conv_count, passes, done, failed, pad_;
};

// Scratch of one step; overlays the tile buffers of the CTA (they are idle while the step runs).
struct StepShared {
  double P[N * N];        // projected covariance of this pass
  double aug[12][26];     // [M | v | HTH] -> [I | y | Z]; column 25 is padding.  Later: (P Jf^T)(0:12, :)
  double Kx[N][12];
  double HTH[144];
  double HTh[12];
  double x[26], xp[26];
  double dx[N], dn[N], dxk[N], dxf[N];
  double J3[2][9];        // A_matrix^T of the two SO3 blocks
  double J2[4];           // S2 block
  double colc[12];        // pivot column of the current elimination step
  int clear, converge, final_pass, singular;
};

// Host executor: items one after the other.
struct SerialExec {
  template <class F> FLIMO_HD void par(int n, F f) { for (int i = 0; i < n; ++i) f(i); }
  template <class F> FLIMO_HD void spread(int n, F f) { for (int i = 0; i < n; ++i) f(i); }
  template <class F> FLIMO_HD void warp0(int n, F f) { for (int i = 0; i < n; ++i) f(i); }
  FLIMO_HD void sync() {}
  FLIMO_HD void warp0_sync() {}
};

#if defined(__CUDACC__)
// Device executor: one CTA of `nt` threads (a multiple of 32, at least 128 for spread()).
struct CtaExec {
  int tid, nt;
  template <class F> __device__ void par(int n, F f) { for (int i = tid; i < n; i += nt) f(i); }
  template <class F> __device__ void spread(int n, F f) { if ((tid & 31) == 0 && (tid >> 5) < n) f(tid >> 5); }   // item j on warp j
  template <class F> __device__ void warp0(int n, F f) { if (tid < 32) for (int i = tid; i < n; i += 32) f(i); }
  __device__ void sync() { __syncthreads(); }
  __device__ void warp0_sync() { if (tid < 32) __syncwarp(); }
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

// The carried state is written by a different CTA every pass: read it past the (non-coherent) L1.
#if defined(__CUDA_ARCH__)
#define FLIMO_LD_CG(p) __ldcg(p)
#else
#define FLIMO_LD_CG(p) (*(p))
#endif

// One pass.  packed96 = the summed measurement (flimo.h layout).  On return st holds the new state and counters;
// st.done says whether this was the last pass (st.P is then the updated covariance).
template <class Ex>
FLIMO_HD inline void iterated_step(Ex& ex, StepShared& s, const UpdInit& in, UpdState& st, const double* packed96) {
  const long long n_rows = (long long)(packed96[90] + 0.5);
  const int iter = FLIMO_LD_CG(&st.iter), conv_count = FLIMO_LD_CG(&st.conv_count), passes = FLIMO_LD_CG(&st.passes);

  // A. unpack the measurement, fetch the states
  ex.par(144, [&](int e) {
    const int i = e / 12, j = e % 12, a = i < j ? i : j, b = i < j ? j : i;
    s.HTH[e] = packed96[a * 12 - (a * (a - 1)) / 2 + (b - a)];
  });
  ex.par(12, [&](int i) { s.HTh[i] = packed96[78 + i]; });
  ex.par(26, [&](int i) {
    s.x[i] = FLIMO_LD_CG(&st.x[i]);
    s.xp[i] = in.x[i];
  });
  ex.par(1, [&](int) { s.singular = 0; });
  ex.sync();

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
  ex.sync();

  // D. the 12x25 system and the fast test of the degeneracy filter
  ex.par(12 * 26, [&](int e) {
    const int r = e / 26, c = e % 26;
    double v;
    if (c < 12) {
      v = (r == c) ? 1.0 : 0.0;
      for (int k = 0; k < 12; ++k) v += s.HTH[r * 12 + k] * (s.P[k * N + c] / in.R);
    } else if (c == 12) {
      v = s.HTh[r];
      for (int k = 0; k < 12; ++k) v += s.HTH[r * 12 + k] * s.dn[k];
    } else if (c < 25) {
      v = s.HTH[r * 12 + (c - 13)];
    } else {
      v = 0.0;
    }
    s.aug[r][c] = v;
  });
  ex.spread(2, [&](int j) {
    if (j == 1) s.clear = (n_rows >= N && all_eigs_above6(s.HTH, fmax(in.D, 1e-3))) ? 1 : 0;
  });
  ex.sync();

  // E. Gauss-Jordan elimination with row pivoting; one column per lane of the first warp
  for (int c = 0; c < 12; ++c) {
    ex.warp0(12, [&](int r) { s.colc[r] = s.aug[r][c]; });
    ex.warp0_sync();
    ex.warp0(25, [&](int l) {
      int p = c;
      double best = fabs(s.colc[c]);
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
  ex.sync();

  // F. dxk = (P(:, 0:12) / R) y - dn
  ex.par(N, [&](int r) {
    double t = 0.0;
    for (int k = 0; k < 12; ++k) t += (s.P[r * N + k] / in.R) * s.aug[k][12];
    s.dxk[r] = t - s.dn[r];
    s.dxf[r] = s.dxk[r];
  });
  ex.sync();

  // G. degeneracy filter (rare: only when the pose block of HTH has a small eigenvalue) and the convergence test
  ex.spread(2, [&](int j) {
    if (j == 0) {
      if (!s.clear) degeneracy_filter(s.HTH, n_rows >= N, in.D, s.dxk, s.dxf);
    } else {
      int conv = 1;
      for (int r = 0; r < N; ++r)
        if (fabs(s.dxk[r]) > in.limit[r]) conv = 0;
      s.converge = conv;
      const int cc = conv_count + conv;
      s.final_pass = (cc > 1 || iter == in.max_iter - 1) ? 1 : 0;
    }
  });
  ex.sync();

  if (s.singular) {     // singular / non-finite system: abandon the update at the propagated state
    ex.par(26, [&](int i) { st.x[i] = in.x[i]; });
    ex.par(N * N, [&](int e) { st.P[e] = in.P[e]; });
    ex.par(1, [&](int) {
      st.failed = 1;
      st.done = 1;
      st.passes = passes + 1;
    });
    ex.sync();
    return;
  }

  // H. x = x (+) dxf
  ex.spread(4, [&](int j) {
    State x;
    x.load(s.x);
    if (j == 0) {
      so3_plus(x.rot, {s.dxf[3], s.dxf[4], s.dxf[5]});
      for (int i = 0; i < 4; ++i) s.x[3 + i] = x.rot[i];
    } else if (j == 1) {
      so3_plus(x.offR, {s.dxf[6], s.dxf[7], s.dxf[8]});
      for (int i = 0; i < 4; ++i) s.x[7 + i] = x.offR[i];
    } else if (j == 2) {
      s2_plus(x.grav, s.dxf[21], s.dxf[22]);
      for (int i = 0; i < 3; ++i) s.x[23 + i] = x.grav[i];
    } else {
      for (int i = 0; i < 3; ++i) {
        s.x[i] = x.pos[i] + s.dxf[i];
        s.x[11 + i] = x.offT[i] + s.dxf[9 + i];
        s.x[14 + i] = x.vel[i] + s.dxf[12 + i];
        s.x[17 + i] = x.bg[i] + s.dxf[15 + i];
        s.x[20 + i] = x.ba[i] + s.dxf[18 + i];
      }
    }
  });
  ex.sync();

  const bool final_pass = s.final_pass != 0;
  if (final_pass) {
    // I1. K_x = (P(:, 0:12) / R) Z and the Jacobians of the last correction
    ex.par(N * 12, [&](int e) {
      const int r = e / 12, c = e % 12;
      double t = 0.0;
      for (int k = 0; k < 12; ++k) t += (s.P[r * N + k] / in.R) * s.aug[k][13 + c];
      s.Kx[r][c] = t;
    });
    ex.spread(3, [&](int j) {
      if (j == 0) {
        store9T(A_matrix({s.dxk[3], s.dxk[4], s.dxk[5]}), s.J3[0]);
      } else if (j == 1) {
        store9T(A_matrix({s.dxk[6], s.dxk[7], s.dxk[8]}), s.J3[1]);
      } else {
        const V3 g = {s.x[23], s.x[24], s.x[25]}, gp = {s.xp[23], s.xp[24], s.xp[25]};
        const Mat<2, 2> J = s2_Nx_yy(g) * s2_Mx(gp, s.dxk[21], s.dxk[22]);
        for (int i = 0; i < 4; ++i) s.J2[i] = J.a[i];
      }
    });
    ex.sync();
    // I2. rows of K_x by Jf (one column per item: no item reads what another writes); (P Jf^T)(0:12, :) into aug
    ex.par(12, [&](int c) {
      double t[3];
      for (int b = 0; b < 2; ++b) {
        const int idx = 3 + 3 * b;
        for (int i = 0; i < 3; ++i)
          t[i] = s.J3[b][i * 3] * s.Kx[idx][c] + s.J3[b][i * 3 + 1] * s.Kx[idx + 1][c] + s.J3[b][i * 3 + 2] * s.Kx[idx + 2][c];
        for (int i = 0; i < 3; ++i) s.Kx[idx + i][c] = t[i];
      }
      const double a = s.J2[0] * s.Kx[21][c] + s.J2[1] * s.Kx[22][c], b2 = s.J2[2] * s.Kx[21][c] + s.J2[3] * s.Kx[22][c];
      s.Kx[21][c] = a;
      s.Kx[22][c] = b2;
    });
    ex.sync();
    double* PJt = &s.aug[0][0];                  // 12 x 23 (276 <= 312 doubles)
    ex.par(12 * N, [&](int e) {
      const int k = e / N, c = e % N, cb = blk_begin(c), cn = blk_size(c);
      double t = 0.0;
      for (int l = 0; l < cn; ++l) t += s.P[k * N + cb + l] * blk_J(s.J3, s.J2, c, cb + l);
      PJt[e] = t;
    });
    ex.sync();
    // I3. P <- Jf P Jf^T - K_x (P Jf^T)(0:12, :)
    ex.par(N * N, [&](int e) {
      const int r = e / N, c = e % N;
      double t = 0.0;
      for (int k = 0; k < 12; ++k) t += s.Kx[r][k] * PJt[k * N + c];
      st.P[e] = congruence(s.P, s.J3, s.J2, r, c) - t;
    });
  }

  // J. write back
  ex.par(26, [&](int i) { st.x[i] = s.x[i]; });
  ex.par(N, [&](int i) { st.last_dx[i] = s.dxk[i]; });
  ex.par(1, [&](int) {
    st.conv_count = conv_count + s.converge;
    st.passes = passes + 1;
    if (final_pass) {
      st.done = 1;
    } else {
      st.iter = iter + 1;
      if (iter + 1 >= in.max_iter) st.done = 1;   // unreachable (the branch above fires at max_iter - 1); defensive
    }
  });
  ex.sync();
}

// begin(): state of a fresh update
FLIMO_HD inline void upd_state_begin(UpdState& st, const UpdInit& in) {
  for (int i = 0; i < 26; ++i) st.x[i] = in.x[i];
  st.iter = -1;
  st.conv_count = 0;
  st.passes = 0;
  st.failed = 0;
  st.done = (st.iter >= in.max_iter) ? 1 : 0;   // max_iter < 0: the reference's loop body never runs
}

// The cooperative step driven from the host, one item at a time: same interface as IteratedUpdate.  Used by the
// CPU tests (flimo_ekf_* with FLIMO_EKF_COOP=1) to pin the device code path's algebra without a GPU.
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
    upd_state_begin(st_, in_);
    for (int i = 0; i < N * N; ++i) st_.P[i] = P529[i];
  }
  bool done() const { return st_.done != 0; }
  bool failed() const { return st_.failed != 0; }
  int passes() const { return st_.passes; }
  void state(double* x26) const { for (int i = 0; i < 26; ++i) x26[i] = st_.x[i]; }
  void end(double* x26, double* P529) const {
    state(x26);
    for (int i = 0; i < N * N; ++i) P529[i] = st_.P[i];
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
    iterated_step(ex, sh_, in_, st_, packed);
    return done();
  }

 private:
  UpdInit in_;
  UpdState st_;
  StepShared sh_;
};

}  // namespace ekf
}  // namespace flimo
