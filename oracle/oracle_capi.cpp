// ORACLE — TEST INFRASTRUCTURE ONLY.  C entry points (ctypes) over the CPU restatement of
// the reference registration path.  Used by tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py as the CHECKER / CPU BASELINE.
// The product library (fast_limo_b200/csrc) never links or calls anything in oracle/.
//
// Parity status: octree PINNED against oracle/_ref (the reference's Octree.hpp compiled unmodified); the rest
// UNPINNED (see ioctree.hpp / plane_match.hpp / ekf.hpp / prep.hpp headers).
#include <cstdint>
#include <cstring>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "ekf.hpp"
#include "ioctree.hpp"
#include "plane_match.hpp"
#include "predict.hpp"
#include "prep.hpp"

using namespace orc;

extern "C" {

struct orc_cfg {
  int32_t k;
  int32_t estimate_extrinsics;
  int32_t num_threads;
  int32_t _pad;
  int64_t max_pc2match;
  int64_t max_matches;
  double max_dist_plane;
  double plane_threshold;
};

static MatchCfg to_cfg(const orc_cfg* c) {
  MatchCfg m;
  m.k = c->k;
  m.estimate_extrinsics = c->estimate_extrinsics;
  m.num_threads = c->num_threads < 1 ? 1 : c->num_threads;
  m.max_pc2match = c->max_pc2match;
  m.max_matches = c->max_matches;
  m.max_dist_plane = c->max_dist_plane;
  m.plane_threshold = c->plane_threshold;
  return m;
}

int orc_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// NB the reference ignores the YAML bucket size (Octree.hpp:178-180 self-assignment): pass 32.
void* orc_map_new(int bucket, float min_extent, int downsample) {
  IOctree* t = new IOctree;
  t->bucket = static_cast<size_t>(bucket);
  t->min_extent = min_extent;
  t->downsample = downsample != 0;
  return t;
}
void orc_map_free(void* m) { delete static_cast<IOctree*>(m); }

// Mapper::add (Mapper.cpp:88-96): first call builds, later calls update.
void orc_map_add(void* m, const float* xyz, size_t n, size_t stride_floats) {
  IOctree* t = static_cast<IOctree*>(m);
  if (n < 1) return;
  if (t->size() == 0)
    t->build(xyz, n, stride_floats);
  else
    t->insert(xyz, n, stride_floats);
}
size_t orc_map_size(void* m) { return static_cast<IOctree*>(m)->size(); }

size_t orc_map_dump(void* m, float* out, size_t cap_pts) {
  std::vector<P3> pts;
  static_cast<IOctree*>(m)->dump(pts);
  const size_t n = pts.size() < cap_pts ? pts.size() : cap_pts;
  for (size_t i = 0; i < n; ++i) {
    out[3 * i] = pts[i].x;
    out[3 * i + 1] = pts[i].y;
    out[3 * i + 2] = pts[i].z;
  }
  return pts.size();
}

void orc_knn(void* m, const float* q, size_t nq, size_t stride, int k, int threads, float* d2, float* nb, int32_t* cnt) {
  const IOctree* t = static_cast<const IOctree*>(m);
#pragma omp parallel for num_threads(threads < 1 ? 1 : threads)
  for (long i = 0; i < static_cast<long>(nq); ++i) {
    std::vector<P3> p(k);
    std::vector<float> d(k);
    const float* qq = q + i * stride;
    const int got = t->knn(P3{qq[0], qq[1], qq[2]}, k, p.data(), d.data());
    cnt[i] = got;
    for (int j = 0; j < k; ++j) {
      d2[i * k + j] = j < got ? d[j] : INFINITY;
      for (int c = 0; c < 3; ++c) nb[(i * k + j) * 3 + c] = j < got ? (&p[j].x)[c] : 0.f;
    }
  }
}

void orc_plane_fit(const float* pts, int k, float out[4]) {
  std::vector<float> A(pts, pts + 3 * k), b(k, -1.0f);
  float x[3];
  colpiv_qr_solve(A.data(), k, b.data(), x);
  const float nrm = std::sqrt(x[0] * x[0] + (x[1] * x[1] + x[2] * x[2]));
  out[0] = x[0] / nrm;
  out[1] = x[1] / nrm;
  out[2] = x[2] / nrm;
  out[3] = static_cast<float>(1.0 / static_cast<double>(nrm));
}

// pcl::transformPointCloud(*pc2match, *final_scan, state.get_RT()) of Localizer.cpp:361 — the same float evaluation of
// T * (x, y, z, 1) as the per-point transform of Mapper::match (Mapper.cpp:71-72).  out: 3 floats per point.
void orc_scan_to_world(const double* state14, const float* scan, size_t n, size_t stride, float* out) {
  FState s;
  s.prepare(state14, state14 + 3, state14 + 7, state14 + 11);
  for (size_t i = 0; i < n; ++i) affine_apply(s.R_wb, s.t_wb, scan + i * stride, out + 3 * i);
}

// One measurement pass (h_share_model): Mapper::match + Localizer::calculate_H + HTH/HTh.
// state14 = pos[3], rot[4] (x,y,z,w), offset_R_L_I[4], offset_T_L_I[3].
// Per-query outputs (may be null) are sized n_q = min(n, max_pc2match).
// H/h (may be null) must hold min(n_valid, max_matches) rows.  Returns the row count.
long orc_match(void* m, const orc_cfg* c, const double* state14, const float* scan, size_t n, size_t stride,
               uint8_t* good, float* plane, float* dist, float* world, float* nn_d2, double* H, double* h,
               double* HTH, double* HTh, int64_t* n_valid) {
  const IOctree* t = static_cast<const IOctree*>(m);
  const MatchCfg cfg = to_cfg(c);
  FState s;
  s.prepare(state14, state14 + 3, state14 + 7, state14 + 11);
  std::vector<PointMatch> all, chosen;
  match_scan(*t, cfg, s, scan, n, stride, all, chosen);
  for (size_t i = 0; i < all.size(); ++i) {
    if (good) good[i] = all[i].good ? 1 : 0;
    if (plane) std::memcpy(plane + 4 * i, all[i].n, 4 * sizeof(float));
    if (dist) dist[i] = all[i].dist;
    if (world) std::memcpy(world + 3 * i, all[i].g, 3 * sizeof(float));
    if (nn_d2)
      for (int j = 0; j < cfg.k; ++j) nn_d2[i * cfg.k + j] = j < all[i].nn_cnt ? all[i].nn_d2[j] : INFINITY;
  }
  if (n_valid) *n_valid = static_cast<int64_t>(chosen.size());
  const long nv = static_cast<long>(chosen.size());
  const long rows = nv > cfg.max_matches ? cfg.max_matches : nv;
  std::vector<double> Hl, hl;
  double* Hp = H;
  double* hp = h;
  if (!Hp) {
    Hl.resize(static_cast<size_t>(rows) * 12 + 12);
    Hp = Hl.data();
  }
  if (!hp) {
    hl.resize(static_cast<size_t>(rows) + 1);
    hp = hl.data();
  }
  jacobian_rows(cfg, s, chosen, Hp, hp);
  if (HTH && HTh) normal_equations(Hp, hp, rows, HTH, HTh);
  return rows;
}

// Whole iterated update (esekf::update_iterated_dyn_share_modified) with the reference's
// measurement model bound to (map, scan).  state26/P529 are updated in place.
// trace (may be null): per pass 26 (state) + 23 (dx_) + 1 (rows) + 144 (HTH) + 12 (HTh) doubles.
int orc_update(void* m, const orc_cfg* c, double* state26, double* P529, int max_iter, const double* limit23,
               double Rn, double D, const float* scan, size_t n, size_t stride, double* trace, int trace_cap) {
  const IOctree* t = static_cast<const IOctree*>(m);
  const MatchCfg cfg = to_cfg(c);
  EkfState x;
  std::memcpy(&x, state26, sizeof(x));
  std::vector<PassTrace> tr;
  MeasModel model = [&](const EkfState& xs, std::vector<double>& H, std::vector<double>& h) -> long {
    FState s;
    s.prepare(xs.pos, xs.rot, xs.offR, xs.offT);
    std::vector<PointMatch> all, chosen;
    match_scan(*t, cfg, s, scan, n, stride, all, chosen);
    const long nv = static_cast<long>(chosen.size());
    const long rows = nv > cfg.max_matches ? cfg.max_matches : nv;
    H.assign(static_cast<size_t>(rows) * 12 + 12, 0.0);
    h.assign(static_cast<size_t>(rows) + 1, 0.0);
    jacobian_rows(cfg, s, chosen, H.data(), h.data());
    return rows;
  };
  const int passes = iterated_update(x, P529, max_iter, limit23, Rn, D, model, &tr);
  std::memcpy(state26, &x, sizeof(x));
  if (trace) {
    const int stride_d = 26 + 23 + 1 + 144 + 12;
    for (int i = 0; i < static_cast<int>(tr.size()) && i < trace_cap; ++i) {
      double* o = trace + static_cast<size_t>(i) * stride_d;
      std::memcpy(o, &tr[i].x_after, 26 * sizeof(double));
      std::memcpy(o + 26, tr[i].dx, 23 * sizeof(double));
      o[49] = static_cast<double>(tr[i].n_rows);
      std::memcpy(o + 50, tr[i].HTH, 144 * sizeof(double));
      std::memcpy(o + 194, tr[i].HTh, 12 * sizeof(double));
    }
  }
  return passes;
}

// Iterated update driven by caller-supplied normal equations is not part of the reference;
// the EKF algebra alone is exposed for unit tests through a synthetic linear model:
// rows H (N x 12), h (N) are fixed for every pass.
int orc_update_fixed(double* state26, double* P529, int max_iter, const double* limit23, double Rn, double D,
                     const double* Hin, const double* hin, long N, double* trace, int trace_cap) {
  EkfState x;
  std::memcpy(&x, state26, sizeof(x));
  std::vector<PassTrace> tr;
  MeasModel model = [&](const EkfState&, std::vector<double>& H, std::vector<double>& h) -> long {
    H.assign(Hin, Hin + N * 12);
    h.assign(hin, hin + N);
    H.resize(static_cast<size_t>(N) * 12 + 12);
    h.resize(static_cast<size_t>(N) + 1);
    return N;
  };
  const int passes = iterated_update(x, P529, max_iter, limit23, Rn, D, model, &tr);
  std::memcpy(state26, &x, sizeof(x));
  if (trace) {
    const int stride_d = 26 + 23 + 1 + 144 + 12;
    for (int i = 0; i < static_cast<int>(tr.size()) && i < trace_cap; ++i) {
      double* o = trace + static_cast<size_t>(i) * stride_d;
      std::memcpy(o, &tr[i].x_after, 26 * sizeof(double));
      std::memcpy(o + 26, tr[i].dx, 23 * sizeof(double));
      o[49] = static_cast<double>(tr[i].n_rows);
      std::memcpy(o + 50, tr[i].HTH, 144 * sizeof(double));
      std::memcpy(o + 194, tr[i].HTh, 12 * sizeof(double));
    }
  }
  return passes;
}

void orc_boxplus(double* state26, const double* d23) {
  EkfState x;
  std::memcpy(&x, state26, sizeof(x));
  state_boxplus(x, d23);
  std::memcpy(state26, &x, sizeof(x));
}
void orc_boxminus(const double* a26, const double* b26, double* d23) {
  EkfState a, b;
  std::memcpy(&a, a26, sizeof(a));
  std::memcpy(&b, b26, sizeof(b));
  state_boxminus(a, b, d23);
}
int orc_invert(double* A, int n) { return invert(A, n) ? 0 : -1; }

// ---- scan preparation (prep.hpp): filters, time sort, deskew, voxel grid ------------------------
// order_out (capacity n) receives the indices into raw of the filtered cloud [sorted by time if sort != 0].
size_t orc_prep_filter_sort(const void* raw, size_t n, const PrepCfg* c, int sort, uint32_t* order_out) {
  const RawPoint* r = static_cast<const RawPoint*>(raw);
  std::vector<uint32_t> kept = prep_filter(r, n, *c);
  if (sort) prep_sort(r, kept, *c);
  std::memcpy(order_out, kept.data(), kept.size() * sizeof(uint32_t));
  return kept.size();
}
void orc_prep_times(const void* raw, const uint32_t* order, size_t m, const PrepCfg* c, double sweep_ref_time, double* t_out) {
  const RawPoint* r = static_cast<const RawPoint*>(raw);
  for (size_t k = 0; k < m; ++k) t_out[k] = prep_point_time(r[order[k]], *c, sweep_ref_time);
}
void orc_prep_deskew(const void* raw, const uint32_t* order, size_t m, const PrepCfg* c, double sweep_ref_time, double offset,
                     const Frame* frames, int n_frames, const float* last_q, const float* last_p, const float* T_l2b,
                     float* out_world4, float* out_xt2_4) {
  std::vector<uint32_t> ord(order, order + m);
  prep_deskew(static_cast<const RawPoint*>(raw), ord, *c, sweep_ref_time, offset, frames, n_frames, last_q, last_p, T_l2b,
              out_world4, out_xt2_4);
}
size_t orc_prep_voxel(const float* in4, size_t n, float leaf, float* out4) {
  const std::vector<float> o = prep_voxel(in4, n, leaf);
  std::memcpy(out4, o.data(), o.size() * sizeof(float));
  return o.size() / 4;
}


// ---- IMU rate (predict.hpp): prediction + propagated states -------------------------------------
void* orc_prop_new(const double* state26, const double* P529) {
  Propagator* p = new Propagator();
  std::memcpy(&p->x, state26, sizeof(EkfState));
  std::memcpy(p->P.data(), P529, 529 * sizeof(double));
  return p;
}
void orc_prop_free(void* p) { delete static_cast<Propagator*>(p); }
void orc_prop_propagate(void* p, double stamp, double dt, const float* lin_accel, const float* ang_vel, const double* cov4) {
  static_cast<Propagator*>(p)->propagate(stamp, dt, lin_accel, ang_vel, cov4);
}
void orc_prop_get(void* p, double* state26, double* P529) {
  Propagator* q = static_cast<Propagator*>(p);
  std::memcpy(state26, &q->x, sizeof(EkfState));
  std::memcpy(P529, q->P.data(), 529 * sizeof(double));
}
// out may be NULL to query the count; returns -1 / 0 / n as Propagator::frames_in_range
long orc_prop_frames(void* p, double start_time, double end_time, Frame* out, size_t cap) {
  std::vector<Frame> fr;
  const long n = static_cast<Propagator*>(p)->frames_in_range(start_time, end_time, fr);
  if (n > 0 && out && cap >= static_cast<size_t>(n)) std::memcpy(out, fr.data(), n * sizeof(Frame));
  return n;
}
// the process model alone, for the finite-difference checks: f (24), df/dx (24 x 23), df/dw (24 x 12)
void orc_process_model(const double* state26, const double* acc, const double* gyro, double* f, double* fx, double* fw) {
  EkfState s;
  std::memcpy(&s, state26, sizeof(EkfState));
  process_f(s, acc, gyro, f);
  process_fx(s, acc, fx);
  process_fw(s, fw);
}

}  // extern "C"
