// ORACLE — TEST INFRASTRUCTURE ONLY.  C entry points over the REFERENCE's own octree
// (/root/reference/include/fast_limo/Objects/Octree.hpp, compiled where it lies, unmodified) so that the
// restatement in oracle/ioctree.hpp can be pinned against the real thing: exact kNN (Octree.hpp:526-599)
// and the incremental insert with its down-sampling rule (Octree.hpp:282-432).
// Only dependency of that header is Eigen::Vector3f; oracle/ref_shim/Eigen/Dense supplies the few
// coefficient-wise operations it uses (Eigen is not installed here).  Built by `make -C oracle ref`
// into oracle/_ref/libref_octree.so; exists only where /root/reference is mounted.
#include <cstddef>
#include <cstring>
#include <vector>

#include "fast_limo/Objects/Octree.hpp"

namespace {
struct P3 {
  float x, y, z;
};
using Cloud = std::vector<P3>;
}  // namespace

extern "C" {

// bucket / downsample / min_extent exactly as Mapper::Mapper + set_config drive it (Mapper.cpp:23-45):
// default-constructed tree, then setBucketSize (a self-assignment in the reference!), setDownsample, setMinExtent.
void* ref_octree_new(int bucket_size, float min_extent, int downsample) {
  auto* t = new fast_limo::octree::Octree();
  t->setBucketSize((size_t)bucket_size);
  t->setDownsample(downsample != 0);
  t->setMinExtent(min_extent);
  return t;
}
void ref_octree_free(void* h) { delete static_cast<fast_limo::octree::Octree*>(h); }

void ref_octree_update(void* h, const float* xyz, size_t n) {
  Cloud c(n);
  for (size_t i = 0; i < n; ++i) c[i] = P3{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
  Cloud* pc = &c;
  static_cast<fast_limo::octree::Octree*>(h)->update(pc);
}
size_t ref_octree_size(void* h) { return static_cast<fast_limo::octree::Octree*>(h)->size(); }

size_t ref_octree_dump(void* h, float* out_xyz, size_t cap) {
  auto pts = static_cast<fast_limo::octree::Octree*>(h)->getData<P3, std::vector<P3>>();
  const size_t n = pts.size() < cap ? pts.size() : cap;
  for (size_t i = 0; i < n; ++i) {
    out_xyz[3 * i] = pts[i].x;
    out_xyz[3 * i + 1] = pts[i].y;
    out_xyz[3 * i + 2] = pts[i].z;
  }
  return pts.size();
}

// k nearest neighbours of each query: out_xyz [nq][k][3], out_d2 [nq][k] (+inf padded), out_cnt [nq]
void ref_octree_knn(void* h, const float* q_xyz, size_t nq, int k, float* out_xyz, float* out_d2, int* out_cnt) {
  auto* t = static_cast<fast_limo::octree::Octree*>(h);
  for (size_t i = 0; i < nq; ++i) {
    std::vector<P3> nb;
    std::vector<float> d;
    t->knn(P3{q_xyz[3 * i], q_xyz[3 * i + 1], q_xyz[3 * i + 2]}, k, nb, d);
    out_cnt[i] = (int)nb.size();
    for (int j = 0; j < k; ++j) {
      const bool have = j < (int)nb.size();
      out_xyz[(i * k + j) * 3] = have ? nb[j].x : 0.f;
      out_xyz[(i * k + j) * 3 + 1] = have ? nb[j].y : 0.f;
      out_xyz[(i * k + j) * 3 + 2] = have ? nb[j].z : 0.f;
      out_d2[i * k + j] = have ? d[j] : __builtin_inff();
    }
  }
}

}  // extern "C"
