// ORACLE — TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// CPU restatement of the reference's incremental octree ("i-Octree"), the map
// container behind fast_limo::Mapper.  Parity status: PINNED against the reference
// itself — oracle/_ref/libref_octree.so is the reference's Octree.hpp compiled where
// it lies, unmodified (make -C oracle ref; Eigen::Vector3f stand-in in ref_shim/), and
// tests/test_ref_octree.py + tests/golden/ref_octree_*.npz show this restatement
// bit-identical to it (sizes after every update, contents, 5-NN distances and order).
// This file follows the reference line by line in behaviour, not in text.
//
// Follows /root/reference/include/fast_limo/Objects/Octree.hpp:
//   Heap              :45-88     -> KnnHeap
//   Octant            :103-132   -> Node
//   ctor / setters    :155-184   -> IOctree()  (NB setBucketSize() at :178-180 is a
//                                   self-assignment, so the bucket stays 32)
//   processPoints     :234-267   -> gather_finite()
//   mortonCode        :269-275   -> octant_of()
//   initialize        :282-298   -> build()
//   createOctant      :301-338   -> make_node()
//   update/expandTree :341-377   -> insert()
//   updateOctant      :380-432   -> insert_into()
//   overlaps          :435-450   -> ball_touches()
//   knn               :526-599   -> knn() / knn_visit()
//
// Arithmetic notes (all float32, no FMA contraction — build with -ffp-contract=off):
//   * squared distance follows Eigen's fixed-size-3 reduction order:
//     d2 = dx*dx + (dy*dy + dz*dz)          (Eigen redux_novec_unroller<0,3>)
//   * child centre = parent centre + (+-0.5f * parent_extent), child extent = 0.5f*extent.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <utility>
#include <vector>

namespace orc {

struct P3 {
  float x, y, z;
};

inline float sqdist3(const P3& a, const P3& b) {
  const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
  const float xx = dx * dx, yy = dy * dy, zz = dz * dz;
  return xx + (yy + zz);
}

// Bounded list of the k best candidates, ascending by squared distance.
// On equal distance the earlier-visited candidate stays in front (Octree.hpp:73,80).
struct KnnHeap {
  struct Item {
    float d2;
    P3 p;
  };
  size_t cap, cnt;
  std::vector<Item> items;  // per-query heap allocation, as in the reference (:58)

  explicit KnnHeap(size_t k) : cap(k), cnt(0), items(k) {
    for (auto& it : items) it.d2 = FLT_MAX;
  }
  bool full() const { return cnt == cap; }
  float worst() const { return full() ? items[cnt - 1].d2 : FLT_MAX; }
  void offer(const P3& p, float d2) {
    if (full() && d2 >= items[cnt - 1].d2) return;
    if (cnt < cap) ++cnt;
    int i = static_cast<int>(cnt) - 1;
    while (i > 0 && items[i - 1].d2 > d2) {
      items[i] = items[i - 1];
      --i;
    }
    items[i].d2 = d2;
    items[i].p = p;
  }
};

struct Node {
  P3 c{0, 0, 0};     // cube centre
  float ext = 0.f;   // half side length
  std::vector<P3> pts;      // payload when leaf
  Node** kid = nullptr;     // 8 slots when interior, nullptr when leaf
  ~Node() {
    if (kid) {
      for (int i = 0; i < 8; ++i) delete kid[i];
      delete[] kid;
    }
  }
  void make_interior() { kid = new Node*[8](); }
};

class IOctree {
 public:
  // Defaults are the reference's effective values (Octree.hpp:155-159).
  size_t bucket = 32;
  float min_extent = 0.2f;
  bool downsample = true;

  IOctree() = default;
  ~IOctree() { delete root_; }
  IOctree(const IOctree&) = delete;
  IOctree& operator=(const IOctree&) = delete;

  size_t size() const { return num_points_; }
  bool empty_tree() const { return root_ == nullptr; }
  const Node* root() const { return root_; }

  void clear() {
    delete root_;
    root_ = nullptr;
  }

  // Octree::initialize — NB num_points_ is NOT reset by the reference's clear().
  void build(const float* xyz, size_t n, size_t stride) {
    clear();
    P3 lo, hi;
    std::vector<P3> pts = gather_finite(xyz, n, stride, lo, hi);
    if (pts.empty()) return;
    const P3 half{0.5f * (hi.x - lo.x), 0.5f * (hi.y - lo.y), 0.5f * (hi.z - lo.z)};
    const P3 c{lo.x + half.x, lo.y + half.y, lo.z + half.z};
    float e = half.x;
    if (half.y > e) e = half.y;
    if (half.z > e) e = half.z;
    root_ = make_node(c, e, pts);
  }

  // Octree::update
  void insert(const float* xyz, size_t n, size_t stride) {
    if (!root_) {
      build(xyz, n, stride);
      return;
    }
    P3 lo, hi;
    std::vector<P3> pts = gather_finite(xyz, n, stride, lo, hi);
    // Deviation (documented): an all-NaN batch makes the reference double the root until
    // its extent overflows to +inf; the oracle returns instead.
    if (pts.empty()) return;
    grow_to(hi);
    grow_to(lo);
    insert_into(root_, pts);
  }

  // Octree::knn (public).  Writes up to k neighbours ascending; returns how many.
  int knn(const P3& q, int k, P3* out_p, float* out_d2) const {
    if (!root_) return 0;
    KnnHeap heap(static_cast<size_t>(k));
    P3 qq = q;
    knn_visit(root_, qq, heap);
    std::vector<KnnHeap::Item> got(heap.items.begin(), heap.items.begin() + heap.cnt);  // get_data() copy (:64-66)
    for (size_t i = 0; i < got.size(); ++i) {
      out_p[i] = got[i].p;
      out_d2[i] = got[i].d2;
    }
    return static_cast<int>(got.size());
  }

  // Depth-first dump in child order 0..7 (Octree::get_points :217-228).
  void dump(std::vector<P3>& out) const { dump_rec(root_, out); }

 private:
  Node* root_ = nullptr;
  size_t num_points_ = 0;

  static int octant_of(const P3& p, const P3& c) {
    int m = 0;
    if (p.x > c.x) m |= 1;
    if (p.y > c.y) m |= 2;
    if (p.z > c.z) m |= 4;
    return m;
  }

  static std::vector<P3> gather_finite(const float* xyz, size_t n, size_t stride, P3& lo, P3& hi) {
    lo = {FLT_MAX, FLT_MAX, FLT_MAX};
    hi = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    std::vector<P3> out;
    out.resize(n);
    size_t m = 0;
    for (size_t i = 0; i < n; ++i) {
      const float* p = xyz + i * stride;
      if (std::isnan(p[0]) || std::isnan(p[1]) || std::isnan(p[2])) continue;
      out[m++] = P3{p[0], p[1], p[2]};
      lo.x = p[0] < lo.x ? p[0] : lo.x;
      lo.y = p[1] < lo.y ? p[1] : lo.y;
      lo.z = p[2] < lo.z ? p[2] : lo.z;
      hi.x = p[0] > hi.x ? p[0] : hi.x;
      hi.y = p[1] > hi.y ? p[1] : hi.y;
      hi.z = p[2] > hi.z ? p[2] : hi.z;
    }
    out.resize(m);
    return out;
  }

  static P3 child_centre(const P3& c, float ext, int i) {
    static const float f[2] = {-0.5f, 0.5f};
    return P3{c.x + f[(i & 1) > 0] * ext, c.y + f[(i & 2) > 0] * ext, c.z + f[(i & 4) > 0] * ext};
  }

  Node* make_node(const P3& c, float ext, const std::vector<P3>& pts) {
    Node* nd = new Node;
    nd->c = c;
    nd->ext = ext;
    if (pts.size() > bucket && ext > 2 * min_extent) {
      nd->make_interior();
      std::vector<std::vector<P3>> part(8);
      for (const P3& p : pts) part[octant_of(p, c)].push_back(p);
      for (int i = 0; i < 8; ++i) {
        if (part[i].empty()) continue;
        nd->kid[i] = make_node(child_centre(c, ext, i), ext * 0.5f, part[i]);
      }
    } else {
      num_points_ += pts.size();
      nd->pts = pts;
    }
    return nd;
  }

  void grow_to(const P3& b) {
    static const float f[2] = {-0.5f, 0.5f};
    for (;;) {
      float ax = std::fabs(b.x - root_->c.x), ay = std::fabs(b.y - root_->c.y), az = std::fabs(b.z - root_->c.z);
      float m = ax;
      if (ay > m) m = ay;
      if (az > m) m = az;
      if (!(m > root_->ext)) break;
      const float pe = 2 * root_->ext;
      const P3 pc{root_->c.x + f[b.x > root_->c.x] * pe, root_->c.y + f[b.y > root_->c.y] * pe,
                  root_->c.z + f[b.z > root_->c.z] * pe};
      Node* up = new Node;
      up->c = pc;
      up->ext = pe;
      up->make_interior();
      up->kid[octant_of(root_->c, pc)] = root_;
      root_ = up;
    }
  }

  void insert_into(Node*& nd, const std::vector<P3>& pts) {
    if (nd->kid == nullptr) {
      if (nd->pts.size() + pts.size() > bucket && nd->ext > 2 * min_extent) {
        num_points_ -= nd->pts.size();
        nd->pts.insert(nd->pts.end(), pts.begin(), pts.end());
        Node* fresh = make_node(nd->c, nd->ext, nd->pts);
        delete nd;
        nd = fresh;
      } else {
        if (downsample && nd->ext <= 2 * min_extent && nd->pts.size() > bucket / 8) return;  // whole sub-batch dropped
        nd->pts.insert(nd->pts.end(), pts.begin(), pts.end());
        num_points_ += pts.size();
      }
      return;
    }
    std::vector<std::vector<P3>> part(8);
    for (const P3& p : pts) part[octant_of(p, nd->c)].push_back(p);
    for (int i = 0; i < 8; ++i) {
      if (part[i].empty()) continue;
      if (nd->kid[i] == nullptr)
        nd->kid[i] = make_node(child_centre(nd->c, nd->ext, i), nd->ext * 0.5f, part[i]);
      else
        insert_into(nd->kid[i], part[i]);
    }
  }

  static bool ball_touches(const Node* nd, const P3& q, float r2) {
    const float dx = std::fabs(q.x - nd->c.x) - nd->ext;
    const float dy = std::fabs(q.y - nd->c.y) - nd->ext;
    const float dz = std::fabs(q.z - nd->c.z) - nd->ext;
    if ((dx > 0 && dx * dx > r2) || (dy > 0 && dy * dy > r2) || (dz > 0 && dz * dz > r2)) return false;
    const int inside_axes = (dx < 0) + (dy < 0) + (dz < 0);
    if (inside_axes > 1) return true;
    const float mx = dx > 0.f ? dx : 0.f, my = dy > 0.f ? dy : 0.f, mz = dz > 0.f ? dz : 0.f;
    return (mx * mx + (my * my + mz * mz)) < r2;
  }

  static bool ball_inside(const Node* nd, const P3& q, float r2) {
    const float dx = nd->ext - std::fabs(q.x - nd->c.x);
    if (dx < 0 || dx * dx < r2) return false;
    const float dy = nd->ext - std::fabs(q.y - nd->c.y);
    if (dy < 0 || dy * dy < r2) return false;
    const float dz = nd->ext - std::fabs(q.z - nd->c.z);
    if (dz < 0 || dz * dz < r2) return false;
    return true;
  }

  // Returns true when the search can stop (k found and their ball lies inside this cube).
  static bool knn_visit(const Node* nd, P3& q, KnnHeap& heap) {
    static const int order[8][7] = {{1, 2, 4, 3, 5, 6, 7}, {0, 3, 5, 2, 4, 7, 6}, {0, 3, 6, 1, 4, 7, 5},
                                    {1, 2, 7, 0, 5, 6, 4}, {0, 5, 6, 1, 2, 7, 3}, {1, 4, 7, 0, 3, 6, 2},
                                    {2, 4, 7, 0, 3, 5, 1}, {3, 5, 6, 1, 2, 4, 0}};
    if (nd->kid == nullptr) {
      for (const P3& p : nd->pts) heap.offer(p, sqdist3(q, p));
      return heap.full() && ball_inside(nd, q, heap.worst());
    }
    const int home = octant_of(q, nd->c);
    if (nd->kid[home] != nullptr && knn_visit(nd->kid[home], q, heap)) return true;
    for (int i = 0; i < 7; ++i) {
      const int c = order[home][i];
      if (nd->kid[c] == nullptr) continue;
      if (heap.full() && !ball_touches(nd->kid[c], q, heap.worst())) continue;
      if (knn_visit(nd->kid[c], q, heap)) return true;
    }
    return heap.full() && ball_inside(nd, q, heap.worst());
  }

  static void dump_rec(const Node* nd, std::vector<P3>& out) {
    if (!nd) return;
    if (!nd->kid) {
      out.insert(out.end(), nd->pts.begin(), nd->pts.end());
      return;
    }
    for (int i = 0; i < 8; ++i) dump_rec(nd->kid[i], out);
  }
};

}  // namespace orc
