// fast_limo::Mapper with the reference's call surface (fast_limo/Modules/Mapper.hpp:29-76) over libflimo_cuda.
//
// The global map lives in B200 device memory (flimo.h); this class owns the handle.  Everything the reference's
// Localizer / use-ikfom.cpp / ROS wrapper call on a Mapper exists with the same signature:
//     getInstance, set_num_threads, set_config, exists, size, last_time, match(State, Ptr&) -> Matches, add(Ptr&, time),
//     public member `matches`.
// match() is the reference-shaped (debug / visualisation) form of a measurement pass: the kernel's per-point records are
// read back and turned into Match objects in scan order (Mapper.cpp:59-86).  The filter itself never needs them — it
// consumes H^T H / H^T h reduced on the device (flimo_update) — so Localizer does not call match() on its hot path.
#pragma once
#include "fast_limo/Common.hpp"
#include "fast_limo/Objects/Match.hpp"
#include "fast_limo/Objects/State.hpp"
#include "fast_limo/Utils/Config.hpp"
#include "flimo.h"

class fast_limo::Mapper {
 public:
  Matches matches;

  Mapper();
  ~Mapper();

  void set_num_threads(int n);                                        // kept for the call surface (the GPU does not use it)
  void set_config(const Config::iKFoM::Mapping& cfg);                 // (re)creates the device handle
  void set_config(const Config::iKFoM::Mapping& cfg, bool estimate_extrinsics, int device);

  bool exists();
  int size();
  double last_time();

  Matches match(State, pcl::PointCloud<PointType>::Ptr&);
  void add(pcl::PointCloud<PointType>::Ptr&, double time);

  flimo_handle gpu() const { return h_; }                              // for Localizer (and tools)

  static Mapper& getInstance() {
    static Mapper* mapper = new Mapper();
    return *mapper;
  }

 private:
  Mapper(const Mapper&) = delete;
  Mapper& operator=(const Mapper&) = delete;
  void check(int rc) const;

  flimo_handle h_ = nullptr;
  Config::iKFoM::Mapping config;
  bool estimate_extrinsics_ = true;
  int device_ = 0;
  int num_threads_ = 1;
};
