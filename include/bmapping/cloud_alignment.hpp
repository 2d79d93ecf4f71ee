// bmapping/cloud_alignment.hpp - stand-in for bmapping/include/bmapping/cloud_alignment.hpp:28-46 for builds WITHOUT
// PCL.  The reference's ScanAlignment wraps pcl::IterativeClosestPoint (cloud_alignment.cpp:37-72,160-223); PCL is
// not part of this repo's hot path and stays on the caller's side of the boundary: bmapping::ParticleFilter only
// needs `bool pclICPWrapper(Transform2D &T, const Transform2D &T_init, const std::vector<float> &)`.
//
//  * ROS build with PCL: keep the reference's own cloud_alignment.hpp / cloud_alignment.cpp (put the reference's
//    bmapping include directory BEFORE this one for that header); its class has the same name and method.
//  * no PCL: this class reports "no match" (SLAM() takes the motion-model branch, particle_filter.cpp:161-176)
//    unless an outcome is injected with setResult() - the seam the parity tests use.
#ifndef B2N_BMAPPING_CLOUD_ALIGNMENT_HPP
#define B2N_BMAPPING_CLOUD_ALIGNMENT_HPP

#include <vector>

#if __has_include(<rigid2d/rigid2d.hpp>)
#include <rigid2d/rigid2d.hpp>
#else
#include "../rigid2d_min/types.hpp"
#endif

#include "sensor_model.hpp"

namespace bmapping
{
using rigid2d::Transform2D;
using rigid2d::Vector2D;

class ScanAlignment
{
public:
  ScanAlignment(const LaserProperties &props, const Transform2D &Trs) : props_(props), Trs_(Trs) {}
  virtual ~ScanAlignment() = default;

  /// reference cloud_alignment.hpp:45-46: T [out] = transform between the previous and the current scan
  virtual bool pclICPWrapper(Transform2D &T, const Transform2D &T_init, const std::vector<float> &beam_length)
  {
    (void)T_init;
    (void)beam_length;
    if (ok_) T = T_;
    return ok_;
  }

  /// not in the reference: fix the outcome of the next pclICPWrapper() calls
  void setResult(bool ok, const Transform2D &T = Transform2D())
  {
    ok_ = ok;
    T_ = T;
  }

private:
  LaserProperties props_;
  Transform2D Trs_;
  bool ok_ = false;
  Transform2D T_;
};
} // namespace bmapping
#endif
