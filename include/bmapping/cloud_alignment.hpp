// bmapping/cloud_alignment.hpp - stand-in for bmapping/include/bmapping/cloud_alignment.hpp:28-46 for builds WITHOUT
// PCL.  The reference's ScanAlignment wraps pcl::IterativeClosestPoint (cloud_alignment.cpp:37-72,160-223); PCL is
// not part of this repo's hot path and stays on the caller's side of the boundary: bmapping::ParticleFilter only
// needs `bool pclICPWrapper(Transform2D &T, const Transform2D &T_init, const std::vector<float> &)`.
//
//  * ROS build with PCL: keep the reference's own cloud_alignment.hpp / cloud_alignment.cpp (put the reference's
//    bmapping include directory BEFORE this one for that header); its class has the same name and method.
//  * no PCL: this class reports "no match" (SLAM() takes the motion-model branch, particle_filter.cpp:161-176)
//    unless an outcome is injected with setResult() - the seam the parity tests use.
//  * no PCL, scan matching wanted: GpuScanAlignment below runs libb2nav's own ICP kernel (b2n_icp_*).
#ifndef B2N_BMAPPING_CLOUD_ALIGNMENT_HPP
#define B2N_BMAPPING_CLOUD_ALIGNMENT_HPP

#include <stdexcept>
#include <vector>

#if __has_include(<rigid2d/rigid2d.hpp>)
#include <rigid2d/rigid2d.hpp>
#else
#include "../rigid2d_min/types.hpp"
#endif

#include "../b2nav.h"
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

protected:
  LaserProperties props_;
  Transform2D Trs_;
  bool ok_ = false;
  Transform2D T_;
};

/// ScanAlignment whose matcher runs on the GPU (b2n_icp_*): libb2nav's own point-to-point ICP
/// with the reference's settings (cloud_alignment.cpp:20-25) and wrapper semantics (:37-72).  Not PCL - see b2nav.h.
class GpuScanAlignment : public ScanAlignment
{
public:
  GpuScanAlignment(const LaserProperties &props, const Transform2D &Trs, int max_beams = 0) : ScanAlignment(props, Trs)
  {
    b2n_icp_params p{};
    p.beam_min = props.beam_min; p.beam_max = props.beam_max; p.beam_delta = props.beam_delta;
    p.range_min = props.range_min; p.range_max = props.range_max;
    p.max_iter = 100; p.max_correspondence_dist = 0.5; p.transformation_epsilon = 1e-8; p.euclidean_fitness_epsilon = 1e-6;
    p.device = -1; p.max_beams = max_beams;
    if (b2n_icp_create(&p, &h_) != B2N_OK) throw std::runtime_error(b2n_last_error());
  }
  ~GpuScanAlignment() override { b2n_icp_destroy(h_); }
  GpuScanAlignment(const GpuScanAlignment &) = delete;
  GpuScanAlignment &operator=(const GpuScanAlignment &) = delete;

  bool pclICPWrapper(Transform2D &T, const Transform2D &T_init, const std::vector<float> &beam_length) override
  {
    const auto g = T_init.displacement();
    const double ti[3] = {g.theta, g.x, g.y};
    double t[3] = {0.0, 0.0, 0.0};
    int ok = 0;
    if (b2n_icp_align(h_, beam_length.data(), (int)beam_length.size(), ti, t, &ok) != B2N_OK) throw std::runtime_error(b2n_last_error());
    if (ok && calls_++ > 0) T = Transform2D(Vector2D(t[1], t[2]), t[0]);   // the first call leaves T untouched (:49-54)
    return ok != 0;
  }

private:
  b2n_icp *h_ = nullptr;
  int calls_ = 0;
};
} // namespace bmapping
#endif
