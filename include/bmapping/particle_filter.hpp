// bmapping/particle_filter.hpp - drop-in for bmapping/include/bmapping/particle_filter.hpp:89-144: same namespace,
// class name, constructor argument list and method signatures, so bmapping/src/turtle_mapping_node.cpp
// (:401-410 construction, :474 SLAM, :479 newMap, :494 getRobotState) compiles against it unchanged and links
// libb2nav.so instead of the bmapping library.
//
// Header-only pimpl over the C ABI in include/b2nav.h: every numeric operation of SLAM() happens in the sm_100a
// kernels (per-particle maps are SoA planes in HBM).  The scan matcher stays a host object (PCL in a ROS build, the
// injectable stand-in of cloud_alignment.hpp otherwise): SLAM() calls scan_matcher.pclICPWrapper() exactly where the
// reference does (particle_filter.cpp:150-153) and hands its verdict and transform to the device.  There is no CPU
// path: construction throws when the library finds no B200.
//
// Error behaviour: the reference throws std::invalid_argument from world2Grid / world2RowMajor / pdfNormal and prints
// "eta is 0"; here B2N_ERR_INVALID_ARGUMENT, B2N_ERR_OFF_MAP and B2N_ERR_NUMERIC are re-thrown as
// std::invalid_argument, every other non-zero status as std::runtime_error.
#ifndef B2N_BMAPPING_PARTICLE_FILTER_HPP
#define B2N_BMAPPING_PARTICLE_FILTER_HPP

#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <random>
#include <vector>

#if __has_include(<rigid2d/diff_drive.hpp>)
#include <rigid2d/diff_drive.hpp>
#include <rigid2d/rigid2d.hpp>
#else
#include "../rigid2d_min/types.hpp"
#endif

#include "../b2nav.h"
#include "cloud_alignment.hpp"
#include "grid_mapper.hpp"
#include "sensor_model.hpp"

namespace bmapping
{
using rigid2d::Pose;
using rigid2d::Transform2D;
using rigid2d::Twist2D;
using rigid2d::Vector2D;

namespace detail
{
inline void check(int rc)
{
  if (rc == B2N_OK) return;
  const std::string text = std::string("libb2nav: ") + b2n_last_error();
  if (rc == B2N_ERR_INVALID_ARGUMENT || rc == B2N_ERR_OFF_MAP || rc == B2N_ERR_NUMERIC) throw std::invalid_argument(text);
  throw std::runtime_error(text);
}

// rigid2d::normalize_angle_PI, rigid2d.hpp:52-64
inline double normalize_angle_PI(double rad)
{
  const double PI = 3.14159265358979323846;
  const double q = std::floor((rad + PI) / (2.0 * PI));
  rad = (rad + PI) - q * 2.0 * PI;
  if (rad < 0) rad += 2.0 * PI;
  return rad - PI;
}
} // namespace detail

class ParticleFilter
{
public:
  /// reference particle_filter.hpp:112-130, particle_filter.cpp:64-138
  ParticleFilter(int num_particles, int k, double srr, double srt, double str, double stt, double motion_noise_theta,
                 double motion_noise_x, double motion_noise_y, double sample_range_theta, double sample_range_x, double sample_range_y,
                 double scan_likelihood_min, double scan_likelihood_max, double pose_likelihood_min, double pose_likelihood_max,
                 ScanAlignment &scan_matcher, const Transform2D &pose, const GridMapper &mapper)
    : scan_matcher_(scan_matcher)
  {
    b2n_pf_params p{};
    const LaserProperties &L = mapper.props;
    p.beam_min = L.beam_min; p.beam_max = L.beam_max; p.beam_delta = L.beam_delta; p.range_min = L.range_min; p.range_max = L.range_max;
    p.z_hit = L.z_hit; p.z_short = L.z_short; p.z_max = L.z_max; p.z_rand = L.z_rand; p.sigma_hit = L.sigma_hit;
    p.resolution = mapper.resolution; p.xmin = mapper.xmin; p.xmax = mapper.xmax; p.ymin = mapper.ymin; p.ymax = mapper.ymax;
    p.num_particles = num_particles; p.k = k;
    p.srr = srr; p.srt = srt; p.str = str; p.stt = stt;
    p.motion_noise_theta = motion_noise_theta; p.motion_noise_x = motion_noise_x; p.motion_noise_y = motion_noise_y;
    p.sample_range_theta = sample_range_theta; p.sample_range_x = sample_range_x; p.sample_range_y = sample_range_y;
    p.scan_likelihood_min = scan_likelihood_min; p.scan_likelihood_max = scan_likelihood_max;
    p.pose_likelihood_min = pose_likelihood_min; p.pose_likelihood_max = pose_likelihood_max;
    const auto d = pose.displacement();
    p.init_pose[0] = d.theta; p.init_pose[1] = d.x; p.init_pose[2] = d.y;   // particle_filter.cpp:133
    p.particle_offset = 0; p.particles_total = num_particles; p.device = -1; p.max_beams = 0;
    detail::check(b2n_pf_create(&p, &h_));
    // like the reference, whose engine is seeded from std::random_device (particle_filter.cpp:17-22); b2n_pf_seed(handle(), ...)
    // makes a run reproducible
    std::random_device rd;
    detail::check(b2n_pf_seed(h_, ((uint64_t)rd() << 32) | (uint64_t)rd(), 0));
    int xs = 0, ys = 0;
    detail::check(b2n_pf_grid_size(h_, &xs, &ys));
    cells_ = (size_t)xs * (size_t)ys;
  }
  ~ParticleFilter() { b2n_pf_destroy(h_); }
  ParticleFilter(const ParticleFilter &) = delete;
  ParticleFilter &operator=(const ParticleFilter &) = delete;

  /// reference particle_filter.cpp:141-251
  void SLAM(const std::vector<float> &scan, const Twist2D &u, const Pose &cur_odom, const Pose &prev_odom)
  {
    // icpInitGuess (particle_filter.cpp:602-612), then the matcher, exactly where the reference calls them
    const double dth = detail::normalize_angle_PI(detail::normalize_angle_PI(cur_odom.theta) - detail::normalize_angle_PI(prev_odom.theta));
    const Transform2D Tinit(Vector2D(cur_odom.x - prev_odom.x, cur_odom.y - prev_odom.y), dth);
    Transform2D Ticp;
    const bool matcher_success = scan_matcher_.pclICPWrapper(Ticp, Tinit, scan);
    const auto t = Ticp.displacement();
    const double twist[3] = {u.w, u.vx, u.vy};
    const double cur[3] = {cur_odom.theta, cur_odom.x, cur_odom.y}, prev[3] = {prev_odom.theta, prev_odom.x, prev_odom.y};
    const double icp[3] = {t.theta, t.x, t.y};
    detail::check(b2n_pf_slam(h_, scan.data(), (int)scan.size(), twist, cur, prev, matcher_success ? 1 : 0, icp));
  }

  /// reference particle_filter.cpp:255-274
  Transform2D getRobotState()
  {
    double pose[3];
    detail::check(b2n_pf_get_robot_state(h_, pose));
    return Transform2D(Vector2D(pose[1], pose[2]), pose[0]);
  }

  /// reference particle_filter.cpp:277-291 (the reference resizes nothing: the caller sizes the vector,
  /// turtle_mapping_node.cpp:441-447; a wrongly sized vector is resized here instead of being overrun)
  void newMap(std::vector<int8_t> &map)
  {
    if (map.size() != cells_) map.resize(cells_);
    detail::check(b2n_pf_new_map(h_, map.data(), map.size()));
  }

  /// not in the reference: the C handle, for the noise seam / taps of include/b2nav.h
  b2n_pf *handle() { return h_; }

private:
  b2n_pf *h_ = nullptr;
  ScanAlignment &scan_matcher_;
  size_t cells_ = 0;
};
} // namespace bmapping
#endif
