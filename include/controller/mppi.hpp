// controller/mppi.hpp - drop-in for the reference header controller/include/controller/mppi.hpp:
// same namespace, class names, constructor argument lists and method signatures
// (CartModel :33, LossFunc :63, MPPI :133-155), so nuturtle_robot/src/mppi_waypoints_node.cpp
// (:186-199 construction, :216/:257 setWaypoint, :265 newControls) compiles against it unchanged
// and links libb2nav.so instead of the controller library.
//
// Header-only pimpl over the C ABI in include/b2nav.h: every numeric operation of newControls()
// happens in the sm_100a kernels.  The reference's private Eigen members (mppi.hpp:169-183) and its
// internally-used kinematicCart/loss/terminalLoss (Eigen::Ref signatures) are not part of what
// callers touch and are not reproduced.  There is no CPU path: construction throws when the
// library finds no B200.
//
// Error behaviour: the reference throws std::out_of_range from LossFunc (std::vector::at) and
// otherwise lets errors escape; here every non-zero b2n status is re-thrown as
// std::invalid_argument (B2N_ERR_INVALID_ARGUMENT) or std::runtime_error (anything else).
#ifndef B2N_CONTROLLER_MPPI_HPP
#define B2N_CONTROLLER_MPPI_HPP

#include <stdexcept>
#include <string>
#include <random>
#include <vector>

#if __has_include(<rigid2d/diff_drive.hpp>)
#include <rigid2d/diff_drive.hpp>
#else
#include "../rigid2d_min/types.hpp"
#endif

#include "../b2nav.h"

namespace controller
{
using rigid2d::Pose;
using rigid2d::WheelVelocities;

namespace detail
{
inline void check(int rc)
{
  if (rc == B2N_OK) return;
  const std::string text = std::string("libb2nav: ") + b2n_last_error();
  if (rc == B2N_ERR_INVALID_ARGUMENT) throw std::invalid_argument(text);
  throw std::runtime_error(text);
}
} // namespace detail

/// controller::CartModel, reference mppi.hpp:31-53 (parameters only; the ODE lives in the kernel)
struct CartModel
{
  CartModel(double wheel_radius, double wheel_base) : wheel_radius(wheel_radius), wheel_base(wheel_base) {}
  double wheel_radius;
  double wheel_base;
};

/// controller::LossFunc, reference mppi.hpp:58-112 (diagonals only; .at() keeps the out_of_range behaviour)
struct LossFunc
{
  LossFunc(std::vector<double> Qdiag, std::vector<double> Rdiag, std::vector<double> P1diag)
  {
    for (int i = 0; i < 3; i++) { Q[i] = Qdiag.at(i); P1[i] = P1diag.at(i); }
    for (int i = 0; i < 2; i++) R[i] = Rdiag.at(i);
  }
  double Q[3], R[2], P1[3];
};

/// controller::MPPI, reference mppi.hpp:121-183
class MPPI
{
public:
  MPPI(const CartModel &cart_model, const LossFunc &loss_func, double lambda, double max_wheel_vel, double ul_var,
       double ur_var, double horizon, double dt, int rollouts)
  {
    b2n_mppi_params p{};
    p.wheel_radius = cart_model.wheel_radius;
    p.wheel_base = cart_model.wheel_base;
    for (int i = 0; i < 3; i++) { p.Q[i] = loss_func.Q[i]; p.P1[i] = loss_func.P1[i]; }
    p.R[0] = loss_func.R[0]; p.R[1] = loss_func.R[1];
    p.lambda = lambda; p.max_wheel_vel = max_wheel_vel; p.ul_var = ul_var; p.ur_var = ur_var;
    p.horizon = horizon; p.dt = dt;
    p.rollouts = rollouts; p.rollout_offset = 0; p.rollouts_total = rollouts; p.device = -1;
    detail::check(b2n_mppi_create(&p, &h_));
    // like the reference, whose engine is seeded from std::random_device (rigid2d/src/rigid2d/utilities.cpp:12-17): every
    // controller draws its own perturbation streams.  b2n_mppi_seed(handle(), ...) makes a run reproducible.
    std::random_device rd;
    detail::check(b2n_mppi_seed(h_, ((uint64_t)rd() << 32) | (uint64_t)rd(), 0));
  }
  ~MPPI() { b2n_mppi_destroy(h_); }
  MPPI(const MPPI &) = delete;
  MPPI &operator=(const MPPI &) = delete;
  MPPI(MPPI &&o) noexcept : h_(o.h_) { o.h_ = nullptr; }

  /// reference mppi.cpp:54-61
  void setInitialControls(double uL, double uR) { detail::check(b2n_mppi_set_initial_controls(h_, uL, uR)); }
  /// reference mppi.cpp:64-69
  void setWaypoint(const Pose &wpt) { detail::check(b2n_mppi_set_waypoint(h_, wpt.x, wpt.y, wpt.theta)); }
  /// reference mppi.cpp:72-140
  WheelVelocities newControls(const Pose &ps)
  {
    WheelVelocities v;
    detail::check(b2n_mppi_new_controls(h_, ps.x, ps.y, ps.theta, &v.ul, &v.ur));
    return v;
  }

  /// not in the reference: the C handle, for the noise seam / taps of include/b2nav.h
  b2n_mppi *handle() { return h_; }

private:
  b2n_mppi *h_ = nullptr;
};
} // namespace controller
#endif
