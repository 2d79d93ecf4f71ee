// rigid2d_min/types.hpp - the three rigid2d value types the two hot-path class surfaces exchange
// with their callers, for builds WITHOUT the reference's rigid2d package on the include path.
//
// In a catkin workspace the real <rigid2d/diff_drive.hpp> / <rigid2d/rigid2d.hpp> are found first
// (include/controller/mppi.hpp and include/bmapping/particle_filter.hpp test for them with
// __has_include) and this file is not used.  Field order follows the reference so aggregate
// initialisation means the same thing: Pose {theta, x, y} (rigid2d/include/rigid2d/diff_drive.hpp:16-21),
// WheelVelocities {ul, ur} (:24-28), Twist2D {w, vx, vy} (rigid2d/include/rigid2d/rigid2d.hpp:157-162).
#ifndef B2N_RIGID2D_MIN_TYPES_HPP
#define B2N_RIGID2D_MIN_TYPES_HPP

#include <cmath>

namespace rigid2d
{
struct Pose
{
  double theta = 0.0, x = 0.0, y = 0.0;
};

struct WheelVelocities
{
  double ul = 0.0, ur = 0.0;
};

struct Twist2D
{
  double w = 0.0, vx = 0.0, vy = 0.0;
};

struct Vector2D
{
  double x = 0.0, y = 0.0;
  Vector2D() = default;
  Vector2D(double x_, double y_) : x(x_), y(y_) {}
};

struct TransformData2D
{
  double theta = 0.0, x = 0.0, y = 0.0;
};

// Only what the particle-filter surface needs: construct from (translation, angle), read it back.
class Transform2D
{
public:
  Transform2D() = default;
  Transform2D(const Vector2D &trans, double radians) : theta_(radians), x_(trans.x), y_(trans.y) {}
  TransformData2D displacement() const
  {
    TransformData2D d;
    d.theta = theta_; d.x = x_; d.y = y_;
    return d;
  }

private:
  double theta_ = 0.0, x_ = 0.0, y_ = 0.0;
};
} // namespace rigid2d
#endif
