// The call sequence of nuturtle_robot/src/mppi_waypoints_node.cpp:186-199,216,265 written against
// include/controller/mppi.hpp (our drop-in header) - compiled and linked with libb2nav.so by
// tests/test_cpp_surface.py.  Prints "ul ur" of a few receding-horizon calls; without a GPU the
// constructor throws and the program prints NO_DEVICE and exits 3.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <controller/mppi.hpp>

int main(int argc, char **argv)
{
  const int rollouts = argc > 1 ? std::atoi(argv[1]) : 128;
  const int calls = argc > 2 ? std::atoi(argv[2]) : 3;
  const unsigned long long seed = argc > 3 ? std::strtoull(argv[3], nullptr, 10) : 42ULL;
  // controller/config/mppi_params.yaml + nuturtle_description/config/diff_params.yaml
  const std::vector<double> Q{1e4, 1e4, 1.0}, R{0.1, 0.1}, P1{1e3, 1e3, 1e3};
  try {
    controller::CartModel cart_model(0.033, 0.16);
    controller::LossFunc loss_func(Q, R, P1);
    controller::MPPI mppi(cart_model, loss_func, 0.01, 6.35495, 0.9, 0.9, 0.5, 0.02, rollouts);
    mppi.setInitialControls(0.0, 0.0);
    b2n_mppi_seed(mppi.handle(), seed, 0);
    rigid2d::Pose wpt;
    wpt.x = 1.0; wpt.y = 0.0; wpt.theta = 1.5707;
    mppi.setWaypoint(wpt);
    rigid2d::Pose pose;
    for (int c = 0; c < calls; c++) {
      rigid2d::WheelVelocities v = mppi.newControls(pose);
      std::printf("%.17g %.17g\n", v.ul, v.ur);
    }
  } catch (const std::runtime_error &e) {
    std::printf("NO_DEVICE %s\n", e.what());
    return 3;
  }
  try {
    controller::LossFunc bad({1.0}, {1.0}, {1.0});   // reference: std::out_of_range from .at()
    return 4;
  } catch (const std::out_of_range &) {
  }
  return 0;
}
