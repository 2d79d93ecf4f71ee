// The call sequence of bmapping/src/turtle_mapping_node.cpp:387-410,474,479,494 written against
// include/bmapping/particle_filter.hpp (our drop-in header) - compiled and linked with libb2nav.so by
// tests/test_cpp_surface.py.  Reads "n_scans n_beams" then per scan: twist(3) cur_odom(3) prev_odom(3) and
// n_beams ranges from stdin; prints per scan the robot state (theta x y) and a checksum of the exported map.
// Without a GPU the constructor throws and the program prints NO_DEVICE and exits 3.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

#include <bmapping/particle_filter.hpp>

int main(int argc, char **argv)
{
  using namespace bmapping;
  using rigid2d::Pose;
  using rigid2d::Transform2D;
  using rigid2d::Twist2D;
  using rigid2d::Vector2D;
  const int num_particles = argc > 1 ? std::atoi(argv[1]) : 8;
  const unsigned long long seed = argc > 2 ? std::strtoull(argv[2], nullptr, 10) : 1ULL;
  const double x0 = argc > 3 ? std::atof(argv[3]) : 0.5, y0 = argc > 4 ? std::atof(argv[4]) : 0.0, th0 = argc > 5 ? std::atof(argv[5]) : 1.5707963267948966;
  try {
    Transform2D Trs;
    // bmapping/launch/slam.launch + nuturtle_robot/config/LDS_01_lidar.yaml
    LaserProperties props(0.0f, 6.28319f, 0.0174533f, 0.12f, 3.5f, 0.95, 0.0, 0.04, 0.01, 0.5);
    GridMapper grid(0.05, -5.0, 5.0, -5.0, 5.0, props, Trs);
    const bool gpu_matcher = argc > 6 && std::atoi(argv[6]) != 0;
    ScanAlignment stub(props, Trs);
    std::unique_ptr<GpuScanAlignment> gpu_aligner;
    if (gpu_matcher) gpu_aligner.reset(new GpuScanAlignment(props, Trs));
    ScanAlignment &aligner = gpu_matcher ? static_cast<ScanAlignment &>(*gpu_aligner) : stub;
    Transform2D robot_pose(Vector2D(x0, y0), th0);
    ParticleFilter pf(num_particles, 50, 0.001, 0.001, 0.001, 0.001, 2e-3, 1e-3, 1e-3, 1e-3, 1e-3, 1e-3, 1.0, 20.0, 1.0, 10.0, aligner,
                      robot_pose, grid);
    b2n_pf_seed(pf.handle(), seed, 0);
    int n_scans = 0, n_beams = 0;
    if (std::scanf("%d %d", &n_scans, &n_beams) != 2) return 0;
    std::vector<int8_t> map;
    for (int s = 0; s < n_scans; s++) {
      Twist2D vb;
      Pose cur, prev;
      if (std::scanf("%lf %lf %lf %lf %lf %lf %lf %lf %lf", &vb.w, &vb.vx, &vb.vy, &cur.theta, &cur.x, &cur.y, &prev.theta, &prev.x, &prev.y) != 9) return 5;
      std::vector<float> scan(n_beams);
      for (auto &r : scan)
        if (std::scanf("%f", &r) != 1) return 5;
      pf.SLAM(scan, vb, cur, prev);
      pf.newMap(map);
      const auto T = pf.getRobotState().displacement();
      long long sum = 0;
      for (size_t i = 0; i < map.size(); i++) sum += (long long)(i % 977 + 1) * map[i];
      std::printf("%.17g %.17g %.17g %lld %zu\n", T.theta, T.x, T.y, sum, map.size());
    }
  } catch (const std::runtime_error &e) {
    std::printf("NO_DEVICE %s\n", e.what());
    return 3;
  }
  return 0;
}
