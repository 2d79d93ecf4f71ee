// bmapping/sensor_model.hpp - drop-in for bmapping/include/bmapping/sensor_model.hpp:20-79 (LaserProperties).
// Only the parameter struct: LaserScanner's geometry (sensor_model.cpp:43-112) runs inside rbpf_update_kernel.
// Same field names, order, defaults and constructor as the reference, so code that fills or reads a
// LaserProperties (turtle_mapping_node.cpp:391-392, cloud_alignment.cpp) compiles unchanged.
#ifndef B2N_BMAPPING_SENSOR_MODEL_HPP
#define B2N_BMAPPING_SENSOR_MODEL_HPP

namespace bmapping
{
struct LaserProperties
{
  // lidar beam
  float beam_min;
  float beam_max;
  float beam_delta;
  float range_min;
  float range_max;
  // mixture weights of the beam model
  double z_hit;
  double z_short;
  double z_max;
  double z_rand;
  // std-dev of the Gaussian for laser hits
  double sigma_hit;

  LaserProperties()
    : beam_min(0.0), beam_max(0.0), beam_delta(0.0), range_min(0.0), range_max(0.0), z_hit(0.25), z_short(0.25), z_max(0.25),
      z_rand(0.25), sigma_hit(1)
  {
  }

  LaserProperties(float beam_min, float beam_max, float beam_delta, float range_min, float range_max, double z_hit, double z_short,
                  double z_max, double z_rand, double sigma_hit)
    : beam_min(beam_min), beam_max(beam_max), beam_delta(beam_delta), range_min(range_min), range_max(range_max), z_hit(z_hit),
      z_short(z_short), z_max(z_max), z_rand(z_rand), sigma_hit(sigma_hit)
  {
  }
};
} // namespace bmapping
#endif
