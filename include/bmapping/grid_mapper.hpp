// bmapping/grid_mapper.hpp - drop-in for bmapping/include/bmapping/grid_mapper.hpp:117-122 (the constructor only).
// In the reference every particle owns a GridMapper (particle_filter.hpp:61-85) and ParticleFilter copies the
// prototype it is given (particle_filter.cpp:125-138).  Here the per-particle maps are SoA planes in HBM owned by the
// filter handle; this class carries the prototype's parameters to bmapping::ParticleFilter.  The map operations
// (likelihoodFieldModel, integrateScan, euclideanSignedDistanceField, gridMap: grid_mapper.cpp:69-435) run in the
// sm_100a kernels of libb2nav and are reachable per particle through the taps of include/b2nav.h.
#ifndef B2N_BMAPPING_GRID_MAPPER_HPP
#define B2N_BMAPPING_GRID_MAPPER_HPP

#include <stdexcept>

#if __has_include(<rigid2d/rigid2d.hpp>)
#include <rigid2d/rigid2d.hpp>
#else
#include "../rigid2d_min/types.hpp"
#endif

#include "sensor_model.hpp"

namespace bmapping
{
using rigid2d::Transform2D;

class GridMapper
{
public:
  /// reference grid_mapper.hpp:121-122, grid_mapper.cpp:37-64.  Trs (robot -> sensor) must be the identity, as in
  /// turtle_mapping_node.cpp:387; the kernels build the map frame end points with it folded in.
  GridMapper(double resolution, double xmin, double xmax, double ymin, double ymax, const LaserProperties &props, const Transform2D &Trs)
    : resolution(resolution), xmin(xmin), xmax(xmax), ymin(ymin), ymax(ymax), props(props)
  {
    const auto d = Trs.displacement();
    if (d.theta != 0.0 || d.x != 0.0 || d.y != 0.0)
      throw std::invalid_argument("bmapping::GridMapper (libb2nav): only Trs = identity is supported");
  }

  double resolution, xmin, xmax, ymin, ymax;
  LaserProperties props;
};
} // namespace bmapping
#endif
