// TEST INFRASTRUCTURE ONLY: PCL is absent here.  The reference header
// bmapping/include/bmapping/cloud_alignment.hpp only needs these names to be declared; the ICP
// itself is NOT restated - oracle/ref_capi.cpp supplies ScanAlignment with an injected result.
#ifndef B2N_ORACLE_PCL_STUB_H
#define B2N_ORACLE_PCL_STUB_H
#include <memory>
#include <vector>
namespace pcl
{
struct PointXYZ { float x, y, z; };
template <class P> struct PointCloud { typedef std::shared_ptr<PointCloud<P>> Ptr; std::vector<P> points; };
}
#endif
