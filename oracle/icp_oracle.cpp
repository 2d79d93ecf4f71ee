// TEST INFRASTRUCTURE ONLY (oracle/).  Not part of the product path.
//
// CPU statement of libb2nav's scan matcher (csrc/icp_kernels.cuh).  PARITY UNPINNED AGAINST THE REFERENCE: the reference
// delegates scan matching to PCL (pcl::IterativeClosestPoint<PointXYZ, PointXYZ>, bmapping/src/bmapping/
// cloud_alignment.cpp:160-223; PCL is absent from /root/reference, not listed in package.xml, ROS Melodic ships 1.8).
// What is restated here is the PUBLISHED structure of that algorithm with the reference's settings
// (cloud_alignment.cpp:20-25,186-190), not PCL's code:
//   * clouds as ScanAlignment::createPointCloud builds them (cloud_alignment.cpp:76-157): float beam angle accumulated
//     with the wrap rule, range gate range_min <= r < range_max, points in float, Trs = identity;
//   * point-to-point ICP: transform the source by the current estimate, nearest target point for every source point
//     (exhaustive search = what a kd-tree returns), drop pairs farther than max_correspondence_dist (0.5 m), fewer than
//     3 pairs = failure, closed-form 2-D rigid alignment (the SVD/Umeyama solution restricted to the plane), compose;
//   * stop like pcl::registration::DefaultConvergenceCriteria with the reference's thresholds: max_iter (100) reached
//     = converged; increment nearly the identity (cos >= 0.99999 and |t|^2 <= transformation_epsilon 1e-8); absolute
//     change of the mean squared pair distance < euclidean_fitness_epsilon (1e-6); relative change < 1e-5.
//   * the wrapper keeps the previous scan and answers the first call with success and an untouched transform
//     (cloud_alignment.cpp:37-72).
// The GPU kernel is checked against THIS file; neither is claimed to reproduce PCL bit for bit.
//
// Known deviations from pcl::IterativeClosestPoint 1.8 as the reference configures it (cloud_alignment.cpp:166-195), listed
// so that whoever pins this against a PCL build knows where to look (the row stays "partial / parity unpinned" until then):
//   1. arithmetic: PCL aligns in float (Eigen::Matrix4f guess from float cos / sin, :172-181; float SVD of the 3x3
//      covariance in TransformationEstimationSVD via Eigen::umeyama); here the pair sums and the closed-form planar solution
//      are double, the points float.  Expect agreement to ~1e-6 in the transform, not bit parity.
//   2. transformation estimation: PCL solves the 3-D problem by SVD; here its restriction to rotations about z (all
//      points have z = 0, so the SVD's answer is that rotation) in closed form: theta = atan2(Sxy - Syx, Sxx + Syy).
//   3. correspondence estimation: PCL queries a FLANN kd-tree (exact nearest neighbour, ties by tree order); here an
//      exhaustive search, ties by lowest target index.  The correspondence distance test is squared distance <= 0.25 m^2
//      in both (CorrespondenceEstimation::determineCorrespondences compares squared distances).
//   4. RANSAC outlier rejection threshold 0.05 (:190): the setter stores a member that IterativeClosestPoint::
//      computeTransformation never reads unless a CorrespondenceRejectorSampleConsensus is added to the rejector list,
//      which the reference does not do; it has no effect there and none here.
//   5. convergence (DefaultConvergenceCriteria): PCL counts consecutive "similar" iterations (max_iterations_similar_
//      transforms_ = 0 by default, so the first similar iteration stops) and tests, in this order, iteration count,
//      the incremental transform (cos of its rotation angle >= rotation_threshold_ = 0.99999 AND squared translation <=
//      translation_threshold_, which IterativeClosestPoint sets to the transformation epsilon itself, 1e-8), then the
//      absolute and the relative MSE change (mse_threshold_absolute_ = the euclidean
//      fitness epsilon, mse_threshold_relative_ = 1e-5).  Restated in that order; hasConverged() is true for all three
//      reasons including "iterations exhausted", which is why the reference's failure branch (:200-204) is in practice
//      only reached through fewer than 3 correspondences.
//   6. the final transform is the product of the increments applied to the guess, accumulated in double here (float in PCL).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace
{
struct IcpParams
{
  float beam_min, beam_max, beam_delta, range_min, range_max;
  int32_t max_iter;
  double max_corr_dist, transform_eps, fitness_eps;
};

struct Icp
{
  IcpParams p;
  std::vector<float> old_scan;
  bool first_scan_received = false;
  int last_iterations = 0, last_pairs = 0;
  double last_mse = 0.0;
};

void cloud(const IcpParams &p, const float *scan, int n, std::vector<float> &xy)
{
  xy.clear();
  float beam_angle = p.beam_min;
  for (int i = 0; i < n; i++) {
    const float range = scan[i];
    if (range >= p.range_min && range < p.range_max) {
      // cloud_alignment.cpp:119-120: `range` is a double there, std::cos / std::sin take the FLOAT angle (cosf / sinf)
      xy.push_back((float)((double)range * (double)std::cos(beam_angle)));
      xy.push_back((float)((double)range * (double)std::sin(beam_angle)));
    }
    beam_angle += p.beam_delta;
    if (p.beam_max < 0.0f && beam_angle <= p.beam_max) beam_angle = p.beam_min;
    else if (p.beam_max >= 0.0f && beam_angle >= p.beam_max) beam_angle = p.beam_min;
  }
}

// returns converged; T (theta, x, y) maps the source cloud onto the target cloud
bool icp(Icp &h, const std::vector<float> &tgt, const std::vector<float> &src0, const double Tinit[3], double Tout[3])
{
  const IcpParams &p = h.p;
  const int ns = (int)src0.size() / 2, nt = (int)tgt.size() / 2;
  h.last_iterations = 0; h.last_pairs = 0; h.last_mse = 0.0;
  if (ns < 3 || nt < 3) return false;
  double c = std::cos(Tinit[0]), s = std::sin(Tinit[0]), tx = Tinit[1], ty = Tinit[2];   // accumulated estimate
  std::vector<double> sx(ns), sy(ns);
  const double max_d2 = p.max_corr_dist * p.max_corr_dist;
  double prev_mse = 1.0e300;
  bool converged = false;
  for (int it = 1; it <= p.max_iter; it++) {
    for (int i = 0; i < ns; i++) {
      const double x = src0[2 * i], y = src0[2 * i + 1];
      sx[i] = c * x - s * y + tx;
      sy[i] = s * x + c * y + ty;
    }
    // correspondences and the sums of the closed-form alignment
    int m = 0;
    double Ssx = 0, Ssy = 0, Stx = 0, Sty = 0, Sdot = 0, Scross = 0, Sd2 = 0;
    for (int i = 0; i < ns; i++) {
      int best = -1;
      double bd = 1.0e300;
      for (int j = 0; j < nt; j++) {
        const double dx = sx[i] - (double)tgt[2 * j], dy = sy[i] - (double)tgt[2 * j + 1];
        const double d2 = dx * dx + dy * dy;
        if (d2 < bd) { bd = d2; best = j; }                 // first minimum wins
      }
      if (bd > max_d2) continue;
      const double qx = tgt[2 * best], qy = tgt[2 * best + 1];
      m++;
      Ssx += sx[i]; Ssy += sy[i]; Stx += qx; Sty += qy;
      Sdot += sx[i] * qx + sy[i] * qy;
      Scross += sx[i] * qy - sy[i] * qx;
      Sd2 += bd;
    }
    h.last_iterations = it; h.last_pairs = m;
    if (m < 3) return false;                                // too few correspondences
    const double inv = 1.0 / m;
    const double mx = Ssx * inv, my = Ssy * inv, qx = Stx * inv, qy = Sty * inv;
    // sum (s - ms).(t - mt) and sum (s - ms) x (t - mt)
    const double a = Sdot - m * (mx * qx + my * qy);
    const double b = Scross - m * (mx * qy - my * qx);
    const double dth = std::atan2(b, a);
    const double dc = std::cos(dth), ds = std::sin(dth);
    const double dtx = qx - (dc * mx - ds * my), dty = qy - (ds * mx + dc * my);
    // compose: T <- dT * T
    const double nc = dc * c - ds * s, nsn = ds * c + dc * s;
    const double ntx = dc * tx - ds * ty + dtx, nty = ds * tx + dc * ty + dty;
    c = nc; s = nsn; tx = ntx; ty = nty;
    const double mse = Sd2 * inv;
    h.last_mse = mse;
    // DefaultConvergenceCriteria, reference thresholds
    if (it >= p.max_iter) { converged = true; break; }
    if (dc >= 0.99999 && dtx * dtx + dty * dty <= p.transform_eps) { converged = true; break; }
    if (std::fabs(mse - prev_mse) < p.fitness_eps) { converged = true; break; }
    if (std::fabs(mse - prev_mse) / prev_mse < 1.0e-5) { converged = true; break; }
    prev_mse = mse;
  }
  Tout[0] = std::atan2(s, c); Tout[1] = tx; Tout[2] = ty;
  return converged;
}
} // namespace

extern "C" {

void *orc_icp_create(const IcpParams *p)
{
  Icp *h = new Icp();
  h->p = *p;
  return h;
}
void orc_icp_destroy(void *h) { delete static_cast<Icp *>(h); }

// pclICPWrapper semantics (cloud_alignment.cpp:37-72): returns 1/0; T is written only on a successful alignment
int orc_icp_align(void *hv, const float *scan, int n, const double Tinit[3], double T[3])
{
  Icp *h = static_cast<Icp *>(hv);
  if (!h->first_scan_received) {
    h->old_scan.assign(scan, scan + n);
    h->first_scan_received = true;
    return 1;
  }
  std::vector<float> tgt, src;
  cloud(h->p, h->old_scan.data(), (int)h->old_scan.size(), tgt);
  cloud(h->p, scan, n, src);
  double out[3];
  if (!icp(*h, tgt, src, Tinit, out)) return 0;
  T[0] = out[0]; T[1] = out[1]; T[2] = out[2];
  h->old_scan.assign(scan, scan + n);
  return 1;
}

void orc_icp_stats(void *hv, int *iterations, int *pairs, double *mse)
{
  Icp *h = static_cast<Icp *>(hv);
  *iterations = h->last_iterations; *pairs = h->last_pairs; *mse = h->last_mse;
}

} // extern "C"
