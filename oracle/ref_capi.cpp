// TEST INFRASTRUCTURE ONLY (oracle/). Not part of the product path; nothing under
// ros-turtlebot-navigation_b200/ or include/ may link or call this.
//
// Thin C API over the UNMODIFIED reference translation units, compiled in place from
// /root/reference by oracle/Makefile into oracle/_ref/libref_nav.so:
//   controller/src/controller/{mppi,rk4}.cpp        rigid2d/src/rigid2d/{rigid2d,diff_drive,utilities}.cpp
//   bmapping/src/bmapping/{particle_filter,grid_mapper,sensor_model}.cpp
// Eigen is replaced by oracle/shim/mini_eigen.hpp and PCL by a declaration-only stub; the ICP
// (bmapping/src/bmapping/cloud_alignment.cpp, needs PCL) is NOT compiled - ScanAlignment is
// defined below with an injected result, which is exactly the seam our C ABI exposes
// (Ticp / icp_ok are inputs at the boundary).
//
// This file contains no reference source text: it only calls the reference classes and copies
// their state out through `#define private public` (layout-neutral).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <iostream>
#include <queue>
#include <random>
#include <sstream>
#include <stdexcept>
#include <unordered_set>
#include <vector>
#include <memory>
#include <algorithm>

#define private public
#include <rigid2d/rigid2d.hpp>
#include <rigid2d/diff_drive.hpp>
#include <rigid2d/utilities.hpp>
#include <controller/mppi.hpp>
#include <bmapping/particle_filter.hpp>
#undef private

// ---------------------------------------------------------------------------------------------
// ScanAlignment with an injected outcome (replaces cloud_alignment.cpp:37-72, which needs PCL)
// ---------------------------------------------------------------------------------------------
namespace
{
bool g_icp_ok = false;
double g_icp_pose[3] = {0.0, 0.0, 0.0};   // theta, x, y
}

namespace bmapping
{
ScanAlignment::ScanAlignment(const LaserProperties &props, const Transform2D &Trs)
    : max_iter_(0), max_correspondence_dist_(0), transform_epsilon_(0), fitness_epsilon_(0), Trs_(Trs),
      beam_min_(props.beam_min), beam_max_(props.beam_max), beam_delta_(props.beam_delta),
      range_min_(props.range_min), range_max_(props.range_max), first_scan_recieved(false)
{
}

bool ScanAlignment::pclICPWrapper(Transform2D &T, const Transform2D &, const std::vector<float> &)
{
  if (g_icp_ok) {
    T = Transform2D(rigid2d::Vector2D(g_icp_pose[1], g_icp_pose[2]), g_icp_pose[0]);
  }
  return g_icp_ok;
}
}

namespace
{
struct SilenceCout
{
  std::streambuf *old;
  std::ostringstream sink;
  SilenceCout() : old(std::cout.rdbuf(sink.rdbuf())) {}
  ~SilenceCout() { std::cout.rdbuf(old); }
};

struct RefMppi
{
  controller::MPPI mppi;
  RefMppi(const controller::CartModel &c, const controller::LossFunc &l, double lambda, double umax,
          double ulv, double urv, double horizon, double dt, int K)
      : mppi(c, l, lambda, umax, ulv, urv, horizon, dt, K) {}
};

struct RefPf
{
  bmapping::LaserProperties props;
  rigid2d::Transform2D Trs;
  bmapping::GridMapper proto;
  bmapping::ScanAlignment matcher;
  std::unique_ptr<bmapping::ParticleFilter> pf;
  RefPf(const bmapping::LaserProperties &p, double res, double xmin, double xmax, double ymin, double ymax)
      : props(p), Trs(), proto(res, xmin, xmax, ymin, ymax, p, Trs), matcher(p, Trs) {}
};

struct RefGrid
{
  bmapping::LaserProperties props;
  rigid2d::Transform2D Trs;
  bmapping::GridMapper grid;
  RefGrid(const bmapping::LaserProperties &p, double res, double xmin, double xmax, double ymin, double ymax)
      : props(p), Trs(), grid(res, xmin, xmax, ymin, ymax, p, Trs) {}
};

bmapping::LaserProperties makeProps(const float lf[5], const double ld[5])
{
  return bmapping::LaserProperties(lf[0], lf[1], lf[2], lf[3], lf[4], ld[0], ld[1], ld[2], ld[3], ld[4]);
}

void dumpGrid(const bmapping::GridMapper &g, double *log_odds, double *prob, double *occ_dist, int *state)
{
  const size_t n = g.map_.size();
  for (size_t i = 0; i < n; i++) {
    if (log_odds) log_odds[i] = g.map_[i].log_odds;
    if (prob) prob[i] = g.map_[i].prob;
    if (occ_dist) occ_dist[i] = g.map_[i].occ_dist;
    if (state) state[i] = g.map_[i].state;
  }
}

int dumpOccOrder(const bmapping::GridMapper &g, int *keys, int cap)
{
  int n = 0;
  for (auto k : g.occ_cells_) { if (n < cap) keys[n] = k; n++; }
  return n;
}
}

extern "C" {

// ----------------------------------------------------------------------------- RNG ----------
void ref_rigid2d_seed(uint64_t seed) { rigid2d::getTwister().seed(seed); }
void ref_bmapping_seed(uint64_t seed) { bmapping::getTwister().seed(seed); }
// n draws of rigid2d::sampleNormalDistribution(mu, sigma) from the global engine
void ref_rigid2d_normals(int n, double mu, double sigma, double *out)
{
  for (int i = 0; i < n; i++) out[i] = rigid2d::sampleNormalDistribution(mu, sigma);
}
void ref_bmapping_std_normals(int n, double *out)
{
  Eigen::VectorXd v = bmapping::sampleStandardNormal(n);
  for (int i = 0; i < n; i++) out[i] = v(i);
}

// --------------------------------------------------------------------------- rigid2d --------
double ref_normalize_angle_pi(double a) { return rigid2d::normalize_angle_PI(a); }
void ref_wheels_to_twist(double base, double radius, double ul, double ur, double out[3])
{
  rigid2d::DiffDrive d(rigid2d::Pose(), base, radius);
  rigid2d::WheelVelocities v; v.ul = ul; v.ur = ur;
  rigid2d::Twist2D t = d.wheelsToTwist(v);
  out[0] = t.w; out[1] = t.vx; out[2] = t.vy;
}
void ref_twist_to_wheels(double base, double radius, double w, double vx, double out[2])
{
  rigid2d::DiffDrive d(rigid2d::Pose(), base, radius);
  rigid2d::Twist2D t; t.w = w; t.vx = vx; t.vy = 0.0;
  rigid2d::WheelVelocities v = d.twistToWheels(t);
  out[0] = v.ul; out[1] = v.ur;
}
// T(theta,x,y).integrateTwist(w,vx,vy) -> (theta,x,y)
void ref_integrate_twist(const double T[3], const double tw[3], double out[3])
{
  rigid2d::Transform2D t(rigid2d::Vector2D(T[1], T[2]), T[0]);
  rigid2d::Twist2D v; v.w = tw[0]; v.vx = tw[1]; v.vy = tw[2];
  rigid2d::TransformData2D d = t.integrateTwist(v).displacement();
  out[0] = d.theta; out[1] = d.x; out[2] = d.y;
}
// DiffDrive: feedforward one twist from a pose, return new pose (theta,x,y)
void ref_feedforward(double base, double radius, const double pose[3], const double tw[3], double out[3])
{
  rigid2d::Pose p; p.theta = pose[0]; p.x = pose[1]; p.y = pose[2];
  rigid2d::DiffDrive d(p, base, radius);
  rigid2d::Twist2D v; v.w = tw[0]; v.vx = tw[1]; v.vy = tw[2];
  d.feedforward(v);
  rigid2d::Pose q = d.pose();
  out[0] = q.theta; out[1] = q.x; out[2] = q.y;
}

// a DiffDrive that lives across calls: the simulated robot / the odometer of the closed-loop tests
// (rigid2d/src/fake_diff_encoders_node.cpp:100-135, bmapping/src/turtle_mapping_node.cpp:456-472)
void *ref_dd_create(const double pose[3], double base, double radius)
{
  rigid2d::Pose p; p.theta = pose[0]; p.x = pose[1]; p.y = pose[2];
  return new rigid2d::DiffDrive(p, base, radius);
}
void ref_dd_destroy(void *h) { delete static_cast<rigid2d::DiffDrive *>(h); }
void ref_dd_feedforward(void *h, double w, double vx)
{
  rigid2d::Twist2D v; v.w = w; v.vx = vx; v.vy = 0.0;
  static_cast<rigid2d::DiffDrive *>(h)->feedforward(v);
}
void ref_dd_update_odometry(void *h, double left, double right, double vel[2])
{
  rigid2d::WheelVelocities v = static_cast<rigid2d::DiffDrive *>(h)->updateOdometry(left, right);
  vel[0] = v.ul; vel[1] = v.ur;
}
// out: pose (theta, x, y), encoders (left, right), wheel velocities (ul, ur)
void ref_dd_state(void *h, double out[7])
{
  rigid2d::DiffDrive *d = static_cast<rigid2d::DiffDrive *>(h);
  rigid2d::Pose p = d->pose();
  rigid2d::WheelEncoders e = d->getEncoders();
  rigid2d::WheelVelocities v = d->wheelVelocities();
  out[0] = p.theta; out[1] = p.x; out[2] = p.y; out[3] = e.left; out[4] = e.right; out[5] = v.ul; out[6] = v.ur;
}

// ----------------------------------------------------------------------------- MPPI ---------
void *ref_mppi_create(double wheel_radius, double wheel_base, const double Q[3], const double R[2],
                      const double P1[3], double lambda, double max_wheel_vel, double ul_var, double ur_var,
                      double horizon, double dt, int rollouts)
{
  controller::CartModel cart(wheel_radius, wheel_base);
  controller::LossFunc loss(std::vector<double>(Q, Q + 3), std::vector<double>(R, R + 2),
                            std::vector<double>(P1, P1 + 3));
  return new RefMppi(cart, loss, lambda, max_wheel_vel, ul_var, ur_var, horizon, dt, rollouts);
}
void ref_mppi_destroy(void *h) { delete static_cast<RefMppi *>(h); }
int ref_mppi_steps(void *h) { return static_cast<RefMppi *>(h)->mppi.steps; }
void ref_mppi_set_initial_controls(void *h, double ul, double ur) { static_cast<RefMppi *>(h)->mppi.setInitialControls(ul, ur); }
void ref_mppi_set_waypoint(void *h, double x, double y, double theta)
{
  rigid2d::Pose p; p.x = x; p.y = y; p.theta = theta;
  static_cast<RefMppi *>(h)->mppi.setWaypoint(p);
}
void ref_mppi_new_controls(void *h, double x, double y, double theta, double *ul, double *ur)
{
  rigid2d::Pose p; p.x = x; p.y = y; p.theta = theta;
  rigid2d::WheelVelocities v = static_cast<RefMppi *>(h)->mppi.newControls(p);
  *ul = v.ul; *ur = v.ur;
}
// u_plan: [2][T]; J (min-subtracted, as the reference leaves it), duL, duR: [T][K] row-major
void ref_mppi_get(void *h, double *u_plan, double *J, double *duL, double *duR)
{
  controller::MPPI &m = static_cast<RefMppi *>(h)->mppi;
  const int T = m.steps, K = m.rollouts;
  if (u_plan) for (int r = 0; r < 2; r++) for (int t = 0; t < T; t++) u_plan[r * T + t] = m.u(r, t);
  for (int t = 0; t < T; t++)
    for (int k = 0; k < K; k++) {
      if (J) J[(size_t)t * K + k] = m.J(t, k);
      if (duL) duL[(size_t)t * K + k] = m.duL(t, k);
      if (duR) duR[(size_t)t * K + k] = m.duR(t, k);
    }
}
// one rollout through the reference integrator and loss: traj [T][3] (x,y,theta), loss [T]
void ref_mppi_rollout(void *h, const double x0[3], const double *u_pert /*[2][T]*/, double *traj, double *loss)
{
  controller::MPPI &m = static_cast<RefMppi *>(h)->mppi;
  const int T = m.steps;
  Eigen::VectorXd x(3); x << x0[0], x0[1], x0[2];
  Eigen::MatrixXd u(2, T);
  for (int t = 0; t < T; t++) { u(0, t) = u_pert[t]; u(1, t) = u_pert[T + t]; }
  Eigen::MatrixXd tr = m.rk4.solve(x, u, m.horizon);
  for (int t = 0; t < T; t++) {
    for (int c = 0; c < 3; c++) traj[t * 3 + c] = tr(c, t);
    loss[t] = m.loss_func.loss(tr.col(t), m.xd, u.col(t));
  }
  loss[T - 1] = m.loss_func.terminalLoss(tr.col(T - 1), m.xd);
}

// --------------------------------------------------------------------- GridMapper (alone) ---
void *ref_grid_create(const float lf[5], const double ld[5], double res, double xmin, double xmax, double ymin, double ymax)
{
  return new RefGrid(makeProps(lf, ld), res, xmin, xmax, ymin, ymax);
}
void ref_grid_destroy(void *h) { delete static_cast<RefGrid *>(h); }
void *ref_grid_clone(void *h) { return new RefGrid(*static_cast<RefGrid *>(h)); }
int ref_grid_size(void *h, int *xs, int *ys)
{
  bmapping::GridMapper &g = static_cast<RefGrid *>(h)->grid;
  if (xs) *xs = g.xsize_; if (ys) *ys = g.ysize_;
  return (int)g.map_.size();
}
// returns 0 ok, 1 when the reference throws (off-map end point etc.)
int ref_grid_likelihood(void *h, const float *scan, int n, const double pose[3], double *p)
{
  try {
    rigid2d::Transform2D T(rigid2d::Vector2D(pose[1], pose[2]), pose[0]);
    *p = static_cast<RefGrid *>(h)->grid.likelihoodFieldModel(std::vector<float>(scan, scan + n), T);
    return 0;
  } catch (std::exception &) { return 1; }
}
int ref_grid_integrate(void *h, const float *scan, int n, const double pose[3])
{
  try {
    rigid2d::Transform2D T(rigid2d::Vector2D(pose[1], pose[2]), pose[0]);
    static_cast<RefGrid *>(h)->grid.integrateScan(std::vector<float>(scan, scan + n), T);
    return 0;
  } catch (std::exception &) { return 1; }
}
void ref_grid_dump(void *h, double *log_odds, double *prob, double *occ_dist, int *state)
{
  dumpGrid(static_cast<RefGrid *>(h)->grid, log_odds, prob, occ_dist, state);
}
int ref_grid_occ_order(void *h, int *keys, int cap) { return dumpOccOrder(static_cast<RefGrid *>(h)->grid, keys, cap); }
int ref_grid_bucket_count(void *h) { return (int)static_cast<RefGrid *>(h)->grid.occ_cells_.bucket_count(); }
void ref_grid_map(void *h, int8_t *out)
{
  std::vector<int8_t> m;
  static_cast<RefGrid *>(h)->grid.gridMap(m);
  std::memcpy(out, m.data(), m.size());
}
// valid end points (map frame) of a scan at a pose: returns count, fills xy[2*i]
int ref_grid_end_points(void *h, const float *scan, int n, const double pose[3], double *xy)
{
  std::vector<rigid2d::Vector2D> pts;
  rigid2d::Transform2D T(rigid2d::Vector2D(pose[1], pose[2]), pose[0]);
  static_cast<RefGrid *>(h)->grid.laserEndPoints(pts, std::vector<float>(scan, scan + n), T);
  for (size_t i = 0; i < pts.size(); i++) { xy[2 * i] = pts[i].x; xy[2 * i + 1] = pts[i].y; }
  return (int)pts.size();
}
// cells the reference marks free for one end point: returns count (or -1 on throw)
int ref_grid_free_cells(void *h, const double pt[2], const double pose[3], int *cells, int cap)
{
  try {
    std::vector<int> idx;
    rigid2d::Transform2D T(rigid2d::Vector2D(pose[1], pose[2]), pose[0]);
    static_cast<RefGrid *>(h)->grid.freeGridIndex(idx, rigid2d::Vector2D(pt[0], pt[1]), T);
    for (size_t i = 0; i < idx.size() && (int)i < cap; i++) cells[i] = idx[i];
    return (int)idx.size();
  } catch (std::exception &) { return -1; }
}

// ------------------------------------------------------------------------ ParticleFilter ----
// pfp: srr,srt,str,stt, motion_noise(theta,x,y), sample_range(theta,x,y), scan_min,scan_max, pose_min,pose_max
void *ref_pf_create(int num_particles, int k, const double pfp[14], const float lf[5], const double ld[5],
                    double res, double xmin, double xmax, double ymin, double ymax, const double init_pose[3])
{
  RefPf *r = new RefPf(makeProps(lf, ld), res, xmin, xmax, ymin, ymax);
  rigid2d::Transform2D T0(rigid2d::Vector2D(init_pose[1], init_pose[2]), init_pose[0]);
  r->pf.reset(new bmapping::ParticleFilter(num_particles, k, pfp[0], pfp[1], pfp[2], pfp[3], pfp[4], pfp[5], pfp[6],
                                           pfp[7], pfp[8], pfp[9], pfp[10], pfp[11], pfp[12], pfp[13],
                                           r->matcher, T0, r->proto));
  return r;
}
void ref_pf_destroy(void *h) { delete static_cast<RefPf *>(h); }
void ref_pf_set_icp(int ok, const double pose[3])
{
  g_icp_ok = ok != 0;
  if (pose) { g_icp_pose[0] = pose[0]; g_icp_pose[1] = pose[1]; g_icp_pose[2] = pose[2]; }
}
// returns 0 ok, 1 when the reference throws; *resampled reports whether "Resampling" was printed
int ref_pf_slam(void *h, const float *scan, int n, const double twist[3], const double cur_odom[3],
                const double prev_odom[3], int *resampled)
{
  RefPf *r = static_cast<RefPf *>(h);
  rigid2d::Twist2D u; u.w = twist[0]; u.vx = twist[1]; u.vy = twist[2];
  rigid2d::Pose c; c.theta = cur_odom[0]; c.x = cur_odom[1]; c.y = cur_odom[2];
  rigid2d::Pose p; p.theta = prev_odom[0]; p.x = prev_odom[1]; p.y = prev_odom[2];
  SilenceCout quiet;
  try {
    r->pf->SLAM(std::vector<float>(scan, scan + n), u, c, p);
  } catch (std::exception &) { return 1; }
  if (resampled) *resampled = quiet.sink.str().find("Resampling") != std::string::npos;
  return 0;
}
int ref_pf_num(void *h) { return static_cast<RefPf *>(h)->pf->num_particles_; }
void ref_pf_get(void *h, double *weights, double *poses /*[N][3] theta,x,y*/, double *prev_poses)
{
  bmapping::ParticleFilter &pf = *static_cast<RefPf *>(h)->pf;
  const int N = pf.num_particles_;
  for (int i = 0; i < N; i++) {
    const bmapping::Particle &p = pf.particle_set_[i];
    if (weights) weights[i] = p.weight;
    for (int c = 0; c < 3; c++) {
      if (poses) poses[i * 3 + c] = p.pose(c);
      if (prev_poses) prev_poses[i * 3 + c] = p.prev_pose(c);
    }
  }
}
void ref_pf_set_weights(void *h, const double *weights)
{
  bmapping::ParticleFilter &pf = *static_cast<RefPf *>(h)->pf;
  for (int i = 0; i < pf.num_particles_; i++) pf.particle_set_[i].weight = weights[i];
}
void ref_pf_grid_dump(void *h, int particle, double *log_odds, double *prob, double *occ_dist, int *state)
{
  dumpGrid(static_cast<RefPf *>(h)->pf->particle_set_[particle].grid, log_odds, prob, occ_dist, state);
}
int ref_pf_grid_occ_order(void *h, int particle, int *keys, int cap)
{
  return dumpOccOrder(static_cast<RefPf *>(h)->pf->particle_set_[particle].grid, keys, cap);
}
void ref_pf_robot_state(void *h, double out[3])
{
  rigid2d::TransformData2D d = static_cast<RefPf *>(h)->pf->getRobotState().displacement();
  out[0] = d.theta; out[1] = d.x; out[2] = d.y;
}
void ref_pf_new_map(void *h, int8_t *out)
{
  std::vector<int8_t> m;
  static_cast<RefPf *>(h)->pf->newMap(m);
  std::memcpy(out, m.data(), m.size());
}
// the reference's own normalise + N_eff test + low-variance walk on the current weights; the
// ancestor of slot m is recovered by tagging each particle's prev_pose(0) with its index first.
int ref_pf_normalize_resample(void *h, int *resampled, int *ancestors)
{
  bmapping::ParticleFilter &pf = *static_cast<RefPf *>(h)->pf;
  SilenceCout quiet;
  const int N = pf.num_particles_;
  for (int i = 0; i < N; i++) pf.particle_set_[i].prev_pose(0) = (double)i;
  pf.normalizeWeights();
  const bool rs = pf.effectiveParticles();
  if (rs) pf.lowVarianceResampling();
  *resampled = rs ? 1 : 0;
  for (int i = 0; i < N; i++) ancestors[i] = (int)pf.particle_set_[i].prev_pose(0);
  return 0;
}

} // extern "C"
