// TEST INFRASTRUCTURE ONLY (oracle/). Not part of the product path: only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
//
// CPU restatement (fp64, Eigen-free, single thread) of controller::MPPI, statement by statement
// in the reference's arithmetic order.  Each block cites the reference lines it follows
// (paths relative to /root/reference).  Pinned by tests/test_oracle_mppi.py against
// oracle/_ref/libref_nav.so (the unmodified reference sources compiled here) and against the
// committed fixtures in tests/golden/ that were generated from it.
//
// Differences from the reference that are deliberate and visible:
//   * the noise source is a seam (mode A = reference engine, mode B = Philox, mode C = caller
//     supplied perturbations) instead of a process-global engine;
//   * every intermediate (states, running loss, cost-to-go, weights) is kept so tests can
//     compare them; the reference discards the trajectories;
//   * an optional obstacle term (SURVEY.md 8d "MPPI obstacle cost (C4)") that the reference
//     does not have; it is off unless a distance field is attached.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "noise.hpp"

namespace
{

struct Params
{
  double wheel_radius, wheel_base;
  double Q[3], R[2], P1[3];
  double lambda, max_wheel_vel, ul_var, ur_var, horizon, dt;
  int rollouts;
};

struct ObstacleField
{
  bool on = false;
  std::vector<float> dist;      // [xsize*ysize], index i*xsize + j  (grid_mapper.cpp:890-898)
  int xsize = 0, ysize = 0;
  double xmin = 0, ymin = 0, xmax = 0, ymax = 0, res = 1;
  double weight = 0, d0 = 0, off_map = 0;
};

struct Mppi
{
  Params p;
  int T, K;
  double xd[3] = {0, 0, 0};
  double uinit[2] = {0, 0};
  std::vector<double> u;        // [2][T]
  std::vector<double> du;       // [K][T][2]  perturbations (L,R)
  std::vector<double> states;   // [K][T][3]
  std::vector<double> loss;     // [T][K]
  std::vector<double> J;        // [T][K] cost-to-go (before the per-step min is subtracted)
  std::vector<double> w;        // [T][K] normalised weights
  int mode = 1;                 // 0 = A (mt19937_64), 1 = B (philox), 2 = C (external)
  orc::RefNormalStream mt;
  uint64_t seed = 0;
  uint32_t call = 0;
  int k_offset = 0;             // global index of local rollout 0 (multi-GPU shards)
  const double *ext = nullptr;
  ObstacleField obs;
};

// controller/include/controller/mppi.hpp:41-48
inline void kinematicCart(const Params &p, const double x[3], const double u[2], double xdot[3])
{
  xdot[0] = (p.wheel_radius / 2.0) * (u[0] + u[1]) * std::cos(x[2]);
  xdot[1] = (p.wheel_radius / 2.0) * (u[0] + u[1]) * std::sin(x[2]);
  xdot[2] = (p.wheel_radius / p.wheel_base) * (u[1] - u[0]);
}

// controller/src/controller/rk4.cpp:95-115
inline void rk4Integrate(const Params &p, double x[3], const double u[2])
{
  const double step = p.dt;
  double k1[3], k2[3], k3[3], k4[3], arg[3];
  kinematicCart(p, x, u, k1);
  for (int c = 0; c < 3; c++) arg[c] = x[c] + step * (0.5 * k1[c]);
  kinematicCart(p, arg, u, k2);
  for (int c = 0; c < 3; c++) arg[c] = x[c] + step * (0.5 * k2[c]);
  kinematicCart(p, arg, u, k3);
  for (int c = 0; c < 3; c++) arg[c] = x[c] + step * k3[c];
  kinematicCart(p, arg, u, k4);
  for (int c = 0; c < 3; c++) x[c] = x[c] + (step / 6.0) * (k1[c] + 2.0 * k2[c] + 2.0 * k3[c] + k4[c]);
}

// diagonal quadratic form evaluated the way (e^T * M * e) is: row vector first, then the dot
inline double quad3(const double e[3], const double m[3])
{
  double s = 0.0;
  for (int c = 0; c < 3; c++) s += (e[c] * m[c]) * e[c];
  return s;
}

// extension, not in the reference: w * max(0, d0 - occ_dist(x,y))^2, cell lookup as
// bmapping/src/bmapping/grid_mapper.cpp:852-887
inline double obstacleCost(const ObstacleField &o, double x, double y)
{
  if (!o.on) return 0.0;
  if (!(x >= o.xmin && x <= o.xmax) || !(y >= o.ymin && y <= o.ymax)) return o.off_map;
  double i = std::floor((x - o.xmin) / o.res);
  if (i == o.xsize) i--;
  double j = std::floor((y - o.ymin) / o.res);
  if (j == o.ysize) j--;
  const double d = (double)o.dist[(size_t)((int)i * o.xsize + (int)j)];
  const double pen = o.d0 - d;
  return pen > 0.0 ? o.weight * pen * pen : 0.0;
}

void newControls(Mppi &m, double px, double py, double ptheta, double *ul, double *ur)
{
  const Params &p = m.p;
  const int T = m.T, K = m.K;
  // mppi.cpp:75-76 : state order (x, y, theta)
  const double x0[3] = {px, py, ptheta};
  // mppi.cpp:176-177 : the generator takes a standard deviation
  const double ul_sig = std::sqrt(p.ul_var), ur_sig = std::sqrt(p.ur_var);

  for (int k = 0; k < K; k++) {                                       // mppi.cpp:81
    double *duk = &m.du[(size_t)k * T * 2];
    // mppi.cpp:84-89,173-184 : draw order L then R, step by step
    for (int i = 0; i < T; i++) {
      if (m.mode == 0) {
        duk[2 * i] = m.mt.normal(0.0, ul_sig);
        duk[2 * i + 1] = m.mt.normal(0.0, ur_sig);
      } else if (m.mode == 1) {
        // one Philox call per PAIR of time steps: binary32 variates (L, R) of step 2j then of step 2j + 1
        float z[4];
        orc::philox_normal_quad_f32(m.seed, orc::DOMAIN_MPPI, m.call, (uint32_t)(m.k_offset + k), (uint32_t)(i >> 1), z);
        duk[2 * i] = (double)z[2 * (i & 1)] * ul_sig;
        duk[2 * i + 1] = (double)z[2 * (i & 1) + 1] * ur_sig;
      } else {
        duk[2 * i] = m.ext[((size_t)k * T + i) * 2];
        duk[2 * i + 1] = m.ext[((size_t)k * T + i) * 2 + 1];
      }
    }
    // mppi.cpp:93-106, rk4.cpp:49-69
    double state[3] = {x0[0], x0[1], x0[2]};
    for (int i = 0; i < T; i++) {
      const double up[2] = {m.u[i] + duk[2 * i], m.u[T + i] + duk[2 * i + 1]};   // NOT clamped
      rk4Integrate(p, state, up);
      double *s = &m.states[((size_t)k * T + i) * 3];
      s[0] = state[0]; s[1] = state[1]; s[2] = state[2];
      const double e[3] = {state[0] - m.xd[0], state[1] - m.xd[1], state[2] - m.xd[2]};   // theta NOT wrapped
      double l;
      if (i < T - 1) {
        // mppi.hpp:87-93
        l = quad3(e, p.Q) + ((up[0] * p.R[0]) * up[0] + (up[1] * p.R[1]) * up[1]);
      } else {
        // mppi.cpp:105, mppi.hpp:100-105 : terminal loss REPLACES the running loss
        l = quad3(e, p.P1);
      }
      l += obstacleCost(m.obs, state[0], state[1]);
      m.loss[(size_t)i * K + k] = l;
    }
  }

  // mppi.cpp:15-25,109 : cost-to-go, from the end
  for (int k = 0; k < K; k++) m.J[(size_t)(T - 1) * K + k] = m.loss[(size_t)(T - 1) * K + k];
  for (int i = T - 2; i >= 0; i--)
    for (int k = 0; k < K; k++) m.J[(size_t)i * K + k] = m.loss[(size_t)i * K + k] + m.J[(size_t)(i + 1) * K + k];

  // mppi.cpp:112-126 : T independent softmaxes over K
  for (int i = 0; i < T; i++) {
    const double *Ji = &m.J[(size_t)i * K];
    double *wi = &m.w[(size_t)i * K];
    double mn = Ji[0];
    for (int k = 1; k < K; k++) if (Ji[k] < mn) mn = Ji[k];
    double sum = 0.0;
    for (int k = 0; k < K; k++) {
      wi[k] = std::exp(((Ji[k] - mn) * -1.0) / p.lambda) + 1e-8;
      sum += wi[k];
    }
    const double inv = 1.0 / sum;
    double aL = 0.0, aR = 0.0;
    for (int k = 0; k < K; k++) {
      wi[k] = wi[k] * inv;
      aL += wi[k] * m.du[((size_t)k * T + i) * 2];
      aR += wi[k] * m.du[((size_t)k * T + i) * 2 + 1];
    }
    m.u[i] += aL;
    m.u[T + i] += aR;
    m.u[i] = std::clamp(m.u[i], -p.max_wheel_vel, p.max_wheel_vel);
    m.u[T + i] = std::clamp(m.u[T + i], -p.max_wheel_vel, p.max_wheel_vel);
  }

  // mppi.cpp:129-137 : emit, shift, re-seed the tail
  *ul = m.u[0];
  *ur = m.u[T];
  for (int i = 0; i + 1 < T; i++) { m.u[i] = m.u[i + 1]; m.u[T + i] = m.u[T + i + 1]; }
  m.u[T - 1] = m.uinit[0];
  m.u[2 * T - 1] = m.uinit[1];
  m.call++;
}

} // namespace

extern "C" {

struct orc_mppi_params
{
  double wheel_radius, wheel_base;
  double Q[3], R[2], P1[3];
  double lambda, max_wheel_vel, ul_var, ur_var, horizon, dt;
  int rollouts;
};

void *orc_mppi_create(const orc_mppi_params *pp)
{
  Mppi *m = new Mppi();
  std::memcpy(&m->p, pp, sizeof(Params));
  m->K = pp->rollouts;
  m->T = static_cast<int>(pp->horizon / pp->dt);          // mppi.cpp:47 (truncation and all)
  const size_t T = m->T, K = m->K;
  m->u.assign(2 * T, 0.0);                                 // mppi.cpp:157-170
  m->du.assign(K * T * 2, 0.0);
  m->states.assign(K * T * 3, 0.0);
  m->loss.assign(T * K, 0.0);
  m->J.assign(T * K, 0.0);
  m->w.assign(T * K, 0.0);
  return m;
}
void orc_mppi_destroy(void *h) { delete static_cast<Mppi *>(h); }
int orc_mppi_steps(void *h) { return static_cast<Mppi *>(h)->T; }

// mppi.cpp:54-61
void orc_mppi_set_initial_controls(void *h, double ul, double ur)
{
  Mppi *m = static_cast<Mppi *>(h);
  m->uinit[0] = ul; m->uinit[1] = ur;
  for (int i = 0; i < m->T; i++) { m->u[i] = ul; m->u[m->T + i] = ur; }
}
// mppi.cpp:64-69
void orc_mppi_set_waypoint(void *h, double x, double y, double theta)
{
  Mppi *m = static_cast<Mppi *>(h);
  m->xd[0] = x; m->xd[1] = y; m->xd[2] = theta;
}
void orc_mppi_noise_mt19937(void *h, uint64_t seed) { Mppi *m = static_cast<Mppi *>(h); m->mode = 0; m->mt.seed(seed); }
void orc_mppi_noise_philox(void *h, uint64_t seed, uint32_t first_call)
{
  Mppi *m = static_cast<Mppi *>(h); m->mode = 1; m->seed = seed; m->call = first_call;
}
void orc_mppi_noise_external(void *h, const double *du) { Mppi *m = static_cast<Mppi *>(h); m->mode = 2; m->ext = du; }
void orc_mppi_set_shard(void *h, int k_offset) { static_cast<Mppi *>(h)->k_offset = k_offset; }
void orc_mppi_set_plan(void *h, const double *u) { Mppi *m = static_cast<Mppi *>(h); std::copy(u, u + 2 * m->T, m->u.begin()); }
void orc_mppi_set_obstacles(void *h, const float *dist, int xsize, int ysize, double xmin, double ymin, double res,
                            double weight, double d0, double off_map)
{
  ObstacleField &o = static_cast<Mppi *>(h)->obs;
  o.on = dist != nullptr;
  if (!o.on) return;
  o.dist.assign(dist, dist + (size_t)xsize * ysize);
  o.xsize = xsize; o.ysize = ysize; o.xmin = xmin; o.ymin = ymin; o.res = res;
  o.xmax = xmin + xsize * res; o.ymax = ymin + ysize * res;
  o.weight = weight; o.d0 = d0; o.off_map = off_map;
}
void orc_mppi_new_controls(void *h, double x, double y, double theta, double *ul, double *ur)
{
  newControls(*static_cast<Mppi *>(h), x, y, theta, ul, ur);
}
// any pointer may be null.  u_plan [2][T]; du [K][T][2]; states [K][T][3]; J, w [T][K]
void orc_mppi_get(void *h, double *u_plan, double *du, double *states, double *J, double *w)
{
  Mppi *m = static_cast<Mppi *>(h);
  if (u_plan) std::copy(m->u.begin(), m->u.end(), u_plan);
  if (du) std::copy(m->du.begin(), m->du.end(), du);
  if (states) std::copy(m->states.begin(), m->states.end(), states);
  if (J) std::copy(m->J.begin(), m->J.end(), J);
  if (w) std::copy(m->w.begin(), m->w.end(), w);
}
// raw noise taps, for pinning the generators themselves
void orc_mt_normals(uint64_t seed, int n, double mu, double sigma, double *out)
{
  orc::RefNormalStream s; s.seed(seed);
  for (int i = 0; i < n; i++) out[i] = s.normal(mu, sigma);
}
void orc_philox_raw(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { orc::philox4x32_10(ctr, key, out); }
void orc_philox_normal_quad_f32(uint64_t seed, uint32_t domain, uint32_t call, uint32_t stream, uint32_t index, float *z)
{
  orc::philox_normal_quad_f32(seed, domain, call, stream, index, z);
}
// the Box-Muller stage alone over a range of first words (ra = (first + i) << 9, every possible radius), one second word
void orc_box_muller_range(uint32_t first, uint32_t count, uint32_t rb, float *z)
{
  for (uint32_t i = 0; i < count; i++) orc::box_muller_f32((first + i) << 9, rb, &z[2 * (size_t)i], &z[2 * (size_t)i + 1]);
}
void orc_philox_normal_pair(uint64_t seed, uint32_t domain, uint32_t call, uint32_t stream, uint32_t index, double *z)
{
  orc::philox_normal_pair(seed, domain, call, stream, index, &z[0], &z[1]);
}

} // extern "C"
