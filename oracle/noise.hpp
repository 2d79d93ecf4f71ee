// TEST INFRASTRUCTURE ONLY (oracle/). Not part of the product path.
//
// The two noise sources the CPU oracle can consume ("noise seam"):
//
//  * RefNormalStream  - "mode A": what the reference really does.  One std::mt19937_64 engine
//    (rigid2d/src/rigid2d/utilities.cpp:12-17, bmapping/src/bmapping/particle_filter.cpp:17-22)
//    and a FRESH std::normal_distribution<double> for every draw (utilities.cpp:20-24,
//    particle_filter.cpp:27-36), i.e. libstdc++'s Marsaglia polar method with the cached second
//    variate thrown away each time.  Restated here without <random>'s distribution so that the
//    algorithm is visible; tests pin it against the compiled reference bit for bit.
//
//  * philox_normal_pair - "mode B": the counter-based generator the GPU kernels use
//    (Philox4x32-10, Salmon et al. SC'11, + Box-Muller), restated independently of
//    ros-turtlebot-navigation_b200/csrc/ so that a bug on either side shows up as a mismatch.
//    Counter layout (also DESIGN.md "Noise"):  ctr = (index, stream, call, domain),
//    key = (seed low 32, seed high 32).
#ifndef B2N_ORACLE_NOISE_HPP
#define B2N_ORACLE_NOISE_HPP

#include <cmath>
#include <cstdint>
#include <cstring>
#include <random>

namespace orc
{

// ------------------------------------------------------------------------------ mode A -------
struct RefNormalStream
{
  std::mt19937_64 eng;   // the engine itself is fully specified by the C++ standard

  void seed(uint64_t s) { eng.seed(s); }

  // std::generate_canonical<double,53> over a 64-bit engine: one engine call, converted to
  // double (round to nearest), divided by 2^64, clamped below 1.
  double canonical()
  {
    const double v = static_cast<double>(eng());
    double r = v / 18446744073709551616.0;
    if (r >= 1.0) r = std::nextafter(1.0, 0.0);
    return r;
  }

  // one draw of N(mu, sigma) from a freshly constructed distribution
  double normal(double mu, double sigma)
  {
    double x, y, r2;
    do {
      x = 2.0 * canonical() - 1.0;
      y = 2.0 * canonical() - 1.0;
      r2 = x * x + y * y;
    } while (r2 > 1.0 || r2 == 0.0);
    const double mult = std::sqrt(-2.0 * std::log(r2) / r2);
    // x*mult would be cached for the next call; the reference never makes one on this object
    return (y * mult) * sigma + mu;
  }
};

// ------------------------------------------------------------------------------ mode B -------
const uint32_t DOMAIN_MPPI = 0x4D505049u;      // "MPPI"
const uint32_t DOMAIN_RBPF = 0x52425046u;      // "RBPF"
const uint32_t STREAM_RESAMPLE = 0xFFFFFFFFu;  // stream id of the single resampling draw

inline void philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4])
{
  uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
  uint32_t k0 = key_in[0], k1 = key_in[1];
  for (int round = 0; round < 10; round++) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// 53 random bits -> (0,1): (m + 0.5) * 2^-53, never 0, never 1
inline double u01_53(uint32_t lo, uint32_t hi)
{
  const uint64_t m = (((uint64_t)hi << 32) | lo) >> 11;
  return ((double)m + 0.5) * (1.0 / 9007199254740992.0);
}

// sin(2*pi*u), cos(2*pi*u) with an exact quadrant reduction first, so that the result keeps
// full RELATIVE accuracy near the zeros (this is what CUDA's sincospi delivers).
inline void sincos_2pi(double u, double *s, double *c)
{
  const double PI = 3.14159265358979323846;
  double t = 4.0 * u;                       // exact; in (0,4)
  double q = std::floor(t + 0.5);           // nearest quarter turn 0..4
  double r = (t - q) * 0.25;                // exact; |r| <= 0.125 turns
  const double a = (2.0 * PI) * r;
  const double sr = std::sin(a), cr = std::cos(a);
  switch (((int)q) & 3) {
    case 0: *s = sr; *c = cr; break;
    case 1: *s = cr; *c = -sr; break;
    case 2: *s = -sr; *c = -cr; break;
    default: *s = -cr; *c = sr; break;
  }
}

// two independent N(0,1) variates for (seed, domain, call, stream, index)
inline void philox_normal_pair(uint64_t seed, uint32_t domain, uint32_t call, uint32_t stream, uint32_t index,
                               double *z0, double *z1)
{
  const uint32_t ctr[4] = {index, stream, call, domain};
  const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t r[4];
  philox4x32_10(ctr, key, r);
  const double u1 = u01_53(r[0], r[1]);
  const double u2 = u01_53(r[2], r[3]);
  const double rad = std::sqrt(-2.0 * std::log(u1));
  double s, c;
  sincos_2pi(u2, &s, &c);
  *z0 = rad * c;
  *z1 = rad * s;
}

// ---- MPPI perturbations: binary32 Box-Muller built from correctly-rounded operations only --------------------
// Restated independently of csrc/common.cuh::box_muller_f32.  Every step is one IEEE binary32 operation
// (fmaf / sqrtf / + - * / on floats; this file is compiled with -ffp-contract=off) or exact integer work, so the
// GPU and this code agree bit for bit.  One Philox call gives four variates: two consecutive time steps x (L, R).
inline void box_muller_f32(uint32_t ra, uint32_t rb, float *z0, float *z1)
{
  const float a = (float)(ra >> 9) + 0.5f;                       // u1 = a * 2^-23 in (0,1)
  int32_t ia;
  std::memcpy(&ia, &a, 4);
  int32_t ix = ia + (0x3f800000 - 0x3f3504f3);
  const int e = (ix >> 23) - 127 - 23;
  ix = (ix & 0x007fffff) + 0x3f3504f3;
  float f;
  std::memcpy(&f, &ix, 4);                                       // a = f * 2^(e+23), f in [sqrt(1/2), sqrt(2))
  const float t = f - 1.0f;                                       // ln f = t P(t), degree-8 fit on [-0.293, 0.414]
  float pl = std::fmaf(t, 0.0874394551f, -0.143773302f);
  pl = std::fmaf(t, pl, 0.149490952f);
  pl = std::fmaf(t, pl, -0.165606961f);
  pl = std::fmaf(t, pl, 0.199569777f);
  pl = std::fmaf(t, pl, -0.250021547f);
  pl = std::fmaf(t, pl, 0.333341837f);
  pl = std::fmaf(t, pl, -0.499999881f);
  pl = std::fmaf(t, pl, 1.0f);
  const float lnf = t * pl;
  const float L = std::fmaf(-1.3862944f, (float)e, -2.0f * lnf); // -2 ln u1
  const float rad = std::sqrt(L);
  const int32_t sv = ((int32_t)(rb << 2)) >> 8;                  // 24 signed bits
  const float y = ((float)sv + 0.5f) * 5.9604645e-08f;           // (-1/2, 1/2), never 0
  const float ang = y * 1.5707964f;
  const float z = ang * ang;
  float ps = std::fmaf(z, -1.9515296e-4f, 8.3321609e-3f);
  ps = std::fmaf(z, ps, -1.6666655e-1f);
  ps = ps * z;
  const float sn = std::fmaf(ang, ps, ang);
  float pc = std::fmaf(z, 2.4433157e-5f, -1.3887316e-3f);
  pc = std::fmaf(z, pc, 4.1666646e-2f);
  pc = std::fmaf(z, pc, -0.5f);
  const float cs = std::fmaf(z, pc, 1.0f);
  float c, d;
  switch (rb >> 30) {
    case 0: c = cs; d = sn; break;
    case 1: c = -sn; d = cs; break;
    case 2: c = -cs; d = -sn; break;
    default: c = sn; d = -cs; break;
  }
  *z0 = rad * c;
  *z1 = rad * d;
}

inline void philox_normal_quad_f32(uint64_t seed, uint32_t domain, uint32_t call, uint32_t stream, uint32_t index, float z[4])
{
  const uint32_t ctr[4] = {index, stream, call, domain};
  const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t r[4];
  philox4x32_10(ctr, key, r);
  box_muller_f32(r[0], r[1], &z[0], &z[1]);
  box_muller_f32(r[2], r[3], &z[2], &z[3]);
}

} // namespace orc
#endif
