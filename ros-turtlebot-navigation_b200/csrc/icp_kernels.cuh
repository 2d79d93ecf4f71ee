// icp_kernels.cuh - libb2nav's scan matcher: point-to-point ICP between two lidar scans on one CTA.
//
// The reference delegates scan matching to PCL (pcl::IterativeClosestPoint, bmapping/src/bmapping/cloud_alignment.cpp:
// 160-223); PCL is not part of /root/reference and cannot be pinned here, so this is the PUBLISHED structure of that
// algorithm with the reference's settings (cloud_alignment.cpp:20-25,186-190), checked against oracle/icp_oracle.cpp -
// not a claim of bit parity with PCL.  Its result feeds the improved-proposal branch of
// ParticleFilter::SLAM (particle_filter.cpp:150-153,178-233).
//
// Work per iteration: every thread owns source points, transforms them by the current estimate, scans ALL target
// points in shared memory for the nearest one (first minimum wins, like a sequential search), gates the pair at
// max_correspondence_dist; seven threads then add the pairs' contributions in point order (the same sequential
// fp64 sums the oracle forms - this translation unit is built with -fmad=false), thread 0 solves the planar
// alignment in closed form, composes it and applies the stopping rules of PCL's DefaultConvergenceCriteria.
#pragma once

#include "common.cuh"

namespace b2n
{

constexpr int kIcpThreads = 512;
constexpr int kIcpMaxPoints = 2048;

struct IcpArgs
{
  const float *tgt, *src;      // [n][2] x, y
  int nt, ns, max_iter;
  double Tinit[3];             // theta, x, y
  double max_d2, transform_eps, fitness_eps;
  double *out;                 // [8]: theta, x, y, converged, iterations, pairs, mse, 0
};

// dynamic shared memory: target points (float2 [nt]) + per-source contributions (8 doubles each)
__host__ __device__ inline size_t icp_smem_bytes(int nt, int ns) { return (size_t)((nt + 1) & ~1) * 8 + (size_t)ns * 8 * 8; }

__global__ void __launch_bounds__(kIcpThreads) icp_align_kernel(const __grid_constant__ IcpArgs a)
{
  extern __shared__ __align__(16) unsigned char smem[];
  float2 *tg = reinterpret_cast<float2 *>(smem);
  double *contrib = reinterpret_cast<double *>(smem + (size_t)((a.nt + 1) & ~1) * 8);   // [ns][8]
  __shared__ double sums[8];
  __shared__ double est[4];      // c, s, tx, ty
  __shared__ int state[2];       // done, converged
  const int tid = threadIdx.x, nt = a.nt, ns = a.ns;
  for (int j = tid; j < nt; j += kIcpThreads) tg[j] = make_float2(a.tgt[2 * j], a.tgt[2 * j + 1]);
  if (tid == 0) {
    est[0] = cos(a.Tinit[0]); est[1] = sin(a.Tinit[0]); est[2] = a.Tinit[1]; est[3] = a.Tinit[2];
    state[0] = 0; state[1] = 0;
  }
  __syncthreads();
  double prev_mse = 1.0e300;
  int it = 0, pairs = 0;
  double mse = 0.0;
  for (it = 1; it <= a.max_iter; it++) {
    const double c = est[0], s = est[1], tx = est[2], ty = est[3];
    for (int i = tid; i < ns; i += kIcpThreads) {
      const double x = (double)a.src[2 * i], y = (double)a.src[2 * i + 1];
      const double sx = c * x - s * y + tx, sy = s * x + c * y + ty;
      int best = -1;
      double bd = 1.0e300;
      for (int j = 0; j < nt; j++) {
        const float2 q = tg[j];
        const double dx = sx - (double)q.x, dy = sy - (double)q.y;
        const double d2 = dx * dx + dy * dy;
        if (d2 < bd) { bd = d2; best = j; }
      }
      double *o = contrib + (size_t)i * 8;
      if (bd > a.max_d2) {
        o[0] = 0.0; o[1] = 0.0; o[2] = 0.0; o[3] = 0.0; o[4] = 0.0; o[5] = 0.0; o[6] = 0.0; o[7] = 0.0;
      } else {
        const double qx = (double)tg[best].x, qy = (double)tg[best].y;
        o[0] = sx; o[1] = sy; o[2] = qx; o[3] = qy; o[4] = sx * qx + sy * qy; o[5] = sx * qy - sy * qx; o[6] = bd; o[7] = 1.0;
      }
    }
    __syncthreads();
    if (tid < 8) {
      // in point order, skipping gated points exactly like the sequential reference loop
      double acc = 0.0;
      for (int i = 0; i < ns; i++)
        if (contrib[(size_t)i * 8 + 7] != 0.0) acc += contrib[(size_t)i * 8 + tid];
      sums[tid] = acc;
    }
    __syncthreads();
    if (tid == 0) {
      const int m = (int)sums[7];
      pairs = m;
      if (m < 3) { state[0] = 1; state[1] = 0; }
      else {
        const double inv = 1.0 / m;
        const double mx = sums[0] * inv, my = sums[1] * inv, qx = sums[2] * inv, qy = sums[3] * inv;
        const double aa = sums[4] - m * (mx * qx + my * qy);
        const double bb = sums[5] - m * (mx * qy - my * qx);
        const double dth = atan2(bb, aa);
        const double dc = cos(dth), ds = sin(dth);
        const double dtx = qx - (dc * mx - ds * my), dty = qy - (ds * mx + dc * my);
        const double nc = dc * c - ds * s, nsn = ds * c + dc * s;
        const double ntx = dc * tx - ds * ty + dtx, nty = ds * tx + dc * ty + dty;
        est[0] = nc; est[1] = nsn; est[2] = ntx; est[3] = nty;
        mse = sums[6] * inv;
        bool conv = false;
        if (it >= a.max_iter) conv = true;
        else if (dc >= 0.99999 && dtx * dtx + dty * dty <= a.transform_eps) conv = true;
        else if (fabs(mse - prev_mse) < a.fitness_eps) conv = true;
        else if (fabs(mse - prev_mse) / prev_mse < 1.0e-5) conv = true;
        prev_mse = mse;
        if (conv) { state[0] = 1; state[1] = 1; }
      }
    }
    __syncthreads();
    if (state[0]) break;
  }
  if (tid == 0) {
    a.out[0] = atan2(est[1], est[0]); a.out[1] = est[2]; a.out[2] = est[3];
    a.out[3] = (double)state[1]; a.out[4] = (double)min(it, a.max_iter); a.out[5] = (double)pairs; a.out[6] = mse; a.out[7] = 0.0;
  }
}

} // namespace b2n
