// mppi_kernels.cuh - the MPPI hot loop of controller::MPPI::newControls() as sm_100a kernels.
//
// Reference path (all under /root/reference): controller/src/controller/mppi.cpp:72-140 (K rollouts,
// loss matrix, cost-to-go, T softmaxes over K, clamp, shift), :173-184 (perturbations),
// controller/src/controller/rk4.cpp:49-69,95-115 (RK4), controller/include/controller/mppi.hpp:41-48
// (diff-drive cart model), :87-105 (running / terminal loss).
//
// Design (DESIGN.md "MPPI"):
//   * ONE WARP PER ROLLOUT.  Lane l owns the S consecutive time steps [l*S, l*S+S).  The cart model's
//     heading rate does not depend on the state, so theta_t is a prefix sum of per-step increments
//     and (x_t, y_t) a prefix sum of increments that depend only on theta_{t-1} and the step's
//     controls: three warp-shuffle scans replace the T-long serial recurrence, and a reverse scan
//     gives the cost-to-go (cumSumCost).  All of it in fp64 registers - the shipped cost weights
//     (Q = 1e4, lambda = 0.01) make the softmax ill-conditioned in anything narrower.
//   * noise is counter-based (Philox4x32-10 keyed by seed, call, rollout, step), generated in
//     registers; nothing is read from HBM except the 2*T plan.
//   * the only mandatory HBM traffic is the fp32 [K][T][3] state tensor: each warp stages its row in
//     shared memory and one lane hands it to the TMA unit (cp.async.bulk shared->global), double
//     buffered so the next rollout's math overlaps the store.
//   * the T softmaxes are carried ONLINE: every lane keeps (min J, sum e, sum e*duL, sum e*duR,
//     sum duL, sum duR) for its S steps across all rollouts of its warp, warps merge through shared
//     memory, each CTA writes one [T][6] partial.  J never makes a round trip through HBM.
//   * mppi_update_kernel merges the partials (from all CTAs, or from all ranks after the allgather),
//     applies the weighted update, clamps, emits the first control and shifts the plan.
#pragma once

#include "common.cuh"

namespace b2n
{

constexpr int kMppiThreads = 256;            // 8 warps = 8 rollouts in flight per CTA
constexpr int kMppiWarps = kMppiThreads / 32;
constexpr int kMppiMaxS = 8;                 // T <= 256

struct MppiArgs
{
  // model, cost, sampling
  double r_half, r_over_L;
  double Q[3], R[2], P1[3];
  double inv_lambda, h, h_sixth, sigL, sigR;
  double x0[3], xd[3];
  int T, K, k_offset;
  uint32_t seed_lo, seed_hi, call;
  int external_noise, capture, tma_store;
  // optional obstacle term (extension, see b2nav.h)
  int obs_on, obs_xsize, obs_ysize;
  double obs_xmin, obs_ymin, obs_xmax, obs_ymax, obs_res, obs_weight, obs_d0, obs_off;
  const float *obs_dist;
  // buffers
  const double *u_plan;    // [2][T]
  float *states;           // [K][T][3]
  const double *ext;       // [K][T][2] or null
  double *J_out;           // [K][T] (capture)
  double *du_out;          // [K][T][2] (capture)
  double *partials;        // [gridDim.x][T][6]
};

struct MppiUpdateArgs
{
  const double *partials;  // [n_partials][T][6]
  int n_partials, T, merge_only;
  double inv_lambda, k_total, umax;
  double uinit[2];
  const double *u_cur;     // [2][T]
  double *u_next;          // [2][T]
  double *out;             // [2] first control of the updated plan
  double *stepstats;       // [T][2] (min J, sum w) for the weights tap
  double *merged;          // [T][6] when merge_only
};

__device__ __forceinline__ double mppi_obstacle_cost(const MppiArgs &a, double x, double y)
{
  if (!(x >= a.obs_xmin && x <= a.obs_xmax) || !(y >= a.obs_ymin && y <= a.obs_ymax)) return a.obs_off;
  double i = floor((x - a.obs_xmin) / a.obs_res);
  if (i == (double)a.obs_xsize) i -= 1.0;
  double j = floor((y - a.obs_ymin) / a.obs_res);
  if (j == (double)a.obs_ysize) j -= 1.0;
  const double d = (double)__ldg(&a.obs_dist[(int)i * a.obs_xsize + (int)j]);
  const double pen = a.obs_d0 - d;
  return pen > 0.0 ? a.obs_weight * pen * pen : 0.0;
}

template <int S>
__global__ void __launch_bounds__(kMppiThreads) mppi_rollout_kernel(const __grid_constant__ MppiArgs a)
{
  constexpr int SLOTS = 32 * S;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float *stage_base = reinterpret_cast<float *>(smem_raw);                                  // [warps][2][SLOTS*3]
  double *cta_acc = reinterpret_cast<double *>(smem_raw + kMppiWarps * 2 * SLOTS * 3 * 4);  // [SLOTS][6]

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int gw = blockIdx.x * kMppiWarps + warp;
  const int nw = gridDim.x * kMppiWarps;
  const int T = a.T;
  float *stage = stage_base + warp * 2 * SLOTS * 3;

  // the lane's slice of the plan, resident for the whole kernel
  double up0[S], up1[S];
  bool act[S];
#pragma unroll
  for (int s = 0; s < S; s++) {
    const int t = lane * S + s;
    act[s] = t < T;
    up0[s] = act[s] ? __ldg(&a.u_plan[t]) : 0.0;
    up1[s] = act[s] ? __ldg(&a.u_plan[T + t]) : 0.0;
  }
  double sin0, cos0;
  sincos(a.x0[2], &sin0, &cos0);

  // online-softmax accumulators of this lane's steps
  double am[S], aS[S], aA[S], aB[S], aDL[S], aDR[S];
#pragma unroll
  for (int s = 0; s < S; s++) {
    am[s] = __longlong_as_double(0x7FF0000000000000LL);
    aS[s] = aA[s] = aB[s] = aDL[s] = aDR[s] = 0.0;
  }

  int buf = 0;
  for (int k = gw; k < a.K; k += nw, buf ^= 1) {
    // ---- perturbed controls (mppi.cpp:84-93,173-184) and per-step kinematic terms ------------
    double duL[S], duR[S], uL[S], uR[S], om[S], vv[S], dth[S];
#pragma unroll
    for (int s = 0; s < S; s++) {
      const int t = lane * S + s;
      double z0 = 0.0, z1 = 0.0;
      if (a.external_noise) {
        if (act[s]) {
          const double2 e = __ldg(reinterpret_cast<const double2 *>(a.ext) + ((size_t)k * T + t));
          z0 = e.x; z1 = e.y;
        }
        duL[s] = z0; duR[s] = z1;
      } else {
        normal_pair(a.seed_lo, a.seed_hi, kDomainMppi, a.call, (uint32_t)(a.k_offset + k), (uint32_t)t, z0, z1);
        duL[s] = act[s] ? z0 * a.sigL : 0.0;
        duR[s] = act[s] ? z1 * a.sigR : 0.0;
      }
      uL[s] = up0[s] + duL[s];            // NOT clamped (mppi.cpp:93)
      uR[s] = up1[s] + duR[s];
      vv[s] = a.r_half * (uL[s] + uR[s]); // mppi.hpp:45-46
      om[s] = a.r_over_L * (uR[s] - uL[s]);   // mppi.hpp:47
      // rk4.cpp:114 on the theta component: (h/6)(k1 + 2k2 + 2k3 + k4) with all four equal to om
      dth[s] = act[s] ? a.h_sixth * (om[s] + 2.0 * om[s] + 2.0 * om[s] + om[s]) : 0.0;
    }

    // ---- theta: prefix sum over the horizon ----------------------------------------------------
    double th_off[S];
    double run = 0.0;
#pragma unroll
    for (int s = 0; s < S; s++) { th_off[s] = run; run += dth[s]; }
    double incl = warp_inclusive_sum(run, lane);
    double excl = __shfl_up_sync(kFullMask, incl, 1);
    if (lane == 0) excl = 0.0;

    // ---- RK4 stage angles: k1 at theta, k2 = k3 at theta + h/2*om, k4 at theta + h*om ----------
    double cm[S], sm[S], ce[S], se[S], tha[S];
#pragma unroll
    for (int s = 0; s < S; s++) {
      tha[s] = a.x0[2] + (excl + th_off[s]);                 // heading at the start of step t
      sincos(tha[s] + a.h * (0.5 * om[s]), &sm[s], &cm[s]);  // rk4.cpp:105-109
      sincos(tha[s] + a.h * om[s], &se[s], &ce[s]);          // rk4.cpp:111-112
    }
    // the k4 angle of step t is the k1 angle of step t+1 to within an ulp: reuse its sin/cos
    double ca_first = __shfl_up_sync(kFullMask, ce[S - 1], 1);
    double sa_first = __shfl_up_sync(kFullMask, se[S - 1], 1);
    if (lane == 0) { ca_first = cos0; sa_first = sin0; }

    double dx[S], dy[S];
#pragma unroll
    for (int s = 0; s < S; s++) {
      const double ca = (s == 0) ? ca_first : ce[s == 0 ? 0 : s - 1];
      const double sa = (s == 0) ? sa_first : se[s == 0 ? 0 : s - 1];
      const double k1x = vv[s] * ca, k23x = vv[s] * cm[s], k4x = vv[s] * ce[s];
      const double k1y = vv[s] * sa, k23y = vv[s] * sm[s], k4y = vv[s] * se[s];
      dx[s] = act[s] ? a.h_sixth * (k1x + 2.0 * k23x + 2.0 * k23x + k4x) : 0.0;   // rk4.cpp:114
      dy[s] = act[s] ? a.h_sixth * (k1y + 2.0 * k23y + 2.0 * k23y + k4y) : 0.0;
    }

    // ---- position: prefix sums -------------------------------------------------------------------
    double px[S], py[S];
    double rx = 0.0, ry = 0.0;
#pragma unroll
    for (int s = 0; s < S; s++) { rx += dx[s]; ry += dy[s]; px[s] = rx; py[s] = ry; }
    double ix = warp_inclusive_sum(rx, lane), iy = warp_inclusive_sum(ry, lane);
    double ex = __shfl_up_sync(kFullMask, ix, 1), ey = __shfl_up_sync(kFullMask, iy, 1);
    if (lane == 0) { ex = 0.0; ey = 0.0; }

    // ---- states after each step, loss (mppi.cpp:99-105, mppi.hpp:87-105) ----------------------
    double X[S], Y[S], TH[S], loss[S];
#pragma unroll
    for (int s = 0; s < S; s++) {
      const int t = lane * S + s;
      X[s] = a.x0[0] + (ex + px[s]);
      Y[s] = a.x0[1] + (ey + py[s]);
      TH[s] = tha[s] + dth[s];
      const double e0 = X[s] - a.xd[0], e1 = Y[s] - a.xd[1], e2 = TH[s] - a.xd[2];   // theta NOT wrapped
      double l;
      if (t < T - 1) {
        l = ((e0 * a.Q[0]) * e0 + (e1 * a.Q[1]) * e1 + (e2 * a.Q[2]) * e2) +
            ((uL[s] * a.R[0]) * uL[s] + (uR[s] * a.R[1]) * uR[s]);
      } else {
        l = (e0 * a.P1[0]) * e0 + (e1 * a.P1[1]) * e1 + (e2 * a.P1[2]) * e2;   // replaces the running loss
      }
      if (a.obs_on) l += mppi_obstacle_cost(a, X[s], Y[s]);
      loss[s] = act[s] ? l : 0.0;
    }

    // ---- cost-to-go: suffix sums (cumSumCost, mppi.cpp:15-25) -----------------------------------
    double J[S];
    double rj = 0.0;
#pragma unroll
    for (int s = S - 1; s >= 0; s--) { rj += loss[s]; J[s] = rj; }
    double ij = warp_inclusive_suffix_sum(rj, lane);
    double ej = __shfl_down_sync(kFullMask, ij, 1);
    if (lane == 31) ej = 0.0;
#pragma unroll
    for (int s = 0; s < S; s++) J[s] += ej;

    // ---- online softmax over rollouts, one accumulator set per owned step ---------------------
#pragma unroll
    for (int s = 0; s < S; s++) {
      if (act[s]) {
        const double d = (am[s] - J[s]) * a.inv_lambda;     // > 0: J is the new minimum
        const double e = exp(-fabs(d));
        const bool newmin = d > 0.0;
        aS[s] = newmin ? fma(aS[s], e, 1.0) : aS[s] + e;
        aA[s] = newmin ? fma(aA[s], e, duL[s]) : fma(e, duL[s], aA[s]);
        aB[s] = newmin ? fma(aB[s], e, duR[s]) : fma(e, duR[s], aB[s]);
        am[s] = fmin(am[s], J[s]);
        aDL[s] += duL[s];
        aDR[s] += duR[s];
      }
    }

    // ---- optional capture for the parity taps ------------------------------------------------------
    if (a.capture) {
#pragma unroll
      for (int s = 0; s < S; s++) {
        const int t = lane * S + s;
        if (act[s]) {
          a.J_out[(size_t)k * T + t] = J[s];
          reinterpret_cast<double2 *>(a.du_out)[(size_t)k * T + t] = make_double2(duL[s], duR[s]);
        }
      }
    }

    // ---- the state tensor: stage the row, hand it to the TMA unit ----------------------------
    float *row = stage + buf * SLOTS * 3;
    if (a.tma_store) {
      if (lane == 0) tma_store_wait_read<1>();   // the store issued two rollouts ago has drained this buffer
      __syncwarp();
    }
#pragma unroll
    for (int s = 0; s < S; s++) {
      const int t = lane * S + s;
      row[t * 3 + 0] = (float)X[s];
      row[t * 3 + 1] = (float)Y[s];
      row[t * 3 + 2] = (float)TH[s];
    }
    float *grow = a.states + (size_t)k * T * 3;
    if (a.tma_store) {
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_1d(grow, row, (uint32_t)(T * 3 * sizeof(float)));
        tma_store_commit();
      }
    } else {
      __syncwarp();
      for (int i = lane; i < T * 3; i += 32) grow[i] = row[i];
      __syncwarp();
    }
  }
  if (a.tma_store && lane == 0) tma_store_wait<0>();

  // ---- merge the warps' accumulators in shared memory, one warp at a time -------------------
  for (int w = 0; w < kMppiWarps; w++) {
    if (warp == w) {
#pragma unroll
      for (int s = 0; s < S; s++) {
        double *c = cta_acc + (lane * S + s) * 6;
        if (w == 0) {
          c[0] = am[s]; c[1] = aS[s]; c[2] = aA[s]; c[3] = aB[s]; c[4] = aDL[s]; c[5] = aDR[s];
        } else {
          const double m0 = c[0], m1 = am[s];
          const double m = fmin(m0, m1);
          // an empty accumulator has min = +inf and zero sums; exp(-inf) = 0 keeps it out
          const double f0 = (m0 == m) ? 1.0 : exp((m - m0) * a.inv_lambda);
          const double f1 = (m1 == m) ? 1.0 : exp((m - m1) * a.inv_lambda);
          c[0] = m;
          c[1] = c[1] * f0 + aS[s] * f1;
          c[2] = c[2] * f0 + aA[s] * f1;
          c[3] = c[3] * f0 + aB[s] * f1;
          c[4] += aDL[s];
          c[5] += aDR[s];
        }
      }
    }
    __syncthreads();
  }
  double *out = a.partials + (size_t)blockIdx.x * T * 6;
  for (int i = threadIdx.x; i < T * 6; i += kMppiThreads) out[i] = cta_acc[i];
}

// One warp per time step: merge the partials, then (unless merge_only) the control update of
// mppi.cpp:112-137 for that step, written one slot to the left (the receding-horizon shift).
__global__ void __launch_bounds__(32) mppi_update_kernel(const MppiUpdateArgs a)
{
  const int t = blockIdx.x;
  const int lane = threadIdx.x;
  const int T = a.T;
  const double inf = __longlong_as_double(0x7FF0000000000000LL);

  double m = inf;
  for (int p = lane; p < a.n_partials; p += 32) m = fmin(m, a.partials[((size_t)p * T + t) * 6]);
  m = warp_min(m);
  double S = 0.0, A = 0.0, B = 0.0, DL = 0.0, DR = 0.0;
  for (int p = lane; p < a.n_partials; p += 32) {
    const double *c = a.partials + ((size_t)p * T + t) * 6;
    const double f = (c[0] == m) ? 1.0 : exp((m - c[0]) * a.inv_lambda);
    S += c[1] * f; A += c[2] * f; B += c[3] * f; DL += c[4]; DR += c[5];
  }
  S = warp_sum(S); A = warp_sum(A); B = warp_sum(B); DL = warp_sum(DL); DR = warp_sum(DR);
  if (lane != 0) return;

  if (a.merge_only) {
    double *o = a.merged + (size_t)t * 6;
    o[0] = m; o[1] = S; o[2] = A; o[3] = B; o[4] = DL; o[5] = DR;
    return;
  }
  // w_k = exp(-(J_k - min)/lambda) + 1e-8, normalised (mppi.cpp:117-118)
  const double sumw = S + a.k_total * 1e-8;
  const double inv = 1.0 / sumw;
  double nl = a.u_cur[t] + (A + 1e-8 * DL) * inv;           // mppi.cpp:120-121
  double nr = a.u_cur[T + t] + (B + 1e-8 * DR) * inv;
  nl = fmin(fmax(nl, -a.umax), a.umax);                     // mppi.cpp:124-125
  nr = fmin(fmax(nr, -a.umax), a.umax);
  if (t == 0) { a.out[0] = nl; a.out[1] = nr; }             // mppi.cpp:129-131
  else { a.u_next[t - 1] = nl; a.u_next[T + t - 1] = nr; }  // mppi.cpp:134
  if (t == T - 1) { a.u_next[T - 1] = a.uinit[0]; a.u_next[2 * T - 1] = a.uinit[1]; }   // mppi.cpp:136-137
  a.stepstats[2 * t] = m;
  a.stepstats[2 * t + 1] = sumw;
}

// parity tap: the normalised weights the reference materialises at mppi.cpp:117-118
__global__ void mppi_weights_kernel(const double *J, const double *stepstats, double *w, int K, int T, double inv_lambda)
{
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)K * T) return;
  const int t = (int)(i % T);
  w[i] = (exp((stepstats[2 * t] - J[i]) * inv_lambda) + 1e-8) / stepstats[2 * t + 1];
}

} // namespace b2n
