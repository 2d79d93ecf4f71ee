// mppi_kernels.cuh - the MPPI hot loop of controller::MPPI::newControls() as sm_100a kernels.
//
// Reference path (all under /root/reference): controller/src/controller/mppi.cpp:72-140 (K rollouts,
// loss matrix, cost-to-go, T softmaxes over K, clamp, shift), :173-184 (perturbations),
// controller/src/controller/rk4.cpp:49-69,95-115 (RK4), controller/include/controller/mppi.hpp:41-48
// (diff-drive cart model), :87-105 (running / terminal loss).
//
// Design (DESIGN.md "MPPI"):
//   * WARP-COOPERATIVE ROLLOUTS.  A group of G lanes (8, 16 or the whole warp, chosen so that every lane
//     owns S = 2, 4 or 8 consecutive time steps) carries one rollout; the cart model's heading rate does not
//     depend on the state, so the T-long serial RK4 recurrence collapses into segmented warp scans: the
//     heading is a prefix sum, its sine / cosine a prefix PRODUCT of per-step rotations (each from a short
//     Taylor series of the half-step angle - no full-range sincos anywhere), the position a prefix sum of
//     increments, the cost-to-go a suffix sum (cumSumCost).  All in fp64 registers - the shipped cost
//     weights (Q = 1e4, lambda = 0.01) make the softmax ill-conditioned in anything narrower.
//   * noise is counter-based (Philox4x32-10 keyed by seed, call, rollout, step pair) and generated in
//     registers: one Philox call feeds four binary32 Box-Muller variates (two steps x two wheels) built
//     from correctly-rounded operations only, so the CPU oracle reproduces them bit for bit.
//   * the only mandatory HBM traffic is the fp32 [K][T][3] state tensor: each warp stages its rows in
//     shared memory and one lane hands them to the TMA unit (cp.async.bulk shared->global), double
//     buffered so the next rollouts' math overlaps the store.
//   * the T softmaxes are carried ONLINE: every lane keeps (min J, sum e, sum e*duL, sum e*duR,
//     sum duL, sum duR) for its S steps across all rollouts it sees, lane groups and warps merge at the
//     end, each CTA writes one [T][6] partial.  J never makes a round trip through HBM.
//   * mppi_update_kernel merges the partials (from all CTAs, or from all ranks after the allgather),
//     applies the weighted update, clamps, emits the first control and shifts the plan.
#pragma once

#include "common.cuh"

namespace b2n
{

constexpr int kMppiThreads = 256;            // 8 warps per CTA
constexpr int kMppiWarps = kMppiThreads / 32;
constexpr int kMppiMaxT = 256;               // 32 lanes x 8 steps

struct MppiArgs
{
  // model, cost, sampling
  double c_v, c_w;            // (r/2)(h/6) and (r/L) h: position and heading increment factors (mppi.hpp:45-47, rk4.cpp:114)
  double Q[3], R[2], P1[3];
  double inv_lambda, sigL, sigR;
  double x0[3], xd[3];
  double cos0, sin0;          // of the start heading x0[2], from the host's libm
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
  double *partials;        // [T][gridDim.x][6]: a step's partials of all CTAs are contiguous for the CTA that merges them
};

struct MppiUpdateArgs
{
  const double *partials;  // partial p of step t at partials + p * p_stride + t * t_stride (doubles):
  int p_stride, t_stride;  //   CTA partials [T][n][6]: (6, 6 n);  gathered per-rank results [n][T][6]: (6 T, 6)
  int n_partials, T, merge_only;
  double inv_lambda, k_total, umax;
  double uinit[2];
  const double *u_cur;     // [2][T]
  double *u_next;          // [2][T]
  double *out;             // [2] first control of the updated plan (mapped pinned host memory)
  unsigned long long *out_seq;   // completion word next to it: set to `seq` after the controls are visible to the host
  unsigned long long seq;
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

// exp(x) for x <= 0.  Below -708 the result is subnormal or zero: added to a softmax sum that already holds the
// minimum's 1.0 it cannot change a bit, so the evaluation is skipped (at the shipped lambda = 0.01 that is almost
// every term).
__device__ __forceinline__ double mppi_exp_neg(double x) { return x > -708.0 ? exp(x) : 0.0; }

// sin and cos of a small angle by Taylor series.  Generic: |d| <= 1/8 with terms to d^9 / d^10 (truncation < 3e-16
// relative), full range falls back to sincos.  ALWAYS_SMALL (the host has proved |d| <= 1/16): terms to d^7 / d^6,
// truncation 4e-17 / 6e-15 absolute.
template <bool ALWAYS_SMALL>
__device__ __forceinline__ void mppi_sincos_small(double d, double &sn, double &cs)
{
  const double z = d * d;
  if (ALWAYS_SMALL) {
    double ps = fma(z, -1.9841269841269841e-04, 8.3333333333333332e-03);    // -1/7!, 1/5!
    ps = fma(z, ps, -1.6666666666666666e-01);                               // -1/3!
    sn = fma(d * z, ps, d);
    double pc = fma(z, -1.3888888888888889e-03, 4.1666666666666664e-02);    // -1/6!, 1/4!
    pc = fma(z, pc, -0.5);
    cs = fma(z, pc, 1.0);
    return;
  }
  if (fabs(d) > 0.125) { sincos(d, &sn, &cs); return; }
  double ps = fma(z, 2.7557319223985893e-06, -1.9841269841269841e-04);   // 1/9!, -1/7!
  ps = fma(z, ps, 8.3333333333333332e-03);                                // 1/5!
  ps = fma(z, ps, -1.6666666666666666e-01);                               // -1/3!
  sn = fma(d * z, ps, d);
  double pc = fma(z, -2.7557319223985888e-07, 2.4801587301587302e-05);    // -1/10!, 1/8!
  pc = fma(z, pc, -1.3888888888888889e-03);                               // -1/6!
  pc = fma(z, pc, 4.1666666666666664e-02);                                // 1/4!
  pc = fma(z, pc, -0.5);
  cs = fma(z, pc, 1.0);
}

// dynamic shared memory: [warps][2][32*S*3] fp32 staging rows, then [S*6][threads] fp64 online-softmax accumulators
// (kept out of the register file so that three CTAs fit an SM; reused as [warps][G*S][6] for the CTA merge)
// (7-warp CTAs stage through ONE buffer so that four of them fit an SM)
__host__ __device__ constexpr int mppi_stage_buffers(int NW) { return NW == 7 ? 1 : 2; }
__host__ __device__ constexpr size_t mppi_rollout_smem(int S, int G, int NW)
{
  return (size_t)NW * mppi_stage_buffers(NW) * 32 * S * 3 * sizeof(float) + (size_t)S * 6 * NW * 32 * sizeof(double);
}

// FAST = the production configuration, decided on the host: own noise, no capture taps, no obstacle term, T == G * S
// (no partially filled lanes), TMA row stores, and half-step heading increments provably inside the Taylor range.
// NW = warps per CTA: 8, or 7 (four CTAs = 28 warps per SM) where that divides the job into equal passes.
template <int S, int G, bool FAST, int NW>
__global__ void __launch_bounds__(NW * 32, (S >= 8 ? 1 : (S >= 4 && NW != 7 ? 3 : 4))) mppi_rollout_kernel(const __grid_constant__ MppiArgs a)
{
  constexpr int kMppiWarps = NW, kMppiThreads = NW * 32;     // shadow the defaults: every layout below follows the CTA shape
  static_assert(S % 2 == 0 && (G == 8 || G == 16 || G == 32), "a Philox call covers two steps; G lanes per rollout");
  constexpr int R = 32 / G;        // rollouts a warp carries at a time
  constexpr int TP = G * S;        // padded horizon
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float *stage_base = reinterpret_cast<float *>(smem_raw);                                       // [warps][2][R*TP*3]
  constexpr int NBUF = mppi_stage_buffers(NW);
  double *acc = reinterpret_cast<double *>(smem_raw + kMppiWarps * NBUF * 32 * S * 3 * 4) + threadIdx.x;   // [S*6][threads]
  double *cta_acc = reinterpret_cast<double *>(smem_raw + kMppiWarps * NBUF * 32 * S * 3 * 4);             // (epilogue)

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int g = lane & (G - 1);    // position inside the rollout's lane group
  const int r = lane / G;          // which of the warp's R rollouts
  const int gw = blockIdx.x * kMppiWarps + warp;
  const int nw = gridDim.x * kMppiWarps;
  const int T = FAST ? TP : a.T;
  const int t0 = g * S;            // first owned step
  const bool ext_noise = !FAST && a.external_noise, capture = !FAST && a.capture, obs_on = !FAST && a.obs_on;
  const bool tma_store = FAST || a.tma_store;
  float *stage = stage_base + warp * NBUF * (R * TP * 3);

  const double sin0 = a.sin0, cos0 = a.cos0;
  const double inf = __longlong_as_double(0x7FF0000000000000LL);

  // online-softmax accumulators of this lane's steps: (min J, sum e, sum e*duL, sum e*duR, sum duL, sum duR) x S
#pragma unroll
  for (int s = 0; s < S; s++) {
    acc[(s * 6 + 0) * kMppiThreads] = inf;
#pragma unroll
    for (int j = 1; j < 6; j++) acc[(s * 6 + j) * kMppiThreads] = 0.0;
  }

  // launched programmatically dependent on the previous call's update kernel: everything above overlapped its tail;
  // the plan it writes is read from here on
  asm volatile("griddepcontrol.wait;" ::: "memory");
  int buf = 0;
  for (int base = gw * R; base < a.K; base += nw * R, buf ^= (NBUF - 1)) {
    const int k = base + r;
    const bool live = k < a.K;

    // ---- perturbed controls (mppi.cpp:84-93,173-184), per-step increments, half-step rotations ----------
    double duL[S], duR[S], cc[S], vh[S], dth[S], cd[S], sd[S];
    double tc = 1.0, ts = 0.0, tth = 0.0;      // this lane's total rotation and heading change
#pragma unroll
    for (int s = 0; s < S; s += 2) {
      float z[4] = {0.f, 0.f, 0.f, 0.f};
      if (!ext_noise)
        normal_quad_f32(a.seed_lo, a.seed_hi, kDomainMppi, a.call, (uint32_t)(a.k_offset + k), (uint32_t)((t0 + s) >> 1), z);
#pragma unroll
      for (int j = 0; j < 2; j++) {
        const int t = t0 + s + j;
        const bool act = FAST || t < T;
        double l = 0.0, rr = 0.0, p0 = 0.0, p1 = 0.0;
        if (act) {
          p0 = __ldg(&a.u_plan[t]);
          p1 = __ldg(&a.u_plan[T + t]);
          if (ext_noise) {
            if (live) {
              const double2 e = __ldg(reinterpret_cast<const double2 *>(a.ext) + ((size_t)k * T + t));
              l = e.x; rr = e.y;
            }
          } else {
            l = (double)z[2 * j] * a.sigL;
            rr = (double)z[2 * j + 1] * a.sigR;
          }
        }
        duL[s + j] = l; duR[s + j] = rr;
        const double uL = p0 + l, uR = p1 + rr;                 // NOT clamped (mppi.cpp:93)
        cc[s + j] = (uL * a.R[0]) * uL + (uR * a.R[1]) * uR;    // u^T R u (mppi.hpp:92)
        vh[s + j] = a.c_v * (uL + uR);                          // (h/6) v, v = (r/2)(uL + uR)   (mppi.hpp:45-46)
        dth[s + j] = a.c_w * (uR - uL);                         // h w, w = (r/L)(uR - uL): rk4.cpp:114 with k1 = k2 = k3 = k4
        mppi_sincos_small<FAST>(0.5 * dth[s + j], sd[s + j], cd[s + j]);
        // product of the lane's HALF-step rotations (squared once below: rotations commute)
        const double nc = tc * cd[s + j] - ts * sd[s + j];
        ts = fma(tc, sd[s + j], ts * cd[s + j]);
        tc = nc;
        tth += dth[s + j];
      }
    }
    {
      // the lane's total rotation = (product of half steps)^2
      const double nc = fma(tc, tc, -(ts * ts));
      ts = 2.0 * (ts * tc);
      tc = nc;
    }

    // ---- segmented inclusive scans over the rollout's G lanes: rotation product and heading sum ----------
#pragma unroll
    for (int d = 1; d < G; d <<= 1) {
      const double oc = __shfl_up_sync(kFullMask, tc, d, G), os = __shfl_up_sync(kFullMask, ts, d, G);
      const double ot = __shfl_up_sync(kFullMask, tth, d, G);
      if (g >= d) {
        const double nc = tc * oc - ts * os;
        ts = fma(tc, os, ts * oc);
        tc = nc;
        tth += ot;
      }
    }
    // exclusive prefix, seeded with the start heading
    double rc = __shfl_up_sync(kFullMask, tc, 1, G), rs = __shfl_up_sync(kFullMask, ts, 1, G);
    double th = __shfl_up_sync(kFullMask, tth, 1, G);
    if (g == 0) { rc = 1.0; rs = 0.0; th = 0.0; }
    {
      const double nc = rc * cos0 - rs * sin0;
      rs = fma(rc, sin0, rs * cos0);
      rc = nc;
    }

    // ---- RK4 position increments: k1 at theta, k2 = k3 at theta + h w / 2, k4 at theta + h w (rk4.cpp:95-115) ----
    double px[S], py[S], TH[S];
    double ax = 0.0, ay = 0.0;
#pragma unroll
    for (int s = 0; s < S; s++) {
      const double mc = rc * cd[s] - rs * sd[s], ms = fma(rc, sd[s], rs * cd[s]);      // mid-step heading
      const double ec = mc * cd[s] - ms * sd[s], es = fma(mc, sd[s], ms * cd[s]);      // end-of-step heading
      // (h/6) v (k1 + 2 k2 + 2 k3 + k4): with the headings theta, theta + d, theta + 2 d the bracket is
      // R (1 + 4 e^{id} + e^{2id}) = R e^{id} (4 + 2 cos d) - the mid-step heading scaled; running sum inside the lane
      const double f = vh[s] * fma(2.0, cd[s], 4.0);
      ax = fma(f, mc, ax);
      ay = fma(f, ms, ay);
      px[s] = ax; py[s] = ay;
      th += dth[s];
      TH[s] = a.x0[2] + th;
      rc = ec; rs = es;
    }

    // ---- position: segmented prefix sums -------------------------------------------------------------------
    double ix = ax, iy = ay;
#pragma unroll
    for (int d = 1; d < G; d <<= 1) {
      const double ox = __shfl_up_sync(kFullMask, ix, d, G), oy = __shfl_up_sync(kFullMask, iy, d, G);
      if (g >= d) { ix += ox; iy += oy; }
    }
    double ex = __shfl_up_sync(kFullMask, ix, 1, G), ey = __shfl_up_sync(kFullMask, iy, 1, G);
    if (g == 0) { ex = 0.0; ey = 0.0; }
    ex += a.x0[0]; ey += a.x0[1];

    // ---- states after each step -> staging row; loss (mppi.cpp:99-105, mppi.hpp:87-105) -------------------
    float *row = stage + buf * (R * TP * 3) + r * (T * 3);
    if (tma_store) {
      if (lane == 0) tma_store_wait_read<NBUF - 1>();   // the store that last used this buffer has drained it
      __syncwarp();
    }
    double J[S];
    double rj = 0.0;
#pragma unroll
    for (int s = S - 1; s >= 0; s--) {
      const int t = t0 + s;
      const double X = ex + px[s], Y = ey + py[s];
      double l = 0.0;
      if (FAST || t < T) {
        row[t * 3 + 0] = (float)X;
        row[t * 3 + 1] = (float)Y;
        row[t * 3 + 2] = (float)TH[s];
        const double e0 = X - a.xd[0], e1 = Y - a.xd[1], e2 = TH[s] - a.xd[2];   // theta NOT wrapped
        if (t < T - 1) l = ((e0 * a.Q[0]) * e0 + (e1 * a.Q[1]) * e1 + (e2 * a.Q[2]) * e2) + cc[s];
        else l = (e0 * a.P1[0]) * e0 + (e1 * a.P1[1]) * e1 + (e2 * a.P1[2]) * e2;   // replaces the running loss
        if (obs_on) l += mppi_obstacle_cost(a, X, Y);
      }
      rj += l;
      J[s] = rj;                                  // suffix sum inside the lane (cumSumCost, mppi.cpp:15-25)
    }

    // ---- cost-to-go: segmented suffix sum over the lanes ---------------------------------------------------
    double ij = rj;
#pragma unroll
    for (int d = 1; d < G; d <<= 1) {
      const double oj = __shfl_down_sync(kFullMask, ij, d, G);
      if (g + d < G) ij += oj;
    }
    double ej = __shfl_down_sync(kFullMask, ij, 1, G);
    if (g == G - 1) ej = 0.0;

    // ---- online softmax over rollouts, one accumulator set per owned step ---------------------------------
#pragma unroll
    for (int s = 0; s < S; s++) {
      const double Js = J[s] + ej;
      if (live && (FAST || t0 + s < T)) {
        double *c = acc + s * 6 * kMppiThreads;
        const double m0 = c[0];
        const double d = (m0 - Js) * a.inv_lambda;     // > 0: J is the new minimum
        const double e = mppi_exp_neg(-fabs(d));
        const bool newmin = d > 0.0;
        // a rollout whose weight underflows against the running minimum leaves the three sums untouched - at the
        // shipped temperature that is nearly every term, and the accumulators stay in shared memory unread
        if (newmin || e != 0.0) {
          const double S0 = c[kMppiThreads], A0 = c[2 * kMppiThreads], B0 = c[3 * kMppiThreads];
          const double scale = newmin ? e : 1.0, add = newmin ? 1.0 : e;
          c[kMppiThreads] = fma(S0, scale, add);
          c[2 * kMppiThreads] = fma(A0, scale, add * duL[s]);
          c[3 * kMppiThreads] = fma(B0, scale, add * duR[s]);
          if (newmin) c[0] = Js;
        }
        c[4 * kMppiThreads] += duL[s];
        c[5 * kMppiThreads] += duR[s];
        if (capture) {
          a.J_out[(size_t)k * T + t0 + s] = Js;
          reinterpret_cast<double2 *>(a.du_out)[(size_t)k * T + t0 + s] = make_double2(duL[s], duR[s]);
        }
      }
    }

    // ---- the state tensor: the warp's rows are contiguous in [K][T][3]; one bulk store ---------------------
    const int nlive = min(R, a.K - base);
    float *gdst = a.states + (size_t)base * T * 3;
    const float *ssrc = stage + buf * (R * TP * 3);
    if (tma_store) {
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_1d(gdst, ssrc, (uint32_t)(nlive * T * 3 * sizeof(float)));
        tma_store_commit();
      }
    } else {
      __syncwarp();
      for (int i = lane; i < nlive * T * 3; i += 32) gdst[i] = ssrc[i];
      __syncwarp();
    }
  }
  // the rollouts are done: let the dependent grid (the update kernel) be scheduled while this CTA merges; it blocks in
  // griddepcontrol.wait until this whole grid has completed and its partials are visible
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (tma_store && lane == 0) tma_store_wait<0>();

  // ---- merge the CTA's accumulator sets per time step: minimum first, then every set rescaled ONCE -----------------
  // (two passes instead of pairwise softmax merges: one exponential per set, all of them independent).  Lane groups
  // of a warp combine by xor-shuffles, warps through shared memory laid out [.][warp][g] so that lanes read
  // consecutive words.
  double am[S], aS[S], aA[S], aB[S], aDL[S], aDR[S];
#pragma unroll
  for (int s = 0; s < S; s++) {
    const double *c = acc + s * 6 * kMppiThreads;
    am[s] = c[0]; aS[s] = c[kMppiThreads]; aA[s] = c[2 * kMppiThreads]; aB[s] = c[3 * kMppiThreads];
    aDL[s] = c[4 * kMppiThreads]; aDR[s] = c[5 * kMppiThreads];
  }
  double *smin = reinterpret_cast<double *>(smem_raw);     // [S][warps][G], over the staging rows
  double *ssum = cta_acc;                                   // [S][warps][5][G], over the accumulator area
  __syncthreads();      // every warp has left the loop and drained its bulk stores: the staging rows are free
#pragma unroll
  for (int s = 0; s < S; s++) {
    double m = am[s];
#pragma unroll
    for (int off = G; off < 32; off <<= 1) m = fmin(m, __shfl_xor_sync(kFullMask, m, off));
    if (r == 0) smin[(s * kMppiWarps + warp) * G + g] = m;
  }
  __syncthreads();      // every thread has read its accumulators: their area may be overwritten from here on
#pragma unroll
  for (int s = 0; s < S; s++) {
    double m = smin[(s * kMppiWarps) * G + g];
#pragma unroll
    for (int w = 1; w < kMppiWarps; w++) m = fmin(m, smin[(s * kMppiWarps + w) * G + g]);
    // an empty set has min = +inf and zero sums
    const double f = (am[s] == inf) ? 0.0 : ((am[s] == m) ? 1.0 : mppi_exp_neg((m - am[s]) * a.inv_lambda));
    double v[5] = {aS[s] * f, aA[s] * f, aB[s] * f, aDL[s], aDR[s]};
#pragma unroll
    for (int j = 0; j < 5; j++) {
#pragma unroll
      for (int off = G; off < 32; off <<= 1) v[j] += __shfl_xor_sync(kFullMask, v[j], off);
      if (r == 0) ssum[((s * kMppiWarps + warp) * 5 + j) * G + g] = v[j];
    }
    am[s] = m;
  }
  __syncthreads();
  if (threadIdx.x < TP) {
    const int ss = threadIdx.x / G, gg = threadIdx.x % G;
    const int t = gg * S + ss;
    if (t < T) {
      double m = smin[(ss * kMppiWarps) * G + gg];
#pragma unroll
      for (int w = 1; w < kMppiWarps; w++) m = fmin(m, smin[(ss * kMppiWarps + w) * G + gg]);
      double v[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int w = 0; w < kMppiWarps; w++)
#pragma unroll
        for (int j = 0; j < 5; j++) v[j] += ssum[((ss * kMppiWarps + w) * 5 + j) * G + gg];
      double *out = a.partials + ((size_t)t * gridDim.x + blockIdx.x) * 6;
      reinterpret_cast<double2 *>(out)[0] = make_double2(m, v[0]);
      reinterpret_cast<double2 *>(out)[1] = make_double2(v[1], v[2]);
      reinterpret_cast<double2 *>(out)[2] = make_double2(v[3], v[4]);
    }
  }
}

// One CTA per time step: merge the partials, then (unless merge_only) the control update of
// mppi.cpp:112-137 for that step, written one slot to the left (the receding-horizon shift).
constexpr int kMppiUpdateThreads = 128;

// merge of the step's partials by one CTA: minimum first, then every partial rescaled once (independent
// exponentials), plain sums.  The result is valid in thread 0.
__device__ __forceinline__ void mppi_block_merge(const double *partials, int n_partials, int p_stride, int t_stride, int t, double inv_lambda, double &m,
                                                 double &S, double &A, double &B, double &DL, double &DR)
{
  constexpr int kPer = 4;                                   // partials held in registers per thread (one L2 round trip)
  __shared__ double red[kMppiUpdateThreads / 32][6];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double inf = __longlong_as_double(0x7FF0000000000000LL);
  const double *base = partials + (size_t)t * t_stride;
  double2 c0[kPer], c1[kPer], c2[kPer];
  m = inf;
#pragma unroll
  for (int i = 0; i < kPer; i++) {
    const int p = threadIdx.x + i * kMppiUpdateThreads;
    c0[i] = make_double2(inf, 0.0); c1[i] = make_double2(0.0, 0.0); c2[i] = make_double2(0.0, 0.0);
    if (p < n_partials) {
      const double2 *c = reinterpret_cast<const double2 *>(base + (size_t)p * p_stride);
      c0[i] = c[0]; c1[i] = c[1]; c2[i] = c[2];
    }
    m = fmin(m, c0[i].x);
  }
  for (int p = threadIdx.x + kPer * kMppiUpdateThreads; p < n_partials; p += kMppiUpdateThreads) m = fmin(m, base[(size_t)p * p_stride]);
  m = warp_min(m);
  if (lane == 0) red[warp][0] = m;
  __syncthreads();
  m = red[0][0];
#pragma unroll
  for (int w = 1; w < kMppiUpdateThreads / 32; w++) m = fmin(m, red[w][0]);
  S = 0.0; A = 0.0; B = 0.0; DL = 0.0; DR = 0.0;
#pragma unroll
  for (int i = 0; i < kPer; i++) {
    if (c0[i].x != inf) {
      const double f = (c0[i].x == m) ? 1.0 : mppi_exp_neg((m - c0[i].x) * inv_lambda);
      S = fma(c0[i].y, f, S); A = fma(c1[i].x, f, A); B = fma(c1[i].y, f, B);
    }
    DL += c2[i].x; DR += c2[i].y;
  }
  for (int p = threadIdx.x + kPer * kMppiUpdateThreads; p < n_partials; p += kMppiUpdateThreads) {
    const double2 *c = reinterpret_cast<const double2 *>(base + (size_t)p * p_stride);
    const double2 d0 = c[0], d1 = c[1], d2 = c[2];
    if (d0.x != inf) {
      const double f = (d0.x == m) ? 1.0 : mppi_exp_neg((m - d0.x) * inv_lambda);
      S = fma(d0.y, f, S); A = fma(d1.x, f, A); B = fma(d1.y, f, B);
    }
    DL += d2.x; DR += d2.y;
  }
  S = warp_sum(S); A = warp_sum(A); B = warp_sum(B); DL = warp_sum(DL); DR = warp_sum(DR);
  __syncthreads();
  if (lane == 0) { red[warp][1] = S; red[warp][2] = A; red[warp][3] = B; red[warp][4] = DL; red[warp][5] = DR; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kMppiUpdateThreads / 32; w++) {
      S += red[w][1]; A += red[w][2]; B += red[w][3]; DL += red[w][4]; DR += red[w][5];
    }
  }
}

// the control update of mppi.cpp:112-137 for step t from the fully merged sums (one thread)
__device__ __forceinline__ void mppi_apply_update(const MppiUpdateArgs &a, int t, double m, double S, double A, double B, double DL, double DR)
{
  const int T = a.T;
  // w_k = exp(-(J_k - min)/lambda) + 1e-8, normalised (mppi.cpp:117-118)
  const double sumw = S + a.k_total * 1e-8;
  const double inv = 1.0 / sumw;
  double nl = a.u_cur[t] + (A + 1e-8 * DL) * inv;           // mppi.cpp:120-121
  double nr = a.u_cur[T + t] + (B + 1e-8 * DR) * inv;
  nl = fmin(fmax(nl, -a.umax), a.umax);                     // mppi.cpp:124-125
  nr = fmin(fmax(nr, -a.umax), a.umax);
  if (t == 0) {                                             // mppi.cpp:129-131
    a.out[0] = nl; a.out[1] = nr;
    if (a.out_seq) {
      // the host polls this word instead of paying a stream synchronisation for 16 bytes
      __threadfence_system();
      asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.out_seq), "l"(a.seq) : "memory");
    }
  }
  else { a.u_next[t - 1] = nl; a.u_next[T + t - 1] = nr; }  // mppi.cpp:134
  if (t == T - 1) { a.u_next[T - 1] = a.uinit[0]; a.u_next[2 * T - 1] = a.uinit[1]; }   // mppi.cpp:136-137
  a.stepstats[2 * t] = m;
  a.stepstats[2 * t + 1] = sumw;
}

__global__ void __launch_bounds__(kMppiUpdateThreads) mppi_update_kernel(const MppiUpdateArgs a)
{
  const int t = blockIdx.x;
  // launched with programmatic stream serialization: wait here until the producing grid has finished and its
  // partials are visible (a no-op for an ordinary launch)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  // the next call's rollout grid may be scheduled now (its prologue overlaps this kernel; it waits before the plan)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  double m, S, A, B, DL, DR;
  mppi_block_merge(a.partials, a.n_partials, a.p_stride, a.t_stride, t, a.inv_lambda, m, S, A, B, DL, DR);
  if (threadIdx.x != 0) return;
  if (a.merge_only) {
    double *o = a.merged + (size_t)t * 6;
    o[0] = m; o[1] = S; o[2] = A; o[3] = B; o[4] = DL; o[5] = DR;
    return;
  }
  mppi_apply_update(a, t, m, S, A, B, DL, DR);
}

// ---- sharded rollouts: merge + exchange + update in ONE kernel over NVLink peer memory ----------------
// Every rank owns an exchange area [2 call parities][nranks][T][12] of 8-byte words; peer[j] is rank j's area mapped
// into this process (CUDA IPC).  CTA t merges this rank's CTA partials for step t and sends the 48-byte result to every
// rank (its own included: one code path) in the low-latency style of NCCL's LL protocol: each 8-byte word carries 4
// bytes of payload and the 32-bit call id, and 8-byte stores are single NVLink transactions, so a word whose upper half
// shows the current call id IS its payload - no fence, no separate flag, one NVLink write latency.  The CTA then spins
// on the 12 x nranks words of its own area, and thread 0 folds the nranks results in rank order (identical on every
// rank, so the plan stays replicated without a broadcast) and applies the update.  No NCCL call, no extra launch.
// A slot of parity p is rewritten at call c + 2 only after its owner finished call c + 1, which needed this rank's
// data of call c + 1, which was sent after this rank finished reading call c: two parities are enough.
constexpr int kMppiMaxRanks = 64;
constexpr int kMppiXchgWords = 12;                // 6 doubles = 12 payload halves

struct MppiXchgArgs
{
  unsigned long long *peer[kMppiMaxRanks];        // rank j's area [2][nranks][T][12]
  int rank, nranks, parity;                       // parity = call number & 1
  uint32_t call_id;                               // call number folded into 1 .. 2^32 - 1: never 0 (the areas start zeroed)
};

__global__ void __launch_bounds__(kMppiUpdateThreads) mppi_exchange_update_kernel(const MppiUpdateArgs a, const __grid_constant__ MppiXchgArgs x)
{
  __shared__ uint32_t mine[kMppiXchgWords];
  __shared__ uint32_t all[kMppiMaxRanks][kMppiXchgWords];
  const int t = blockIdx.x, T = a.T;
  const int par = x.parity;
  asm volatile("griddepcontrol.wait;" ::: "memory");                 // programmatic dependent launch, as in mppi_update_kernel
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  double m, S, A, B, DL, DR;
  mppi_block_merge(a.partials, a.n_partials, a.p_stride, a.t_stride, t, a.inv_lambda, m, S, A, B, DL, DR);
  if (threadIdx.x == 0) {
    const double v[6] = {m, S, A, B, DL, DR};
#pragma unroll
    for (int i = 0; i < 6; i++) {
      mine[2 * i] = (uint32_t)__double2loint(v[i]);
      mine[2 * i + 1] = (uint32_t)__double2hiint(v[i]);
    }
  }
  __syncthreads();
  const int n_words = x.nranks * kMppiXchgWords;
  for (int i = threadIdx.x; i < n_words; i += kMppiUpdateThreads) {
    const int r = i / kMppiXchgWords, w = i % kMppiXchgWords;
    // word w of this rank's slot in rank r's area
    unsigned long long *dst = x.peer[r] + (((size_t)par * x.nranks + x.rank) * T + t) * kMppiXchgWords + w;
    const unsigned long long packed = ((unsigned long long)x.call_id << 32) | mine[w];
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(dst), "l"(packed) : "memory");
  }
  for (int i = threadIdx.x; i < n_words; i += kMppiUpdateThreads) {
    const int r = i / kMppiXchgWords, w = i % kMppiXchgWords;
    const unsigned long long *src = x.peer[x.rank] + (((size_t)par * x.nranks + r) * T + t) * kMppiXchgWords + w;
    unsigned long long got;
    unsigned polls = 0;
    do {
      asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(got) : "l"(src) : "memory");
      if (++polls == (1u << 27)) __trap();          // about a minute without the peer's word: fail loudly, not silently
    } while ((uint32_t)(got >> 32) != x.call_id);
    all[r][w] = (uint32_t)got;
  }
  __syncthreads();
  // fold the nranks results in rank order (identical on every rank).  The rescale factors are independent of one
  // another: thread r computes rank r's (one exponential each, side by side), thread 0 then runs the ordered sums -
  // the same operations in the same order as a serial fold, without nranks exponentials back to back
  __shared__ double fac[kMppiMaxRanks];
  const double inf = __longlong_as_double(0x7FF0000000000000LL);
  auto val = [&](int r, int i) { return __hiloint2double((int)all[r][2 * i + 1], (int)all[r][2 * i]); };
  m = inf;
  for (int r = 0; r < x.nranks; r++) m = fmin(m, val(r, 0));
  if ((int)threadIdx.x < x.nranks) {
    const double mr = val((int)threadIdx.x, 0);
    fac[threadIdx.x] = (mr == inf) ? 0.0 : (mr == m) ? 1.0 : mppi_exp_neg((m - mr) * a.inv_lambda);
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  S = A = B = DL = DR = 0.0;
  for (int r = 0; r < x.nranks; r++) {
    if (val(r, 0) != inf) {
      const double f = fac[r];
      S = fma(val(r, 1), f, S); A = fma(val(r, 2), f, A); B = fma(val(r, 3), f, B);
    }
    DL += val(r, 4); DR += val(r, 5);
  }
  mppi_apply_update(a, t, m, S, A, B, DL, DR);
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
