// mppi_kernels.cuh - the MPPI hot loop of controller::MPPI::newControls() as ONE sm_100a kernel per call.
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
//   * noise is counter-based (Philox4x32-10 keyed by seed, call, rollout, step pair): one Philox call feeds four binary32
//     Box-Muller variates (two steps x two wheels) built from correctly-rounded operations only, so the CPU oracle
//     reproduces them bit for bit.  The production variants LOAD them: mppi_noise_kernel draws a call's variates two calls
//     ahead, in the tail of an earlier call; the generic variant (taps, caller-supplied noise, ragged horizons) draws them
//     in the loop.
//   * the only mandatory HBM traffic is the fp32 [K][T][3] state tensor: each warp stages its rows in
//     shared memory and one lane hands them to the TMA unit (cp.async.bulk shared->global), so the next rollouts' math
//     overlaps the store.
//   * the T softmaxes are carried ONLINE, (min J, sum e, sum e*duL, sum e*duR) per lane and step in shared memory.
//     Production variants park the cost-to-go of three passes by step and update the sets once per batch - minimum first,
//     then every weight independently through the SFU in binary32 (the exponent difference is formed in fp64), straight-
//     line code; the generic variant updates rollout by rollout in fp64 behind a threshold test on the high word of J.
//     J never makes a round trip through HBM.
//   * THE CALL IS ONE LAUNCH.  The grid is the rollout CTAs followed by T MERGER CTAs.  Every rollout CTA merges its warps
//     into one [T][6] partial and sends it to the mergers as TAGGED WORDS (32 bytes = two doubles in four 8-byte units, each
//     4 bytes of payload next to the call's tag: a unit that shows the tag is its payload, so nothing has to be ordered -
//     no release / acquire pair, no fence; NCCL's LL protocol on L2).  The merger CTAs carry the highest block indices, so
//     they take the SM slots the first rollout CTAs leave; merger t merges step t's partials of all CTAs (minimum first,
//     every partial rescaled once, fixed order: bit-reproducible), exchanges the result with the other ranks over NVLink
//     peer memory when the job is sharded (same words, system scope), applies the update of mppi.cpp:112-137, leaves the
//     plan's step as a tagged word for the next call's CTAs and publishes the first control in mapped pinned host memory.
//     A call queued behind an uncollected call waits for those words and for the variates' count, not for the grids in
//     front of it (MppiArgs::skip_wait).  (A two-level "last arriver merges" tree inside the rollout CTAs was built first
//     and measured 3x slower: two fence + atomic + dependent-load rounds in series on ONE CTA against T CTAs side by side.)
#pragma once

#include "common.cuh"

namespace b2n
{

constexpr int kMppiMaxT = 256;               // 32 lanes x 8 steps
constexpr int kMppiMaxRanks = 64;
constexpr int kMppiXchgWords = 12;           // 6 doubles = 12 payload halves

struct MppiLL;

struct MppiArgs
{
  // model, cost, sampling
  double c_v, c_w;            // (r/2)(h/6) and (r/L) h: position and heading increment factors (mppi.hpp:45-47, rk4.cpp:114)
  double Q[3], R[2], P1[3];
  double inv_lambda, sigL, sigR;
  double cut_lambda;          // a rollout whose cost-to-go exceeds the running minimum by more than this has weight 0 in fp64
  double x0[3], xd[3];
  double cos0, sin0;          // of the start heading x0[2], from the host's libm
  // polynomial coefficients, passed as parameters so that DFMA takes them from the constant bank (as immediates every
  // use costs two uniform-register moves): sin / cos Taylor terms
  double ks3, ks5, ks7, ks9, kc2, kc4, kc6, kc8, kc10;
  int T, K, k_offset;
  int thr_on;                 // every cost weight is >= 0, so J >= 0 and the high words of J order like J (threshold test below)
  uint32_t call;
  uint32_t key0[10], key1[10];   // Philox round keys: seed lo / hi + round * Weyl constants
  int external_noise, capture, tma_store;
  // optional obstacle term (extension, see b2nav.h)
  int obs_on, obs_xsize, obs_ysize;
  double obs_xmin, obs_ymin, obs_xmax, obs_ymax, obs_res, obs_inv_res, obs_weight, obs_d0, obs_off;
  const float *obs_dist;
  int obs_ti0, obs_tj0;       // first cell of the tile staged in shared memory (kMppiObsTile square), -1: no tile
  // buffers
  const double *u_plan;    // [2][T]
  float *states;           // [K][T][3]
  const double *ext;       // [K][T][2] or null
  const float4 *zbuf;      // [K][T/2] the call's standard normals (zL_t, zR_t, zL_t+1, zR_t+1), drawn ahead by mppi_noise_kernel (FAST variants)
  double *J_out;           // [K][T] (capture)
  double *du_out;          // [K][T][2] (capture)
  double *partials;        // [T][gridDim.x][6]: a step's partials of all CTAs are contiguous for the CTA that merges them
  // ---- fused tail: merge tree, exchange, update (mppi.cpp:112-137) ----
  int tail;                // 0: stop after the CTA partials (NCCL transport, bench hook b2n_mppi_time_rollout)
  int n_roll;              // rollout CTAs (block indices below this); with tail: gridDim.x = n_roll + T, the rest are mergers
  // hand-over inside the fused call and from call to call WITHOUT fences or counters: doubles travel in pairs as 32-byte
  // words of four 8-byte units (4 bytes of payload, the call's tag).  8-byte units are single transactions, so a unit that
  // shows the tag IS its payload; the reader spins on the words it needs and nothing else is ordered (the LL protocol of
  // NCCL, on L2)
  MppiLL *ll_partials;     // [T][3][n_roll] this call's CTA partials, tagged `tag`: merger CTA t reads step t's
  const MppiLL *ll_plan;   // [T] the current plan (uL_t, uR_t) as tagged words, valid when plan_tag != 0 (the call before was a fused one)
  MppiLL *ll_plan_next;    // [T] the plan this call writes, tagged `tag`
  uint32_t tag, plan_tag;  // never 0
  // a hint, not a synchronisation: warps that have sent their partial words count themselves in with a relaxed reduction
  // (nothing waits for it), and a merger CTA watches this one word instead of spinning on n_roll x 96 bytes; the words
  // themselves are still checked by their tags when they are read
  // pipelined calls: the next call does not wait for this grid (and everything in front of it) to COMPLETE.  What it reads
  // from earlier grids it waits for by itself: the plan through its tagged words, the variates through the count of noise-
  // kernel CTAs that have finished (a release reduction each; monotonic over the handle's noise launches)
  int skip_wait;                     // 1: no griddepcontrol.wait (fused call behind a fused call, no device-resident input besides these two)
  const unsigned long long *z_ready;
  unsigned long long z_need;
  unsigned long long *arrive;        // monotonic over the handle's fused calls
  unsigned long long arrive_need;    // its value when this call's are all in: n_roll x ceil(T / 32) per fused call
  double k_total, umax;
  double uinit[2];
  double *u_next;          // [2][T]
  double *out;             // [2] first control of the updated plan (mapped pinned host memory)
  unsigned long long *out_seq;   // completion word next to it: set to `seq` after the controls are visible to the host
  unsigned long long seq;
  double *stepstats;       // [T][2] (min J, sum w) for the weights tap
  double *merged;          // [T][6] this rank's merged sums (tap)
  unsigned long long *dbg;       // [grid + T][kMppiDbgSlots] stage timestamps (globaltimer ns) of every CTA's thread 0, tuning runs only (null: off)
  // sharded job: peer-memory exchange areas [2 parities][nranks][T][3 tagged 32-byte words], see mppi_exchange_step()
  int rank, nranks, parity;
  uint32_t call_id;
  unsigned long long *peer[kMppiMaxRanks];
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

// minimum of two costs (never NaN): a compare and a select, where fmin spends half a dozen instructions on NaN rules
__device__ __forceinline__ double mppi_min(double a, double b) { return b < a ? b : a; }

// A load of the variates' buffer.  Through the read-only path (ld.global.nc) by default: it runs beside the load / store unit,
// which the shared-memory traffic of the loop keeps busy (measured on one box: rollout phase 12.62 us against 12.76 us with
// plain loads, a queued call 16.37 against 16.53 us).  The path is specified for data that nobody writes while the kernel
// lives; what makes it safe here: the buffer a call reads was filled two calls earlier by a kernel that has completed
// before this grid reads a single word of it (griddepcontrol.wait, or the acquire load of the finished-CTA count, both of
// which also drop what the SM's L1 holds), the kernels that run beside this one write the OTHER two buffers of the rotation,
// and no CTA touches the buffer before that wait.  -DB2N_MPPI_Z_PLAIN_LOADS selects ordinary loads.
__device__ __forceinline__ float4 mppi_z_load(const float4 *p)
{
#ifdef B2N_MPPI_Z_PLAIN_LOADS
  return *p;
#else
  return __ldg(p);
#endif
}

// tagged 32-byte words (see MppiArgs::ll_partials): two doubles, each 4-byte half next to the tag in its own 8-byte unit.
// One 256-bit store writes a whole 32-byte sector (a narrower store would leave a partially valid sector in L2, and the
// reader's load would then wait for the rest of it from DRAM)
struct __align__(32) MppiLL { unsigned long long w[4]; };

__device__ __forceinline__ void mppi_ll_store(MppiLL *dst, double v0, double v1, uint32_t tag)
{
  const unsigned long long tg = (unsigned long long)tag << 32;
  asm volatile("st.relaxed.gpu.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(dst), "l"(tg | (uint32_t)__double2loint(v0)),
               "l"(tg | (uint32_t)__double2hiint(v0)), "l"(tg | (uint32_t)__double2loint(v1)), "l"(tg | (uint32_t)__double2hiint(v1)) : "memory");
}
__device__ __forceinline__ bool mppi_ll_load(const MppiLL *src, uint32_t tag, double &v0, double &v1)
{
  unsigned long long a, b, c, d;
  asm volatile("ld.relaxed.gpu.global.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(src) : "memory");
  v0 = __hiloint2double((int)(uint32_t)b, (int)(uint32_t)a);
  v1 = __hiloint2double((int)(uint32_t)d, (int)(uint32_t)c);
  return ((uint32_t)(a >> 32) == tag) & ((uint32_t)(b >> 32) == tag) & ((uint32_t)(c >> 32) == tag) & ((uint32_t)(d >> 32) == tag);
}
__device__ __forceinline__ void mppi_ll_wait(const MppiLL *src, uint32_t tag, double &v0, double &v1)
{
  unsigned spins = 0;
  while (!mppi_ll_load(src, tag, v0, v1))
    if (++spins == (1u << 27)) __trap();            // about a minute without the word: fail loudly, not silently
}

constexpr int kMppiObsTile = 32;             // cells per side of the obstacle-field tile staged through TMA

// cell index of a coordinate exactly as the oracle computes it, floor((x - xmin) / res): the rare path of
// mppi_obstacle_cost, kept out of line (an fp64 division is some forty instructions)
__device__ __noinline__ int mppi_obstacle_cell_exact(double x, double xmin, double res) { return (int)floor((x - xmin) / res); }

// the obstacle term of one state (extension, b2nav.h): the distance comes from the tile in shared memory when the cell
// is inside it (every state of a horizon lies within a few cells of the start pose), from global memory otherwise
__device__ __forceinline__ double mppi_obstacle_cost(const MppiArgs &a, const float *tile, double x, double y)
{
  const bool inside = (x >= a.obs_xmin) & (x <= a.obs_xmax) & (y >= a.obs_ymin) & (y <= a.obs_ymax);
  // floor((x - xmin) / res) on the fp64 add / multiply pipe only (the division and the double -> integer conversions run
  // at a fraction of its rate): the product with 1 / res is within a few ulps of the quotient, so unless it lands within
  // 1e-6 of an integer both floors agree; the floor itself comes from the round-to-nearest of adding 1.5 * 2^52 (the
  // integer is then the sum's low word), one less when that rounded up
  const double kMagic = 6755399441055744.0;
  const double qx = (x - a.obs_xmin) * a.obs_inv_res, qy = (y - a.obs_ymin) * a.obs_inv_res;
  const double tx = qx + kMagic, ty = qy + kMagic;
  const double rx = tx - kMagic, ry = ty - kMagic;          // rint(qx), rint(qy)
  int ii = __double2loint(tx) - (rx > qx ? 1 : 0), jj = __double2loint(ty) - (ry > qy ? 1 : 0);
  const bool near = (fabs(qx - rx) < 1e-6) | (fabs(qy - ry) < 1e-6);     // (a product of at most a few thousand cells is off by < 1e-12)
  if (near && inside) {
    ii = mppi_obstacle_cell_exact(x, a.obs_xmin, a.obs_res);
    jj = mppi_obstacle_cell_exact(y, a.obs_ymin, a.obs_res);
  }
  ii = max(0, min(ii, a.obs_xsize - 1));                    // x == xmax lands in cell xsize: the reference steps back (grid_mapper.cpp:866-875)
  jj = max(0, min(jj, a.obs_ysize - 1));                    // (and an outside state reads a valid address; its value is not used)
  const unsigned ti = (unsigned)(ii - a.obs_ti0), tj = (unsigned)(jj - a.obs_tj0);
  float df;
  if (a.obs_ti0 >= 0 && ti < (unsigned)kMppiObsTile && tj < (unsigned)kMppiObsTile) df = tile[ti * kMppiObsTile + tj];
  else df = __ldg(&a.obs_dist[(size_t)ii * a.obs_ysize + jj]);       // row-major [xsize][ysize]
  const double pen = a.obs_d0 - (double)df;
  const double cost = pen > 0.0 ? a.obs_weight * pen * pen : 0.0;
  return inside ? cost : a.obs_off;
}

// exp(x) for x <= 0.  Below the cut the result cannot change a softmax sum that already holds the minimum's 1.0, so
// the evaluation is skipped (at the shipped lambda = 0.01 that is almost every term).
//   exact (generic kernel variant, update kernels): libdevice exp, cut at -708 (subnormal results dropped).
//   fast  (production variant): 2^x split into an integer and a fraction in [-1/2, 1/2]; the fraction goes through the
//          SFU (ex2.approx.f32, relative error 2^-22), the integer into the exponent field.  2.5e-7 relative on a softmax
//          term, against the 1e-5 the contract allows on weights and controls; cut at -700 so the exponent never
//          reaches the subnormal range.
template <bool FASTEXP>
__device__ __forceinline__ double mppi_exp_neg(double x)
{
  if (!FASTEXP) return x > -708.0 ? exp(x) : 0.0;
  if (!(x > -700.0)) return 0.0;
  const double t = x * 1.4426950408889634;                 // log2(e)
  const double tn = t + 6755399441055744.0;                // 1.5 * 2^52: the low word now holds rint(t)
  const int n = __double2loint(tn);
  const float f = (float)(t - (tn - 6755399441055744.0));  // in [-1/2, 1/2]
  float p;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(f));
  const double pd = (double)p;                             // in [0.707, 1.415]
  return __hiloint2double(__double2hiint(pd) + (n << 20), __double2loint(pd));
}

// sin and cos of a small angle by Taylor series.  Generic: |d| <= 1/8 with terms to d^9 / d^10 (truncation < 3e-16
// relative), full range falls back to sincos.  ALWAYS_SMALL (the host has proved |d| <= 1/16): terms to d^7 / d^6,
// truncation 4e-17 / 6e-15 absolute.
template <bool ALWAYS_SMALL>
__device__ __forceinline__ void mppi_sincos_small(const MppiArgs &a, double d, double &sn, double &cs)
{
  const double z = d * d;
  if (ALWAYS_SMALL) {
    double ps = fma(z, a.ks7, a.ks5);    // -1/7!, 1/5!
    ps = fma(z, ps, a.ks3);              // -1/3!
    sn = fma(d * z, ps, d);
    double pc = fma(z, a.kc6, a.kc4);    // -1/6!, 1/4!
    pc = fma(z, pc, a.kc2);              // -1/2
    cs = fma(z, pc, 1.0);
    return;
  }
  if (fabs(d) > 0.125) { sincos(d, &sn, &cs); return; }
  double ps = fma(z, a.ks9, a.ks7);      // 1/9!, -1/7!
  ps = fma(z, ps, a.ks5);
  ps = fma(z, ps, a.ks3);
  sn = fma(d * z, ps, d);
  double pc = fma(z, a.kc10, a.kc8);     // -1/10!, 1/8!
  pc = fma(z, pc, a.kc6);
  pc = fma(z, pc, a.kc4);
  pc = fma(z, pc, a.kc2);
  cs = fma(z, pc, 1.0);
}

// ---- segmented warp scans over the G lanes of a rollout -----------------------------------------------------------
// A lane outside a scan step (its partner would lie in another rollout's segment) keeps its value.  For the additive scans
// that is x = fma(other, m, x) with m = 1.0 or 0.0 built by ONE select on the high word: written as `if (g >= d) x += other`
// it compiles to DADD + two FSEL on the 32-bit halves + the moves that re-pair them (549 instructions per pass against 516).
// (The masks in a shared-memory table gave 513 instructions and a SLOWER loop: their loads sit on the scans' dependent chain.
// A version that took shfl.sync's "source lane in range" predicate and predicated the dependent fp64 arithmetic on it in
// inline PTX was built and measured too: ptxas turns the predicated fp64 operations back into FSEL pairs and the unpack /
// repack around the asm costs a hundred register moves per pass.)
// 1.0 where the condition holds, else 0.0: a select on the high word (x = fma(other, m, x) then is the guarded addition in
// one instruction, bit for bit: the product is exact and x + 0 = x)
__device__ __forceinline__ double mppi_mask(bool on) { return __hiloint2double(on ? 0x3FF00000 : 0, 0); }

template <int G>
__device__ __forceinline__ void scan_rot_up(double &tc, double &ts, double &tth, int d, int g)
{
  const double oc = __shfl_up_sync(kFullMask, tc, d, G), os = __shfl_up_sync(kFullMask, ts, d, G);
  const double ot = __shfl_up_sync(kFullMask, tth, d, G);
  tth = fma(ot, mppi_mask(g >= d), tth);
  if (g >= d) {
    const double nc = fma(tc, oc, -(ts * os));
    ts = fma(tc, os, ts * oc);
    tc = nc;
  }
}

template <int G>
__device__ __forceinline__ void scan_add2_up(double &x, double &y, int d, int g)
{
  const double ox = __shfl_up_sync(kFullMask, x, d, G), oy = __shfl_up_sync(kFullMask, y, d, G);
  const double m = mppi_mask(g >= d);
  x = fma(ox, m, x); y = fma(oy, m, y);
}

template <int G>
__device__ __forceinline__ void scan_add_down(double &x, int d, int g)
{
  const double ox = __shfl_down_sync(kFullMask, x, d, G);
  x = fma(ox, mppi_mask(g + d < G), x);
}

// value of the neighbouring lane inside the segment (delta 1), or `edge` at the segment boundary
template <int G, bool UP>
__device__ __forceinline__ double shift1(double v, double edge, int g)
{
  const double o = UP ? __shfl_up_sync(kFullMask, v, 1, G) : __shfl_down_sync(kFullMask, v, 1, G);
  return (UP ? g == 0 : g == G - 1) ? edge : o;
}

// ---- noise ----------------------------------------------------------------------------------------------------------
// normal_quad_f32 of common.cuh with the Philox round keys taken from the kernel parameters (constant bank) instead of
// being re-derived from the seed every pass
__device__ __forceinline__ void mppi_normal_quad(const MppiArgs &a, uint32_t stream, uint32_t index, float z[4])
{
  uint32_t c0 = index, c1 = stream, c2 = a.call, c3 = kDomainMppi;
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ a.key0[r], n2 = hi0 ^ c3 ^ a.key1[r];
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
  }
  box_muller_f32(c0, c1, z[0], z[1]);
  box_muller_f32(c2, c3, z[2], z[3]);
}

constexpr int kMppiDbgSlots = 24;
// SM clock of warp 0 at a point inside its second pass (slots 8 and up).  Compiled in only with -DB2N_MPPI_PHASE_CLOCKS (the
// tuning build behind the phase table of DESIGN.md 3.3): eleven predicated stores and clock reads per pass otherwise sit in
// the issue-bound loop of every call.
__device__ __forceinline__ void mppi_clock(const MppiArgs &a, bool first, int j)
{
#ifdef B2N_MPPI_PHASE_CLOCKS
  if (a.dbg && first && threadIdx.x == 0) a.dbg[(size_t)blockIdx.x * kMppiDbgSlots + j] = (unsigned long long)clock64();
#else
  (void)a; (void)first; (void)j;
#endif
}

__device__ __forceinline__ void mppi_stamp(const MppiArgs &a, int j)
{
  if (a.dbg && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t) :: "memory");      // (the clobber keeps it on its side of a barrier)
    unsigned long long *row = a.dbg + (size_t)blockIdx.x * kMppiDbgSlots;
    if (j == 0) row[11] = row[0];           // the previous call's loop start / update: the period of pipelined calls
    if (j == 6) row[12] = row[6];
    row[j] = t;
  }
}

// ---- merge of n accumulator sets per time step by one CTA ----------------------------------------------------------------
// Every set is (min J, sum e, sum e*duL, sum e*duR, sum duL, sum duR) with the three weighted sums relative to the set's
// own minimum.  Thread x handles step t = x mod TP2 and the sets q, q + Q, q + 2Q, ... (q = x / TP2, Q = NT / TP2, TP2 =
// the horizon rounded up to a power of two): minimum first, then every set rescaled ONCE (independent exponentials) and
// summed in index order, then the Q partial results of a step summed in q order - a fixed order, whatever the arrival
// order of the sets' producers.  `load(t, i, v)` fetches set i of step t, `load_min(t, i)` its minimum only.
// scratch: NT * 6 doubles of shared memory.  The result is valid in threads x < T (t = x).
template <int NT, bool FASTEXP, class LoadMin, class Load>
__device__ __forceinline__ void mppi_merge_sets(int T, int TP2, int n, double inv_lambda, double *scratch, LoadMin load_min, Load load, double out[6])
{
  constexpr int kU = 4;       // sets fetched per round trip: their loads are issued together (an L2 round trip each otherwise)
  const double inf = __longlong_as_double(0x7FF0000000000000LL);
  const int x = threadIdx.x;
  const int t = x & (TP2 - 1), q = x / TP2, Q = NT / TP2;
  double *smin = scratch, *ssum = scratch + NT;           // [Q][TP2], [Q][TP2][5]
  double m = inf;
  if (t < T) {
    for (int i0 = q; i0 < n; i0 += kU * Q) {
      double mv[kU];
#pragma unroll
      for (int u = 0; u < kU; u++) mv[u] = (i0 + u * Q < n) ? load_min(t, i0 + u * Q) : inf;
#pragma unroll
      for (int u = 0; u < kU; u++) m = mppi_min(m, mv[u]);
    }
  }
  smin[q * TP2 + t] = m;
  __syncthreads();
  m = smin[t];
  for (int qq = 1; qq < Q; qq++) m = mppi_min(m, smin[qq * TP2 + t]);
  double s[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  if (t < T) {
    for (int i0 = q; i0 < n; i0 += kU * Q) {
      double v[kU][6];
#pragma unroll
      for (int u = 0; u < kU; u++) {
        if (i0 + u * Q < n) load(t, i0 + u * Q, v[u]);
        else { v[u][0] = inf; v[u][4] = 0.0; v[u][5] = 0.0; }
      }
#pragma unroll
      for (int u = 0; u < kU; u++) {
        if (v[u][0] != inf) {      // an empty set has min = +inf and zero sums
          const double f = (v[u][0] == m) ? 1.0 : mppi_exp_neg<FASTEXP>((m - v[u][0]) * inv_lambda);
          s[0] = fma(v[u][1], f, s[0]); s[1] = fma(v[u][2], f, s[1]); s[2] = fma(v[u][3], f, s[2]);
        }
        s[3] += v[u][4]; s[4] += v[u][5];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 5; j++) ssum[(q * TP2 + t) * 5 + j] = s[j];
  __syncthreads();
  out[0] = m;
  if (q == 0) {
    for (int qq = 1; qq < Q; qq++) {
#pragma unroll
      for (int j = 0; j < 5; j++) s[j] += ssum[(qq * TP2 + t) * 5 + j];
    }
  }
#pragma unroll
  for (int j = 0; j < 5; j++) out[1 + j] = s[j];
  __syncthreads();        // the scratch area may be reused by the next merge
}

// the control update of mppi.cpp:112-137 for step t from the fully merged sums (one thread)
template <class A>
__device__ __forceinline__ double2 mppi_apply_update(const A &a, double ul_cur, double ur_cur, int t, double m, double S, double Aw, double Bw, double DL, double DR)
{
  const int T = a.T;
  // w_k = exp(-(J_k - min)/lambda) + 1e-8, normalised (mppi.cpp:117-118)
  const double sumw = S + a.k_total * 1e-8;
  const double inv = 1.0 / sumw;
  double nl = ul_cur + (Aw + 1e-8 * DL) * inv;              // mppi.cpp:120-121
  double nr = ur_cur + (Bw + 1e-8 * DR) * inv;
  nl = fmin(fmax(nl, -a.umax), a.umax);                     // mppi.cpp:124-125
  nr = fmin(fmax(nr, -a.umax), a.umax);
  if (t == 0) {                                             // mppi.cpp:129-131
    if (a.out_seq) {
      // the host polls these words instead of paying a stream synchronisation for 16 bytes.  Four 8-byte words, each a
      // 4-byte half of a control next to the low half of the call's sequence number: a word that shows the sequence number
      // IS its payload, so nothing has to be ordered and no system-scope fence sits on the call's critical path
      const unsigned long long tag = (a.seq & 0xFFFFFFFFull) << 32;
      unsigned long long *w = reinterpret_cast<unsigned long long *>(a.out);
      asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(w + 0), "l"(tag | (unsigned)__double2loint(nl)) : "memory");
      asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(w + 1), "l"(tag | (unsigned)__double2hiint(nl)) : "memory");
      asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(w + 2), "l"(tag | (unsigned)__double2loint(nr)) : "memory");
      asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(w + 3), "l"(tag | (unsigned)__double2hiint(nr)) : "memory");
    } else {
      a.out[0] = nl; a.out[1] = nr;
    }
  }
  else { a.u_next[t - 1] = nl; a.u_next[T + t - 1] = nr; }  // mppi.cpp:134
  if (t == T - 1) { a.u_next[T - 1] = a.uinit[0]; a.u_next[2 * T - 1] = a.uinit[1]; }   // mppi.cpp:136-137
  if (a.stepstats) {            // (min J, sum w) per step: only the weights tap reads them
    a.stepstats[2 * t] = m;
    a.stepstats[2 * t + 1] = sumw;
  }
  return make_double2(nl, nr);
}

// ---- sharded rollouts: the exchange over NVLink peer memory ---------------------------------------------
// Every rank owns an exchange area [2 call parities][nranks][T][3] of tagged 32-byte words (MppiLL); peer[j] is rank j's
// area mapped into this process (CUDA IPC).  Merger CTA t sends this rank's merged sums of step t to every rank (its own
// included: one code path) in the low-latency style of NCCL's LL protocol: each 8-byte unit carries 4 bytes of payload
// and the 32-bit call id, and 8-byte units are single NVLink transactions, so a unit whose upper half shows the current
// call id IS its payload - no fence, no separate flag, one NVLink write latency.  Lane r of the CTA's first warp then spins
// on rank r's three words in this rank's own area and the warp folds the nranks results with shuffles - a fixed tree, the
// same on every rank, so the plan stays replicated bit for bit without a broadcast.  No NCCL call, no extra launch.
// A slot of parity p is rewritten at call c + 2 only after its owner finished call c + 1, which needed this rank's
// data of call c + 1, which was sent after this rank finished reading call c: two parities are enough.
__device__ __forceinline__ void mppi_ll_store_sys(MppiLL *dst, double v0, double v1, uint32_t tag)
{
  const unsigned long long tg = (unsigned long long)tag << 32;
  asm volatile("st.relaxed.sys.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(dst), "l"(tg | (uint32_t)__double2loint(v0)),
               "l"(tg | (uint32_t)__double2hiint(v0)), "l"(tg | (uint32_t)__double2loint(v1)), "l"(tg | (uint32_t)__double2hiint(v1)) : "memory");
}
__device__ __forceinline__ bool mppi_ll_load_sys(const MppiLL *src, uint32_t tag, double &v0, double &v1)
{
  unsigned long long a, b, c, d;
  asm volatile("ld.relaxed.sys.global.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(src) : "memory");
  v0 = __hiloint2double((int)(uint32_t)b, (int)(uint32_t)a);
  v1 = __hiloint2double((int)(uint32_t)d, (int)(uint32_t)c);
  return ((uint32_t)(a >> 32) == tag) & ((uint32_t)(b >> 32) == tag) & ((uint32_t)(c >> 32) == tag) & ((uint32_t)(d >> 32) == tag);
}

template <int NT, bool FASTEXP>
__device__ __forceinline__ void mppi_exchange_step(const MppiArgs &a, int t, double &m, double &S, double &A, double &B, double &DL, double &DR)
{
  __shared__ double mine[6];
  const int T = a.T, par = a.parity, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mine[0] = m; mine[1] = S; mine[2] = A; mine[3] = B; mine[4] = DL; mine[5] = DR; }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * a.nranks; i += NT) {
    const int r = i / 3, j = i - 3 * r;
    // word j of this rank's slot in rank r's area
    MppiLL *dst = reinterpret_cast<MppiLL *>(a.peer[r]) + (((size_t)par * a.nranks + a.rank) * T + t) * 3 + j;
    mppi_ll_store_sys(dst, mine[2 * j], mine[2 * j + 1], a.call_id);
  }
  if (threadIdx.x >= 32) return;
  const double inf = __longlong_as_double(0x7FF0000000000000LL);
  m = inf; S = A = B = DL = DR = 0.0;
  for (int r0 = 0; r0 < a.nranks; r0 += 32) {
    const int r = r0 + lane;
    double v[6] = {inf, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (r < a.nranks) {
      const MppiLL *src = reinterpret_cast<const MppiLL *>(a.peer[a.rank]) + (((size_t)par * a.nranks + r) * T + t) * 3;
      unsigned polls = 0;
      for (;;) {
        bool ok = mppi_ll_load_sys(src, a.call_id, v[0], v[1]);
        ok &= mppi_ll_load_sys(src + 1, a.call_id, v[2], v[3]);
        ok &= mppi_ll_load_sys(src + 2, a.call_id, v[4], v[5]);
        if (ok) break;
        if (++polls == (1u << 27)) __trap();          // about a minute without the peer's words: fail loudly, not silently
      }
    }
    __syncwarp();
    // this group of (up to) 32 ranks into the running result: minimum first, every rank rescaled once
    const double mn = mppi_min(m, warp_min(v[0]));
    if (mn != inf) {
      const double f = (v[0] == inf) ? 0.0 : (v[0] == mn) ? 1.0 : mppi_exp_neg<FASTEXP>((mn - v[0]) * a.inv_lambda);
      const double fo = (m == inf) ? 0.0 : (m == mn) ? 1.0 : mppi_exp_neg<FASTEXP>((mn - m) * a.inv_lambda);
      S = fma(S, fo, warp_sum(v[1] * f)); A = fma(A, fo, warp_sum(v[2] * f)); B = fma(B, fo, warp_sum(v[3] * f));
    }
    DL += warp_sum(v[4]); DR += warp_sum(v[5]);
    m = mn;
  }
  mppi_stamp(a, 7);
}

// merge of one step's partials by one CTA of NT threads: minimum first, then every partial rescaled once (independent
// exponentials), plain sums in a fixed order.  partial p of step t at partials + p * p_stride + t * t_stride (doubles).
// The result is valid in thread 0.
// LL: the partials are the tagged words of MppiArgs::ll_partials ([T][6][n]); every thread spins on its own until they show
// the call's tag - the wait for the rollout CTAs and the load are one and the same L2 round trip.
template <int NT, bool FASTEXP, bool LL = false>
__device__ __forceinline__ void mppi_block_merge(const double *partials, int n_partials, int p_stride, int t_stride, int t, double inv_lambda, double &m,
                                                 double &S, double &A, double &B, double &DL, double &DR, const MppiLL *ll = nullptr, uint32_t tag = 0,
                                                 unsigned long long *stamp = nullptr)
{
  constexpr int kPer = 2;                                   // partials held in registers per thread (one L2 round trip)
  constexpr int NWARP = NT / 32;
  __shared__ double red[NWARP][6];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double inf = __longlong_as_double(0x7FF0000000000000LL);
  const double *base = partials + (size_t)t * t_stride;
  double2 c0[kPer], c1[kPer], c2[kPer];
  m = inf;
#pragma unroll
  for (int i = 0; i < kPer; i++) {
    const int p = threadIdx.x + i * NT;
    c0[i] = make_double2(inf, 0.0); c1[i] = make_double2(0.0, 0.0); c2[i] = make_double2(0.0, 0.0);
    if (!LL && p < n_partials) {
      const double2 *c = reinterpret_cast<const double2 *>(base + (size_t)p * p_stride);
      c0[i] = __ldcg(c); c1[i] = __ldcg(c + 1); c2[i] = __ldcg(c + 2);
    }
  }
  const MppiLL *lbase = ll + (size_t)t * 3 * n_partials;    // word j of partial p at lbase[j * n + p]: a warp's loads are contiguous
  if (LL) {
    // all words of the thread's partials in flight together, again until every one of them shows the tag
    unsigned spins = 0;
    for (;;) {
      bool ok = true;
#pragma unroll
      for (int i = 0; i < kPer; i++) {
        const int p = threadIdx.x + i * NT;
        if (p < n_partials) {
          const MppiLL *w = lbase + p;
          ok &= mppi_ll_load(w, tag, c0[i].x, c0[i].y);
          ok &= mppi_ll_load(w + n_partials, tag, c1[i].x, c1[i].y);
          ok &= mppi_ll_load(w + 2 * (size_t)n_partials, tag, c2[i].x, c2[i].y);
        }
      }
      if (ok) break;
      if (++spins == (1u << 27)) __trap();
    }
  }
#pragma unroll
  for (int i = 0; i < kPer; i++) m = mppi_min(m, c0[i].x);
  for (int p = threadIdx.x + kPer * NT; p < n_partials; p += NT) {
    double v0, v1;
    if (LL) mppi_ll_wait(lbase + p, tag, v0, v1);
    else v0 = __ldcg(base + (size_t)p * p_stride);
    m = mppi_min(m, v0);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) m = mppi_min(m, __shfl_xor_sync(kFullMask, m, d));
  if (lane == 0) red[warp][0] = m;
  __syncthreads();
  m = red[0][0];
#pragma unroll
  for (int w = 1; w < NWARP; w++) m = mppi_min(m, red[w][0]);
  if (stamp && threadIdx.x == 0) {            // tuning runs: every thread's partials are in
    unsigned long long now;       // (the operand ties the read to data from behind the barrier: BAR.SYNC.DEFER_BLOCKING lets a warp run on)
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now) : "d"(m) : "memory");
    *stamp = now;
  }
  S = 0.0; A = 0.0; B = 0.0; DL = 0.0; DR = 0.0;
#pragma unroll
  for (int i = 0; i < kPer; i++) {
    if (c0[i].x != inf) {
      const double f = (c0[i].x == m) ? 1.0 : mppi_exp_neg<FASTEXP>((m - c0[i].x) * inv_lambda);
      S = fma(c0[i].y, f, S); A = fma(c1[i].x, f, A); B = fma(c1[i].y, f, B);
    }
    DL += c2[i].x; DR += c2[i].y;
  }
  for (int p = threadIdx.x + kPer * NT; p < n_partials; p += NT) {
    double2 d0, d1, d2;
    if (LL) {
      const MppiLL *w = lbase + p;
      mppi_ll_wait(w, tag, d0.x, d0.y);
      mppi_ll_wait(w + n_partials, tag, d1.x, d1.y);
      mppi_ll_wait(w + 2 * (size_t)n_partials, tag, d2.x, d2.y);
    } else {
      const double2 *c = reinterpret_cast<const double2 *>(base + (size_t)p * p_stride);
      d0 = __ldcg(c); d1 = __ldcg(c + 1); d2 = __ldcg(c + 2);
    }
    if (d0.x != inf) {
      const double f = (d0.x == m) ? 1.0 : mppi_exp_neg<FASTEXP>((m - d0.x) * inv_lambda);
      S = fma(d0.y, f, S); A = fma(d1.x, f, A); B = fma(d1.y, f, B);
    }
    DL += d2.x; DR += d2.y;
  }
  S = warp_sum(S); A = warp_sum(A); B = warp_sum(B); DL = warp_sum(DL); DR = warp_sum(DR);
  // (columns 1..5: the minima in column 0 may still be read by slower warps)
  if (lane == 0) { red[warp][1] = S; red[warp][2] = A; red[warp][3] = B; red[warp][4] = DL; red[warp][5] = DR; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < NWARP; w++) {
      S += red[w][1]; A += red[w][2]; B += red[w][3]; DL += red[w][4]; DR += red[w][5];
    }
  }
}

// a merger CTA of the fused call: step t = blockIdx.x - n_roll
template <int NT, bool FASTEXP>
__device__ __forceinline__ void mppi_merger(const MppiArgs &a)
{
  const int t = (int)blockIdx.x - a.n_roll;
  // this step of the current plan: fetched now, used after the merge (the plan does not change during the call; a merger
  // can be resident before the previous call has written it: the tagged word says when it is there, as for the rollout CTAs)
  double ul_cur = 0.0, ur_cur = 0.0;
  if (threadIdx.x == 0) {
    if (a.plan_tag) mppi_ll_wait(a.ll_plan + t, a.plan_tag, ul_cur, ur_cur);
    else { ul_cur = a.u_plan[t]; ur_cur = a.u_plan[a.T + t]; }
  }
  // the next grid in the stream (the kernel that draws the next call's variates) may take the SMs the rollout CTAs leave
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  mppi_stamp(a, 4);
  if (threadIdx.x == 0) {
    unsigned long long seen;
    unsigned spins = 0;
    for (;;) {
      asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(a.arrive) : "memory");
      if (seen >= a.arrive_need) break;
      if (++spins == (1u << 28)) __trap();
    }
  }
  __syncthreads();
  // every rollout CTA's partial of this step (sent, by the count; each word is checked by its tag)
  double m, S, A, B, DL, DR;
  mppi_block_merge<NT, FASTEXP, true>(nullptr, a.n_roll, 0, 0, t, a.inv_lambda, m, S, A, B, DL, DR, a.ll_partials, a.tag,
                                      a.dbg ? a.dbg + (size_t)blockIdx.x * kMppiDbgSlots + 3 : nullptr);
  mppi_stamp(a, 5);
  if (a.nranks > 1) mppi_exchange_step<NT, FASTEXP>(a, t, m, S, A, B, DL, DR);
  if (threadIdx.x != 0) return;
  const double2 u = mppi_apply_update(a, ul_cur, ur_cur, t, m, S, A, B, DL, DR);
  // this step of the plan for the next call's CTAs (one slot to the left, as u_next): tagged words, nothing to order
  const int T = a.T;
  if (t > 0) mppi_ll_store(a.ll_plan_next + t - 1, u.x, u.y, a.tag);
  if (t == T - 1) mppi_ll_store(a.ll_plan_next + T - 1, a.uinit[0], a.uinit[1], a.tag);
  mppi_stamp(a, 6);
}

// dynamic shared memory of the rollout kernel:
//   [warps][1 or 2][32*S*3] fp32 staging rows of the state tensor (after the loop: merge scratch, [threads*6] doubles)
//   [S*4][threads]     online-softmax accumulators (min J, sum e, sum e*duL, sum e*duR) per owned step: fp64 (generic), or fp64 min +
//                      three binary32 sums (production)
//   [S/2][threads]     float4 sums of the binary32 variates (sum duL, sum duR before the scaling by sigma)
//   [2][G*S]           the control plan
//   [warps][3][32*S]   fp64 cost-to-go of the last passes' rollouts, by step: the softmax accumulators are updated in
//                      batches (production variants)
//   [S*2][threads]     fp64 sum duL / sum duR (generic variant only)
//   [tile][tile]       fp32 obstacle-field tile (obstacle variants only)
// accumulator set of one step in shared memory: generic (min J, sum e, sum e*duL, sum e*duR) fp64; production min J fp64 and the
// three sums binary32 (a lane's set holds a few rollouts: 1e-7 relative, merged in fp64 from the CTA level on)
__host__ __device__ constexpr size_t mppi_acc_bytes(bool fast) { return fast ? 8 + 3 * 4 : 4 * 8; }
constexpr int kMppiBatch = 3;        // passes whose cost-to-go is buffered before the softmax accumulators are touched (production variants)
__host__ __device__ constexpr size_t mppi_dz_bytes(int S) { return (size_t)S * 8; }
// staging buffers per warp: ONE (a pass is microseconds, the TMA unit drains a row in a fraction of that: the wait before the
// next pass's first store is over by then) - except where one buffer would be smaller than the merge scratch that reuses it
__host__ __device__ constexpr int mppi_stage_buffers(int S) { return S >= 4 ? 1 : 2; }
__host__ __device__ constexpr size_t mppi_rollout_smem(int S, int G, int NW, bool obs, bool fast)
{
  return (size_t)NW * mppi_stage_buffers(S) * 32 * S * 3 * sizeof(float) + (size_t)S * NW * 32 * mppi_acc_bytes(fast) + (size_t)NW * 32 * mppi_dz_bytes(S) +
         (size_t)2 * G * S * sizeof(double) + (fast ? (size_t)kMppiBatch * NW * 32 * S * sizeof(double) : (size_t)S * 2 * NW * 32 * sizeof(double)) +
         (obs ? (size_t)kMppiObsTile * kMppiObsTile * sizeof(float) + 16 : 0);
}
// CTAs per SM the register budget is cut for: 8-warp CTAs run three to an SM (80 registers), 10-warp CTAs two and 20-warp CTAs
// one (20 warps per SM, 96 registers: no spills in the rollout loop)
__host__ __device__ constexpr int mppi_min_blocks(int S, int NW) { return NW >= 16 ? 1 : (NW >= 10 ? 2 : (S >= 8 ? 1 : 3)); }

// FAST = the production configuration, decided on the host: own noise DRAWN AHEAD by mppi_noise_kernel, no capture taps,
// T == G * S (no partially filled lanes), TMA row stores, and half-step heading increments provably inside the Taylor range.  OBS adds the obstacle term
// (the generic variant takes it from a.obs_on).  NW = warps per CTA.
template <int S, int G, bool FAST, bool OBS, int NW>
__global__ void __launch_bounds__(NW * 32, mppi_min_blocks(S, NW)) mppi_rollout_kernel(const __grid_constant__ MppiArgs a)
{
  constexpr int NT = NW * 32;
  static_assert(S % 2 == 0 && (G == 8 || G == 16 || G == 32), "a Philox call covers two steps; G lanes per rollout");
  if (a.tail && (int)blockIdx.x >= a.n_roll) {
    mppi_merger<NT, FAST>(a);
    return;
  }
  constexpr int R = 32 / G;        // rollouts a warp carries at a time
  constexpr int TP = G * S;        // padded horizon
  constexpr int NBUF = mppi_stage_buffers(S);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr size_t kStageBytes = (size_t)NW * NBUF * 32 * S * 3 * 4, kAccBytes = (size_t)S * NT * mppi_acc_bytes(FAST), kDzBytes = (size_t)NT * mppi_dz_bytes(S),
                   kPlanBytes = (size_t)2 * TP * 8, kDaccBytes = FAST ? (size_t)kMppiBatch * NT * S * 8 : (size_t)S * 2 * NT * 8;
  static_assert((size_t)NT * 48 <= kStageBytes, "the merge scratch lives in the staging rows");
  float *stage_base = reinterpret_cast<float *>(smem_raw);                                       // [warps][2][R*TP*3]
  double *scratch = reinterpret_cast<double *>(smem_raw);                                        // [NT*6], after the loop
  double *acc_base = reinterpret_cast<double *>(smem_raw + kStageBytes);                         // [S*4][threads]
  double *acc = acc_base + threadIdx.x;
  float4 *dz4 = reinterpret_cast<float4 *>(smem_raw + kStageBytes + kAccBytes);                  // [S/2][threads]
  double *plan = reinterpret_cast<double *>(smem_raw + kStageBytes + kAccBytes + kDzBytes);      // [2][TP]
  double *dacc = reinterpret_cast<double *>(smem_raw + kStageBytes + kAccBytes + kDzBytes + kPlanBytes);     // [S*2][threads], generic only
  double *jbuf = dacc;                                                                                      // [warps][kMppiBatch][R][TP], production only
  float *tile = reinterpret_cast<float *>(smem_raw + kStageBytes + kAccBytes + kDzBytes + kPlanBytes + kDaccBytes);
  uint64_t *tile_bar = reinterpret_cast<uint64_t *>(tile + kMppiObsTile * kMppiObsTile);

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  mppi_stamp(a, 13);               // (tuning runs) the CTA's first instruction
  const int g = lane & (G - 1);    // position inside the rollout's lane group
  const int r = lane / G;          // which of the warp's R rollouts
  // warp-major numbering of the grid's warps: the last, partly filled round of passes then spreads over ALL CTAs (the first
  // warps of each) instead of filling the first CTAs - two CTAs share an SM, and an SM holding two full ones sets the pace
  const int gw = warp * a.n_roll + (int)blockIdx.x;
  const int nw = a.n_roll * NW;
  const int T = FAST ? TP : a.T;
  const int t0 = g * S;            // first owned step
  const bool ext_noise = !FAST && a.external_noise, capture = !FAST && a.capture, obs_on = OBS || (!FAST && a.obs_on);
  const bool tma_store = FAST || a.tma_store;
  float *stage = stage_base + warp * NBUF * (R * TP * 3);

  const double sin0 = a.sin0, cos0 = a.cos0;
  const double inf = __longlong_as_double(0x7FF0000000000000LL);

  // online-softmax accumulators of this lane's steps: (min J, sum e, sum e*duL, sum e*duR) x S in shared memory, with
  // the high word of (min J + cut) in a register: a cost-to-go above it cannot contribute
  float *accf = reinterpret_cast<float *>(acc_base + (size_t)S * NT) + threadIdx.x;      // production: [S*3][threads] behind the [S] minima
#pragma unroll
  for (int s = 0; s < S; s++) {
    if (FAST) {
      acc[s * NT] = inf;
#pragma unroll
      for (int j = 0; j < 3; j++) accf[(s * 3 + j) * NT] = 0.f;
    } else {
      acc[(s * 4 + 0) * NT] = inf;
#pragma unroll
      for (int j = 1; j < 4; j++) acc[(s * 4 + j) * NT] = 0.0;
    }
  }
  int thr_hi[S];
  // sum duL, sum duR take part in the update only through the +1e-8 weight floor (mppi.cpp:117): the production variant
  // sums the binary32 variates themselves in shared memory (one 16-byte load / store per two steps; scaled by sigma once,
  // at the end), the generic one the fp64 perturbations
  double sDL[FAST ? 1 : S], sDR[FAST ? 1 : S];
#pragma unroll
  for (int s = 0; s < S; s++) {
    thr_hi[s] = 0x7FF00000;
    if (!FAST) { sDL[s] = 0.0; sDR[s] = 0.0; }
  }
#pragma unroll
  for (int s = 0; s < S; s += 2) dz4[(s / 2) * NT + threadIdx.x] = make_float4(0.f, 0.f, 0.f, 0.f);

  // the variates of a call do not depend on the plan.  Production: mppi_noise_kernel drew them AHEAD (right behind the
  // previous call, off this call's critical path) and the loop loads them, 16 bytes per two steps.  Otherwise they are drawn
  // in the loop, the first pass's before the wait below so that it overlaps the previous call's tail
  constexpr bool z_ahead = FAST;
  float zq[S * 2];
  int base = gw * R;
  if (!ext_noise && !z_ahead && base < a.K) {
#pragma unroll
    for (int s = 0; s < S; s += 2) mppi_normal_quad(a, (uint32_t)(a.k_offset + base + r), (uint32_t)((t0 + s) >> 1), &zq[2 * s]);
  }

  // launched programmatically dependent on the previous call: everything above overlapped its tail; the plan it writes
  // is read from here on
  if (!a.skip_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
  else if (z_ahead) {
    // not the grids in front, only what this call reads from them: the variates (drawn two calls ahead: long there) and,
    // below, the plan.  One thread watches the count (a single word polled by every warp of every CTA would starve the
    // reductions that advance it); its acquire load and the barrier order the CTA's ordinary loads behind the noise kernel
    if (threadIdx.x == 0) {
      unsigned long long seen;
      unsigned spins = 0;
      for (;;) {
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(a.z_ready) : "memory");
        if (seen >= a.z_need) break;
        if (++spins == (1u << 26)) __trap();
        __nanosleep(200);
      }
    }
    __syncthreads();
  }
  // the variates of the first pass: in flight while this CTA waits for the plan
  float4 zn[FAST ? S / 2 : 1];
  if (z_ahead && base < a.K) {
    const float4 *zr = a.zbuf + ((size_t)min(base + r, a.K - 1) * (TP / 2) + (t0 >> 1));
#pragma unroll
    for (int s = 0; s < S; s += 2) zn[s / 2] = mppi_z_load(zr + s / 2);
  }
  // the obstacle-field tile around the start pose (occupancy-grid tiles staged through TMA, north_star): rows of the tile
  // are contiguous in the field, one bulk copy each, all completing on one mbarrier.  After the wait: the field may have
  // been written by the kernel that precedes this one in the stream (b2n_pf_write_distance_field)
  if (obs_on && a.obs_ti0 >= 0 && threadIdx.x == 0) {
    mbar_init(tile_bar, 1);
    fence_mbar_init();
    mbar_expect_tx(tile_bar, kMppiObsTile * kMppiObsTile * sizeof(float));
    for (int i = 0; i < kMppiObsTile; i++)
      tma_load_1d(tile + i * kMppiObsTile, a.obs_dist + (size_t)(a.obs_ti0 + i) * a.obs_ysize + a.obs_tj0, kMppiObsTile * sizeof(float), tile_bar);
  }
  // the plan.  Behind a fused call: its tagged words, each thread spinning on its own until the merger CTA of that step has
  // written it (the grid in front of this one may be the noise kernel; the call in front of THAT is normally long done).
  // Otherwise the plain array, ordered by the stream
  for (int t = threadIdx.x; t < TP; t += NT) {
    double ul = 0.0, ur = 0.0;
    if (t < T) {
      if (a.plan_tag) mppi_ll_wait(a.ll_plan + t, a.plan_tag, ul, ur);
      else { ul = a.u_plan[t]; ur = a.u_plan[T + t]; }
    }
    plan[t] = ul; plan[TP + t] = ur;
  }
  __syncthreads();
  if (obs_on && a.obs_ti0 >= 0) mbar_wait(tile_bar, 0);

  mppi_stamp(a, 0);
  // ---- production variants: the T softmaxes in BATCHES -------------------------------------------------------------------
  // A lane's accumulator set sees only K / (warps x R) rollouts per call (2.8 at C2): updated rollout by rollout, most
  // updates are "a new minimum of a nearly empty set" - a serialised, branchy block per owned step and pass, with the
  // warp walking through all of them because some lane always needs each.  Instead the cost-to-go of kMppiBatch passes is
  // parked in shared memory BY STEP (which also transposes it: lane g simulated steps gS .. gS + S - 1 but keeps the
  // accumulators of steps g, G + g, 2G + g, ...), and the set is then updated once per batch: minimum over the old set and
  // the candidates first, every exponential independent of the others, straight-line code for all lanes.  The variates of
  // the candidates are read again from the buffer the noise kernel filled (L1 / L2 hits).
  int nb = 0;
  int kb[kMppiBatch];
#pragma unroll
  for (int p = 0; p < kMppiBatch; p++) kb[p] = 0;
  auto flush = [&](int np) {
    __syncwarp();
    const double *jb = jbuf + (size_t)warp * kMppiBatch * R * TP + (size_t)r * TP;
    const double k2 = a.inv_lambda * 1.4426950408889634;       // weights as 2^((M - J) k2)
#pragma unroll
    for (int s = 0; s < S; s++) {
      const int ts = s * G + g;                       // the step this accumulator set belongs to
      const double m0 = acc[s * NT];
      const float S0 = accf[(s * 3) * NT], A0 = accf[(s * 3 + 1) * NT], B0 = accf[(s * 3 + 2) * NT];
      double Jc[kMppiBatch];
      float2 zc[kMppiBatch];
#pragma unroll
      for (int p = 0; p < kMppiBatch; p++) {
        const int kk = kb[p] + r;
        const bool on = p < np && kk < a.K;
        Jc[p] = on ? jb[(size_t)p * R * TP + ts] : inf;
        const float4 q = mppi_z_load(a.zbuf + ((size_t)(on ? kk : 0) * (TP / 2) + (ts >> 1)));
        zc[p] = (g & 1) ? make_float2(q.z, q.w) : make_float2(q.x, q.y);
      }
      double M = m0;
#pragma unroll
      for (int p = 0; p < kMppiBatch; p++) M = Jc[p] < M ? Jc[p] : M;      // (no NaN among costs: a plain compare, not fmin's six instructions)
      // weight of a member = 2^((M - J) k2) through the SFU in binary32: the exponent difference is formed in fp64 (that is
      // where the conditioning sits), the weight itself needs the 1e-5 of the contract, not 1e-16.  A member that is not
      // there (J = +inf) gets 2^-inf = 0, weights below 2^-126 flush to zero: no branch anywhere.  (All absent: M = +inf
      // would give inf - inf; the reference point is then irrelevant.)
      const double Mr = (M == inf) ? 0.0 : M;
      float f0;
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(f0) : "f"((float)((Mr - m0) * k2)));
      float Sn = S0 * f0, An = A0 * f0, Bn = B0 * f0;
#pragma unroll
      for (int p = 0; p < kMppiBatch; p++) {
        float e;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"((float)((Mr - Jc[p]) * k2)));
        Sn += e;
        An = fmaf(e, zc[p].x, An);
        Bn = fmaf(e, zc[p].y, Bn);
      }
      acc[s * NT] = M;
      accf[(s * 3) * NT] = Sn; accf[(s * 3 + 1) * NT] = An; accf[(s * 3 + 2) * NT] = Bn;
    }
    __syncwarp();
  };
  int buf = 0;
  bool have_noise = true;
  for (; base < a.K; base += nw * R, buf ^= (NBUF - 1)) {
    const int k = base + r;
    const bool live = k < a.K;
    const bool first = base == gw * R + nw * R;      // (the SECOND pass: thresholds set, steady state)
    mppi_clock(a, first, 8);

    // ---- perturbed controls (mppi.cpp:84-93,173-184) and the lane's S steps integrated in the lane's OWN frame -----
    // (heading 0 and position 0 at the lane's first step; the scans below place the lane in the world).  RK4 with the
    // control held over the step (rk4.cpp:95-115): k1 at the step's heading, k2 = k3 at heading + h w / 2, k4 at
    // heading + h w, so with d = h w / 2 the bracket (k1 + 2 k2 + 2 k3 + k4) is R (1 + 4 e^{id} + e^{2id}) =
    // R e^{id} (4 + 2 cos d): the mid-step heading, scaled.  sin d, cos d come from a short Taylor series.
    if (z_ahead) {
      // (loaded one pass ahead, below: an L2 round trip is most of a thousand cycles)
#pragma unroll
      for (int s = 0; s < S; s += 2) { zq[2 * s] = zn[s / 2].x; zq[2 * s + 1] = zn[s / 2].y; zq[2 * s + 2] = zn[s / 2].z; zq[2 * s + 3] = zn[s / 2].w; }
    } else if (!have_noise && !ext_noise) {
#pragma unroll
      for (int s = 0; s < S; s += 2) mppi_normal_quad(a, (uint32_t)(a.k_offset + k), (uint32_t)((t0 + s) >> 1), &zq[2 * s]);
    }
    have_noise = false;
    mppi_clock(a, first, 9);
    if (FAST && live) {
#pragma unroll
      for (int s = 0; s < S; s += 2) {
        float4 d = dz4[(s / 2) * NT + threadIdx.x];
        d.x += zq[2 * s]; d.y += zq[2 * s + 1]; d.z += zq[2 * s + 2]; d.w += zq[2 * s + 3];
        dz4[(s / 2) * NT + threadIdx.x] = d;
      }
    }
    mppi_clock(a, first, 10);
    double duL[FAST ? 1 : S], duR[FAST ? 1 : S];   // the production variant re-derives the perturbation from its variate where needed
    double cc[S], lx[S], ly[S], thc[S];
    double tc = 1.0, ts = 0.0, tth = 0.0;      // rotation and heading change of the lane so far
    double ax = 0.0, ay = 0.0;                 // displacement in the lane's frame so far
#pragma unroll
    for (int s = 0; s < S; s++) {
      const int t = t0 + s;
      const bool act = FAST || t < T;
      double l = 0.0, rr = 0.0;
      const double p0 = plan[t], p1 = plan[TP + t];
      if (act) {
        if (ext_noise) {
          if (live) {
            const double2 e = __ldg(reinterpret_cast<const double2 *>(a.ext) + ((size_t)k * T + t));
            l = e.x; rr = e.y;
          }
        } else {
          l = (double)zq[2 * s] * a.sigL;
          rr = (double)zq[2 * s + 1] * a.sigR;
        }
      }
      if (!FAST) { duL[s] = l; duR[s] = rr; }
      const double uL = p0 + l, uR = p1 + rr;                 // NOT clamped (mppi.cpp:93)
      cc[s] = (uL * a.R[0]) * uL + (uR * a.R[1]) * uR;        // u^T R u (mppi.hpp:92)
      const double vh = a.c_v * (uL + uR);                    // (h/6) v, v = (r/2)(uL + uR)   (mppi.hpp:45-46)
      const double dth = a.c_w * (uR - uL);                   // h w, w = (r/L)(uR - uL): rk4.cpp:114 with k1 = k2 = k3 = k4
      double sd, cd;
      mppi_sincos_small<FAST>(a, 0.5 * dth, sd, cd);
      const double mc = fma(tc, cd, -(ts * sd)), ms = fma(tc, sd, ts * cd);      // mid-step heading
      const double f = vh * fma(2.0, cd, 4.0);
      ax = fma(f, mc, ax);
      ay = fma(f, ms, ay);
      lx[s] = ax; ly[s] = ay;
      tc = fma(mc, cd, -(ms * sd)); ts = fma(mc, sd, ms * cd);                   // end-of-step heading
      tth += dth;
      thc[s] = tth;
    }

    mppi_clock(a, first, 11);
    // ---- segmented inclusive scans over the rollout's G lanes: rotation product and heading sum ----------
#pragma unroll
    for (int d = 1; d < G; d <<= 1) scan_rot_up<G>(tc, ts, tth, d, g);
    // exclusive prefix, seeded with the start heading: the world heading at the lane's first step
    double rc = shift1<G, true>(tc, 1.0, g), rs = shift1<G, true>(ts, 0.0, g);
    const double th = shift1<G, true>(tth, 0.0, g) + a.x0[2];
    {
      const double nc = fma(rc, cos0, -(rs * sin0));
      rs = fma(rc, sin0, rs * cos0);
      rc = nc;
    }

    mppi_clock(a, first, 12);
    // ---- position: the lane's displacement turned into the world frame, segmented prefix sums ----------------------
    double ix = fma(rc, ax, -(rs * ay)), iy = fma(rs, ax, rc * ay);
#pragma unroll
    for (int d = 1; d < G; d <<= 1) scan_add2_up<G>(ix, iy, d, g);
    const double ex = shift1<G, true>(ix, 0.0, g) + a.x0[0], ey = shift1<G, true>(iy, 0.0, g) + a.x0[1];

    mppi_clock(a, first, 13);
    // ---- states after each step -> staging row; loss (mppi.cpp:99-105, mppi.hpp:87-105) -------------------
    float *row = stage + buf * (R * TP * 3) + r * (T * 3);
    if (tma_store) {
      if (lane == 0) tma_store_wait_read<NBUF - 1>();   // the store that last used this buffer has drained it
      __syncwarp();
    }
    mppi_clock(a, first, 14);
    double J[S];
    double rj = 0.0;
    float st[S * 3];
#pragma unroll
    for (int s = S - 1; s >= 0; s--) {
      const int t = t0 + s;
      const double X = fma(rc, lx[s], fma(-rs, ly[s], ex)), Y = fma(rs, lx[s], fma(rc, ly[s], ey));
      const double TH = th + thc[s];
      double l = 0.0;
      st[s * 3 + 0] = (float)X; st[s * 3 + 1] = (float)Y; st[s * 3 + 2] = (float)TH;
      if (FAST || t < T) {
        const double e0 = X - a.xd[0], e1 = Y - a.xd[1], e2 = TH - a.xd[2];   // theta NOT wrapped
        if (t < T - 1) l = ((e0 * a.Q[0]) * e0 + (e1 * a.Q[1]) * e1 + (e2 * a.Q[2]) * e2) + cc[s];
        else l = (e0 * a.P1[0]) * e0 + (e1 * a.P1[1]) * e1 + (e2 * a.P1[2]) * e2;   // replaces the running loss
        if (obs_on) l += mppi_obstacle_cost(a, tile, X, Y);
      }
      rj += l;
      J[s] = rj;                                  // suffix sum inside the lane (cumSumCost, mppi.cpp:15-25)
    }
    if (FAST) {
      // the lane's S states are 12 S contiguous bytes of the row: vector stores
      if ((S * 3) % 4 == 0) {
        float4 *dst = reinterpret_cast<float4 *>(row + t0 * 3);
#pragma unroll
        for (int i = 0; i < S * 3 / 4; i++) dst[i] = make_float4(st[4 * i], st[4 * i + 1], st[4 * i + 2], st[4 * i + 3]);
      } else {
        float2 *dst = reinterpret_cast<float2 *>(row + t0 * 3);
#pragma unroll
        for (int i = 0; i < S * 3 / 2; i++) dst[i] = make_float2(st[2 * i], st[2 * i + 1]);
      }
    } else {
#pragma unroll
      for (int s = 0; s < S; s++)
        if (t0 + s < T) { row[(t0 + s) * 3 + 0] = st[s * 3]; row[(t0 + s) * 3 + 1] = st[s * 3 + 1]; row[(t0 + s) * 3 + 2] = st[s * 3 + 2]; }
    }

    mppi_clock(a, first, 15);
    if (z_ahead && base + nw * R < a.K) {
      // the next pass's variates: in flight during the scan and the softmax below
      const float4 *zr = a.zbuf + ((size_t)min(base + nw * R + r, a.K - 1) * (TP / 2) + (t0 >> 1));
#pragma unroll
      for (int s = 0; s < S; s += 2) zn[s / 2] = mppi_z_load(zr + s / 2);
    }
    // ---- cost-to-go: segmented suffix sum over the lanes ---------------------------------------------------
    double ij = rj;
#pragma unroll
    for (int d = 1; d < G; d <<= 1) scan_add_down<G>(ij, d, g);
    const double ej = shift1<G, false>(ij, 0.0, g);

    mppi_clock(a, first, 16);
    // ---- the T softmaxes over rollouts (mppi.cpp:112-121), carried online -------------------------------------------------
    if (FAST) {
      // park this pass's cost-to-go by step; the accumulators are updated once per kMppiBatch passes (flush above)
      double *jb = jbuf + ((size_t)(warp * kMppiBatch + nb) * R + r) * TP;
#pragma unroll
      for (int s = 0; s < S; s += 2) *reinterpret_cast<double2 *>(jb + t0 + s) = make_double2(J[s] + ej, J[s + 1] + ej);
#pragma unroll
      for (int p = 0; p < kMppiBatch; p++)
        if (nb == p) kb[p] = base;
      nb++;
      if (nb == kMppiBatch) { flush(nb); nb = 0; }
    } else {
#pragma unroll
      for (int s = 0; s < S; s++) {
        const double Js = J[s] + ej;
        if (live && t0 + s < T) {
          sDL[s] += duL[s]; sDR[s] += duR[s];
          // J >= 0, so the high words order like the values: one integer compare against the register threshold decides
          // whether this rollout can matter for the step
          if (__double2hiint(Js) <= thr_hi[s]) {
            double *c = acc + s * 4 * NT;
            const double m0 = c[0];
            const double d = (m0 - Js) * a.inv_lambda;     // > 0: J is the new minimum
            const bool newmin = d > 0.0;
            const double e = mppi_exp_neg<false>(-fabs(d));
            if (newmin || e != 0.0) {
              const double l = duL[s], rr = duR[s];
              if (newmin) {
                // the sums are rescaled to the new minimum
                c[NT] = fma(c[NT], e, 1.0);
                c[2 * NT] = fma(c[2 * NT], e, l);
                c[3 * NT] = fma(c[3 * NT], e, rr);
                c[0] = Js;
                if (a.thr_on) thr_hi[s] = __double2hiint(Js + a.cut_lambda) + 1;
              } else {
                c[NT] += e;
                c[2 * NT] = fma(e, l, c[2 * NT]);
                c[3 * NT] = fma(e, rr, c[3 * NT]);
              }
            }
          }
          if (capture) {
            a.J_out[(size_t)k * T + t0 + s] = Js;
            reinterpret_cast<double2 *>(a.du_out)[(size_t)k * T + t0 + s] = make_double2(duL[s], duR[s]);
          }
        }
      }
    }

    mppi_clock(a, first, 17);
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
    mppi_clock(a, first, 18);
  }
  if (FAST && nb) flush(nb);
  // the rollouts are done: let the dependent grid (the next call) be scheduled while this CTA merges; it blocks in
  // griddepcontrol.wait until this whole grid has completed and the plan is visible
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // the TMA unit has READ the staging rows (they are reused below); the writes themselves complete with the grid - a CTA
  // does not have to sit out their latency
  if (tma_store && lane == 0) tma_store_wait_read<0>();
  __syncthreads();      // every warp has left the loop and its bulk stores have drained the staging rows
  mppi_stamp(a, 1);

  // ---- merge the CTA's accumulator sets per time step -----------------------------------------------------------------
  if (!FAST) {
#pragma unroll
    for (int s = 0; s < S; s++) {
      dacc[(s * 2 + 0) * NT + threadIdx.x] = sDL[s];
      dacc[(s * 2 + 1) * NT + threadIdx.x] = sDR[s];
    }
    __syncthreads();
  }
  // power of two >= T, at most NT
  int TP2 = 1;
  while (TP2 < T) TP2 <<= 1;
  const int x = threadIdx.x;
  // set i of step t = the accumulators of lane group (warp i / R, rollout slot i % R) for step t: in the production
  // variants lane t mod G holds them in slot t / G (transposed, see the softmax above), the sums of the variates stay with
  // the lane that simulated the step (lane t / S, slot t mod S)
  auto grp = [&](int i) { return (i / R) * 32 + (i % R) * G; };
  const float *accf_base = reinterpret_cast<const float *>(acc_base + (size_t)S * NT);
  auto cta_min = [&](int t, int i) { return FAST ? acc_base[(t / G) * NT + grp(i) + t % G] : acc_base[((t % S) * 4) * NT + grp(i) + t / S]; };
  auto cta_set = [&](int t, int i, double v[6]) {
    const int tid_d = grp(i) + t / S, sd = t % S;
    if (FAST) {
      const int tid_a = grp(i) + t % G, sa = t / G;
      v[0] = acc_base[sa * NT + tid_a];
      v[1] = (double)accf_base[(sa * 3) * NT + tid_a];
      v[2] = (double)accf_base[(sa * 3 + 1) * NT + tid_a];                 // sums over the VARIATES: scaled by sigma once, after the merge
      v[3] = (double)accf_base[(sa * 3 + 2) * NT + tid_a];
      const float2 z = reinterpret_cast<const float2 *>(dz4 + (sd / 2) * NT + tid_d)[sd & 1];
      v[4] = (double)z.x; v[5] = (double)z.y;
    } else {
#pragma unroll
      for (int j = 0; j < 4; j++) v[j] = acc_base[(sd * 4 + j) * NT + tid_d];
      v[4] = dacc[(sd * 2) * NT + tid_d]; v[5] = dacc[(sd * 2 + 1) * NT + tid_d];
    }
  };
  double v[6];
  mppi_merge_sets<NT, FAST>(T, TP2, NW * R, a.inv_lambda, scratch, cta_min, cta_set, v);
  if (FAST) { v[2] *= a.sigL; v[3] *= a.sigR; v[4] *= a.sigL; v[5] *= a.sigR; }
  const int n_cta = a.n_roll;
  if (a.dbg && x == 0) {          // tuning runs: the CTA's sets are merged (tied to the result)
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now) : "d"(v[1]) : "memory");
    a.dbg[(size_t)blockIdx.x * kMppiDbgSlots + 3] = now;
  }
  if (x < T) {
    if (a.tail) {
      // to merger CTA x, which spins on these words: no count, no fence
      MppiLL *w = a.ll_partials + (size_t)x * 3 * n_cta + blockIdx.x;
#pragma unroll
      for (int j = 0; j < 3; j++) mppi_ll_store(w + (size_t)j * n_cta, v[2 * j], v[2 * j + 1], a.tag);
      if (lane == 0) asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(a.arrive), "l"(1ull) : "memory");
    }
    // the plain copy: the update kernel of the NCCL transport and the partials tap read it after the grid
    double *o = a.partials + ((size_t)x * n_cta + blockIdx.x) * 6;
    reinterpret_cast<double2 *>(o)[0] = make_double2(v[0], v[1]);
    reinterpret_cast<double2 *>(o)[1] = make_double2(v[2], v[3]);
    reinterpret_cast<double2 *>(o)[2] = make_double2(v[4], v[5]);
  }
  mppi_stamp(a, 2);
  if (a.dbg && a.tail && x == 0) {          // tuning runs: when this thread's last word has landed in L2 and come back
    double r0, r1;
    mppi_ll_wait(a.ll_partials + 2 * (size_t)n_cta + blockIdx.x, a.tag, r0, r1);
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now) : "d"(r0) : "memory");
    a.dbg[(size_t)blockIdx.x * kMppiDbgSlots + 7] = now;
  }
}

// ---- the variates of the NEXT call, drawn ahead ------------------------------------------------------------------------
// One thread per Philox call = two time steps of one rollout: [K][T/2] float4 (zL_t, zR_t, zL_t+1, zR_t+1), the values
// mppi_normal_quad would produce inside the rollout loop, bit for bit.  Launched right behind a call's kernel with
// programmatic dependent launch: its CTAs fill the SMs that the call's early finishers leave (the tail of the call - the
// last passes, the merge tree - runs on a few CTAs), and in a control loop it runs while the host turns the pose around.
struct MppiNoiseArgs
{
  unsigned long long *z_ready;   // CTAs of noise kernels that have finished (MppiArgs::z_ready)
  float4 *zbuf;
  int K, half_T, k_offset;
  int half_shift;                // log2(half_T) when it is a power of two (the usual horizons), else -1: spares a 64-bit division per element
  uint32_t call;
  uint32_t key0[10], key1[10];
};

__global__ void __launch_bounds__(256) mppi_noise_kernel(const __grid_constant__ MppiNoiseArgs n)
{
  // the kernel behind this one (the next call) may be scheduled as soon as SMs free up; it waits for this grid's completion
  // (or, queued, for the count at the end of this kernel) and, through the plan's tagged words, for the call in front of
  // this grid.  No wait here: the buffer written was last read three calls ago, and a grid that blocks while resident could
  // starve another rank sharing the GPU
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const size_t total = (size_t)n.K * n.half_T;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t k = n.half_shift >= 0 ? (uint32_t)(i >> n.half_shift) : (uint32_t)(i / n.half_T);
    const uint32_t idx = (uint32_t)(i - (size_t)k * n.half_T);
    uint32_t c0 = idx, c1 = (uint32_t)n.k_offset + k, c2 = n.call, c3 = kDomainMppi;
#pragma unroll
    for (int r = 0; r < 10; r++) {
      const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      const uint32_t n0 = hi1 ^ c1 ^ n.key0[r], n2 = hi0 ^ c3 ^ n.key1[r];
      c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    }
    float4 z;
    box_muller_f32(c0, c1, z.x, z.y);
    box_muller_f32(c2, c3, z.z, z.w);
    n.zbuf[i] = z;
  }
  // count this CTA's share in (a release by one thread after the barrier covers every thread's stores)
  __syncthreads();
  if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(n.z_ready), "l"(1ull) : "memory");
}

// ---- the update as a separate kernel: the NCCL transport of a sharded job and the partials tap -------------------------
// One CTA per time step: merge the partials, then (unless merge_only) the control update of mppi.cpp:112-137 for that
// step, written one slot to the left (the receding-horizon shift).
constexpr int kMppiUpdateThreads = 128;

__global__ void __launch_bounds__(kMppiUpdateThreads) mppi_update_kernel(const MppiUpdateArgs a)
{
  const int t = blockIdx.x;
  // launched with programmatic stream serialization: wait here until the producing grid has finished and its
  // partials are visible (a no-op for an ordinary launch)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  // the next call's rollout grid may be scheduled now (its prologue overlaps this kernel; it waits before the plan)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  double m, S, A, B, DL, DR;
  mppi_block_merge<kMppiUpdateThreads, false>(a.partials, a.n_partials, a.p_stride, a.t_stride, t, a.inv_lambda, m, S, A, B, DL, DR);
  if (threadIdx.x != 0) return;
  if (a.merge_only) {
    double *o = a.merged + (size_t)t * 6;
    o[0] = m; o[1] = S; o[2] = A; o[3] = B; o[4] = DL; o[5] = DR;
    return;
  }
  mppi_apply_update(a, a.u_cur[t], a.u_cur[a.T + t], t, m, S, A, B, DL, DR);
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
