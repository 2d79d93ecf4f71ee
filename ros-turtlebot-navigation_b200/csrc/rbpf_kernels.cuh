// rbpf_kernels.cuh - bmapping::ParticleFilter::SLAM() as sm_100a kernels (fallback and improved-proposal
// branches, map integration, distance field, normalise / resample, map export).
//
// Reference path (all under /root/reference): bmapping/src/bmapping/particle_filter.cpp:141-251 (SLAM),
// :295-322 (motion model), :383-437 (odometry likelihood), :442-500 (normalise, N_eff, low-variance walk),
// :504-599 (improved proposal); bmapping/src/bmapping/grid_mapper.cpp:69-133 (likelihood field), :140-182
// (integrateScan), :272-435 (brushfire distance field), :438-546 (cell state + occupied set), :549-898
// (Bresenham, indexing), :185-226 (map export); bmapping/src/bmapping/sensor_model.cpp:43-112 (end points).
//
// This translation unit is compiled with -fmad=false: every a*b+c below is two IEEE operations, like the
// reference's x86-64 build, because cell indices and resampling ancestors have to come out bit-exact.
//
// Design (DESIGN.md "RBPF"):
//   * ONE WARP PER PARTICLE.  Lanes stride the beams for end points and likelihood terms; the product of the
//     per-beam terms is then taken in beam order (the reference's rounding) by a redundant-lane loop.
//   * per-particle map = SoA planes in HBM: log-odds fp64 [N][G], squared cell distance to the claimed obstacle
//     u32 [N][G] (occ_dist = sqrt(d2) * resolution exactly; the likelihood term of every possible d2 is a
//     host-built table, its head staged into shared memory by TMA together with the beam sin/cos table),
//     plus the occupied set in libstdc++ unordered_set layout (node list + bucket heads, u16) because its
//     ITERATION ORDER seeds the distance transform.
//   * rays are applied in beam order, 32 cells of a ray at a time (cells of one ray are distinct; the k-th cell
//     of the reference's Bresenham variants has a closed form), hash events replayed in lane order.
//   * the distance field is the reference's order-dependent brushfire: one serial chain of G heap steps per particle
//     (libstdc++ push_heap / pop_heap mechanics, ties included).  One CTA per SM keeps 28 particles in flight (a single
//     wave at 4096 particles): the heaps fill the shared memory, the visited bitmaps live in TENSOR MEMORY
//     (tcgen05.ld / tcgen05.st as a scratchpad); a warp carries one particle, or two / four in lane groups under SIMT
//     divergence; lanes 0-3 of a group test the four neighbours.  Particles whose occupied set did not change are skipped.
//   * normalise / N_eff / low-variance walk run in the reference's sequential fp64 order on one warp (weights streamed
//     32 at a time); resampling copies are one gather kernel - across GPUs it reads the ancestor's planes straight from
//     the peer's HBM over NVLink (CUDA IPC mappings).
#pragma once

#include "common.cuh"

namespace b2n
{

constexpr uint16_t kNil16 = 0xFFFFu;
constexpr uint32_t kD2Unreached = 0xFFFFFFFFu;   // cell never claimed: occ_dist = max_occ_dist (grid_mapper.cpp:49,58)
constexpr int kPfWarpsPerCta = 4;
constexpr int kPfStatusOffMap = 1;               // reference: world2Grid / world2RowMajor throw
constexpr int kPfStatusNumeric = 2;              // reference: "eta is 0" / zero variance in pdfNormal

// bucket counts libstdc++'s _Prime_rehash_policy walks through when keys arrive one at a time
__constant__ uint32_t kBucketChainDev[19] = {13, 29, 59, 127, 257, 541, 1109, 2357, 5087, 10273, 20753, 42043, 85229,
                                             172933, 351061, 712697, 1447153, 2938679, 5967347};

struct PfParticle
{
  double pose[3], prev_pose[3];   // theta, x, y (particle_filter.cpp:133)
  double weight;
  uint32_t n_occ, bucket_count, next_resize;
  int32_t chain;
  uint32_t occ_dirty;             // the occupied set changed since the distance field was last grown from it
};

struct PfPlanes
{
  double *log_odds;    // [N][G]
  uint32_t *d2;        // [N][G]
  uint16_t *nxt;       // [N][nxt_stride], slot G = before-begin
  uint16_t *bkt;       // [N][bkt_stride]
  PfParticle *meta;    // [N]
};

struct PfConst
{
  int N, G, xsize, ysize;
  int gstride;                     // plane stride per particle (G rounded up to a multiple of 4)
  int nxt_stride, bkt_stride;
  int cell_radius;                 // grid_mapper.cpp:50
  double xmin, xmax, ymin, ymax, res;
  double range_min, range_max;     // floats of LaserProperties, promoted as the comparison in sensor_model.cpp:81 does
  double t_occ, t_free;            // log-odds thresholds equivalent to prob >= 0.9 / prob <= 0.35 under the host's exp()
  double d_free, d_occ;            // log_odds_free_ - log_odds_prior_, log_odds_occ_ - log_odds_prior_
  const double *beam_cs;           // [max_beams][2] cos, sin of the accumulated beam angle (host libm)
  const double *pz_table;          // [pz_n] z_hit * pdfNormal(sqrt(d2) * res, sigma^2) + z_rand / z_max; last = unreached
  int pz_n, pz_stage;              // table length; entries staged in shared memory
  double sig[3];                   // sqrt of the motion-noise variances (Cholesky of a diagonal matrix)
  // improved proposal
  int k_samples;
  double sig_mode[3];
  double srr, srt, str, stt;
  double scan_min, scan_max, pose_min, pose_max;
};

struct PfCall
{
  const float *scan;
  int n_beams;
  double u_w, u_vx;                // body twist of the scan interval
  double cur_od[3], prev_od[3];
  int icp_ok;
  double icp[3];                   // theta, x, y of Ticp
  uint32_t seed_lo, seed_hi, call;
  int particle_offset;
  const double *ext;               // external standard normals [N][ext_per] + 1, or null
  int ext_per;
  int *status;                     // OR of kPfStatus*
};

// ---- small helpers -------------------------------------------------------------------------------------------
__device__ __forceinline__ bool pf_almost_zero(double d) { return fabs(d - 0.0) < 1.0e-12; }   // rigid2d.hpp:24-27

// rigid2d.hpp:52-64
__device__ __forceinline__ double pf_normalize_angle_pi(double rad)
{
  const double PI = 3.14159265358979323846;
  const double q = floor((rad + PI) / (2.0 * PI));
  rad = (rad + PI) - q * 2.0 * PI;
  if (rad < 0) rad += 2.0 * PI;
  return rad - PI;
}

// grid_mapper.cpp:810-887 on one coordinate pair; false where the reference throws
__device__ __forceinline__ bool pf_world2grid(const PfConst &c, double x, double y, int &gi, int &gj)
{
  if (!(x >= c.xmin && x <= c.xmax)) return false;
  if (!(y >= c.ymin && y <= c.ymax)) return false;
  double i = floor((x - c.xmin) / c.res);
  if (i == (double)c.xsize) i -= 1.0;
  double j = floor((y - c.ymin) / c.res);
  if (j == (double)c.ysize) j -= 1.0;
  gi = (int)i; gj = (int)j;
  return true;
}

// one standard normal of particle `gid` (draw number d of this call); every lane computes the same value
__device__ __forceinline__ double pf_std_normal(const PfCall &q, int local, int gid, int d)
{
  if (q.ext) return q.ext[(size_t)local * q.ext_per + d];
  double z0, z1;
  normal_pair(q.seed_lo, q.seed_hi, kDomainRbpf, q.call, (uint32_t)gid, (uint32_t)(d >> 1), z0, z1);
  return (d & 1) ? z1 : z0;
}

// ---- the occupied set: std::unordered_set<int> in libstdc++ layout ------------------------------------------------
struct OccSetDev
{
  uint16_t *nxt, *bkt;
  uint32_t G, bucket_count, count, next_resize;
  int chain;
  uint32_t dirty;                 // any insert / erase (a rehash only happens inside an insert)
};

// _Hashtable::_M_rehash_aux (unique keys).  Executed by every lane with identical data (same addresses, same
// values), the bucket clear is shared out over the lanes.
__device__ __forceinline__ void occ_rehash(OccSetDev &s, uint32_t n, int lane)
{
  for (uint32_t b = lane; b < n; b += 32) s.bkt[b] = kNil16;
  __syncwarp();
  if (lane == 0) {
    uint32_t p = s.nxt[s.G];
    s.nxt[s.G] = kNil16;
    uint32_t bbegin = 0;
    while (p != kNil16) {
      const uint32_t next = s.nxt[p];
      const uint32_t b = p % n;
      const uint32_t before = s.bkt[b];
      if (before == kNil16) {
        const uint32_t first = s.nxt[s.G];
        s.nxt[p] = (uint16_t)first;
        s.nxt[s.G] = (uint16_t)p;
        s.bkt[b] = (uint16_t)s.G;
        if (first != kNil16) s.bkt[bbegin] = (uint16_t)p;
        bbegin = b;
      } else {
        s.nxt[p] = s.nxt[before];
        s.nxt[before] = (uint16_t)p;
      }
      p = next;
    }
  }
  s.bucket_count = n;
  __syncwarp();
}

// _M_insert_unique_node for a key that is NOT present (caller checked membership through the log-odds)
__device__ __forceinline__ void occ_insert(OccSetDev &s, uint32_t key, int lane)
{
  if (s.count + 1 > s.next_resize) {   // _Prime_rehash_policy::_M_need_rehash, max load factor 1
    const uint32_t floor_bkts = s.next_resize ? (s.count + 1) : max(s.count + 1, 11u);
    if (floor_bkts >= s.bucket_count) {
      s.chain++;
      const uint32_t n = kBucketChainDev[s.chain];
      occ_rehash(s, n, lane);
      s.next_resize = n;
    } else {
      s.next_resize = s.bucket_count;
    }
  }
  if (lane == 0) {
    const uint32_t b = key % s.bucket_count;
    const uint32_t before = s.bkt[b];
    if (before != kNil16) {
      s.nxt[key] = s.nxt[before];
      s.nxt[before] = (uint16_t)key;
    } else {
      const uint32_t first = s.nxt[s.G];
      s.nxt[key] = (uint16_t)first;
      s.nxt[s.G] = (uint16_t)key;
      if (first != kNil16) s.bkt[first % s.bucket_count] = (uint16_t)key;
      s.bkt[b] = (uint16_t)s.G;
    }
  }
  s.count++;
  __syncwarp();
}

// _M_erase(bkt, prev, node) for a key that IS present
__device__ __forceinline__ void occ_erase(OccSetDev &s, uint32_t key, int lane)
{
  if (lane == 0) {
    const uint32_t b = key % s.bucket_count;
    const uint32_t head = s.bkt[b];
    uint32_t prev = head;
    while (s.nxt[prev] != key) prev = s.nxt[prev];
    const uint32_t next = s.nxt[key];
    if (prev == head) {
      if (next == kNil16 || next % s.bucket_count != b) {
        if (next != kNil16) s.bkt[next % s.bucket_count] = (uint16_t)head;
        s.bkt[b] = kNil16;
      }
    } else if (next != kNil16) {
      const uint32_t nb = next % s.bucket_count;
      if (nb != b) s.bkt[nb] = (uint16_t)prev;
    }
    s.nxt[prev] = (uint16_t)next;
    s.nxt[key] = kNil16;
  }
  s.count--;
  __syncwarp();
}

// ---- end points and likelihood terms of one pose (sensor_model.cpp:43-112, grid_mapper.cpp:69-133) ------------------
// Lanes stride the beams.  ep[b] = (i << 16 | j) of the end-point cell, or 0xFFFFFFFF for a gated-out beam;
// pz[b] = likelihood term (only when want_pz).  Returns false (on every lane) if a valid end point is off the map.
__device__ __forceinline__ bool pf_end_points(const PfConst &c, const float *scan, int n, const double *s_beam, const double *s_pz,
                                              const uint32_t *d2_plane, double th, double x, double y, uint32_t *ep, double *pz,
                                              bool want_pz, int lane)
{
  // Transform2D(Vector2D(x, y), theta) then pose * Trs with Trs = identity (rigid2d.cpp:154-166,221-231):
  // x = c*0 - s*0 + x, y = s*0 + c*0 + y, theta += 0, cos/sin recomputed from the same theta
  double st, ct;
  sincos(th, &st, &ct);
  const double tx = ct * 0.0 - st * 0.0 + x;
  const double ty = st * 0.0 + ct * 0.0 + y;
  bool ok = true;
  for (int b = lane; b < n; b += 32) {
    const double range = (double)scan[b];
    uint32_t e = 0xFFFFFFFFu;
    if (range >= c.range_min && range < c.range_max) {
      const double px = range * s_beam[2 * b], py = range * s_beam[2 * b + 1];   // sensor_model.cpp:9-16
      const double wx = ct * px - st * py + tx;                                    // rigid2d.cpp:163-164
      const double wy = st * px + ct * py + ty;
      int gi, gj;
      if (pf_world2grid(c, wx, wy, gi, gj)) {
        e = ((uint32_t)gi << 16) | (uint32_t)gj;
        if (want_pz) {
          const uint32_t d2 = d2_plane[gi * c.xsize + gj];
          const int t = (int)min(d2, (uint32_t)(c.pz_n - 1));
          pz[b] = t < c.pz_stage ? s_pz[t] : __ldg(&c.pz_table[t]);
        }
      } else {
        ok = false;
        e = 0xFFFFFFFEu;   // valid beam, off the map
      }
    }
    ep[b] = e;
  }
  __syncwarp();
  return __all_sync(kFullMask, ok);
}

// p = 1; for every valid beam in order: p *= pz (grid_mapper.cpp:88-128).  Redundant on all lanes.
__device__ __forceinline__ double pf_ordered_product(const uint32_t *ep, const double *pz, int n)
{
  double p = 1.0;
  for (int b = 0; b < n; b++)
    if (ep[b] < 0xFFFFFFFEu) p *= pz[b];
  return p;
}

// ---- integrateScan without the distance field (grid_mapper.cpp:140-178, 438-546, 549-807) -------------------------
__device__ __forceinline__ void pf_integrate_rays(const PfConst &c, const uint32_t *ep, int n, int i0, int j0, double *L,
                                                  OccSetDev &occ, int lane)
{
  const int xs = c.xsize;
  for (int b = 0; b < n; b++) {
    const uint32_t e = ep[b];
    if (e >= 0xFFFFFFFEu) continue;
    const int i1 = (int)(e >> 16), j1 = (int)(e & 0xFFFFu);
    const int dx = i1 - i0, dy = j1 - j0;
    const int adx = abs(dx), ady = abs(dy);
    // the k-th free cell of freeGridIndex, by direction class
    int len, kind;          // kind 0: axis / diagonal walk from the robot cell; 1: lineLow; 2: lineHigh
    int sx = 0, sy = 0;     // step of kind 0
    int ax = 0, ay = 0, inc = 1, dmaj = 1, dmin = 0;
    if (dx == 0) { kind = 0; len = ady; sy = dy < 0 ? -1 : 1; }
    else if (dy == 0) { kind = 0; len = adx; sx = dx < 0 ? -1 : 1; }
    else if (ady < adx) {
      // start cell pushed explicitly, then lineLow from the end with the smaller x (ctr == 0 skipped)
      kind = 1; len = adx; dmaj = adx; dmin = ady;
      if (i0 > i1) { ax = i1; ay = j1; inc = (j0 - j1) < 0 ? -1 : 1; }
      else { ax = i0; ay = j0; inc = dy < 0 ? -1 : 1; }
    } else if (ady > adx) {
      kind = 2; len = ady; dmaj = ady; dmin = adx;
      if (j0 > j1) { ax = i1; ay = j1; inc = (i0 - i1) < 0 ? -1 : 1; }
      else { ax = i0; ay = j0; inc = dx < 0 ? -1 : 1; }
    } else { kind = 0; len = adx; sx = dx < 0 ? -1 : 1; sy = dy < 0 ? -1 : 1; }

    for (int base = 0; base < len; base += 32) {
      const int k = base + lane;
      const bool has = k < len;
      uint32_t cell = 0;
      bool erase = false;
      if (has) {
        int ci, cj;
        if (kind == 0 || k == 0) { ci = i0 + sx * k; cj = j0 + sy * k; }
        else {
          // Bresenham error term in closed form: minor-axis increments before step k
          const int a = 2 * dmin * k - dmaj;
          const int nk = a > 0 ? (a + 2 * dmaj - 1) / (2 * dmaj) : 0;
          if (kind == 1) { ci = ax + k; cj = ay + inc * nk; }
          else { ci = ax + inc * nk; cj = ay + k; }
        }
        cell = (uint32_t)(ci * xs + cj);
        const double lo = L[cell];
        const double ln = lo + c.d_free;                      // grid_mapper.cpp:163
        L[cell] = ln;
        erase = (lo >= c.t_occ) && !(ln >= c.t_occ);          // occupied -> unknown band: leaves occ_cells_
      }
      unsigned m = __ballot_sync(kFullMask, erase);
      while (m) {
        const int src = __ffs(m) - 1;
        const uint32_t key = __shfl_sync(kFullMask, cell, src);
        occ_erase(occ, key, lane);
        occ.dirty = 1u;
        m &= m - 1;
      }
    }
    __syncwarp();
    // the end-point cell (grid_mapper.cpp:172-176)
    const uint32_t cell = (uint32_t)(i1 * xs + j1);
    const double lo = L[cell];
    const double ln = lo + c.d_occ;
    __syncwarp();
    if (lane == 0) L[cell] = ln;
    if (!(lo >= c.t_occ) && (ln >= c.t_occ)) { occ_insert(occ, cell, lane); occ.dirty = 1u; }
    __syncwarp();
  }
}

// shared-memory tables staged by TMA: beam cos/sin and the head of the likelihood table
__device__ __forceinline__ void pf_stage_tables(const PfConst &c, int n_beams, double *s_beam, double *s_pz, uint64_t *bar)
{
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t b0 = (uint32_t)n_beams * 16u, b1 = (uint32_t)c.pz_stage * 8u;
    mbar_expect_tx(bar, b0 + b1);
    tma_load_1d(s_beam, c.beam_cs, b0, bar);
    tma_load_1d(s_pz, c.pz_table, b1, bar);
  }
  mbar_wait(bar, 0);
}

// dynamic shared memory layout shared by the per-particle kernels:
//   [bar 16 B][s_beam n*16][s_pz pz_stage*8][per warp: pz n*8, ep n*4]
__host__ __device__ inline size_t pf_smem_bytes(int n_beams, int pz_stage, int warps)
{
  const size_t n = (size_t)((n_beams + 3) & ~3);
  return 16 + n * 16 + (size_t)pz_stage * 8 + (size_t)warps * (n * 8 + n * 4);
}

// ---- SLAM, motion-model branch: sample, weight, integrate (particle_filter.cpp:158-176, 235-239) -------------------
// mode 0: full step; mode 1: likelihood only at the current poses into out[] (tap, nothing is modified)
__global__ void __launch_bounds__(kPfWarpsPerCta * 32) rbpf_update_kernel(const __grid_constant__ PfConst c, const PfPlanes pl,
                                                                           const __grid_constant__ PfCall q, int mode, double *out)
{
  extern __shared__ __align__(128) unsigned char smem[];
  const int n = q.n_beams;
  const int npad = (n + 3) & ~3;
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
  double *s_beam = reinterpret_cast<double *>(smem + 16);
  double *s_pz = s_beam + 2 * npad;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double *w_pz = s_pz + c.pz_stage + (size_t)warp * npad;
  uint32_t *w_ep = reinterpret_cast<uint32_t *>(s_pz + c.pz_stage + (size_t)kPfWarpsPerCta * npad) + (size_t)warp * npad;

  pf_stage_tables(c, n, s_beam, s_pz, bar);

  const int p = blockIdx.x * kPfWarpsPerCta + warp;
  if (p >= c.N) return;
  PfParticle *me = pl.meta + p;
  double th = me->pose[0], x = me->pose[1], y = me->pose[2];
  const uint32_t n_occ = me->n_occ;
  const uint32_t *d2_plane = pl.d2 + (size_t)p * c.gstride;

  if (mode == 0) {
    // sampleMotionModel (particle_filter.cpp:295-322); L * z for the diagonal motion noise is sig[i] * z[i]
    const int gid = q.particle_offset + p;
    const double w0 = c.sig[0] * pf_std_normal(q, p, gid, 0);
    const double w1 = c.sig[1] * pf_std_normal(q, p, gid, 1);
    const double w2 = c.sig[2] * pf_std_normal(q, p, gid, 2);
    if (lane == 0) { me->prev_pose[0] = th; me->prev_pose[1] = x; me->prev_pose[2] = y; }
    if (pf_almost_zero(q.u_w)) {
      th = pf_normalize_angle_pi(th + w0);
      x += q.u_vx * cos(th) + w1;
      y += q.u_vx * sin(th) + w2;
    } else {
      th = pf_normalize_angle_pi(th + q.u_w + w0);
      x += (-q.u_vx / q.u_w) * sin(th) + (q.u_vx / q.u_w) * sin(th + q.u_w) + w1;
      y += (q.u_vx / q.u_w) * cos(th) - (q.u_vx / q.u_w) * cos(th + q.u_w) + w2;
    }
  }

  const bool on_map = pf_end_points(c, q.scan, n, s_beam, s_pz, d2_plane, th, x, y, w_ep, w_pz, n_occ > 0, lane);
  // likelihoodFieldModel returns 1.0 before looking at any end point when the map has no obstacle (grid_mapper.cpp:94-98)
  double lik = 1.0;
  if (n_occ > 0) {
    if (!on_map) { if (lane == 0) atomicOr(q.status, kPfStatusOffMap); return; }
    lik = pf_ordered_product(w_ep, w_pz, n);
  }
  if (mode == 1) { if (lane == 0) out[p] = lik; return; }

  int i0, j0;
  const bool pose_on_map = pf_world2grid(c, x, y, i0, j0);
  if (!on_map || !pose_on_map) { if (lane == 0) atomicOr(q.status, kPfStatusOffMap); return; }

  OccSetDev occ;
  occ.nxt = pl.nxt + (size_t)p * c.nxt_stride; occ.bkt = pl.bkt + (size_t)p * c.bkt_stride;
  occ.G = (uint32_t)c.G; occ.bucket_count = me->bucket_count; occ.count = n_occ; occ.next_resize = me->next_resize;
  occ.chain = me->chain; occ.dirty = me->occ_dirty;
  pf_integrate_rays(c, w_ep, n, i0, j0, pl.log_odds + (size_t)p * c.gstride, occ, lane);

  if (lane == 0) {
    me->pose[0] = th; me->pose[1] = x; me->pose[2] = y;
    me->weight = me->weight * lik;                         // particle_filter.cpp:175
    me->n_occ = occ.count; me->bucket_count = occ.bucket_count; me->next_resize = occ.next_resize; me->chain = occ.chain;
    me->occ_dirty = occ.dirty;
  }
}

// ---- improved proposal (particle_filter.cpp:178-233, 504-599, 383-437) followed by the same map integration ------
__device__ __forceinline__ bool pf_pdf_normal(double a, double b, double &out)
{
  const double PI = 3.14159265358979323846;
  if (pf_almost_zero(b)) return false;                      // grid_mapper.cpp:20-23 throws
  const double sqrt_inv = 1.0 / sqrt(2.0 * PI * b);
  const double var = -0.5 * (a * a) / b;
  out = sqrt_inv * exp(var);
  return true;
}

__device__ __forceinline__ bool pf_pose_likelihood_odom(const PfConst &c, const double *cur, const double *prev, const double *co,
                                                        const double *po, double &out)
{
  const double rot1 = atan2(co[2] - po[2], co[1] - po[1]) - po[0];
  const double trans = sqrt((co[1] - po[1]) * (co[1] - po[1]) + (co[2] - po[2]) * (co[2] - po[2]));
  const double rot2 = pf_normalize_angle_pi(pf_normalize_angle_pi(co[0]) - pf_normalize_angle_pi(po[0]) - rot1);
  const double rot1_hat = atan2(cur[2] - prev[2], cur[1] - prev[1]) - prev[0];
  const double trans_hat = sqrt((cur[1] - prev[1]) * (cur[1] - prev[1]) + (cur[2] - prev[2]) * (cur[2] - prev[2]));
  const double rot2_hat = pf_normalize_angle_pi(pf_normalize_angle_pi(cur[0]) - pf_normalize_angle_pi(prev[0]) - rot1_hat);
  const double temp1 = c.srr * rot1_hat * rot1_hat + c.srt * trans_hat * trans_hat;
  const double temp2 = c.str * trans_hat * trans_hat + c.stt * rot1_hat * rot1_hat + c.stt * rot2_hat * rot2_hat;
  const double temp3 = c.srr * rot2_hat * rot2_hat + c.srt * trans_hat * trans_hat;
  double p1, p2, p3;
  if (!pf_pdf_normal(pf_normalize_angle_pi(pf_normalize_angle_pi(rot1) - pf_normalize_angle_pi(rot1_hat)), temp1, p1)) return false;
  if (!pf_pdf_normal(trans - trans_hat, temp2, p2)) return false;
  if (!pf_pdf_normal(pf_normalize_angle_pi(pf_normalize_angle_pi(rot2) - pf_normalize_angle_pi(rot2_hat)), temp3, p3)) return false;
  out = p1 * p2 * p3;
  return true;
}

// Eigen LLT (unblocked Cholesky, stops at the first non-positive pivot) then mu + L * z
__device__ __forceinline__ void pf_sample_gaussian3(const double mu[3], double a[3][3], const double z[3], double out[3])
{
  for (int k = 0; k < 3; k++) {
    double x = a[k][k];
    for (int r = 0; r < k; r++) x -= a[k][r] * a[k][r];
    if (x <= 0.0) break;
    x = sqrt(x);
    a[k][k] = x;
    for (int i = k + 1; i < 3; i++) {
      double s = 0.0;
      for (int r = 0; r < k; r++) s += a[i][r] * a[k][r];
      a[i][k] = (a[i][k] - s) / x;
    }
  }
  for (int i = 0; i < 3; i++) {
    const double l0 = a[i][0], l1 = i >= 1 ? a[i][1] : 0.0, l2 = i >= 2 ? a[i][2] : 0.0;
    out[i] = mu[i] + ((l0 * z[0] + l1 * z[1]) + l2 * z[2]);
  }
}

__global__ void __launch_bounds__(kPfWarpsPerCta * 32) rbpf_proposal_kernel(const __grid_constant__ PfConst c, const PfPlanes pl,
                                                                             const __grid_constant__ PfCall q, double *samples)
{
  extern __shared__ __align__(128) unsigned char smem[];
  const int n = q.n_beams;
  const int npad = (n + 3) & ~3;
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
  double *s_beam = reinterpret_cast<double *>(smem + 16);
  double *s_pz = s_beam + 2 * npad;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double *w_pz = s_pz + c.pz_stage + (size_t)warp * npad;
  uint32_t *w_ep = reinterpret_cast<uint32_t *>(s_pz + c.pz_stage + (size_t)kPfWarpsPerCta * npad) + (size_t)warp * npad;

  pf_stage_tables(c, n, s_beam, s_pz, bar);

  const int p = blockIdx.x * kPfWarpsPerCta + warp;
  if (p >= c.N) return;
  PfParticle *me = pl.meta + p;
  const int gid = q.particle_offset + p;
  const double pose[3] = {me->pose[0], me->pose[1], me->pose[2]};
  const double prev[3] = {me->prev_pose[0], me->prev_pose[1], me->prev_pose[2]};
  const uint32_t n_occ = me->n_occ;
  const uint32_t *d2_plane = pl.d2 + (size_t)p * c.gstride;
  double *my_samples = samples + (size_t)p * c.k_samples * 4;     // x[3], likelihood

  // T_x = T(pose) * Ticp (particle_filter.cpp:181-183, rigid2d.cpp:221-231)
  double sp, cp;
  sincos(pose[0], &sp, &cp);
  const double mx = cp * q.icp[1] - sp * q.icp[2] + pose[1];
  const double my = sp * q.icp[1] + cp * q.icp[2] + pose[2];
  const double mth = pose[0] + q.icp[0];

  // sampleMode + gaussianProposal first pass
  double mu[3] = {0.0, 0.0, 0.0}, eta = 0.0;
  bool bad_map = false, bad_num = false;
  for (int s = 0; s < c.k_samples; s++) {
    double xj[3];
    xj[0] = pf_normalize_angle_pi(mth + c.sig_mode[0] * pf_std_normal(q, p, gid, 3 * s + 0));
    xj[1] = mx + c.sig_mode[1] * pf_std_normal(q, p, gid, 3 * s + 1);
    xj[2] = my + c.sig_mode[2] * pf_std_normal(q, p, gid, 3 * s + 2);
    const bool on_map = pf_end_points(c, q.scan, n, s_beam, s_pz, d2_plane, xj[0], xj[1], xj[2], w_ep, w_pz, n_occ > 0, lane);
    double p_scan = 1.0;
    if (n_occ > 0) {
      if (!on_map) { bad_map = true; break; }
      p_scan = pf_ordered_product(w_ep, w_pz, n);
    }
    __syncwarp();
    double p_pose;
    if (!pf_pose_likelihood_odom(c, xj, prev, q.cur_od, q.prev_od, p_pose)) { bad_num = true; break; }
    p_scan = fmin(fmax(p_scan, c.scan_min), c.scan_max);           // std::clamp, :548-549
    p_pose = fmin(fmax(p_pose, c.pose_min), c.pose_max);
    const double lik = p_scan * p_pose;
    if (lane == 0) { my_samples[4 * s + 0] = xj[0]; my_samples[4 * s + 1] = xj[1]; my_samples[4 * s + 2] = xj[2]; my_samples[4 * s + 3] = lik; }
    for (int i = 0; i < 3; i++) mu[i] += xj[i] * lik;
    eta += lik;
  }
  if (!bad_map && !bad_num && pf_almost_zero(eta)) bad_num = true;      // "eta is 0", :577-580
  if (bad_map || bad_num) { if (lane == 0) atomicOr(q.status, bad_map ? kPfStatusOffMap : kPfStatusNumeric); return; }
  __syncwarp();
  for (int i = 0; i < 3; i++) mu[i] /= eta;
  mu[0] = pf_normalize_angle_pi(mu[0]);
  double sigma[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int s = 0; s < c.k_samples; s++) {
    const double d[3] = {my_samples[4 * s + 0] - mu[0], my_samples[4 * s + 1] - mu[1], my_samples[4 * s + 2] - mu[2]};
    const double lik = my_samples[4 * s + 3];
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) sigma[a][b] += (d[a] * d[b]) * lik;
  }
  for (int a = 0; a < 3; a++)
    for (int b = 0; b < 3; b++) sigma[a][b] /= eta;
  const int d0 = 3 * c.k_samples;
  const double z[3] = {pf_std_normal(q, p, gid, d0), pf_std_normal(q, p, gid, d0 + 1), pf_std_normal(q, p, gid, d0 + 2)};
  double np[3];
  pf_sample_gaussian3(mu, sigma, z, np);                               // :214

  // integrateScan at the new pose (:235-239)
  const bool on_map = pf_end_points(c, q.scan, n, s_beam, s_pz, d2_plane, np[0], np[1], np[2], w_ep, w_pz, false, lane);
  int i0, j0;
  const bool pose_on_map = pf_world2grid(c, np[1], np[2], i0, j0);
  if (!on_map || !pose_on_map) { if (lane == 0) atomicOr(q.status, kPfStatusOffMap); return; }
  OccSetDev occ;
  occ.nxt = pl.nxt + (size_t)p * c.nxt_stride; occ.bkt = pl.bkt + (size_t)p * c.bkt_stride;
  occ.G = (uint32_t)c.G; occ.bucket_count = me->bucket_count; occ.count = n_occ; occ.next_resize = me->next_resize;
  occ.chain = me->chain; occ.dirty = me->occ_dirty;
  pf_integrate_rays(c, w_ep, n, i0, j0, pl.log_odds + (size_t)p * c.gstride, occ, lane);
  if (lane == 0) {
    for (int i = 0; i < 3; i++) { me->prev_pose[i] = pose[i]; me->pose[i] = np[i]; }
    me->weight = me->weight * eta;                                      // :231
    me->n_occ = occ.count; me->bucket_count = occ.bucket_count; me->next_resize = occ.next_resize; me->chain = occ.chain;
    me->occ_dirty = occ.dirty;
  }
}

// ---- distance field: the reference's brushfire (grid_mapper.cpp:272-435) ---------------------------------------------
// The result depends on the ORDER in which a libstdc++ binary heap releases equal keys and on the pop-after-push quirk
// of grid_mapper.cpp:399-431, so each particle's field is one serial chain of about G heap operations: the kernel is
// bound by the latency of one chain step, not by bandwidth, and all that can be done is (a) run every particle's chain
// at the same time and (b) make a step short.
//   (a) ONE CTA PER SM with up to 28 warps, one particle per warp: 148 x 28 = 4144 chains in flight, a single wave for
//       the 4096 particles of BASELINE configs[2].  That leaves 8 KB of shared memory per warp, which is spent on the
//       heap alone; the visited bitmap (G bits = 5 KB at 200 x 200) lives in TENSOR MEMORY: each warp owns a
//       32-lane x 73-column slice of the SM's 256 KB of TMEM (word w of the bitmap = lane w % 32 of column w / 32),
//       read with tcgen05.ld (two adjacent columns cover the four neighbours of a cell), updated in registers and
//       written back with tcgen05.st.  No tensor-core instruction is involved; TMEM is used as a scratchpad.
//   (b) heap entry = d2 << 32 | ci << 24 | cj << 16 | si << 8 | sj (cell and claiming obstacle as byte coordinates: no
//       division to decode); entry i lives in slot i + 1 so that both children of a node come in one 16-byte load;
//       the comparator of grid_mapper.hpp:104-111 (a.occ_dist > b.occ_dist) is the comparison of the d2 words;
//       every lane runs the heap code redundantly (broadcast loads, same-value stores), lanes 0-3 test the four
//       neighbours; entries beyond the shared-memory capacity spill to a per-warp global area.
constexpr int kDfMaxWarps = 28;
constexpr int kDfStatusHeapOverflow = 4;

__device__ __forceinline__ void tmem_ld2(uint32_t taddr, uint32_t &r0, uint32_t &r1)
{
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];\n" : "=r"(r0), "=r"(r1) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_st2(uint32_t taddr, uint32_t r0, uint32_t r1)
{
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};\n" ::"r"(taddr), "r"(r0), "r"(r1) : "memory");
}
__device__ __forceinline__ void tmem_st1(uint32_t taddr, uint32_t r0)
{
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};\n" ::"r"(taddr), "r"(r0) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

// shared-memory accessors on 32-bit shared addresses (one instruction each, no generic-pointer arithmetic)
__device__ __forceinline__ uint2 lds64(uint32_t addr)
{
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];\n" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr)
{
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts64(uint32_t addr, uint2 v)
{
  asm volatile("st.shared.v2.u32 [%0], {%1, %2};\n" ::"r"(addr), "r"(v.x), "r"(v.y) : "memory");
}

// Heap entry as two words: .y = key (squared cell distance), .x = ci << 24 | cj << 16 | si << 8 | sj.
// Entry i lives in shared slot i + 1 (byte address sb + 8 (i + 1)) while i < hcap, else in the global spill area.
struct HeapDev
{
  uint32_t sb;             // shared byte address of slot 0
  uint2 *g;                // global spill
  int hcap, gcap, len;
  bool overflow;
  __device__ __forceinline__ uint2 get(int i) const { return i < hcap ? lds64(sb + 8u * (uint32_t)(i + 1)) : g[min(i - hcap, gcap - 1)]; }
  __device__ __forceinline__ void set(int i, uint2 v)
  {
    if (i < hcap) sts64(sb + 8u * (uint32_t)(i + 1), v);
    else if (i - hcap < gcap) g[i - hcap] = v;
    else overflow = true;
  }
  // std::__push_heap(first, hole, top = 0, value), any storage
  __device__ __forceinline__ void sift_up(int hole, uint2 value)
  {
    while (hole > 0) {
      const int parent = (hole - 1) >> 1;
      const uint2 pe = get(parent);
      if (pe.y > value.y) { set(hole, pe); hole = parent; }
      else break;
    }
    set(hole, value);
  }
  // push while the whole heap (including the new entry) is in shared memory; returns true if the entry reached the root
  __device__ __forceinline__ bool push_shared(uint2 e)
  {
    int hole = len++;
    while (hole > 0) {
      const int parent = (hole - 1) >> 1;
      const uint2 pe = lds64(sb + 8u * (uint32_t)(parent + 1));
      if (pe.y > e.y) { sts64(sb + 8u * (uint32_t)(hole + 1), pe); hole = parent; }
      else break;
    }
    sts64(sb + 8u * (uint32_t)(hole + 1), e);
    return hole == 0;
  }
  __device__ __forceinline__ void push_any(uint2 e) { sift_up(len, e); len++; }
  // std::pop_heap + pop_back = __adjust_heap(first, 0, len - 1, last value), len >= 2, whole heap in shared memory.
  // The walk is carried on byte addresses: a = address of slot `second` (= entry second - 1); its children pair
  // (entries 2 second + 1, 2 second + 2) sits in slots 2 second + 2, 2 second + 3 = address 2 a - (sb - 16).
  __device__ __forceinline__ void pop_shared()
  {
    const int L = --len;                                   // index of the last entry, >= 1
    const uint2 value = lds64(sb + 8u * (uint32_t)(L + 1));
    const uint32_t alim = sb + 8u * (uint32_t)((L - 1) >> 1);
    const uint32_t K = sb - 16u;
    uint32_t a = sb;                                       // slot of index `second` = 0 ... the hole is entry `second`
    uint2 above = make_uint2(0u, 0u);                      // the entry now sitting in the hole's parent
    // (tried and measured slower in every lane-group configuration: loading the children pairs of both candidates one
    // level ahead - 3 % faster with one chain per SM, slower with 28: the walk is bound by its dependent instructions,
    // not by the shared-memory latency; eleven fully unrolled, predicated levels instead of the loop - 42.9 vs 39.7 ms;
    // keeping the root in a register across steps with a peeled first level - 41.0 vs 39.7 ms; a speculative walk that
    // takes log2(GL) levels per round - every lane of the group loads the children pair of one node below the hole, a
    // ballot of the winners tells each lane whether it is on the path, the path's lanes move their winners up at once -
    // bit-exact, but 42.8 vs 40.7 ms: three rounds of LDS.128 + two ballots + two shuffles cost more than nine levels
    // of eleven instructions; waiting for the bitmap's tcgen05.st only before the next step's tcgen05.ld - no change)
    while (a < alim) {
      const uint32_t hole_a = a + 8u;
      a = 2u * a - K;
      const uint4 pr = lds128(a);                          // entries second - 1 (x, y) and second (z, w)
      const bool left = pr.w > pr.y;                       // comp(right, left): take the left child
      above.x = left ? pr.x : pr.z;
      above.y = min(pr.y, pr.w);
      a -= left ? 8u : 0u;
      sts64(hole_a, above);
    }
    int second = (int)((a - sb) >> 3);
    if ((L & 1) == 0 && second == ((L - 2) >> 1)) {
      const uint32_t hole_a = a + 8u;
      second = 2 * second + 1;                             // the only child
      a = sb + 8u * (uint32_t)second;
      above = lds64(a + 8u);
      sts64(hole_a, above);
    }
    // __push_heap from the leaf: the parent of the hole is the entry just moved there
    if (second > 0 && above.y > value.y) sift_up(second, value);
    else sts64(a + 8u, value);
  }
  __device__ __forceinline__ void pop_any()
  {
    if (len > 1) {
      const int L = len - 1;
      const uint2 value = get(L);
      int hole = 0, second = 0;
      while (second < (L - 1) / 2) {
        second = 2 * (second + 1);
        const uint2 r = get(second), l = get(second - 1);
        uint2 mv = r;
        if (r.y > l.y) { second--; mv = l; }
        set(hole, mv);
        hole = second;
      }
      if ((L & 1) == 0 && second == (L - 2) / 2) {
        second = 2 * (second + 1);
        set(hole, get(second - 1));
        hole = second - 1;
      }
      sift_up(hole, value);
    }
    len--;
  }
};

// dynamic shared memory of the distance-field kernel: [warps][hcap + 2] heap slots (+ [warps][words] visited bitmap
// when it is not kept in tensor memory)
__host__ __device__ inline size_t pf_df_smem_bytes(int G, int hcap, int warps, bool tmem_marks)
{
  return (size_t)warps * ((size_t)(hcap + 2) * 8 + (tmem_marks ? 0 : (size_t)((G + 31) / 32) * 4));
}

struct PfDfArgs
{
  int hcap, gcap, warps, cols_per_warp, skip_clean;
  unsigned long long *spill;   // [gridDim.x * warps][gcap]
  unsigned long long *stats;
  int *status;
};

template <bool TMEM_MARKS>
__global__ void __launch_bounds__(kDfMaxWarps * 32, 1) rbpf_distance_field_kernel(const __grid_constant__ PfConst c, const PfPlanes pl,
                                                                                   const PfDfArgs d)
{
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint32_t tmem_base_slot;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int words = (c.G + 31) / 32;
  const int xs = c.xsize, ys = c.ysize, R = c.cell_radius;

  uint32_t tbase = 0;
  if (TMEM_MARKS) {
    if (warp == 0) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_slot)), "n"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    // this warp's slice: its lane quadrant (the hardware lets warp w touch lanes 32 (w % 4) .. + 31 only), its columns
    tbase = tmem_base_slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * d.cols_per_warp);
  }

  HeapDev H;
  H.sb = smem_u32(smem) + (uint32_t)warp * (uint32_t)(d.hcap + 2) * 8u;
  H.g = reinterpret_cast<uint2 *>(d.spill) + ((size_t)blockIdx.x * d.warps + warp) * d.gcap;
  H.hcap = d.hcap; H.gcap = d.gcap; H.overflow = false;
  uint32_t *marked = TMEM_MARKS ? nullptr : reinterpret_cast<uint32_t *>(smem + (size_t)d.warps * (d.hcap + 2) * 8) + (size_t)warp * words;
  unsigned long long iters = 0, skipped = 0;
  int heap_max = 0;
  // lanes 0..3 test (i-1, j), (i, j-1), (i+1, j), (i, j+1)  (:401-427)
  const int dI = lane == 0 ? -1 : lane == 2 ? 1 : 0, dJ = lane == 1 ? -1 : lane == 3 ? 1 : 0;
  const int dIdx = dI * xs + dJ;
  const uint32_t dEntry = ((uint32_t)dI << 24) + ((uint32_t)dJ << 16);   // added to ci << 24 | cj << 16 (no carry crosses: in bounds)
  const bool tester = lane < 4;
  const int R2 = R * R;

  for (int p = blockIdx.x * d.warps + warp; p < c.N; p += gridDim.x * d.warps) {
    PfParticle *me = pl.meta + p;
    if (me->n_occ == 0) continue;                                        // grid_mapper.cpp:335-338
    // the field is a deterministic function of the occupied set's iteration order (cells nothing reaches keep their
    // value): an untouched set would reproduce the field that is already there
    if (d.skip_clean && me->occ_dirty == 0) { skipped++; continue; }
    __syncwarp();
    if (lane == 0) me->occ_dirty = 0;
    const uint16_t *nxt = pl.nxt + (size_t)p * c.nxt_stride;
    uint32_t *d2p = pl.d2 + (size_t)p * c.gstride;
    const int cols = (words + 31) / 32;
    if (TMEM_MARKS) {
      for (int col = 0; col <= cols; col++) tmem_st1(tbase + col, 0u);
      tmem_wait_st();
    } else {
      for (int w = lane; w < words; w += 32) marked[w] = 0;
      __syncwarp();
    }
    H.len = 0;
    // seeds in the iteration order of occ_cells_ (:348-361); all of distance 0, so push_heap leaves them where they land
    {
      uint32_t col_cur = 0xFFFFFFFFu, wv = 0;     // TMEM: the bitmap column being filled (this lane's word of it)
      for (uint32_t key = nxt[c.G]; key != kNil16; key = nxt[key]) {
        const uint32_t ki = key / (uint32_t)xs, kj = key - ki * (uint32_t)xs;
        if (TMEM_MARKS) {
          const uint32_t w = key >> 5, col = w >> 5;
          if (col != col_cur) {
            if (col_cur != 0xFFFFFFFFu) { tmem_st1(tbase + col_cur, wv); tmem_wait_st(); }
            uint32_t dummy;
            tmem_ld2(tbase + col, wv, dummy);
            col_cur = col;
          }
          if ((uint32_t)lane == (w & 31u)) wv |= 1u << (key & 31);
        } else if (lane == 0) marked[key >> 5] |= 1u << (key & 31);
        if (lane == 0) d2p[key] = 0;
        H.set(H.len, make_uint2((ki << 24) | (kj << 16) | (ki << 8) | kj, 0u));
        H.len++;
      }
      if (TMEM_MARKS && col_cur != 0xFFFFFFFFu) { tmem_st1(tbase + col_cur, wv); tmem_wait_st(); }
    }
    __syncwarp();
    heap_max = max(heap_max, H.len);
    uint32_t it = 0;
    while (H.len > 0) {
      const uint2 top = lds64(H.sb + 8u);                                 // Q.top(), :399 (entry 0 is always in shared memory)
      const int ci = (int)(top.x >> 24), cj = (int)((top.x >> 16) & 0xFFu), si = (int)((top.x >> 8) & 0xFFu), sj = (int)(top.x & 0xFFu);
      const int ni = ci + dI, nj = cj + dJ;
      const bool inb = tester && (unsigned)ni < (unsigned)xs && (unsigned)nj < (unsigned)ys;
      const int idx0 = ci * xs + cj;                                      // grid2RowMajor
      const int idx = inb ? idx0 + dIdx : idx0;
      uint32_t r0 = 0, r1 = 0, colA = 0, word;
      if (TMEM_MARKS) {
        colA = (uint32_t)max(idx0 - xs, 0) >> 10;                        // columns colA, colA + 1 hold all four neighbours
        tmem_ld2(tbase + colA, r0, r1);
        const uint32_t w = (uint32_t)idx >> 5;
        const uint32_t v0 = __shfl_sync(kFullMask, r0, w & 31u), v1 = __shfl_sync(kFullMask, r1, w & 31u);
        word = ((w >> 5) != colA) ? v1 : v0;
      } else word = marked[idx >> 5];
      const int di = ni - si, dj = nj - sj;
      const int d2 = di * di + dj * dj;
      // unmarked, inside distances_.at()'s range, not farther than cell_radius_
      const bool valid = inb && !((word >> (idx & 31)) & 1u) && d2 <= R2 && max(abs(di), abs(dj)) < R;
      if (valid) d2p[idx] = (uint32_t)d2;
      const uint2 entry = make_uint2((top.x & 0xFFFFu) + (top.x & 0xFFFF0000u) + dEntry, (uint32_t)d2);
      unsigned m = __ballot_sync(kFullMask, valid);
      if (m) {
        const bool shared_ok = H.len + 4 <= H.hcap;
        do {
          const int from = __ffs(m) - 1;
          m &= m - 1;
          const uint2 e = make_uint2(__shfl_sync(kFullMask, entry.x, from), __shfl_sync(kFullMask, entry.y, from));
          const int eidx = __shfl_sync(kFullMask, idx, from);
          if (TMEM_MARKS) {
            const uint32_t w = (uint32_t)eidx >> 5;
            if ((uint32_t)lane == (w & 31u)) {
              if ((w >> 5) != colA) r1 |= 1u << (eidx & 31); else r0 |= 1u << (eidx & 31);
            }
          } else if (lane == 0) marked[eidx >> 5] |= 1u << (eidx & 31);
          if (shared_ok) H.push_shared(e);
          else H.push_any(e);
        } while (m);
        if (TMEM_MARKS) { tmem_st2(tbase + colA, r0, r1); tmem_wait_st(); }
        else __syncwarp();
        heap_max = max(heap_max, H.len);
      }
      // Q.pop() pops whatever is on top NOW (a pushed cell may be nearer than the current top), :431
      if (H.len >= 2 && H.len <= H.hcap) H.pop_shared();
      else H.pop_any();
      it++;
    }
    iters += it;
  }
  if (d.stats && lane == 0) { atomicAdd(&d.stats[0], iters); atomicMax(&d.stats[1], (unsigned long long)heap_max); atomicAdd(&d.stats[2], skipped); }
  if (H.overflow && lane == 0) atomicOr(d.status, kDfStatusHeapOverflow);
  if (TMEM_MARKS) {
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base_slot), "n"(512));
  }
}

// ---- the same brushfire with SEVERAL particles per warp -------------------------------------------------------------
// The single-particle-per-warp kernel above is bound by instruction issue: 28 warps per SM each spend ~250 issue
// slots per heap step on work that only one lane's worth of data needs.  Here a warp carries NG = 32 / GL particles,
// one per group of GL lanes (8 or 16): the groups run their heap loops under SIMT divergence (different trip counts
// are masked, the common iterations issue once for all groups), so an issue slot serves up to NG particles, and the
// 28 chains of an SM need only 28 / NG warps.  The visited bitmaps stay in tensor memory; tcgen05.ld/st are warp-wide
// with a uniform address, so the two columns a group needs are staged through a small shared-memory window
// ([NG][2][32] words per warp): one tcgen05.ld per group, testers read and atomicOr their word there, dirty
// windows go back with one tcgen05.st.
__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr)
{
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];\n" : "=r"(r) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
  return r;
}

__host__ __device__ inline size_t pf_dfg_smem_bytes(int hcap, int warps, int ng)
{
  // heaps, then per warp the staged bitmap windows [ng][2][32] words and the candidate entries [ng][4] x 8 bytes
  return (size_t)warps * ng * (size_t)(hcap + 2) * 8 + (size_t)warps * ng * (64 * 4 + 4 * 8);
}

template <int GL>
__global__ void __launch_bounds__(kDfMaxWarps * 32, 1) rbpf_distance_field_groups_kernel(const __grid_constant__ PfConst c, const PfPlanes pl,
                                                                                          const PfDfArgs d)
{
  constexpr int NG = 32 / GL;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint32_t tmem_base_slot;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = lane / GL, sl = lane % GL;
  const int slots = d.warps * NG, slot = warp * NG + q;
  const int words = (c.G + 31) / 32, cols = (words + 31) / 32;
  const int CG = cols + 1;                                               // TMEM columns per particle (one spare for the x2 loads)
  const int xs = c.xsize, ys = c.ysize, R = c.cell_radius;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t twarp = tmem_base_slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * d.cols_per_warp);

  HeapDev H;
  H.sb = smem_u32(smem) + (uint32_t)slot * (uint32_t)(d.hcap + 2) * 8u;
  H.g = reinterpret_cast<uint2 *>(d.spill) + ((size_t)blockIdx.x * slots + slot) * d.gcap;
  H.hcap = d.hcap; H.gcap = d.gcap; H.overflow = false; H.len = 0;
  uint32_t *stage = reinterpret_cast<uint32_t *>(smem + (size_t)slots * (d.hcap + 2) * 8) + (size_t)warp * NG * (64 + 8);   // [NG][2][32]
  uint32_t *my_stage = stage + q * 64;
  // candidate entries of the group's four testers, handed to the whole group through shared memory (a shuffle with a
  // per-group mask inside the divergent push loop costs a dozen instructions of mask checking)
  const uint32_t cand = smem_u32(stage + NG * 64) + (uint32_t)q * 32u;
  unsigned long long iters = 0, skipped = 0;
  int heap_max = 0;
  // lanes 0..3 of a group test (i-1, j), (i, j-1), (i+1, j), (i, j+1)  (:401-427)
  const int dI = sl == 0 ? -1 : sl == 2 ? 1 : 0, dJ = sl == 1 ? -1 : sl == 3 ? 1 : 0;
  const int dIdx = dI * xs + dJ;
  const uint32_t dEntry = ((uint32_t)dI << 24) + ((uint32_t)dJ << 16);
  const bool tester = sl < 4;
  const int R2 = R * R;

  for (int p0 = blockIdx.x * slots; p0 < c.N; p0 += gridDim.x * slots) {
    const int p = p0 + slot;
    bool act = p < c.N && pl.meta[min(p, c.N - 1)].n_occ != 0;           // grid_mapper.cpp:335-338
    // an untouched occupied set would reproduce the field that is already there (see the single-particle kernel)
    if (act && d.skip_clean && pl.meta[p].occ_dirty == 0) { act = false; if (sl == 0) skipped++; }
    __syncwarp();
    if (act && sl == 0) pl.meta[p].occ_dirty = 0;
    const uint16_t *nxt = pl.nxt + (size_t)min(p, c.N - 1) * c.nxt_stride;
    uint32_t *d2p = pl.d2 + (size_t)min(p, c.N - 1) * c.gstride;
    uint32_t *d2n = d2p + dIdx;
    for (int col = 0; col < NG * CG; col++) tmem_st1(twarp + col, 0u);
    tmem_wait_st();
    H.len = 0;
    // seeds in the iteration order of occ_cells_ (:348-361), the groups in lockstep; all of distance 0, so push_heap
    // leaves them where they land
    {
      uint32_t key = act ? (uint32_t)nxt[c.G] : (uint32_t)kNil16;
      while (__any_sync(kFullMask, key != kNil16)) {
        const bool has = key != kNil16;
        if (has) {
          const uint32_t ki = key / (uint32_t)xs, kj = key - ki * (uint32_t)xs;
          if (sl == 0) d2p[key] = 0;
          H.set(H.len, make_uint2((ki << 24) | (kj << 16) | (ki << 8) | kj, 0u));
          H.len++;
        }
#pragma unroll
        for (int qq = 0; qq < NG; qq++) {
          const uint32_t kq = __shfl_sync(kFullMask, has ? key : 0xFFFFFFFFu, qq * GL);
          if (kq != 0xFFFFFFFFu) {
            const uint32_t ta = twarp + (uint32_t)(qq * CG) + (kq >> 10);
            uint32_t v = tmem_ld1(ta);
            if ((uint32_t)lane == ((kq >> 5) & 31u)) v |= 1u << (kq & 31u);
            tmem_st1(ta, v);
            tmem_wait_st();
          }
        }
        if (has) key = nxt[key];
      }
    }
    heap_max = max(heap_max, H.len);
    uint32_t it = 0;
    while (__any_sync(kFullMask, H.len > 0)) {
      const bool on = H.len > 0;
      __syncwarp();                                                       // heap and window accesses of the previous step are done in every lane
      const uint2 top = lds64(H.sb + 8u);                                 // Q.top(), :399
      const int ci = (int)(top.x >> 24), cj = (int)((top.x >> 16) & 0xFFu), si = (int)((top.x >> 8) & 0xFFu), sj = (int)(top.x & 0xFFu);
      const int ni = ci + dI, nj = cj + dJ;
      const bool inb = on && tester && (unsigned)ni < (unsigned)xs && (unsigned)nj < (unsigned)ys;
      const int idx0 = ci * xs + cj;                                      // grid2RowMajor
      const int idx = inb ? idx0 + dIdx : idx0;
      const uint32_t colA = on ? ((uint32_t)max(idx0 - xs, 0) >> 10) : 0u;   // columns colA, colA + 1 hold all four neighbours
      uint32_t cA[NG], a0[NG], a1[NG];
#pragma unroll
      for (int qq = 0; qq < NG; qq++) {
        cA[qq] = __shfl_sync(kFullMask, colA, qq * GL);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];\n" : "=r"(a0[qq]), "=r"(a1[qq]) : "r"(twarp + (uint32_t)(qq * CG) + cA[qq]));
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");      // one wait for all the groups' loads
#pragma unroll
      for (int qq = 0; qq < NG; qq++) {
        stage[qq * 64 + lane] = a0[qq];
        stage[qq * 64 + 32 + lane] = a1[qq];
      }
      __syncwarp();
      const uint32_t w = (uint32_t)idx >> 5;
      uint32_t *wp = my_stage + (((w >> 5) != colA) ? 32 : 0) + (w & 31u);
      const uint32_t word = *wp;
      const int di = ni - si, dj = nj - sj;
      const int d2 = di * di + dj * dj;
      // unmarked, inside distances_.at()'s range, not farther than cell_radius_
      const bool valid = inb && !((word >> (idx & 31)) & 1u) && d2 <= R2 && max(abs(di), abs(dj)) < R;
      if (valid) {
        d2n[idx0] = (uint32_t)d2;
        atomicOr(wp, 1u << (idx & 31));
        sts64(cand + 8u * (uint32_t)sl, make_uint2(top.x + dEntry, (uint32_t)d2));
      }
      const unsigned vmask = __ballot_sync(kFullMask, valid);
      if (vmask) {
        __syncwarp();
#pragma unroll
        for (int qq = 0; qq < NG; qq++) {
          if ((vmask >> (qq * GL)) & 0xFu) tmem_st2(twarp + (uint32_t)(qq * CG) + cA[qq], stage[qq * 64 + lane], stage[qq * 64 + 32 + lane]);
        }
        tmem_wait_st();
      }
      unsigned m = (vmask >> (q * GL)) & 0xFu;
      if (m) {
        const bool shared_ok = H.len + 4 <= H.hcap;
        do {
          const int from = __ffs(m) - 1;
          m &= m - 1;
          const uint2 e = lds64(cand + 8u * (uint32_t)from);
          if (shared_ok) H.push_shared(e);
          else H.push_any(e);
        } while (m);
        heap_max = max(heap_max, H.len);
      }
      // Q.pop() pops whatever is on top NOW (a pushed cell may be nearer than the current top), :431
      if (on) {
        if (H.len >= 2 && H.len <= H.hcap) H.pop_shared();
        else H.pop_any();
        it++;
      }
    }
    iters += it;
  }
  if (d.stats && sl == 0) { atomicAdd(&d.stats[0], iters); atomicMax(&d.stats[1], (unsigned long long)heap_max); atomicAdd(&d.stats[2], skipped); }
  if (H.overflow && sl == 0) atomicOr(d.status, kDfStatusHeapOverflow);
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base_slot), "n"(512));
}

// ---- normalise, N_eff, low-variance walk (particle_filter.cpp:442-500) -----------------------------------------------
struct PfResample
{
  double *w;            // [n_total] gathered weights in global particle order, normalised in place
  double *cum;          // [n_total] cumulative sums of the normalised weights, in index order
  int32_t *ancestors;   // [n_total]
  int *info;            // [0] = N_eff as printed, [1] = resampled
};

// normalizeWeights + effectiveParticles + lowVarianceResampling (particle_filter.cpp:442-500) on the gathered weights of
// ALL ranks.  Bit-exact ancestors need the reference's sequential fp64 order in three places: the sum, the sum of squares
// and the cumulative sum c the walk compares against - each a chain of dependent additions (about 8 cycles an element on
// this part) that no amount of parallelism shortens.  So the kernel makes sure the chains wait for nothing else, and that
// nothing else is serial:
//   * one CTA; thread 0 runs the sum and then the sum-of-squares chain, thread 32 the cumulative-sum chain next to it,
//     both out of shared memory; every other thread streams the weights in (a chunk ahead, double buffered), divides them
//     by the sum in parallel (the division is per element, not part of a chain) and writes them back;
//   * the walk itself is NOT a chain: sample m takes the first i with U_m <= c_i (clamped to N - 1 when the weights run
//     out).  c is non-decreasing (weights are products of positive likelihoods) and so is U_m = r + m / (N - 1), so the
//     reference's nested loops (:484-497), which resume the search at the previous sample's i, find exactly that i - and
//     every m can search for it independently (rbpf_ancestors_kernel, a binary search per sample over the stored c).
constexpr int kPfMaxRanks = 64;
constexpr int kNormThreads = 256;
constexpr int kNormChunk = 2048;       // weights per buffer

__global__ void __launch_bounds__(kNormThreads) rbpf_normalize_kernel(PfResample r, int n_total)
{
  __shared__ double buf[2][kNormChunk];
  __shared__ double s_sum;
  const int tid = threadIdx.x;
  const int n_chunks = (n_total + kNormChunk - 1) / kNormChunk;
  // chunk c into its buffer, by the `nload` threads that are not running a chain at the moment (lrank = 0 .. nload - 1)
  auto load_chunk = [&](int c, bool normalise, double sum, int lrank, int nload) {
    double *b = buf[c & 1];
    for (int i = lrank; i < kNormChunk; i += nload) {
      const int idx = c * kNormChunk + i;
      double v = 0.0;
      if (idx < n_total) {
        v = r.w[idx];
        if (normalise) { v = v / sum; r.w[idx] = v; }                      // :451
      }
      b[i] = v;
    }
  };
  // ---- sum in index order (:446-449) --------------------------------------------------------------------------------------
  load_chunk(0, false, 0.0, tid, kNormThreads);
  __syncthreads();
  double sum = 0.0;
  for (int c = 0; c < n_chunks; c++) {
    if (tid == 0) {
      const double *b = buf[c & 1];
      const int cnt = min(kNormChunk, n_total - c * kNormChunk);
      int i = 0;
      for (; i + 8 <= cnt; i += 8) {
        const double v0 = b[i], v1 = b[i + 1], v2 = b[i + 2], v3 = b[i + 3], v4 = b[i + 4], v5 = b[i + 5], v6 = b[i + 6], v7 = b[i + 7];
        sum += v0; sum += v1; sum += v2; sum += v3; sum += v4; sum += v5; sum += v6; sum += v7;
      }
      for (; i < cnt; i++) sum += b[i];
    } else if (c + 1 < n_chunks) {
      load_chunk(c + 1, false, 0.0, tid - 1, kNormThreads - 1);
    }
    __syncthreads();
  }
  if (tid == 0) s_sum = sum;
  __syncthreads();
  sum = s_sum;
  // ---- w /= sum; sum of squares (:451-457, 463) and, side by side, the cumulative sum of the walk (:481,492) ------------------
  double sq = 0.0, cacc = 0.0;
  load_chunk(0, true, sum, tid, kNormThreads);
  __syncthreads();
  for (int c = 0; c < n_chunks; c++) {
    const double *b = buf[c & 1];
    const int cnt = min(kNormChunk, n_total - c * kNormChunk);
    if (tid == 0) {
      int i = 0;
      for (; i + 4 <= cnt; i += 4) {
        const double v0 = b[i], v1 = b[i + 1], v2 = b[i + 2], v3 = b[i + 3];
        const double p0 = v0 * v0, p1 = v1 * v1, p2 = v2 * v2, p3 = v3 * v3;    // std::pow(w, 2)
        sq += p0; sq += p1; sq += p2; sq += p3;
      }
      for (; i < cnt; i++) { const double v = b[i]; sq += v * v; }
    } else if (tid == 32) {
      double *out = r.cum + (size_t)c * kNormChunk;
      int i = 0;
      for (; i + 4 <= cnt; i += 4) {
        const double v0 = b[i], v1 = b[i + 1], v2 = b[i + 2], v3 = b[i + 3];
        cacc += v0; const double c0 = cacc;                                // c = w_0, then c += w_i
        cacc += v1; const double c1 = cacc;
        cacc += v2; const double c2 = cacc;
        cacc += v3;
        out[i] = c0; out[i + 1] = c1; out[i + 2] = c2; out[i + 3] = cacc;
      }
      for (; i < cnt; i++) { cacc += b[i]; out[i] = cacc; }
    } else if (c + 1 < n_chunks) {
      load_chunk(c + 1, true, sum, tid - (tid > 32 ? 2 : 1), kNormThreads - 2);
    }
    __syncthreads();
  }
  if (tid == 0) {
    const int neff = (int)(1.0 / sq);                                      // :463-464
    r.info[0] = neff; r.info[1] = neff < (n_total / 2) ? 1 : 0;
  }
}

// the ancestors: identity when the filter does not resample, else for every sample m the first i with U_m <= c_i
__global__ void __launch_bounds__(256) rbpf_ancestors_kernel(PfResample r, int n_total, const __grid_constant__ PfCall q)
{
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n_total) return;
  if (!r.info[1]) { r.ancestors[m] = m; return; }
  double z;
  if (q.ext) z = q.ext[(size_t)q.ext_per];                                 // caller passes a pointer to the last variate
  else { double z1; normal_pair(q.seed_lo, q.seed_hi, kDomainRbpf, q.call, kStreamResample, 0u, z, z1); }
  const double rr = z / (double)n_total;                                   // :475-476
  const double step = 1.0 / (n_total - 1);
  const double U = rr + (double)(m * step);                                // :485
  // smallest i in [0, N - 1] with !(U > c_i); N - 1 when there is none (:486-491)
  int lo = 0, hi = n_total - 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (U > r.cum[mid]) lo = mid + 1; else hi = mid;
  }
  r.ancestors[m] = lo;
}

__global__ void rbpf_gather_weights_kernel(const PfParticle *meta, double *w, int n)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) w[i] = meta[i].weight;
}

// after normalisation every particle's weight is the normalised one (resampled or not)
__global__ void rbpf_scatter_weights_kernel(PfParticle *meta, const double *w, int n)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) meta[i].weight = w[i];
}

// particle m of dst <- particle src_index[m] of src (skipped when negative), plane by plane with 16-byte accesses;
// blockIdx.y = m.  src_index holds LOCAL particle indices (the host turns global ancestors into them).
__global__ void __launch_bounds__(256) rbpf_copy_particles_kernel(const __grid_constant__ PfConst c, const PfPlanes src, const PfPlanes dst,
                                                                   const int32_t *src_index)
{
  const int m = blockIdx.y;
  const int a = src_index[m];
  if (a < 0) return;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  const uint4 *s0 = reinterpret_cast<const uint4 *>(src.log_odds + (size_t)a * c.gstride);
  uint4 *d0 = reinterpret_cast<uint4 *>(dst.log_odds + (size_t)m * c.gstride);
  for (int i = tid; i < c.gstride / 2; i += nth) d0[i] = s0[i];
  const uint4 *s1 = reinterpret_cast<const uint4 *>(src.d2 + (size_t)a * c.gstride);
  uint4 *d1 = reinterpret_cast<uint4 *>(dst.d2 + (size_t)m * c.gstride);
  for (int i = tid; i < c.gstride / 4; i += nth) d1[i] = s1[i];
  const uint4 *s2 = reinterpret_cast<const uint4 *>(src.nxt + (size_t)a * c.nxt_stride);
  uint4 *d2 = reinterpret_cast<uint4 *>(dst.nxt + (size_t)m * c.nxt_stride);
  for (int i = tid; i < c.nxt_stride / 8; i += nth) d2[i] = s2[i];
  const int used = (int)((src.meta[a].bucket_count + 7) / 8);
  const uint4 *s3 = reinterpret_cast<const uint4 *>(src.bkt + (size_t)a * c.bkt_stride);
  uint4 *d3 = reinterpret_cast<uint4 *>(dst.bkt + (size_t)m * c.bkt_stride);
  for (int i = tid; i < used; i += nth) d3[i] = s3[i];
  if (tid == 0) dst.meta[m] = src.meta[a];
}

// Resampling across GPUs over NVLink peer memory: slot m of dst <- particle src_index[m] of rank src_rank[m], read straight
// out of that rank's planes (mapped with CUDA IPC; sets[r] = rank r's OLD set, the own rank included, so local and
// remote ancestors are one code path and one launch).  The copied particle's weight is taken from the normalised global
// weight vector every rank holds after the allgather - the remote meta may not have received it yet.
__global__ void __launch_bounds__(256) rbpf_copy_particles_p2p_kernel(const __grid_constant__ PfConst c, const PfPlanes *sets, const PfPlanes dst,
                                                                       const int32_t *ancestors, int slot_offset, int n_local, const double *w_all)
{
  const int m = blockIdx.y;
  const int ga = ancestors[slot_offset + m];                 // global ancestor
  const PfPlanes src = sets[ga / n_local];
  const int a = ga % n_local;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  const uint4 *s0 = reinterpret_cast<const uint4 *>(src.log_odds + (size_t)a * c.gstride);
  uint4 *d0 = reinterpret_cast<uint4 *>(dst.log_odds + (size_t)m * c.gstride);
  for (int i = tid; i < c.gstride / 2; i += nth) d0[i] = s0[i];
  const uint4 *s1 = reinterpret_cast<const uint4 *>(src.d2 + (size_t)a * c.gstride);
  uint4 *d1 = reinterpret_cast<uint4 *>(dst.d2 + (size_t)m * c.gstride);
  for (int i = tid; i < c.gstride / 4; i += nth) d1[i] = s1[i];
  const uint4 *s2 = reinterpret_cast<const uint4 *>(src.nxt + (size_t)a * c.nxt_stride);
  uint4 *d2 = reinterpret_cast<uint4 *>(dst.nxt + (size_t)m * c.nxt_stride);
  for (int i = tid; i < c.nxt_stride / 8; i += nth) d2[i] = s2[i];
  const PfParticle meta = src.meta[a];
  const int used = (int)((meta.bucket_count + 7) / 8);
  const uint4 *s3 = reinterpret_cast<const uint4 *>(src.bkt + (size_t)a * c.bkt_stride);
  uint4 *d3 = reinterpret_cast<uint4 *>(dst.bkt + (size_t)m * c.bkt_stride);
  for (int i = tid; i < used; i += nth) d3[i] = s3[i];
  if (tid == 0) {
    PfParticle q = meta;
    q.weight = w_all[ga];
    dst.meta[m] = q;
  }
}

// ---- getRobotState / newMap (particle_filter.cpp:255-291, grid_mapper.cpp:185-226) ----------------------------------
// first particle with the largest weight, strict > starting from 0.0 (so index 0 when nothing is positive)
__global__ void __launch_bounds__(1024) rbpf_best_kernel(const PfParticle *meta, int n, int *best, double *best_pose_weight)
{
  __shared__ double sw[32];
  __shared__ int si[32];
  double bw = 0.0;
  int bi = 0x7FFFFFFF;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double w = meta[i].weight;
    if (w > bw) { bw = w; bi = i; }
  }
  for (int d = 16; d > 0; d >>= 1) {
    const double ow = __shfl_xor_sync(kFullMask, bw, d);
    const int oi = __shfl_xor_sync(kFullMask, bi, d);
    if (ow > bw || (ow == bw && oi < bi)) { bw = ow; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { sw[threadIdx.x >> 5] = bw; si[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x < 32) {
    bw = threadIdx.x < (blockDim.x >> 5) ? sw[threadIdx.x] : 0.0;
    bi = threadIdx.x < (blockDim.x >> 5) ? si[threadIdx.x] : 0x7FFFFFFF;
    for (int d = 16; d > 0; d >>= 1) {
      const double ow = __shfl_xor_sync(kFullMask, bw, d);
      const int oi = __shfl_xor_sync(kFullMask, bi, d);
      if (ow > bw || (ow == bw && oi < bi)) { bw = ow; bi = oi; }
    }
    if (threadIdx.x == 0) {
      const int b = (bi == 0x7FFFFFFF) ? 0 : bi;
      best[0] = b; best[1] = 0;
      best_pose_weight[0] = meta[b].pose[0]; best_pose_weight[1] = meta[b].pose[1]; best_pose_weight[2] = meta[b].pose[2];
      best_pose_weight[3] = bw;
    }
  }
}

// the same argmax over a SHARDED filter: every rank holds the normalised weights of all ranks (the allgather of the last
// SLAM call) and the ancestor list, so the weight of global slot m after that call is w[anc[m]] (w[m] when the filter did
// not resample) - no communication.  best = (index in the owner's set, owner rank); the pose is then read from the owner's
// memory (CUDA IPC mapping, the same one the resampling copies use).
__global__ void __launch_bounds__(1024) rbpf_best_global_kernel(const double *w, const int32_t *anc, int resampled, int n_total, int n_local, int *best,
                                                                double *best_weight)
{
  __shared__ double sw[32];
  __shared__ int si[32];
  double bw = 0.0;
  int bi = 0x7FFFFFFF;
  for (int i = threadIdx.x; i < n_total; i += blockDim.x) {
    const double v = w[resampled ? anc[i] : i];
    if (v > bw) { bw = v; bi = i; }
  }
  for (int d = 16; d > 0; d >>= 1) {
    const double ow = __shfl_xor_sync(kFullMask, bw, d);
    const int oi = __shfl_xor_sync(kFullMask, bi, d);
    if (ow > bw || (ow == bw && oi < bi)) { bw = ow; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { sw[threadIdx.x >> 5] = bw; si[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x < 32) {
    bw = threadIdx.x < (blockDim.x >> 5) ? sw[threadIdx.x] : 0.0;
    bi = threadIdx.x < (blockDim.x >> 5) ? si[threadIdx.x] : 0x7FFFFFFF;
    for (int d = 16; d > 0; d >>= 1) {
      const double ow = __shfl_xor_sync(kFullMask, bw, d);
      const int oi = __shfl_xor_sync(kFullMask, bi, d);
      if (ow > bw || (ow == bw && oi < bi)) { bw = ow; bi = oi; }
    }
    if (threadIdx.x == 0) {
      const int b = (bi == 0x7FFFFFFF) ? 0 : bi;
      best[0] = b % n_local; best[1] = b / n_local;
      best_weight[3] = bw;
    }
  }
}

__global__ void rbpf_fetch_pose_kernel(const PfPlanes *sets, const int *which, double *best_pose_weight)
{
  const PfParticle *p = sets[which[1]].meta + which[0];
  best_pose_weight[0] = p->pose[0]; best_pose_weight[1] = p->pose[1]; best_pose_weight[2] = p->pose[2];
}

// distance field of one particle as fp32 metres (Cell::occ_dist = sqrt(d2) * resolution, max_occ_dist where never reached)
__global__ void rbpf_export_distance_kernel(const __grid_constant__ PfConst c, const PfPlanes *sets, const int *which, double max_occ_dist,
                                            float *out)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.G) return;
  const uint32_t v = sets[which[1]].d2[(size_t)which[0] * c.gstride + i];       // which = (index in its owner's set, owner rank)
  out[i] = (float)(v == kD2Unreached ? max_occ_dist : sqrt((double)v) * c.res);
}

// occupancy export of one particle: prob from the log-odds exactly as updateCellState left it, transposed output
__global__ void rbpf_export_map_kernel(const __grid_constant__ PfConst c, const PfPlanes *sets, const int *which, int8_t *out)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.G) return;
  const double l = sets[which[1]].log_odds[(size_t)which[0] * c.gstride + i];   // which = (index in its owner's set, owner rank)
  const int row = i / c.xsize, col = i - row * c.xsize;
  const int idx = col * c.xsize + row;                                      // grid_mapper.cpp:195-197
  int8_t v;
  if (l >= c.t_occ) v = 100;                                                // prob := 1
  else if (l <= c.t_free) v = 0;                                            // prob := 0
  else {
    const double prob = 1 - (1 / (1 + exp(l)));                             // grid_mapper.hpp:27-30
    v = (prob == 0.5) ? (int8_t)-1 : (int8_t)(prob * 100);
  }
  out[idx] = v;
}

} // namespace b2n
