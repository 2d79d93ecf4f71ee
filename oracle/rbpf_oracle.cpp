// TEST INFRASTRUCTURE ONLY (oracle/). Not part of the product path: only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
//
// CPU restatement (fp64, Eigen-free, single thread) of bmapping::ParticleFilter::SLAM() and what
// it stands on: LaserScanner::laserEndPoints, GridMapper::{likelihoodFieldModel, integrateScan,
// freeGridIndex, updateCellState, updateCellHash, euclideanSignedDistanceField, gridMap} and the
// filter's motion model, improved proposal, normalisation and low-variance resampling - statement
// by statement in the reference's arithmetic order.  Each block cites the reference lines it follows
// (paths relative to /root/reference).  Pinned BIT-EXACTLY by tests/test_oracle_rbpf.py against
// oracle/_ref/libref_nav.so (the unmodified reference sources compiled here) and against the
// committed fixtures in tests/golden/ generated from it.
//
// Two pieces of the C++ standard library decide the reference's RESULTS and are therefore restated
// explicitly instead of being used (tests pin both against the real containers):
//   * OccSet  - iteration order of std::unordered_set<int> (libstdc++ _Hashtable: singly linked node
//     list, bucket -> before-node, prime bucket growth).  That order seeds the distance transform
//     (grid_mapper.cpp:348-361).
//   * MinHeap - std::priority_queue = std::push_heap / std::pop_heap (libstdc++ __push_heap,
//     __adjust_heap).  Its tie order decides which source claims a cell (grid_mapper.cpp:397-431).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <queue>
#include <unordered_set>
#include <vector>

#include "noise.hpp"

namespace
{

const double PI = 3.14159265358979323846;   // rigid2d/include/rigid2d/rigid2d.hpp:17

// rigid2d.hpp:24-27
inline bool almost_equal(double a, double b, double eps = 1.0e-12) { return std::fabs(a - b) < eps; }

// rigid2d.hpp:52-64
inline double normalize_angle_PI(double rad)
{
  const double q = std::floor((rad + PI) / (2.0 * PI));
  rad = (rad + PI) - q * 2.0 * PI;
  if (rad < 0) rad += 2.0 * PI;
  return rad - PI;
}

// grid_mapper.hpp:27-38
inline double logOdds2Prob(double l) { return 1 - (1 / (1 + std::exp(l))); }
inline double prob2LogOdds(double p) { return std::log(p / (1 - p)); }

// grid_mapper.cpp:18-28; returns false where the reference throws
inline bool pdfNormal(double a, double b, double *out)
{
  if (almost_equal(b, 0.0)) return false;
  const double sqrt_inv = 1.0 / std::sqrt(2.0 * PI * b);
  const double var = -0.5 * (a * a) / b;
  *out = sqrt_inv * std::exp(var);
  return true;
}

// ------------------------------------------------------------------------------------------------
// OccSet: std::unordered_set<int> as libstdc++ lays it out (bits/hashtable.h, hashtable_policy.h,
// src/c++11/hashtable_c++0x.cc).  hash(int) = value; bucket = hash % bucket_count; max load factor 1.
// ------------------------------------------------------------------------------------------------
// bucket counts the growth policy walks through when elements arrive one at a time:
// first insert -> _M_next_bkt(max(11 + 1, 2)) = 13, afterwards _M_next_bkt(2 * bucket_count)
const uint32_t kBucketChain[] = {13, 29, 59, 127, 257, 541, 1109, 2357, 5087, 10273, 20753, 42043, 85229, 172933,
                                 351061, 712697, 1447153, 2938679, 5967347};

struct OccSet
{
  int G = 0;                       // keys are cell indices in [0, G); node G is _M_before_begin
  std::vector<int32_t> nxt;        // [G+1] successor in the node list, -1 = null
  std::vector<uint8_t> in;         // membership (the reference uses find())
  std::vector<int32_t> bkt;        // bucket -> node BEFORE the bucket's first node, -1 = empty bucket
  uint32_t bucket_count = 1;       // libstdc++ starts with the single in-object bucket
  uint32_t count = 0;
  uint32_t next_resize = 0;
  int chain = -1;                  // index into kBucketChain of the current bucket_count

  void init(int cells)
  {
    G = cells;
    nxt.assign(G + 1, -1);
    in.assign(G, 0);
    bkt.assign(1, -1);
    bucket_count = 1; count = 0; next_resize = 0; chain = -1;
  }
  bool contains(int key) const { return in[key] != 0; }

  // _Hashtable::_M_rehash_aux(n, true_type)
  void rehash(uint32_t n)
  {
    std::vector<int32_t> nb(n, -1);
    int32_t p = nxt[G];
    nxt[G] = -1;
    uint32_t bbegin_bkt = 0;
    while (p >= 0) {
      const int32_t next = nxt[p];
      const uint32_t b = (uint32_t)p % n;
      if (nb[b] < 0) {
        nxt[p] = nxt[G];
        nxt[G] = p;
        nb[b] = G;
        if (nxt[p] >= 0) nb[bbegin_bkt] = p;
        bbegin_bkt = b;
      } else {
        nxt[p] = nxt[nb[b]];
        nxt[nb[b]] = p;
      }
      p = next;
    }
    bkt.swap(nb);
    bucket_count = n;
  }

  // _M_insert_unique_node: _Prime_rehash_policy::_M_need_rehash, then _M_insert_bucket_begin
  void insert(int key)
  {
    if (in[key]) return;
    if (count + 1 > next_resize) {
      const double min_bkts = (double)std::max<uint32_t>(count + 1, next_resize ? 0 : 11);   // max load factor 1
      if (min_bkts >= bucket_count) {
        chain++;
        rehash(kBucketChain[chain]);
        next_resize = kBucketChain[chain];
      } else {
        next_resize = bucket_count;
      }
    }
    const uint32_t b = (uint32_t)key % bucket_count;
    if (bkt[b] >= 0) {
      nxt[key] = nxt[bkt[b]];
      nxt[bkt[b]] = key;
    } else {
      nxt[key] = nxt[G];
      nxt[G] = key;
      if (nxt[key] >= 0) bkt[(uint32_t)nxt[key] % bucket_count] = key;
      bkt[b] = G;
    }
    in[key] = 1;
    count++;
  }

  // _M_erase(bkt, prev, n) with _M_remove_bucket_begin
  void erase(int key)
  {
    if (!in[key]) return;
    const uint32_t b = (uint32_t)key % bucket_count;
    int32_t prev = bkt[b];
    while (nxt[prev] != key) prev = nxt[prev];
    const int32_t next = nxt[key];
    if (prev == bkt[b]) {
      // key is the first node of its bucket
      const bool next_other = next < 0 || (uint32_t)next % bucket_count != b;
      if (next_other) {
        if (next >= 0) bkt[(uint32_t)next % bucket_count] = bkt[b];
        bkt[b] = -1;
      }
    } else if (next >= 0) {
      const uint32_t nb = (uint32_t)next % bucket_count;
      if (nb != b) bkt[nb] = prev;
    }
    nxt[prev] = next;
    nxt[key] = -1;
    in[key] = 0;
    count--;
  }

  template <class F>
  void for_each(F f) const
  {
    for (int32_t p = nxt[G]; p >= 0; p = nxt[p]) f(p);
  }
};

// ------------------------------------------------------------------------------------------------
// MinHeap: std::priority_queue<Cell, std::vector<Cell>, CompareDistance> (grid_mapper.hpp:104-111:
// comp(a, b) = a.occ_dist > b.occ_dist) through libstdc++'s __push_heap / __adjust_heap.
// ------------------------------------------------------------------------------------------------
struct HeapCell
{
  double occ_dist;
  int i, j, src_i, src_j;
};

struct MinHeap
{
  std::vector<HeapCell> v;
  size_t max_size = 0;
  static bool comp(const HeapCell &a, const HeapCell &b) { return a.occ_dist > b.occ_dist; }

  bool empty() const { return v.empty(); }
  const HeapCell &top() const { return v.front(); }

  void push_heap_at(size_t hole, size_t top_index, const HeapCell &value)
  {
    size_t parent = (hole - 1) / 2;
    while (hole > top_index && comp(v[parent], value)) {
      v[hole] = v[parent];
      hole = parent;
      parent = (hole - 1) / 2;
    }
    v[hole] = value;
  }
  void push(const HeapCell &c)
  {
    v.push_back(c);
    const HeapCell value = v.back();
    push_heap_at(v.size() - 1, 0, value);
    max_size = std::max(max_size, v.size());
  }
  void pop()
  {
    if (v.size() > 1) {
      const size_t len = v.size() - 1;
      const HeapCell value = v[len];
      v[len] = v[0];
      size_t hole = 0, second = 0;
      while (second < (len - 1) / 2) {
        second = 2 * (second + 1);
        if (comp(v[second], v[second - 1])) second--;
        v[hole] = v[second];
        hole = second;
      }
      if ((len & 1) == 0 && second == (len - 2) / 2) {
        second = 2 * (second + 1);
        v[hole] = v[second - 1];
        hole = second - 1;
      }
      push_heap_at(hole, 0, value);
    }
    v.pop_back();
  }
};

// ------------------------------------------------------------------------------------------------
struct Vec2 { double x, y; };

// rigid2d::Transform2D as far as this path uses it (rigid2d/src/rigid2d/rigid2d.cpp:154-166,221-231)
struct Tf
{
  double theta = 0, c = 1, s = 0, x = 0, y = 0;
  Tf() {}
  Tf(double px, double py, double th) : theta(th), c(std::cos(th)), s(std::sin(th)), x(px), y(py) {}
  Vec2 apply(Vec2 v) const { return Vec2{c * v.x - s * v.y + x, s * v.x + c * v.y + y}; }
  Tf &mul(const Tf &r)
  {
    x = c * r.x - s * r.y + x;
    y = s * r.x + c * r.y + y;
    theta += r.theta;
    c = std::cos(theta);
    s = std::sin(theta);
    return *this;
  }
};

struct Laser
{
  float beam_min, beam_max, beam_delta, range_min, range_max;
  double z_hit, z_short, z_max, z_rand, sigma_hit;
};

struct GridStats
{
  uint64_t ray_cells = 0, esdf_iterations = 0, heap_max = 0, esdf_pushes = 0;
};

struct Grid
{
  Laser L;
  double prior = 0.5, prob_occ = 0.90, prob_free = 0.35;          // grid_mapper.cpp:42-44
  double l_prior, l_occ, l_free;                                  // :45-47
  double res, max_occ_dist = 10.0;                                // :48-49
  unsigned int cell_radius;                                       // :50
  double xmin, xmax, ymin, ymax;
  int xsize, ysize;                                               // :55-56
  std::vector<double> log_odds, prob, occ_dist;                   // Cell fields, grid_mapper.hpp:65-101
  std::vector<int> state;
  OccSet occ;
  GridStats stats;

  static unsigned int mapSize(double lo, double hi, double r) { return (unsigned int)std::ceil((hi - lo) / r); }   // :31-34

  void init(const Laser &laser, double r, double x0, double x1, double y0, double y1)
  {
    L = laser; res = r; xmin = x0; xmax = x1; ymin = y0; ymax = y1;
    l_prior = prob2LogOdds(prior); l_occ = prob2LogOdds(prob_occ); l_free = prob2LogOdds(prob_free);
    cell_radius = mapSize(0.0, max_occ_dist, res);
    xsize = (int)mapSize(xmin, xmax, res);
    ysize = (int)mapSize(ymin, ymax, res);
    const size_t n = (size_t)xsize * ysize;
    log_odds.assign(n, l_prior); prob.assign(n, prior); occ_dist.assign(n, max_occ_dist); state.assign(n, -1);   // :58
    occ.init((int)n);
  }

  int grid2RowMajor(int i, int j) const { return i * xsize + j; }   // :890-898

  // grid_mapper.cpp:810-849; false where the reference throws
  bool world2Grid(double x, double y, int *gi, int *gj) const
  {
    if (!(x >= xmin && x <= xmax)) return false;
    if (!(y >= ymin && y <= ymax)) return false;
    int i = (int)std::floor((x - xmin) / res);
    if (i == xsize) i--;
    int j = (int)std::floor((y - ymin) / res);
    if (j == ysize) j--;
    *gi = i; *gj = j;
    return true;
  }
  // grid_mapper.cpp:852-887 (floor kept in double, then passed to grid2RowMajor(int,int))
  bool world2RowMajor(double x, double y, unsigned int *idx) const
  {
    if (!(x >= xmin && x <= xmax)) return false;
    if (!(y >= ymin && y <= ymax)) return false;
    double i = std::floor((x - xmin) / res);
    if (i == xsize) i--;
    double j = std::floor((y - ymin) / res);
    if (j == ysize) j--;
    *idx = (unsigned int)grid2RowMajor((int)i, (int)j);
    return true;
  }

  // sensor_model.cpp:43-112 (Trs = identity as in turtle_mapping_node.cpp:397: Tms = pose * Trs)
  void laserEndPoints(std::vector<Vec2> &pts, const float *scan, int n, const Tf &pose) const
  {
    Tf Tms = pose;
    Tms.mul(Tf());
    double beam_angle = L.beam_min;
    for (int i = 0; i < n; i++) {
      const double range = scan[i];
      if (range >= L.range_min && range < L.range_max) {
        Vec2 p{range * std::cos(beam_angle), range * std::sin(beam_angle)};   // :9-16
        pts.push_back(Tms.apply(p));
      }
      beam_angle += L.beam_delta;
      if (L.beam_max < 0.0 && beam_angle <= L.beam_max) beam_angle = L.beam_min;
      else if (L.beam_max >= 0.0 && beam_angle >= L.beam_max) beam_angle = L.beam_min;
    }
  }

  // grid_mapper.cpp:69-133; returns 0 ok, 1 where the reference throws
  int likelihoodFieldModel(const float *scan, int n, const Tf &pose, double *out) const
  {
    const double var_hit = L.sigma_hit * L.sigma_hit;
    std::vector<Vec2> pts;
    laserEndPoints(pts, scan, n, pose);
    double p = 1.0;
    if (occ.count == 0) { *out = p; return 0; }
    for (const Vec2 &pt : pts) {
      double pz = 0.0;
      unsigned int idx;
      if (!world2RowMajor(pt.x, pt.y, &idx)) return 1;
      const double z = occ_dist[idx];
      double g;
      if (!pdfNormal(z, var_hit, &g)) return 1;
      pz += L.z_hit * g;
      pz += L.z_rand / L.z_max;
      p *= pz;
    }
    *out = p;
    return 0;
  }

  // grid_mapper.cpp:480-546
  void updateCellHash(int st, int index)
  {
    if (st == 1) occ.insert(index);
    else occ.erase(index);
  }
  // grid_mapper.cpp:438-477
  void updateCellState(int idx)
  {
    const double p = logOdds2Prob(log_odds[idx]);
    if (p == prior) { state[idx] = -1; prob[idx] = prior; updateCellHash(-1, idx); }
    else if (p >= prob_occ) { state[idx] = 1; prob[idx] = 1; updateCellHash(1, idx); }
    else if (p <= prob_free) { state[idx] = 0; prob[idx] = 0; }
    else { state[idx] = -1; prob[idx] = p; updateCellHash(-1, idx); }
  }

  // grid_mapper.cpp:707-807
  void lineLow(std::vector<int> &out, int x0, int y0, int x1, int y1) const
  {
    int dx = x1 - x0, dy = y1 - y0, yi = 1;
    if (dy < 0) { yi = -1; dy = -dy; }
    int D = 2 * dy - dx, y = y0, ctr = 0;
    for (int x = x0; x < x1; x++) {
      if (ctr != 0) out.push_back(grid2RowMajor(x, y));
      if (D > 0) { y += yi; D -= 2 * dx; }
      D += 2 * dy;
      ctr++;
    }
  }
  void lineHigh(std::vector<int> &out, int x0, int y0, int x1, int y1) const
  {
    int dx = x1 - x0, dy = y1 - y0, xi = 1;
    if (dx < 0) { xi = -1; dx = -dx; }
    int D = 2 * dx - dy, x = x0, ctr = 0;
    for (int y = y0; y < y1; y++) {
      if (ctr != 0) out.push_back(grid2RowMajor(x, y));
      if (D > 0) { x += xi; D -= 2 * dy; }
      D += 2 * dx;
      ctr++;
    }
  }
  void lineDiag(std::vector<int> &out, int x0, int y0, int x1, int y1) const
  {
    const int dx = x1 - x0, dy = y1 - y0;
    const int xi = dx < 0 ? -1 : 1, yi = dy < 0 ? -1 : 1;
    int x = x0, y = y0;
    while (x != x1 && y != y1) {
      out.push_back(grid2RowMajor(x, y));
      x += xi; y += yi;
    }
  }
  // grid_mapper.cpp:549-704; false where the reference throws
  bool freeGridIndex(std::vector<int> &out, const Vec2 &point, const Tf &pose) const
  {
    int x0, y0, x1, y1;
    if (!world2Grid(pose.x, pose.y, &x0, &y0)) return false;
    if (!world2Grid(point.x, point.y, &x1, &y1)) return false;
    const int dx = x1 - x0, dy = y1 - y0;
    if (dx == 0) {
      if (dy < 0) for (int y = y0; y > y1; y--) out.push_back(grid2RowMajor(x0, y));
      else for (int y = y0; y < y1; y++) out.push_back(grid2RowMajor(x0, y));
    } else if (dy == 0) {
      if (dx < 0) for (int x = x0; x > x1; x--) out.push_back(grid2RowMajor(x, y0));
      else for (int x = x0; x < x1; x++) out.push_back(grid2RowMajor(x, y0));
    } else if (std::abs(dy) < std::abs(dx)) {
      out.push_back(grid2RowMajor(x0, y0));
      if (x0 > x1) lineLow(out, x1, y1, x0, y0);
      else lineLow(out, x0, y0, x1, y1);
    } else if (std::abs(dy) > std::abs(dx)) {
      out.push_back(grid2RowMajor(x0, y0));
      if (y0 > y1) lineHigh(out, x1, y1, x0, y0);
      else lineHigh(out, x0, y0, x1, y1);
    } else {
      lineDiag(out, x0, y0, x1, y1);
    }
    return true;
  }

  // grid_mapper.cpp:272-329
  void enqueueCell(int i, int j, int src_i, int src_j, MinHeap &Q, std::vector<uint8_t> &marked)
  {
    const int idx = grid2RowMajor(i, j);
    if (marked[idx]) return;
    const unsigned int di = (unsigned int)std::abs(i - src_i), dj = (unsigned int)std::abs(j - src_j);
    if (di >= cell_radius || dj >= cell_radius) return;            // distances_.at() throws -> caught -> return
    const double dist = std::sqrt((double)(di * di + dj * dj));    // :257-269 (unsigned arithmetic)
    if (dist > cell_radius) return;
    occ_dist[idx] = dist * res;
    Q.push(HeapCell{occ_dist[idx], i, j, src_i, src_j});
    stats.esdf_pushes++;
    marked[idx] = 1;
  }
  // grid_mapper.cpp:333-435
  void distanceField()
  {
    if (occ.count == 0) return;
    std::vector<uint8_t> marked((size_t)xsize * ysize, 0);
    MinHeap Q;
    occ.for_each([&](int key) {
      occ_dist[key] = 0.0;
      marked[key] = 1;
      Q.push(HeapCell{0.0, key / xsize, key % xsize, key / xsize, key % xsize});
    });
    while (!Q.empty()) {
      const HeapCell c = Q.top();
      if (c.i > 0) enqueueCell(c.i - 1, c.j, c.src_i, c.src_j, Q, marked);
      if (c.j > 0) enqueueCell(c.i, c.j - 1, c.src_i, c.src_j, Q, marked);
      if (c.i < xsize - 1) enqueueCell(c.i + 1, c.j, c.src_i, c.src_j, Q, marked);
      if (c.j < ysize - 1) enqueueCell(c.i, c.j + 1, c.src_i, c.src_j, Q, marked);
      Q.pop();                                                       // pops the CURRENT top, :431
      stats.esdf_iterations++;
    }
    stats.heap_max = std::max<uint64_t>(stats.heap_max, Q.max_size);
  }

  // grid_mapper.cpp:140-182; 0 ok, 1 where the reference throws (state is left as the reference leaves it)
  int integrateScan(const float *scan, int n, const Tf &pose)
  {
    std::vector<Vec2> pts;
    laserEndPoints(pts, scan, n, pose);
    for (size_t b = 0; b < pts.size(); b++) {
      std::vector<int> free_index;
      if (!freeGridIndex(free_index, pts[b], pose)) return 1;
      for (int idx : free_index) {
        log_odds[idx] += l_free - l_prior;
        updateCellState(idx);
      }
      stats.ray_cells += free_index.size() + 1;
      unsigned int idx;
      if (!world2RowMajor(pts[b].x, pts[b].y, &idx)) return 1;
      log_odds[idx] += l_occ - l_prior;
      updateCellState((int)idx);
    }
    distanceField();
    return 0;
  }

  // grid_mapper.cpp:185-226
  void gridMap(int8_t *map) const
  {
    const size_t n = log_odds.size();
    for (size_t i = 0; i < n; i++) {
      const size_t row = i / xsize, col = i % xsize, idx = col * xsize + row;
      const double p = prob[i];
      if (p == prior) map[idx] = -1;
      else if (p >= prob_occ) map[idx] = 100;
      else if (p <= prob_free) map[idx] = 0;
      else map[idx] = (int8_t)(p * 100);
    }
  }
};

// ------------------------------------------------------------------------------------------------
struct Particle
{
  double weight;
  Grid grid;
  double pose[3], prev_pose[3];   // theta, x, y (particle_filter.cpp:133)
};

struct PfParams
{
  Laser laser;
  double res, xmin, xmax, ymin, ymax;
  int num_particles, k;
  double srr, srt, str, stt;
  double motion_noise[3], sample_range[3];
  double scan_min, scan_max, pose_min, pose_max;
  double init_pose[3];
};

struct Pf
{
  PfParams P;
  int N;
  std::vector<Particle> set;
  double normal_sqrd_sum = 0.0;
  // noise seam
  int mode = 1;                 // 0 = A (mt19937_64, the reference's draw order), 1 = B (philox), 2 = C (external)
  orc::RefNormalStream mt;
  uint64_t seed = 0;
  uint32_t call = 0;
  int particle_offset = 0;
  const double *ext = nullptr;  // [N][draws_per_particle] then 1 resampling draw
  int ext_per_particle = 0;
  // per-call draw bookkeeping
  int cur_particle = 0, cur_draw = 0;
  // outcome of the last call
  int last_neff = 0, last_resampled = 0;
  std::vector<int> ancestors;

  // one standard normal, particle_filter.cpp:25-36
  double stdNormal()
  {
    double z;
    if (mode == 0) {
      z = mt.normal(0.0, 1.0);
    } else if (mode == 2) {
      z = cur_particle < 0 ? ext[(size_t)N * ext_per_particle] : ext[(size_t)cur_particle * ext_per_particle + cur_draw];
    } else {
      double zz[2];
      const uint32_t stream = cur_particle < 0 ? orc::STREAM_RESAMPLE : (uint32_t)(particle_offset + cur_particle);
      orc::philox_normal_pair(seed, orc::DOMAIN_RBPF, call, stream, (uint32_t)(cur_draw / 2), &zz[0], &zz[1]);
      z = zz[cur_draw & 1];
    }
    cur_draw++;
    return z;
  }

  // cov.llt().matrixL() for a 3x3 (Eigen LLT, unblocked Cholesky; stops at the first non-positive pivot)
  static void cholesky3(const double a_in[3][3], double l[3][3])
  {
    double a[3][3];
    std::memcpy(a, a_in, sizeof(a));
    for (int k = 0; k < 3; k++) {
      double x = a[k][k];
      for (int q = 0; q < k; q++) x -= a[k][q] * a[k][q];
      if (x <= 0.0) break;
      x = std::sqrt(x);
      a[k][k] = x;
      for (int i = k + 1; i < 3; i++) {
        double s = 0.0;
        for (int q = 0; q < k; q++) s += a[i][q] * a[k][q];
        a[i][k] = (a[i][k] - s) / x;
      }
    }
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) l[i][j] = j <= i ? a[i][j] : 0.0;
  }
  // L * z, full 3x3 product in Eigen's order (sum over columns, left to right)
  static void lmul(const double l[3][3], const double z[3], double out[3])
  {
    for (int i = 0; i < 3; i++) out[i] = (l[i][0] * z[0] + l[i][1] * z[1]) + l[i][2] * z[2];
  }
  // particle_filter.cpp:39-60
  void sampleMultivariate(const double mu[3], const double cov[3][3], double out[3])
  {
    double z[3];
    for (int i = 0; i < 3; i++) z[i] = stdNormal();
    double l[3][3], w[3];
    cholesky3(cov, l);
    lmul(l, z, w);
    for (int i = 0; i < 3; i++) out[i] = mu ? mu[i] + w[i] : w[i];
  }

  void create(const PfParams &p)
  {
    P = p; N = p.num_particles;
    Grid proto;
    proto.init(p.laser, p.res, p.xmin, p.xmax, p.ymin, p.ymax);
    set.clear();
    set.reserve(N);
    for (int i = 0; i < N; i++) {          // particle_filter.cpp:125-138
      Particle q;
      q.weight = 1.0 / N;
      q.grid = proto;
      for (int c = 0; c < 3; c++) q.pose[c] = q.prev_pose[c] = p.init_pose[c];
      set.push_back(q);
    }
    ancestors.assign(N, 0);
    for (int i = 0; i < N; i++) ancestors[i] = i;
  }

  // particle_filter.cpp:295-322
  void sampleMotionModel(const double u[3] /*w,vx,vy*/, double pose[3])
  {
    double cov[3][3] = {{P.motion_noise[0], 0, 0}, {0, P.motion_noise[1], 0}, {0, 0, P.motion_noise[2]}};
    double w[3];
    sampleMultivariate(nullptr, cov, w);
    const double uw = u[0], vx = u[1];
    if (almost_equal(uw, 0.0)) {
      pose[0] = normalize_angle_PI(pose[0] + w[0]);
      pose[1] += vx * std::cos(pose[0]) + w[1];
      pose[2] += vx * std::sin(pose[0]) + w[2];
    } else {
      pose[0] = normalize_angle_PI(pose[0] + uw + w[0]);
      pose[1] += (-vx / uw) * std::sin(pose[0]) + (vx / uw) * std::sin(pose[0] + uw) + w[1];
      pose[2] += (vx / uw) * std::cos(pose[0]) - (vx / uw) * std::cos(pose[0] + uw) + w[2];
    }
  }

  // particle_filter.cpp:383-437; false where pdfNormal throws
  bool poseLikelihoodOdom(const double cur[3], const double prev[3], const double co[3], const double po[3], double *out) const
  {
    const double a1 = P.srr, a2 = P.srt, a3 = P.str, a4 = P.stt;
    const double rot1 = std::atan2(co[2] - po[2], co[1] - po[1]) - po[0];
    const double trans = std::sqrt(std::pow(co[1] - po[1], 2) + std::pow(co[2] - po[2], 2));
    const double rot2 = normalize_angle_PI(normalize_angle_PI(co[0]) - normalize_angle_PI(po[0]) - rot1);
    const double rot1_hat = std::atan2(cur[2] - prev[2], cur[1] - prev[1]) - prev[0];
    const double trans_hat = std::sqrt(std::pow(cur[1] - prev[1], 2) + std::pow(cur[2] - prev[2], 2));
    const double rot2_hat = normalize_angle_PI(normalize_angle_PI(cur[0]) - normalize_angle_PI(prev[0]) - rot1_hat);
    const double temp1 = a1 * rot1_hat * rot1_hat + a2 * trans_hat * trans_hat;
    const double temp2 = a3 * trans_hat * trans_hat + a4 * rot1_hat * rot1_hat + a4 * rot2_hat * rot2_hat;
    const double temp3 = a1 * rot2_hat * rot2_hat + a2 * trans_hat * trans_hat;
    double p1, p2, p3;
    if (!pdfNormal(normalize_angle_PI(normalize_angle_PI(rot1) - normalize_angle_PI(rot1_hat)), temp1, &p1)) return false;
    if (!pdfNormal(trans - trans_hat, temp2, &p2)) return false;
    if (!pdfNormal(normalize_angle_PI(normalize_angle_PI(rot2) - normalize_angle_PI(rot2_hat)), temp3, &p3)) return false;
    *out = p1 * p2 * p3;
    return true;
  }

  // particle_filter.cpp:442-458
  void normalizeWeights()
  {
    double sum = 0.0;
    for (const Particle &p : set) sum += p.weight;
    normal_sqrd_sum = 0.0;
    for (Particle &p : set) {
      p.weight /= sum;
      normal_sqrd_sum += std::pow(p.weight, 2);
    }
  }
  // particle_filter.cpp:461-465
  bool effectiveParticles()
  {
    last_neff = (int)(1.0 / normal_sqrd_sum);
    return last_neff < (N / 2);
  }
  // particle_filter.cpp:468-500
  void lowVarianceResampling()
  {
    std::vector<Particle> temp;
    cur_particle = -1; cur_draw = 0;
    const double v = stdNormal();
    const double r = v / (double)N;
    double c = set.at(0).weight;
    int i = 0;
    for (int m = 0; m < N; m++) {
      const double U = r + (double)(m * (1.0 / (N - 1)));
      while (U > c) {
        i++;
        if (i > N - 1) { i = N - 1; break; }
        c += set.at(i).weight;
      }
      temp.push_back(set.at(i));
      ancestors[m] = i;
    }
    set = temp;
  }

  // particle_filter.cpp:141-251.  0 ok, 1 off-map (std::invalid_argument from the grid), 2 "eta is 0" / zero variance
  int slam(const float *scan, int n, const double u[3], const double cur_od[3], const double prev_od[3], int icp_ok,
           const double icp_pose[3])
  {
    for (int i = 0; i < N; i++) ancestors[i] = i;
    last_resampled = 0;
    Tf Ticp;
    if (icp_ok) Ticp = Tf(icp_pose[1], icp_pose[2], icp_pose[0]);
    for (int pi = 0; pi < N; pi++) {
      Particle &q = set[pi];
      cur_particle = pi; cur_draw = 0;
      if (!icp_ok) {
        for (int c = 0; c < 3; c++) q.prev_pose[c] = q.pose[c];
        sampleMotionModel(u, q.pose);
        const Tf T_pose(q.pose[1], q.pose[2], q.pose[0]);
        double lik;
        if (q.grid.likelihoodFieldModel(scan, n, T_pose, &lik)) return 1;
        q.weight *= lik;
      } else {
        Tf T_x(q.pose[1], q.pose[2], q.pose[0]);
        T_x.mul(Ticp);                                               // :181-183
        // sampleMode, :504-519
        const double mu0[3] = {T_x.theta, T_x.x, T_x.y};
        double cov[3][3] = {{P.sample_range[0], 0, 0}, {0, P.sample_range[1], 0}, {0, 0, P.sample_range[2]}};
        std::vector<double> samples((size_t)P.k * 3);
        for (int s = 0; s < P.k; s++) {
          double x[3];
          sampleMultivariate(mu0, cov, x);
          x[0] = normalize_angle_PI(x[0]);
          for (int c = 0; c < 3; c++) samples[s * 3 + c] = x[c];
        }
        // gaussianProposal, :522-599
        double mu[3] = {0, 0, 0}, sigma[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, eta = 0.0;
        std::vector<double> lik(P.k);
        for (int s = 0; s < P.k; s++) {
          const double *xj = &samples[s * 3];
          const Tf Txj(xj[1], xj[2], xj[0]);
          double p_scan, p_pose;
          if (q.grid.likelihoodFieldModel(scan, n, Txj, &p_scan)) return 1;
          if (!poseLikelihoodOdom(xj, q.prev_pose, cur_od, prev_od, &p_pose)) return 2;
          p_scan = std::clamp(p_scan, P.scan_min, P.scan_max);
          p_pose = std::clamp(p_pose, P.pose_min, P.pose_max);
          const double p = p_scan * p_pose;
          lik[s] = p;
          for (int c = 0; c < 3; c++) mu[c] += xj[c] * p;
          eta += p;
        }
        if (almost_equal(eta, 0.0)) return 2;
        for (int c = 0; c < 3; c++) mu[c] /= eta;
        mu[0] = normalize_angle_PI(mu[0]);
        for (int s = 0; s < P.k; s++) {
          const double *xj = &samples[s * 3];
          const double d[3] = {xj[0] - mu[0], xj[1] - mu[1], xj[2] - mu[2]};
          for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) sigma[a][b] += (d[a] * d[b]) * lik[s];
        }
        for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) sigma[a][b] /= eta;
        double np[3];
        sampleMultivariate(mu, sigma, np);                           // :214
        for (int c = 0; c < 3; c++) { q.prev_pose[c] = q.pose[c]; q.pose[c] = np[c]; }
        q.weight *= eta;
      }
      const Tf Pp(q.pose[1], q.pose[2], q.pose[0]);
      if (q.grid.integrateScan(scan, n, Pp)) return 1;               // :237-239
    }
    normalizeWeights();
    if (effectiveParticles()) {
      last_resampled = 1;
      lowVarianceResampling();
    }
    call++;
    return 0;
  }

  int best() const   // particle_filter.cpp:255-274
  {
    double w = 0.0;
    int idx = 0;
    for (int i = 0; i < N; i++) if (set[i].weight > w) { w = set[i].weight; idx = i; }
    return idx;
  }
};

} // namespace

extern "C" {

struct orc_pf_params
{
  float beam_min, beam_max, beam_delta, range_min, range_max;
  double z_hit, z_short, z_max, z_rand, sigma_hit;
  double resolution, xmin, xmax, ymin, ymax;
  int32_t num_particles, k;
  double srr, srt, str, stt;
  double motion_noise[3], sample_range[3];
  double scan_min, scan_max, pose_min, pose_max;
  double init_pose[3];
};

void *orc_pf_create(const orc_pf_params *q)
{
  PfParams p;
  p.laser = Laser{q->beam_min, q->beam_max, q->beam_delta, q->range_min, q->range_max, q->z_hit, q->z_short, q->z_max, q->z_rand, q->sigma_hit};
  p.res = q->resolution; p.xmin = q->xmin; p.xmax = q->xmax; p.ymin = q->ymin; p.ymax = q->ymax;
  p.num_particles = q->num_particles; p.k = q->k;
  p.srr = q->srr; p.srt = q->srt; p.str = q->str; p.stt = q->stt;
  for (int i = 0; i < 3; i++) { p.motion_noise[i] = q->motion_noise[i]; p.sample_range[i] = q->sample_range[i]; p.init_pose[i] = q->init_pose[i]; }
  p.scan_min = q->scan_min; p.scan_max = q->scan_max; p.pose_min = q->pose_min; p.pose_max = q->pose_max;
  Pf *f = new Pf();
  f->create(p);
  return f;
}
void orc_pf_destroy(void *h) { delete static_cast<Pf *>(h); }
void orc_pf_noise_mt19937(void *h, uint64_t seed) { Pf *f = static_cast<Pf *>(h); f->mode = 0; f->mt.seed(seed); }
void orc_pf_noise_philox(void *h, uint64_t seed, uint32_t first_call) { Pf *f = static_cast<Pf *>(h); f->mode = 1; f->seed = seed; f->call = first_call; }
void orc_pf_noise_external(void *h, const double *z, int per_particle) { Pf *f = static_cast<Pf *>(h); f->mode = 2; f->ext = z; f->ext_per_particle = per_particle; }
void orc_pf_set_shard(void *h, int particle_offset) { static_cast<Pf *>(h)->particle_offset = particle_offset; }
int orc_pf_grid_size(void *h, int *xs, int *ys)
{
  Pf *f = static_cast<Pf *>(h);
  if (xs) *xs = f->set[0].grid.xsize;
  if (ys) *ys = f->set[0].grid.ysize;
  return f->set[0].grid.xsize * f->set[0].grid.ysize;
}
int orc_pf_slam(void *h, const float *scan, int n, const double twist[3], const double cur_odom[3], const double prev_odom[3],
                int icp_ok, const double icp_pose[3])
{
  return static_cast<Pf *>(h)->slam(scan, n, twist, cur_odom, prev_odom, icp_ok, icp_pose);
}
void orc_pf_get(void *h, double *weights, double *poses, double *prev_poses)
{
  Pf *f = static_cast<Pf *>(h);
  for (int i = 0; i < f->N; i++) {
    if (weights) weights[i] = f->set[i].weight;
    for (int c = 0; c < 3; c++) {
      if (poses) poses[i * 3 + c] = f->set[i].pose[c];
      if (prev_poses) prev_poses[i * 3 + c] = f->set[i].prev_pose[c];
    }
  }
}
void orc_pf_set_weights(void *h, const double *w) { Pf *f = static_cast<Pf *>(h); for (int i = 0; i < f->N; i++) f->set[i].weight = w[i]; }
void orc_pf_set_poses(void *h, const double *p)
{
  Pf *f = static_cast<Pf *>(h);
  for (int i = 0; i < f->N; i++) for (int c = 0; c < 3; c++) f->set[i].pose[c] = p[i * 3 + c];
}
void orc_pf_get_resample(void *h, int *neff, int *resampled, int *ancestors)
{
  Pf *f = static_cast<Pf *>(h);
  if (neff) *neff = f->last_neff;
  if (resampled) *resampled = f->last_resampled;
  if (ancestors) std::copy(f->ancestors.begin(), f->ancestors.end(), ancestors);
}
// normalise + N_eff test + walk on the current weights (particle_filter.cpp:244-249)
int orc_pf_normalize_resample(void *h, int *resampled, int *ancestors)
{
  Pf *f = static_cast<Pf *>(h);
  for (int i = 0; i < f->N; i++) f->ancestors[i] = i;
  f->normalizeWeights();
  f->last_resampled = f->effectiveParticles() ? 1 : 0;
  if (f->last_resampled) f->lowVarianceResampling();
  f->call++;
  if (resampled) *resampled = f->last_resampled;
  if (ancestors) std::copy(f->ancestors.begin(), f->ancestors.end(), ancestors);
  return 0;
}
void orc_pf_robot_state(void *h, double out[3])
{
  Pf *f = static_cast<Pf *>(h);
  const int b = f->best();
  for (int c = 0; c < 3; c++) out[c] = f->set[b].pose[c];
}
void orc_pf_new_map(void *h, int8_t *out) { Pf *f = static_cast<Pf *>(h); f->set[f->best()].grid.gridMap(out); }

// ---- per-particle grid access (a filter with one particle doubles as a stand-alone GridMapper) ----
void orc_pf_grid_dump(void *h, int particle, double *log_odds, double *prob, double *occ_dist, int *state)
{
  const Grid &g = static_cast<Pf *>(h)->set[particle].grid;
  const size_t n = g.log_odds.size();
  for (size_t i = 0; i < n; i++) {
    if (log_odds) log_odds[i] = g.log_odds[i];
    if (prob) prob[i] = g.prob[i];
    if (occ_dist) occ_dist[i] = g.occ_dist[i];
    if (state) state[i] = g.state[i];
  }
}
int orc_pf_grid_occ_order(void *h, int particle, int *keys, int cap)
{
  const Grid &g = static_cast<Pf *>(h)->set[particle].grid;
  int n = 0;
  g.occ.for_each([&](int k) { if (n < cap) keys[n] = k; n++; });
  return n;
}
int orc_pf_grid_bucket_count(void *h, int particle) { return (int)static_cast<Pf *>(h)->set[particle].grid.occ.bucket_count; }
int orc_pf_grid_likelihood(void *h, int particle, const float *scan, int n, const double pose[3], double *p)
{
  return static_cast<Pf *>(h)->set[particle].grid.likelihoodFieldModel(scan, n, Tf(pose[1], pose[2], pose[0]), p);
}
int orc_pf_grid_integrate(void *h, int particle, const float *scan, int n, const double pose[3])
{
  return static_cast<Pf *>(h)->set[particle].grid.integrateScan(scan, n, Tf(pose[1], pose[2], pose[0]));
}
int orc_pf_grid_end_points(void *h, int particle, const float *scan, int n, const double pose[3], double *xy)
{
  std::vector<Vec2> pts;
  static_cast<Pf *>(h)->set[particle].grid.laserEndPoints(pts, scan, n, Tf(pose[1], pose[2], pose[0]));
  for (size_t i = 0; i < pts.size(); i++) { xy[2 * i] = pts[i].x; xy[2 * i + 1] = pts[i].y; }
  return (int)pts.size();
}
int orc_pf_grid_free_cells(void *h, int particle, const double pt[2], const double pose[3], int *cells, int cap)
{
  std::vector<int> idx;
  if (!static_cast<Pf *>(h)->set[particle].grid.freeGridIndex(idx, Vec2{pt[0], pt[1]}, Tf(pose[1], pose[2], pose[0]))) return -1;
  for (size_t i = 0; i < idx.size() && (int)i < cap; i++) cells[i] = idx[i];
  return (int)idx.size();
}
void orc_pf_grid_map(void *h, int particle, int8_t *out) { static_cast<Pf *>(h)->set[particle].grid.gridMap(out); }
// ray cells, brushfire iterations, pushes and the largest heap seen so far by one particle's grid
void orc_pf_grid_stats(void *h, int particle, uint64_t out[4])
{
  const GridStats &s = static_cast<Pf *>(h)->set[particle].grid.stats;
  out[0] = s.ray_cells; out[1] = s.esdf_iterations; out[2] = s.esdf_pushes; out[3] = s.heap_max;
}

// ---- self-tests of the two standard-library restatements against the real containers ------------
// random insert/erase traffic; returns the number of steps after which the iteration orders differed (0 = never)
int orc_selftest_occset(uint64_t seed, int keyspace, int steps)
{
  std::mt19937_64 g(seed);
  std::unordered_set<int> real;
  OccSet emu;
  emu.init(keyspace);
  for (int s = 1; s <= steps; s++) {
    const int key = (int)(g() % (uint64_t)keyspace);
    // bias towards growth so that several rehashes happen
    if (g() % 4 != 0) { if (real.find(key) == real.end()) real.insert(key); emu.insert(key); }
    else { if (real.find(key) != real.end()) real.erase(key); emu.erase(key); }
    if (s % 16 == 0 || s == steps) {
      std::vector<int> a, b;
      for (int k : real) a.push_back(k);
      emu.for_each([&](int k) { b.push_back(k); });
      if (a != b || real.bucket_count() != emu.bucket_count) return s;
    }
    if (s % 64 == 0) {   // copies keep order and bucket count (resampling copies whole particles)
      std::unordered_set<int> copy(real);
      std::vector<int> a, b;
      for (int k : copy) a.push_back(k);
      for (int k : real) b.push_back(k);
      if (a != b || copy.bucket_count() != real.bucket_count()) return -s;
      real = copy;
    }
  }
  return 0;
}
// random push/pop traffic with heavy ties; returns the first step at which top() differed (0 = never)
int orc_selftest_heap(uint64_t seed, int steps, int distinct)
{
  struct C { double occ_dist; int tag; };
  struct Cmp { bool operator()(const C &a, const C &b) const { return a.occ_dist > b.occ_dist; } };
  std::mt19937_64 g(seed);
  std::priority_queue<C, std::vector<C>, Cmp> real;
  MinHeap emu;
  for (int s = 1; s <= steps; s++) {
    if (real.empty() || g() % 5 < 3) {
      const double d = (double)(g() % (uint64_t)distinct) * 0.05;
      real.push(C{d, s});
      emu.push(HeapCell{d, s, 0, 0, 0});
    } else {
      real.pop();
      emu.pop();
    }
    if (real.size() != emu.v.size()) return s;
    if (!real.empty() && (real.top().tag != emu.top().i || real.top().occ_dist != emu.top().occ_dist)) return s;
  }
  while (!real.empty()) {
    if (real.top().tag != emu.top().i) return steps + 1;
    real.pop(); emu.pop();
  }
  return 0;
}

} // extern "C"
