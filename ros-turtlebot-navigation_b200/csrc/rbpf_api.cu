// rbpf_api.cu - extern "C" RBPF entry points of libb2nav (see include/b2nav.h).
// Host side of bmapping::ParticleFilter (reference: bmapping/src/bmapping/particle_filter.cpp:64-138,141-291),
// bmapping::GridMapper's constructor constants (bmapping/src/bmapping/grid_mapper.cpp:37-64) and
// bmapping::LaserScanner's beam-angle sequence (bmapping/src/bmapping/sensor_model.cpp:66-108).
//
// Compiled with -fmad=false (see rbpf_kernels.cuh).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include <nccl.h>

#include "rbpf_kernels.cuh"

using namespace b2n;

struct b2n_pf
{
  b2n_pf_params p;
  int N = 0, n_total = 0, offset = 0, device = 0, n_sm = 0;
  int max_beams = 0;
  PfConst c;
  PfPlanes set[2];
  int cur = 0;
  double l_prior = 0.0, max_occ_dist = 10.0;

  cudaStream_t own_stream = nullptr, stream = nullptr;
  float *d_scan = nullptr, *h_scan = nullptr;
  double *d_beam_cs = nullptr, *d_pz = nullptr;
  int *d_status = nullptr;            // [0] status bits, [1] N_eff, [2] resampled, [3] best particle's index in its owner's set, [4] owner rank;
                                      // [8 ..] every rank's status bits (sharded: gathered next to the weights)
  PfPlanes *d_own_sets = nullptr;     // device copy of set[0], set[1]
  bool global_w_valid = false;        // d_w / d_anc describe the particle set (set by SLAM, cleared by the weight tap)
  int *h_status = nullptr;            // pinned mirror
  double *d_w = nullptr;              // [n_total]
  double *d_cum = nullptr;            // [n_total] cumulative normalised weights (the walk's c)
  int32_t *d_anc = nullptr;           // [n_total]
  std::vector<int32_t> h_anc;
  double *d_ext = nullptr;
  size_t ext_cap = 0, ext_count = 0;
  bool ext_armed = false;
  double *d_samples = nullptr;        // [N][k][4]
  unsigned long long *d_spill = nullptr, *d_stats = nullptr;
  // distance-field launch shape (configure_df): one CTA per SM, df_warps particles in flight per CTA
  int df_grid = 0, df_warps = 1, df_hcap = 0, df_hcap_request = 0, df_gcap = 0, df_cols = 0;
  int df_gl_active = 32;
  int df_gl = 16;                     // lanes per particle in the distance-field kernel: 32 (one particle per warp), 16 or 8
  bool df_tmem = true, df_tmem_active = true;
  bool df_skip_clean = true;          // do not regrow a distance field whose occupied set did not change (B2N_PF_DF_ALWAYS=1: always)
  size_t df_smem = 0, spill_entries = 0;
  size_t smem_optin = 0;
  double *d_best = nullptr, *h_best = nullptr;   // pose[3], weight
  int8_t *d_map = nullptr;
  double *d_lik = nullptr;
  int32_t *d_idx = nullptr;           // [2][N] copy lists of a cross-rank resampling
  // peer-memory migration (CUDA IPC): every rank's plane allocations of both sets mapped into this process
  bool p2p_ready = false;
  std::vector<void *> peer_ptrs;      // [nranks][2 sets][5 planes], own entries = own pointers
  PfPlanes *d_peer_sets = nullptr;    // device: [2 sets][nranks]
  int last_migrated_in = 0, last_migrated_out = 0;

  uint64_t seed = 0;
  uint32_t call = 0;
  int last_neff = 0, last_resampled = 0;
  uint64_t launches = 0;

  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;

  bool timing = false;
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  float last_ms[3] = {0.f, 0.f, 0.f};   // update, distance field, normalise + resample
};

namespace
{

// grid_mapper.hpp:27-38 with the HOST's libm: the same functions the reference evaluates
double logOdds2Prob(double l) { return 1 - (1 / (1 + std::exp(l))); }
double prob2LogOdds(double p) { return std::log(p / (1 - p)); }

// doubles ordered as integers, for bisection over representable values
int64_t ordered(double d) { int64_t i; std::memcpy(&i, &d, 8); return i < 0 ? (int64_t)0x8000000000000000LL - i : i; }
double unordered(int64_t i) { if (i < 0) i = (int64_t)0x8000000000000000LL - i; double d; std::memcpy(&d, &i, 8); return d; }

// smallest double l in [lo, hi] with pred(l) true, pred monotone false -> true
template <class F>
double first_true(double lo, double hi, F pred)
{
  int64_t a = ordered(lo), b = ordered(hi);
  while (a < b) {
    const int64_t m = (a >> 1) + (b >> 1) + (a & b & 1);   // no overflow: the span exceeds INT64_MAX
    if (pred(unordered(m))) b = m; else a = m + 1;
  }
  return unordered(a);
}

unsigned int mapSize(double lower, double upper, double res) { return static_cast<unsigned int>(std::ceil((upper - lower) / res)); }

int set_device(const b2n_pf *h)
{
  B2N_CUDA(cudaSetDevice(h->device));
  return B2N_OK;
}

void free_planes(PfPlanes &s)
{
  cudaFree(s.log_odds); cudaFree(s.d2); cudaFree(s.nxt); cudaFree(s.bkt); cudaFree(s.meta);
  s = PfPlanes{};
}

__global__ void pf_init_kernel(const __grid_constant__ PfConst c, PfPlanes pl, double l_prior, double weight, double th, double x, double y)
{
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  const size_t cells = (size_t)c.N * c.gstride;
  for (size_t i = tid; i < cells; i += nth) { pl.log_odds[i] = l_prior; pl.d2[i] = kD2Unreached; }
  for (size_t i = tid; i < (size_t)c.N * c.nxt_stride; i += nth) pl.nxt[i] = kNil16;
  for (size_t i = tid; i < (size_t)c.N * c.bkt_stride; i += nth) pl.bkt[i] = kNil16;
  for (size_t i = tid; i < (size_t)c.N; i += nth) {
    PfParticle q;
    q.pose[0] = q.prev_pose[0] = th; q.pose[1] = q.prev_pose[1] = x; q.pose[2] = q.prev_pose[2] = y;
    q.weight = weight;
    q.n_occ = 0; q.bucket_count = 1; q.next_resize = 0; q.chain = -1;   // an empty std::unordered_set
    q.occ_dirty = 0;
    pl.meta[i] = q;
  }
}

int read_meta(b2n_pf *h, std::vector<PfParticle> &m)
{
  m.resize(h->N);
  B2N_CUDA(cudaStreamSynchronize(h->stream));
  B2N_CUDA(cudaMemcpy(m.data(), h->set[h->cur].meta, sizeof(PfParticle) * h->N, cudaMemcpyDeviceToHost));
  return B2N_OK;
}

PfCall make_call(b2n_pf *h, int n_beams)
{
  PfCall q;
  std::memset(&q, 0, sizeof(q));
  q.scan = h->d_scan; q.n_beams = n_beams;
  q.seed_lo = (uint32_t)h->seed; q.seed_hi = (uint32_t)(h->seed >> 32); q.call = h->call;
  q.particle_offset = h->offset;
  q.status = h->d_status;
  return q;
}

int upload_scan(b2n_pf *h, const float *scan, int n_beams)
{
  B2N_REQUIRE(scan && n_beams > 0, B2N_ERR_INVALID_ARGUMENT, "empty scan");
  B2N_REQUIRE(n_beams <= h->max_beams, B2N_ERR_INVALID_ARGUMENT, "scan has %d beams, handle was created for %d", n_beams, h->max_beams);
  std::memcpy(h->h_scan, scan, sizeof(float) * n_beams);
  B2N_CUDA(cudaMemcpyAsync(h->d_scan, h->h_scan, sizeof(float) * n_beams, cudaMemcpyHostToDevice, h->stream));
  return B2N_OK;
}

// Where every slot of THIS rank takes its particle from after a resampling that every rank computed identically
// (SURVEY.md 8e): local ancestors are copied on the device; an ancestor living on another rank is received ONCE per
// (ancestor, destination rank) into the first slot that wants it and fanned out from there; symmetric send list.
// Order of the exchange for a pair (s -> t): ascending global ancestor index - both sides derive it from the same
// ancestor vector, so sends and receives match without any negotiation.
struct MigrationPlan
{
  std::vector<int32_t> copy1;       // [n_local] local source index in the OLD set, -1 = not a local copy
  std::vector<int32_t> copy2;       // [n_local] local slot in the NEW set to copy from (second pass), -1 = none
  std::vector<int32_t> recv_slot, recv_anc, recv_rank;   // receive into local slot <- global ancestor on rank
  std::vector<int32_t> send_idx, send_rank;              // send local particle index -> rank
};

void plan_migration(const int32_t *anc, int n_total, int n_local, int rank, int nranks, MigrationPlan &pl)
{
  pl.copy1.assign(n_local, -1); pl.copy2.assign(n_local, -1);
  pl.recv_slot.clear(); pl.recv_anc.clear(); pl.recv_rank.clear(); pl.send_idx.clear(); pl.send_rank.clear();
  const int off = rank * n_local;
  // receives: my slots whose ancestor is remote; ancestors are non-decreasing in m (low-variance walk), so equal
  // ancestors are adjacent, but do not rely on it: remember the first slot per remote ancestor
  std::vector<int32_t> first_slot(n_total, -1);
  for (int m = 0; m < n_local; m++) {
    const int a = anc[off + m];
    const int s = a / n_local;
    if (s == rank) { pl.copy1[m] = a - off; continue; }
    if (first_slot[a] < 0) first_slot[a] = m;
    else pl.copy2[m] = first_slot[a];
  }
  for (int s = 0; s < nranks; s++) {
    if (s == rank) continue;
    for (int a = s * n_local; a < (s + 1) * n_local; a++)
      if (first_slot[a] >= 0) { pl.recv_slot.push_back(first_slot[a]); pl.recv_anc.push_back(a); pl.recv_rank.push_back(s); }
  }
  // sends: for every other rank t, my particles that some slot of t wants, ascending
  std::vector<char> wanted(n_local);
  for (int t = 0; t < nranks; t++) {
    if (t == rank) continue;
    std::fill(wanted.begin(), wanted.end(), 0);
    for (int m = t * n_local; m < (t + 1) * n_local; m++) {
      const int a = anc[m];
      if (a / n_local == rank) wanted[a - off] = 1;
    }
    for (int i = 0; i < n_local; i++)
      if (wanted[i]) { pl.send_idx.push_back(i); pl.send_rank.push_back(t); }
  }
}

// Launch shape of the distance-field kernel (rbpf_kernels.cuh): one CTA per SM, one particle per warp, as many warps
// as it takes to hold every particle of this handle in flight at once (at most 28 = 4144 chains on 148 SMs); the heap
// of each warp gets an equal share of the SM's shared memory, the visited bitmap goes to tensor memory when each
// warp's 32-lane slice of the 512 columns can hold it.
int configure_df(b2n_pf *h)
{
  const PfConst &c = h->c;
  const int words = (c.G + 31) / 32;
  const int cols_needed = (words + 31) / 32 + 1;
  const size_t budget = h->smem_optin - 2048;     // static shared memory and alignment slack
  const int per_sm = std::max(1, std::min(kDfMaxWarps, (h->N + h->n_sm - 1) / h->n_sm));   // particles in flight per SM
  int gl = h->df_tmem ? h->df_gl : 32;
  if (gl != 32) {
    // several particles per warp (rbpf_distance_field_groups_kernel)
    const int ng = 32 / gl;
    const int W = (per_sm + ng - 1) / ng;
    const int cols_per_warp = ng * cols_needed;
    int hcap = 0;
    const size_t stage_bytes = (size_t)W * ng * (64 * 4 + 4 * 8);
    if (((W + 3) / 4) * cols_per_warp <= 512 && budget > stage_bytes + 1024) {
      hcap = (int)((budget - stage_bytes) / ((size_t)W * ng) / 8) - 2;
      hcap = std::min(hcap & ~1, 16384);
    }
    if (hcap >= 64) {
      if (h->df_hcap_request > 0) hcap = std::min(hcap, h->df_hcap_request);
      h->df_warps = W; h->df_hcap = hcap; h->df_cols = cols_per_warp; h->df_gl_active = gl;
      h->df_grid = std::max(1, std::min(h->n_sm, (h->N + W * ng - 1) / (W * ng)));
      h->df_gcap = std::min(c.G, 16384);
      h->df_smem = pf_dfg_smem_bytes(hcap, W, ng);
      h->df_tmem_active = true;
      const size_t need = (size_t)h->df_grid * W * ng * h->df_gcap;
      if (need > h->spill_entries) {
        cudaFree(h->d_spill); h->d_spill = nullptr; h->spill_entries = 0;
        B2N_CUDA(cudaMalloc(&h->d_spill, need * sizeof(unsigned long long)));
        h->spill_entries = need;
      }
      return B2N_OK;
    }
  }
  // one particle per warp (rbpf_distance_field_kernel)
  int W = per_sm;
  bool tmem = h->df_tmem;
  int hcap = 0, cols = 0;
  for (;; W--) {
    cols = 512 / ((W + 3) / 4);
    const bool t = tmem && cols_needed <= cols;
    const size_t per_warp = budget / W;
    const size_t marks = t ? 0 : (size_t)words * 4;
    hcap = per_warp > marks + 16 ? (int)((per_warp - marks) / 8) - 2 : 0;
    hcap = std::min(hcap & ~1, 16384);
    if (hcap >= 64 || W == 1) { tmem = t; break; }
  }
  B2N_REQUIRE(hcap >= 2, B2N_ERR_UNSUPPORTED, "the distance-field kernel does not fit this map in shared memory");
  if (h->df_hcap_request > 0) hcap = std::min(hcap, h->df_hcap_request);
  h->df_warps = W; h->df_hcap = hcap; h->df_cols = cols; h->df_gl_active = 32;
  h->df_grid = std::max(1, std::min(h->n_sm, (h->N + W - 1) / W));
  h->df_gcap = std::min(c.G, 16384);
  h->df_smem = pf_df_smem_bytes(c.G, hcap, W, tmem);
  h->df_tmem_active = tmem;
  const size_t need = (size_t)h->df_grid * W * h->df_gcap;
  if (need > h->spill_entries) {
    cudaFree(h->d_spill); h->d_spill = nullptr; h->spill_entries = 0;
    B2N_CUDA(cudaMalloc(&h->d_spill, need * sizeof(unsigned long long)));
    h->spill_entries = need;
  }
  return B2N_OK;
}

// Everything create() derives on the host from the constructor arguments: geometry, log-odds steps and thresholds,
// the beam cos/sin table and the likelihood table.  No device needed (b2n_pf_host_tables exposes it to CPU tests).
void host_tables(const b2n_pf_params &p, int xsize, int ysize, long long G, int max_beams, PfConst &c, double &l_prior,
                 std::vector<double> &beam, std::vector<double> &pz)
{
  const double max_occ_dist = 10.0;                          // grid_mapper.cpp:49
  // ---- constants of GridMapper's constructor (grid_mapper.cpp:37-64) ----------------------------------------
  std::memset(&c, 0, sizeof(c));
  c.N = p.num_particles; c.G = (int)G; c.xsize = xsize; c.ysize = ysize;
  c.gstride = (int)((G + 3) & ~3LL);
  c.nxt_stride = (int)((G + 1 + 7) & ~7LL);
  int chain = 0;
  static const uint32_t kChain[] = {13, 29, 59, 127, 257, 541, 1109, 2357, 5087, 10273, 20753, 42043, 85229};
  while (kChain[chain] < (uint32_t)G) chain++;              // bucket_count never exceeds the first chain value >= element count
  c.bkt_stride = (int)((kChain[chain] + 7) & ~7u);
  c.cell_radius = (int)mapSize(0.0, max_occ_dist, p.resolution);
  c.xmin = p.xmin; c.xmax = p.xmax; c.ymin = p.ymin; c.ymax = p.ymax; c.res = p.resolution;
  c.range_min = (double)p.range_min; c.range_max = (double)p.range_max;
  const double prior = 0.5, prob_occ = 0.90, prob_free = 0.35;
  l_prior = prob2LogOdds(prior);
  c.d_free = prob2LogOdds(prob_free) - l_prior;
  c.d_occ = prob2LogOdds(prob_occ) - l_prior;
  // updateCellState (grid_mapper.cpp:438-477) classifies on logOdds2Prob(l); the same decisions in log-odds space,
  // with thresholds found by bisection over the host's own exp() so that ties (a single hit gives prob == 0.9
  // exactly) fall on the reference's side
  c.t_occ = first_true(-40.0, 40.0, [&](double l) { return logOdds2Prob(l) >= prob_occ; });
  const double first_above_free = first_true(-40.0, 40.0, [&](double l) { return !(logOdds2Prob(l) <= prob_free); });
  c.t_free = std::nextafter(first_above_free, -1.0e300);
  for (int i = 0; i < 3; i++) {
    const double mv[3] = {p.motion_noise_theta, p.motion_noise_x, p.motion_noise_y};
    const double sv[3] = {p.sample_range_theta, p.sample_range_x, p.sample_range_y};
    c.sig[i] = mv[i] > 0.0 ? std::sqrt(mv[i]) : 0.0;        // Eigen LLT stops at a non-positive pivot: L stays 0
    c.sig_mode[i] = sv[i] > 0.0 ? std::sqrt(sv[i]) : 0.0;
  }
  // LLT of a diagonal matrix stops at the FIRST non-positive pivot and leaves the later rows untouched (= the
  // variances themselves on the diagonal); reproduce that corner for completeness
  {
    const double mv[3] = {p.motion_noise_theta, p.motion_noise_x, p.motion_noise_y};
    const double sv[3] = {p.sample_range_theta, p.sample_range_x, p.sample_range_y};
    bool stop = false;
    for (int i = 0; i < 3; i++) { if (!(mv[i] > 0.0)) stop = true; if (stop) c.sig[i] = mv[i]; }
    stop = false;
    for (int i = 0; i < 3; i++) { if (!(sv[i] > 0.0)) stop = true; if (stop) c.sig_mode[i] = sv[i]; }
  }
  c.k_samples = p.k;
  c.srr = p.srr; c.srt = p.srt; c.str = p.str; c.stt = p.stt;
  c.scan_min = p.scan_likelihood_min; c.scan_max = p.scan_likelihood_max;
  c.pose_min = p.pose_likelihood_min; c.pose_max = p.pose_likelihood_max;

  // ---- beam table: the accumulated angle of sensor_model.cpp:66-108 and its host cos/sin ----------------------
  beam.assign(2 * (size_t)max_beams, 0.0);
  {
    double beam_angle = p.beam_min;
    for (int i = 0; i < max_beams; i++) {
      beam[2 * i] = std::cos(beam_angle);
      beam[2 * i + 1] = std::sin(beam_angle);
      beam_angle += p.beam_delta;
      if (p.beam_max < 0.0 && beam_angle <= p.beam_max) beam_angle = p.beam_min;
      else if (p.beam_max >= 0.0 && beam_angle >= p.beam_max) beam_angle = p.beam_min;
    }
  }
  // ---- likelihood table over every distance the brushfire can produce (grid_mapper.cpp:18-28,101-128) ---------
  const int R2 = c.cell_radius * c.cell_radius;
  c.pz_n = R2 + 2;
  c.pz_stage = std::min(2048, c.pz_n & ~1);
  pz.assign((size_t)c.pz_n, 0.0);
  {
    const double PI = 3.14159265358979323846;
    const double var_hit = p.sigma_hit * p.sigma_hit;
    auto term = [&](double z) {
      const double sqrt_inv = 1.0 / std::sqrt(2.0 * PI * var_hit);
      const double var = -0.5 * (z * z) / var_hit;
      double pzv = 0.0;
      pzv += p.z_hit * (sqrt_inv * std::exp(var));
      pzv += p.z_rand / p.z_max;
      return pzv;
    };
    for (int d2 = 0; d2 <= R2; d2++) pz[d2] = term(std::sqrt((double)d2) * p.resolution);
    pz[R2 + 1] = term(max_occ_dist);
  }

}

// Resampling across ranks: device copies for local ancestors, one ncclSend/ncclRecv group for the particles that
// change GPU (five contiguous pieces each: log-odds, squared distances, occupied-set links, bucket heads, meta), then a
// second device pass fans received particles out to the other slots that chose the same ancestor.
int migrate_particles(b2n_pf *h, const PfPlanes &src, const PfPlanes &dst)
{
  const PfConst &c = h->c;
  if (h->p2p_ready) {
    // every slot reads its ancestor where it lives: local HBM or a peer's HBM over NVLink; the allgather of the weights
    // that preceded the walk is the barrier that makes every rank's old set final, the next one protects it from reuse
    int in = 0;
    for (int m = 0; m < h->N; m++) in += (h->h_anc[h->offset + m] / h->N) != h->rank;
    h->last_migrated_in = in; h->last_migrated_out = 0;
    dim3 grid(8, h->N);
    rbpf_copy_particles_p2p_kernel<<<grid, 256, 0, h->stream>>>(c, h->d_peer_sets + (size_t)h->cur * h->nranks, dst, h->d_anc, h->offset, h->N, h->d_w);
    B2N_CUDA(cudaGetLastError());
    h->launches++;
    return B2N_OK;
  }
  MigrationPlan mp;
  plan_migration(h->h_anc.data(), h->n_total, h->N, h->rank, h->nranks, mp);
  if (!h->d_idx) B2N_CUDA(cudaMalloc(&h->d_idx, sizeof(int32_t) * 2 * (size_t)h->N));
  B2N_CUDA(cudaMemcpyAsync(h->d_idx, mp.copy1.data(), sizeof(int32_t) * h->N, cudaMemcpyHostToDevice, h->stream));
  B2N_CUDA(cudaMemcpyAsync(h->d_idx + h->N, mp.copy2.data(), sizeof(int32_t) * h->N, cudaMemcpyHostToDevice, h->stream));
  dim3 grid(8, h->N);
  rbpf_copy_particles_kernel<<<grid, 256, 0, h->stream>>>(c, src, dst, h->d_idx);
  B2N_CUDA(cudaGetLastError());
  h->launches++;
  const size_t n_send = mp.send_idx.size(), n_recv = mp.recv_slot.size();
  h->last_migrated_in = (int)n_recv; h->last_migrated_out = (int)n_send;
  if (n_send + n_recv > 0) {
#define B2N_NCCL(expr)                                                                             \
  do {                                                                                             \
    ncclResult_t r__ = (expr);                                                                     \
    B2N_REQUIRE(r__ == ncclSuccess, B2N_ERR_COMM, "%s: %s", #expr, ncclGetErrorString(r__));       \
  } while (0)
    B2N_NCCL(ncclGroupStart());
    for (size_t i = 0; i < n_send; i++) {
      const size_t a = (size_t)mp.send_idx[i];
      const int t = mp.send_rank[i];
      B2N_NCCL(ncclSend(src.log_odds + a * c.gstride, (size_t)c.gstride, ncclDouble, t, h->comm, h->stream));
      B2N_NCCL(ncclSend(src.d2 + a * c.gstride, (size_t)c.gstride, ncclUint32, t, h->comm, h->stream));
      B2N_NCCL(ncclSend(src.nxt + a * c.nxt_stride, (size_t)c.nxt_stride * 2, ncclUint8, t, h->comm, h->stream));
      B2N_NCCL(ncclSend(src.bkt + a * c.bkt_stride, (size_t)c.bkt_stride * 2, ncclUint8, t, h->comm, h->stream));
      B2N_NCCL(ncclSend(src.meta + a, sizeof(PfParticle), ncclUint8, t, h->comm, h->stream));
    }
    for (size_t i = 0; i < n_recv; i++) {
      const size_t m = (size_t)mp.recv_slot[i];
      const int s = mp.recv_rank[i];
      B2N_NCCL(ncclRecv(dst.log_odds + m * c.gstride, (size_t)c.gstride, ncclDouble, s, h->comm, h->stream));
      B2N_NCCL(ncclRecv(dst.d2 + m * c.gstride, (size_t)c.gstride, ncclUint32, s, h->comm, h->stream));
      B2N_NCCL(ncclRecv(dst.nxt + m * c.nxt_stride, (size_t)c.nxt_stride * 2, ncclUint8, s, h->comm, h->stream));
      B2N_NCCL(ncclRecv(dst.bkt + m * c.bkt_stride, (size_t)c.bkt_stride * 2, ncclUint8, s, h->comm, h->stream));
      B2N_NCCL(ncclRecv(dst.meta + m, sizeof(PfParticle), ncclUint8, s, h->comm, h->stream));
    }
    B2N_NCCL(ncclGroupEnd());
#undef B2N_NCCL
  }
  // fan-out of received particles inside the new set (slots are distinct from their sources)
  rbpf_copy_particles_kernel<<<grid, 256, 0, h->stream>>>(c, dst, dst, h->d_idx + h->N);
  B2N_CUDA(cudaGetLastError());
  h->launches++;
  return B2N_OK;
}

} // namespace

extern "C" {

int b2n_pf_create(const b2n_pf_params *params, b2n_pf **out)
{
  B2N_REQUIRE(params && out, B2N_ERR_INVALID_ARGUMENT, "b2n_pf_create: null argument");
  *out = nullptr;
  const b2n_pf_params &p = *params;
  B2N_REQUIRE(p.num_particles > 0, B2N_ERR_INVALID_ARGUMENT, "num_particles must be positive (got %d)", p.num_particles);
  B2N_REQUIRE(p.resolution > 0.0 && p.xmax > p.xmin && p.ymax > p.ymin, B2N_ERR_INVALID_ARGUMENT, "bad map geometry");
  B2N_REQUIRE(p.sigma_hit != 0.0 && std::fabs(p.sigma_hit * p.sigma_hit) >= 1.0e-12, B2N_ERR_INVALID_ARGUMENT,
              "sigma_hit^2 is 0 (reference: pdfNormal throws)");
  B2N_REQUIRE(p.z_max != 0.0, B2N_ERR_INVALID_ARGUMENT, "z_max must be non-zero (it divides z_rand, grid_mapper.cpp:121)");
  B2N_REQUIRE(p.motion_noise_theta >= 0 && p.motion_noise_x >= 0 && p.motion_noise_y >= 0 && p.sample_range_theta >= 0 &&
                  p.sample_range_x >= 0 && p.sample_range_y >= 0,
              B2N_ERR_INVALID_ARGUMENT, "variances must be non-negative");
  B2N_REQUIRE(p.k >= 0, B2N_ERR_INVALID_ARGUMENT, "k must be non-negative");
  const int xsize = (int)mapSize(p.xmin, p.xmax, p.resolution), ysize = (int)mapSize(p.ymin, p.ymax, p.resolution);
  // the reference decodes keys with xsize_ for both axes (grid_mapper.cpp:352-353): only square maps are meaningful
  B2N_REQUIRE(xsize == ysize, B2N_ERR_UNSUPPORTED, "map is %d x %d cells; the reference's distance field assumes a square map", xsize, ysize);
  const long long G = (long long)xsize * ysize;
  B2N_REQUIRE(G >= 4 && G <= 65534, B2N_ERR_UNSUPPORTED, "map has %lld cells; this build keeps cell ids in 16 bits (max 65534)", G);

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    set_error("b2n_pf_create: no CUDA device (libb2nav has no CPU path)");
    return B2N_ERR_CUDA;
  }
  b2n_pf *h = new (std::nothrow) b2n_pf();
  B2N_REQUIRE(h, B2N_ERR_CUDA, "out of host memory");
  h->p = p;
  h->N = p.num_particles;
  h->n_total = p.particles_total > 0 ? p.particles_total : p.num_particles;
  h->offset = p.particle_offset;
  h->max_beams = p.max_beams > 0 ? p.max_beams : 1024;
  if (p.device >= 0) h->device = p.device; else cudaGetDevice(&h->device);

#define B2N_TRY(expr)                                                                          \
  do {                                                                                         \
    cudaError_t e__ = (expr);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      set_error("%s: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__);         \
      b2n_pf_destroy(h);                                                                       \
      return B2N_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

  std::vector<double> beam, pz;
  B2N_TRY(cudaSetDevice(h->device));
  cudaDeviceProp prop;
  B2N_TRY(cudaGetDeviceProperties(&prop, h->device));
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; libb2nav is built for sm_100a only", h->device, prop.major, prop.minor);
    b2n_pf_destroy(h);
    return B2N_ERR_CUDA;
  }
  h->n_sm = prop.multiProcessorCount;
  B2N_TRY(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
  h->stream = h->own_stream;

  host_tables(p, xsize, ysize, G, h->max_beams, h->c, h->l_prior, beam, pz);
  PfConst &c = h->c;

  B2N_TRY(cudaMalloc(&h->d_beam_cs, beam.size() * sizeof(double)));
  B2N_TRY(cudaMemcpy(h->d_beam_cs, beam.data(), beam.size() * sizeof(double), cudaMemcpyHostToDevice));
  B2N_TRY(cudaMalloc(&h->d_pz, pz.size() * sizeof(double)));
  B2N_TRY(cudaMemcpy(h->d_pz, pz.data(), pz.size() * sizeof(double), cudaMemcpyHostToDevice));
  c.beam_cs = h->d_beam_cs; c.pz_table = h->d_pz;

  for (int s = 0; s < 2; s++) {
    B2N_TRY(cudaMalloc(&h->set[s].log_odds, (size_t)h->N * c.gstride * sizeof(double)));
    B2N_TRY(cudaMalloc(&h->set[s].d2, (size_t)h->N * c.gstride * sizeof(uint32_t)));
    B2N_TRY(cudaMalloc(&h->set[s].nxt, (size_t)h->N * c.nxt_stride * sizeof(uint16_t)));
    B2N_TRY(cudaMalloc(&h->set[s].bkt, (size_t)h->N * c.bkt_stride * sizeof(uint16_t)));
    B2N_TRY(cudaMalloc(&h->set[s].meta, (size_t)h->N * sizeof(PfParticle)));
  }
  // initParticleSet (particle_filter.cpp:125-138): weight 1/N over the WHOLE filter, pose = init pose
  pf_init_kernel<<<h->n_sm * 4, 256, 0, h->stream>>>(c, h->set[0], h->l_prior, 1.0 / h->n_total, p.init_pose[0], p.init_pose[1], p.init_pose[2]);
  B2N_TRY(cudaGetLastError());
  h->launches++;

  B2N_TRY(cudaMalloc(&h->d_scan, sizeof(float) * h->max_beams));
  B2N_TRY(cudaMallocHost(&h->h_scan, sizeof(float) * h->max_beams));
  B2N_TRY(cudaMalloc(&h->d_status, (8 + kPfMaxRanks) * sizeof(int)));
  B2N_TRY(cudaMemsetAsync(h->d_status, 0, (8 + kPfMaxRanks) * sizeof(int), h->stream));
  B2N_TRY(cudaMalloc(&h->d_own_sets, 2 * sizeof(PfPlanes)));
  B2N_TRY(cudaMemcpyAsync(h->d_own_sets, h->set, 2 * sizeof(PfPlanes), cudaMemcpyHostToDevice, h->stream));
  B2N_TRY(cudaMallocHost(&h->h_status, (8 + kPfMaxRanks) * sizeof(int)));
  B2N_TRY(cudaMalloc(&h->d_w, sizeof(double) * h->n_total));
  B2N_TRY(cudaMalloc(&h->d_cum, sizeof(double) * h->n_total));
  B2N_TRY(cudaMalloc(&h->d_anc, sizeof(int32_t) * h->n_total));
  h->h_anc.assign(h->n_total, 0);
  for (int i = 0; i < h->n_total; i++) h->h_anc[i] = i;
  B2N_TRY(cudaMalloc(&h->d_samples, sizeof(double) * 4 * (size_t)h->N * std::max(1, p.k)));
  B2N_TRY(cudaMalloc(&h->d_best, 4 * sizeof(double)));
  B2N_TRY(cudaMallocHost(&h->h_best, 4 * sizeof(double)));
  B2N_TRY(cudaMalloc(&h->d_map, (size_t)G));
  B2N_TRY(cudaMalloc(&h->d_lik, sizeof(double) * h->N));
  B2N_TRY(cudaMalloc(&h->d_stats, 4 * sizeof(unsigned long long)));
  B2N_TRY(cudaMemsetAsync(h->d_stats, 0, 4 * sizeof(unsigned long long), h->stream));

  // distance-field launch shape
  h->smem_optin = prop.sharedMemPerBlockOptin;
  // the limit is per function and process-wide: always the device maximum, never lowered by another handle
  B2N_TRY(cudaFuncSetAttribute(rbpf_distance_field_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin - 1024));
  B2N_TRY(cudaFuncSetAttribute(rbpf_distance_field_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin - 1024));
  B2N_TRY(cudaFuncSetAttribute(rbpf_distance_field_groups_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin - 1024));
  B2N_TRY(cudaFuncSetAttribute(rbpf_distance_field_groups_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin - 1024));
  if (const char *env = std::getenv("B2N_PF_DF_SMEM_MARKS")) h->df_tmem = env[0] != '1';
  if (const char *env = std::getenv("B2N_PF_DF_ALWAYS")) h->df_skip_clean = env[0] != '1';
  if (const char *env = std::getenv("B2N_PF_DF_LANES")) { const int v = std::atoi(env); if (v == 8 || v == 16 || v == 32) h->df_gl = v; }
  if (configure_df(h) != B2N_OK) { b2n_pf_destroy(h); return B2N_ERR_CUDA; }
  if (pf_smem_bytes(h->max_beams, c.pz_stage, kPfWarpsPerCta) > prop.sharedMemPerBlockOptin) {
    set_error("max_beams = %d needs more shared memory than the device has", h->max_beams);
    b2n_pf_destroy(h);
    return B2N_ERR_UNSUPPORTED;
  }
  B2N_TRY(cudaFuncSetAttribute(rbpf_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin));
  B2N_TRY(cudaFuncSetAttribute(rbpf_proposal_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin));
  B2N_TRY(cudaStreamSynchronize(h->stream));
#undef B2N_TRY
  *out = h;
  return B2N_OK;
}

void b2n_pf_destroy(b2n_pf *h)
{
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->comm) ncclCommDestroy(h->comm);
  if (h->p2p_ready) {
    for (int r = 0; r < h->nranks; r++)
      if (r != h->rank)
        for (int k = 0; k < 10; k++)
          if (h->peer_ptrs[(size_t)r * 10 + k]) cudaIpcCloseMemHandle(h->peer_ptrs[(size_t)r * 10 + k]);
  }
  cudaFree(h->d_peer_sets);
  for (auto &e : h->ev) if (e) cudaEventDestroy(e);
  free_planes(h->set[0]); free_planes(h->set[1]);
  cudaFree(h->d_scan); cudaFree(h->d_beam_cs); cudaFree(h->d_pz); cudaFree(h->d_status); cudaFree(h->d_own_sets); cudaFree(h->d_w); cudaFree(h->d_cum); cudaFree(h->d_anc);
  cudaFree(h->d_ext); cudaFree(h->d_samples); cudaFree(h->d_spill); cudaFree(h->d_stats); cudaFree(h->d_best); cudaFree(h->d_map);
  cudaFree(h->d_lik); cudaFree(h->d_idx);
  if (h->h_scan) cudaFreeHost(h->h_scan);
  if (h->h_status) cudaFreeHost(h->h_status);
  if (h->h_best) cudaFreeHost(h->h_best);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  cudaGetLastError();
  delete h;
}

int b2n_pf_seed(b2n_pf *h, uint64_t seed, uint32_t first_call)
{
  B2N_REQUIRE(h, B2N_ERR_INVALID_ARGUMENT, "null handle");
  h->seed = seed; h->call = first_call;
  return B2N_OK;
}

int b2n_pf_set_noise(b2n_pf *h, const double *z, size_t count)
{
  B2N_REQUIRE(h && z, B2N_ERR_INVALID_ARGUMENT, "null argument");
  const size_t a = (size_t)h->N * 3 + 1, b = (size_t)h->N * 3 * ((size_t)h->p.k + 1) + 1;
  B2N_REQUIRE(count == a || count == b, B2N_ERR_INVALID_ARGUMENT,
              "noise count %zu; expected 3*N+1 = %zu (motion-model branch) or 3*(k+1)*N+1 = %zu (proposal branch)", count, a, b);
  if (int rc = set_device(h)) return rc;
  if (count > h->ext_cap) {
    cudaFree(h->d_ext); h->d_ext = nullptr; h->ext_cap = 0;
    B2N_CUDA(cudaMalloc(&h->d_ext, count * sizeof(double)));
    h->ext_cap = count;
  }
  B2N_CUDA(cudaMemcpyAsync(h->d_ext, z, count * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  B2N_CUDA(cudaStreamSynchronize(h->stream));
  h->ext_count = count;
  h->ext_armed = true;
  return B2N_OK;
}

int b2n_pf_slam(b2n_pf *h, const float *scan, int n_beams, const double twist[3], const double cur_odom[3], const double prev_odom[3],
                int icp_ok, const double icp_pose[3])
{
  B2N_REQUIRE(h && twist && cur_odom && prev_odom, B2N_ERR_INVALID_ARGUMENT, "null argument");
  B2N_REQUIRE(!icp_ok || icp_pose, B2N_ERR_INVALID_ARGUMENT, "icp_ok set without icp_pose");
  B2N_REQUIRE(!icp_ok || h->p.k > 0, B2N_ERR_INVALID_ARGUMENT, "proposal branch needs k > 0 mode samples");
  if (int rc = set_device(h)) return rc;
  if (int rc = upload_scan(h, scan, n_beams)) return rc;
  const PfConst &c = h->c;
  PfPlanes &pl = h->set[h->cur];
  const int per = icp_ok ? 3 * (h->p.k + 1) : 3;
  if (h->ext_armed)
    B2N_REQUIRE(h->ext_count == (size_t)h->N * per + 1, B2N_ERR_INVALID_ARGUMENT, "noise armed for the other branch (%zu values, this call needs %zu)",
                h->ext_count, (size_t)h->N * per + 1);

  PfCall q = make_call(h, n_beams);
  q.u_w = twist[0]; q.u_vx = twist[1];
  for (int i = 0; i < 3; i++) { q.cur_od[i] = cur_odom[i]; q.prev_od[i] = prev_odom[i]; q.icp[i] = icp_ok ? icp_pose[i] : 0.0; }
  q.icp_ok = icp_ok;
  q.ext = h->ext_armed ? h->d_ext : nullptr;
  q.ext_per = per;

  B2N_CUDA(cudaMemsetAsync(h->d_status, 0, 4 * sizeof(int), h->stream));
  if (h->timing) B2N_CUDA(cudaEventRecord(h->ev[0], h->stream));
  const int ctas = (h->N + kPfWarpsPerCta - 1) / kPfWarpsPerCta;
  const size_t smem = pf_smem_bytes(n_beams, c.pz_stage, kPfWarpsPerCta);
  if (!icp_ok) rbpf_update_kernel<<<ctas, kPfWarpsPerCta * 32, smem, h->stream>>>(c, pl, q, 0, nullptr);
  else rbpf_proposal_kernel<<<ctas, kPfWarpsPerCta * 32, smem, h->stream>>>(c, pl, q, h->d_samples);
  B2N_CUDA(cudaGetLastError());
  h->launches++;
  if (h->timing) { B2N_CUDA(cudaEventRecord(h->ev[1], h->stream)); B2N_CUDA(cudaEventRecord(h->ev[2], h->stream)); }

  // euclideanSignedDistanceField at the end of every integrateScan (grid_mapper.cpp:181)
  {
    PfDfArgs d;
    d.hcap = h->df_hcap; d.gcap = h->df_gcap; d.warps = h->df_warps; d.cols_per_warp = h->df_cols; d.skip_clean = h->df_skip_clean ? 1 : 0;
    d.spill = h->d_spill; d.stats = h->d_stats; d.status = h->d_status;
    if (h->df_gl_active == 8) rbpf_distance_field_groups_kernel<8><<<h->df_grid, h->df_warps * 32, h->df_smem, h->stream>>>(c, pl, d);
    else if (h->df_gl_active == 16) rbpf_distance_field_groups_kernel<16><<<h->df_grid, h->df_warps * 32, h->df_smem, h->stream>>>(c, pl, d);
    else if (h->df_tmem_active) rbpf_distance_field_kernel<true><<<h->df_grid, h->df_warps * 32, h->df_smem, h->stream>>>(c, pl, d);
    else rbpf_distance_field_kernel<false><<<h->df_grid, h->df_warps * 32, h->df_smem, h->stream>>>(c, pl, d);
  }
  B2N_CUDA(cudaGetLastError());
  h->launches++;
  if (h->timing) { B2N_CUDA(cudaEventRecord(h->ev[3], h->stream)); B2N_CUDA(cudaEventRecord(h->ev[4], h->stream)); }

  // normalizeWeights + effectiveParticles + lowVarianceResampling (particle_filter.cpp:244-249)
  double *w_local = h->d_w + (h->nranks > 1 ? h->offset : 0);
  rbpf_gather_weights_kernel<<<(h->N + 255) / 256, 256, 0, h->stream>>>(pl.meta, w_local, h->N);
  B2N_CUDA(cudaGetLastError());
  h->launches++;
  if (h->nranks > 1) {
    // weights of all ranks in global particle order: one allgather; every rank then runs the identical walk
    // ... and, in the same group, every rank's status bits: a rank that left the map (or ran out of heap) makes EVERY rank
    // return the error below, so that none of them enters the migration alone and `cur` / `call` stay in step
    ncclResult_t r = ncclGroupStart();
    if (r == ncclSuccess) r = ncclAllGather(w_local, h->d_w, (size_t)h->N, ncclDouble, h->comm, h->stream);
    if (r == ncclSuccess) r = ncclAllGather(h->d_status, h->d_status + 8, 1, ncclInt32, h->comm, h->stream);
    if (r == ncclSuccess) r = ncclGroupEnd();
    B2N_REQUIRE(r == ncclSuccess, B2N_ERR_COMM, "ncclAllGather: %s", ncclGetErrorString(r));
  }
  PfResample rs;
  rs.w = h->d_w; rs.cum = h->d_cum; rs.ancestors = h->d_anc; rs.info = h->d_status + 1;
  PfCall qn = q;
  qn.ext = h->ext_armed ? h->d_ext + (size_t)h->N * per : nullptr;
  qn.ext_per = 0;
  rbpf_normalize_kernel<<<1, kNormThreads, 0, h->stream>>>(rs, h->n_total);
  rbpf_ancestors_kernel<<<(h->n_total + 255) / 256, 256, 0, h->stream>>>(rs, h->n_total, qn);
  B2N_CUDA(cudaGetLastError());
  h->launches += 2;
  rbpf_scatter_weights_kernel<<<(h->N + 255) / 256, 256, 0, h->stream>>>(pl.meta, w_local, h->N);
  B2N_CUDA(cudaGetLastError());
  h->launches++;
  B2N_CUDA(cudaMemcpyAsync(h->h_status, h->d_status, (8 + (h->nranks > 1 ? h->nranks : 0)) * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  B2N_CUDA(cudaStreamSynchronize(h->stream));
  h->ext_armed = false;
  int status = h->h_status[0];
  for (int r = 0; r < (h->nranks > 1 ? h->nranks : 0); r++) status |= h->h_status[8 + r];
  h->global_w_valid = false;
  if (status & kPfStatusOffMap) {
    set_error("a beam end point or a particle pose left the map (reference: world2Grid / world2RowMajor throw, grid_mapper.cpp:817-825,854-862)");
    return B2N_ERR_OFF_MAP;
  }
  if (status & kDfStatusHeapOverflow) {
    set_error("distance field: the brushfire heap outgrew its %d + %d entries", h->df_hcap, h->df_gcap);
    return B2N_ERR_UNSUPPORTED;
  }
  if (status & kPfStatusNumeric) {
    set_error("eta is 0 or a zero variance reached pdfNormal (reference: particle_filter.cpp:577-580, grid_mapper.cpp:20-23)");
    return B2N_ERR_NUMERIC;
  }
  h->last_neff = h->h_status[1];
  h->last_resampled = h->h_status[2];
  if (h->last_resampled) {
    B2N_CUDA(cudaMemcpyAsync(h->h_anc.data(), h->d_anc, sizeof(int32_t) * h->n_total, cudaMemcpyDeviceToHost, h->stream));
    B2N_CUDA(cudaStreamSynchronize(h->stream));
    PfPlanes &dst = h->set[h->cur ^ 1];
    dim3 grid(8, h->N);
    if (h->nranks > 1) {
      if (int rc = migrate_particles(h, pl, dst)) return rc;
    } else {
      rbpf_copy_particles_kernel<<<grid, 256, 0, h->stream>>>(c, pl, dst, h->d_anc);
      B2N_CUDA(cudaGetLastError());
      h->launches++;
    }
    h->cur ^= 1;
  } else {
    for (int i = 0; i < h->n_total; i++) h->h_anc[i] = i;
  }
  if (h->timing) {
    B2N_CUDA(cudaEventRecord(h->ev[5], h->stream));
    B2N_CUDA(cudaStreamSynchronize(h->stream));
    cudaEventElapsedTime(&h->last_ms[0], h->ev[0], h->ev[1]);
    cudaEventElapsedTime(&h->last_ms[1], h->ev[2], h->ev[3]);
    cudaEventElapsedTime(&h->last_ms[2], h->ev[4], h->ev[5]);
  }
  h->call++;
  h->global_w_valid = true;
  return B2N_OK;
}

int b2n_pf_likelihoods(b2n_pf *h, const float *scan, int n_beams, double *out, size_t count)
{
  B2N_REQUIRE(h && out, B2N_ERR_INVALID_ARGUMENT, "null argument");
  B2N_REQUIRE(count == (size_t)h->N, B2N_ERR_INVALID_ARGUMENT, "count %zu, expected N = %d", count, h->N);
  if (int rc = set_device(h)) return rc;
  if (int rc = upload_scan(h, scan, n_beams)) return rc;
  PfCall q = make_call(h, n_beams);
  B2N_CUDA(cudaMemsetAsync(h->d_status, 0, 4 * sizeof(int), h->stream));
  const int ctas = (h->N + kPfWarpsPerCta - 1) / kPfWarpsPerCta;
  rbpf_update_kernel<<<ctas, kPfWarpsPerCta * 32, pf_smem_bytes(n_beams, h->c.pz_stage, kPfWarpsPerCta), h->stream>>>(h->c, h->set[h->cur], q, 1, h->d_lik);
  B2N_CUDA(cudaGetLastError());
  h->launches++;
  B2N_CUDA(cudaMemcpyAsync(h->h_status, h->d_status, 4 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  B2N_CUDA(cudaStreamSynchronize(h->stream));
  if (h->h_status[0] & kPfStatusOffMap) {
    set_error("a beam end point left the map (reference: world2RowMajor throws, grid_mapper.cpp:854-862)");
    return B2N_ERR_OFF_MAP;
  }
  B2N_CUDA(cudaMemcpy(out, h->d_lik, sizeof(double) * h->N, cudaMemcpyDeviceToHost));
  return B2N_OK;
}

// sharded filter: every rank has reached this point of its stream (one small all-reduce)
static int pf_barrier(b2n_pf *h)
{
  if (h->nranks <= 1) return B2N_OK;
  ncclResult_t r = ncclAllReduce(h->d_status + 5, h->d_status + 6, 1, ncclInt32, ncclSum, h->comm, h->stream);
  B2N_REQUIRE(r == ncclSuccess, B2N_ERR_COMM, "ncclAllReduce: %s", ncclGetErrorString(r));
  return B2N_OK;
}

// the sets the best particle may live in: this rank's current set, or (sharded) every rank's current set
static const PfPlanes *best_sets(const b2n_pf *h)
{
  return h->nranks > 1 ? h->d_peer_sets + (size_t)h->cur * h->nranks : h->d_own_sets + h->cur;
}

// argmax of the weights over the WHOLE filter (particle_filter.cpp:255-274): d_status[3..4] = (index, owner), d_best = pose, weight
static int run_best(b2n_pf *h)
{
  if (h->nranks > 1) {
    B2N_REQUIRE(h->p2p_ready, B2N_ERR_UNSUPPORTED,
                "the best particle of a sharded filter may live on another GPU: b2n_pf_p2p_init (peer memory) is required");
    B2N_REQUIRE(h->global_w_valid, B2N_ERR_UNSUPPORTED, "sharded filter: the weights changed since the last SLAM() (test tap)");
    // every rank's SLAM() - its resampling copies included - is complete before anyone reads a peer's set
    if (int rc = pf_barrier(h)) return rc;
    rbpf_best_global_kernel<<<1, 1024, 0, h->stream>>>(h->d_w, h->d_anc, h->last_resampled, h->n_total, h->N, h->d_status + 3, h->d_best);
    rbpf_fetch_pose_kernel<<<1, 1, 0, h->stream>>>(best_sets(h), h->d_status + 3, h->d_best);
    B2N_CUDA(cudaGetLastError());
    h->launches += 2;
    return B2N_OK;
  }
  rbpf_best_kernel<<<1, 1024, 0, h->stream>>>(h->set[h->cur].meta, h->N, h->d_status + 3, h->d_best);
  B2N_CUDA(cudaGetLastError());
  h->launches++;
  return B2N_OK;
}

int b2n_pf_get_robot_state(b2n_pf *h, double pose[3])
{
  B2N_REQUIRE(h && pose, B2N_ERR_INVALID_ARGUMENT, "null argument");
  if (int rc = set_device(h)) return rc;
  if (int rc = run_best(h)) return rc;
  if (int rc = pf_barrier(h)) return rc;      // ... and every rank has read before anyone's next SLAM() moves the particles
  B2N_CUDA(cudaMemcpyAsync(h->h_best, h->d_best, 4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  B2N_CUDA(cudaStreamSynchronize(h->stream));
  pose[0] = h->h_best[0]; pose[1] = h->h_best[1]; pose[2] = h->h_best[2];
  return B2N_OK;
}

int b2n_pf_new_map(b2n_pf *h, int8_t *out, size_t count)
{
  B2N_REQUIRE(h && out, B2N_ERR_INVALID_ARGUMENT, "null argument");
  B2N_REQUIRE(count == (size_t)h->c.G, B2N_ERR_INVALID_ARGUMENT, "map count %zu, expected %d cells", count, h->c.G);
  if (int rc = set_device(h)) return rc;
  if (int rc = run_best(h)) return rc;
  rbpf_export_map_kernel<<<(h->c.G + 255) / 256, 256, 0, h->stream>>>(h->c, best_sets(h), h->d_status + 3, h->d_map);
  B2N_CUDA(cudaGetLastError());
  h->launches++;
  if (int rc = pf_barrier(h)) return rc;
  B2N_CUDA(cudaStreamSynchronize(h->stream));
  B2N_CUDA(cudaMemcpy(out, h->d_map, count, cudaMemcpyDeviceToHost));
  return B2N_OK;
}

int b2n_pf_grid_size(const b2n_pf *h, int *xsize, int *ysize)
{
  B2N_REQUIRE(h, B2N_ERR_INVALID_ARGUMENT, "null handle");
  if (xsize) *xsize = h->c.xsize;
  if (ysize) *ysize = h->c.ysize;
  return B2N_OK;
}

int b2n_pf_get_weights(b2n_pf *h, double *out, size_t count)
{
  B2N_REQUIRE(h && out, B2N_ERR_INVALID_ARGUMENT, "null argument");
  B2N_REQUIRE(count == (size_t)h->N, B2N_ERR_INVALID_ARGUMENT, "count %zu, expected N = %d", count, h->N);
  if (int rc = set_device(h)) return rc;
  std::vector<PfParticle> m;
  if (int rc = read_meta(h, m)) return rc;
  for (int i = 0; i < h->N; i++) out[i] = m[i].weight;
  return B2N_OK;
}

int b2n_pf_set_weights(b2n_pf *h, const double *w, size_t count)
{
  B2N_REQUIRE(h && w, B2N_ERR_INVALID_ARGUMENT, "null argument");
  B2N_REQUIRE(count == (size_t)h->N, B2N_ERR_INVALID_ARGUMENT, "count %zu, expected N = %d", count, h->N);
  if (int rc = set_device(h)) return rc;
  std::vector<PfParticle> m;
  if (int rc = read_meta(h, m)) return rc;
  for (int i = 0; i < h->N; i++) m[i].weight = w[i];
  B2N_CUDA(cudaMemcpy(h->set[h->cur].meta, m.data(), sizeof(PfParticle) * h->N, cudaMemcpyHostToDevice));
  h->global_w_valid = false;
  return B2N_OK;
}

int b2n_pf_get_poses(b2n_pf *h, double *poses, double *prev_poses, size_t count)
{
  B2N_REQUIRE(h, B2N_ERR_INVALID_ARGUMENT, "null handle");
  B2N_REQUIRE(count == (size_t)h->N * 3, B2N_ERR_INVALID_ARGUMENT, "count %zu, expected 3*N = %d", count, 3 * h->N);
  if (int rc = set_device(h)) return rc;
  std::vector<PfParticle> m;
  if (int rc = read_meta(h, m)) return rc;
  for (int i = 0; i < h->N; i++)
    for (int k = 0; k < 3; k++) {
      if (poses) poses[3 * i + k] = m[i].pose[k];
      if (prev_poses) prev_poses[3 * i + k] = m[i].prev_pose[k];
    }
  return B2N_OK;
}

int b2n_pf_set_poses(b2n_pf *h, const double *poses, size_t count)
{
  B2N_REQUIRE(h && poses, B2N_ERR_INVALID_ARGUMENT, "null argument");
  B2N_REQUIRE(count == (size_t)h->N * 3, B2N_ERR_INVALID_ARGUMENT, "count %zu, expected 3*N = %d", count, 3 * h->N);
  if (int rc = set_device(h)) return rc;
  std::vector<PfParticle> m;
  if (int rc = read_meta(h, m)) return rc;
  for (int i = 0; i < h->N; i++)
    for (int k = 0; k < 3; k++) m[i].pose[k] = poses[3 * i + k];
  B2N_CUDA(cudaMemcpy(h->set[h->cur].meta, m.data(), sizeof(PfParticle) * h->N, cudaMemcpyHostToDevice));
  return B2N_OK;
}

int b2n_pf_get_resample(b2n_pf *h, int *neff, int *resampled, int32_t *ancestors, size_t count)
{
  B2N_REQUIRE(h, B2N_ERR_INVALID_ARGUMENT, "null handle");
  if (neff) *neff = h->last_neff;
  if (resampled) *resampled = h->last_resampled;
  if (ancestors) {
    B2N_REQUIRE(count == (size_t)h->n_total, B2N_ERR_INVALID_ARGUMENT, "count %zu, expected particles_total = %d", count, h->n_total);
    std::copy(h->h_anc.begin(), h->h_anc.end(), ancestors);
  }
  return B2N_OK;
}

// the normalise / N_eff / walk part alone on the current weights (tests, kernel timing)
int b2n_pf_normalize_resample(b2n_pf *h)
{
  B2N_REQUIRE(h, B2N_ERR_INVALID_ARGUMENT, "null handle");
  B2N_REQUIRE(h->nranks == 1, B2N_ERR_UNSUPPORTED, "single-rank tap");
  if (int rc = set_device(h)) return rc;
  PfPlanes &pl = h->set[h->cur];
  const PfConst &c = h->c;
  if (h->ext_armed) B2N_REQUIRE(h->ext_count == (size_t)h->N * 3 + 1, B2N_ERR_INVALID_ARGUMENT, "noise must hold 3*N+1 values");
  PfCall q = make_call(h, 0);
  q.ext = h->ext_armed ? h->d_ext + (size_t)h->N * 3 : nullptr;
  q.ext_per = 0;
  rbpf_gather_weights_kernel<<<(h->N + 255) / 256, 256, 0, h->stream>>>(pl.meta, h->d_w, h->N);
  PfResample rs;
  rs.w = h->d_w; rs.cum = h->d_cum; rs.ancestors = h->d_anc; rs.info = h->d_status + 1;
  rbpf_normalize_kernel<<<1, kNormThreads, 0, h->stream>>>(rs, h->n_total);
  rbpf_ancestors_kernel<<<(h->n_total + 255) / 256, 256, 0, h->stream>>>(rs, h->n_total, q);
  rbpf_scatter_weights_kernel<<<(h->N + 255) / 256, 256, 0, h->stream>>>(pl.meta, h->d_w, h->N);
  B2N_CUDA(cudaGetLastError());
  h->launches += 4;
  B2N_CUDA(cudaMemcpyAsync(h->h_status, h->d_status, 4 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  B2N_CUDA(cudaMemcpyAsync(h->h_anc.data(), h->d_anc, sizeof(int32_t) * h->n_total, cudaMemcpyDeviceToHost, h->stream));
  B2N_CUDA(cudaStreamSynchronize(h->stream));
  h->ext_armed = false;
  h->last_neff = h->h_status[1];
  h->last_resampled = h->h_status[2];
  if (h->last_resampled) {
    dim3 grid(8, h->N);
    rbpf_copy_particles_kernel<<<grid, 256, 0, h->stream>>>(c, pl, h->set[h->cur ^ 1], h->d_anc);
    B2N_CUDA(cudaGetLastError());
    h->launches++;
    h->cur ^= 1;
  }
  h->call++;
  return B2N_OK;
}

int b2n_pf_get_grid(b2n_pf *h, int particle, double *log_odds, double *occ_dist, int8_t *state, size_t count)
{
  B2N_REQUIRE(h, B2N_ERR_INVALID_ARGUMENT, "null handle");
  B2N_REQUIRE(particle >= 0 && particle < h->N, B2N_ERR_INVALID_ARGUMENT, "particle %d out of range", particle);
  B2N_REQUIRE(count == (size_t)h->c.G, B2N_ERR_INVALID_ARGUMENT, "count %zu, expected %d cells", count, h->c.G);
  if (int rc = set_device(h)) return rc;
  B2N_CUDA(cudaStreamSynchronize(h->stream));
  const PfConst &c = h->c;
  std::vector<double> l(count);
  std::vector<uint32_t> d2(count);
  B2N_CUDA(cudaMemcpy(l.data(), h->set[h->cur].log_odds + (size_t)particle * c.gstride, count * sizeof(double), cudaMemcpyDeviceToHost));
  B2N_CUDA(cudaMemcpy(d2.data(), h->set[h->cur].d2 + (size_t)particle * c.gstride, count * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < count; i++) {
    if (log_odds) log_odds[i] = l[i];
    if (occ_dist) occ_dist[i] = d2[i] == kD2Unreached ? h->max_occ_dist : std::sqrt((double)d2[i]) * c.res;
    if (state) state[i] = l[i] >= c.t_occ ? 1 : (l[i] <= c.t_free ? 0 : -1);
  }
  return B2N_OK;
}

int b2n_pf_get_occ_order(b2n_pf *h, int particle, int32_t *keys, size_t cap, int *n_occ)
{
  B2N_REQUIRE(h && n_occ, B2N_ERR_INVALID_ARGUMENT, "null argument");
  B2N_REQUIRE(particle >= 0 && particle < h->N, B2N_ERR_INVALID_ARGUMENT, "particle %d out of range", particle);
  if (int rc = set_device(h)) return rc;
  B2N_CUDA(cudaStreamSynchronize(h->stream));
  const PfConst &c = h->c;
  std::vector<uint16_t> nxt((size_t)c.G + 1);
  B2N_CUDA(cudaMemcpy(nxt.data(), h->set[h->cur].nxt + (size_t)particle * c.nxt_stride, nxt.size() * sizeof(uint16_t), cudaMemcpyDeviceToHost));
  int n = 0;
  for (uint32_t k = nxt[c.G]; k != kNil16 && n <= c.G; k = nxt[k]) {
    if (keys && (size_t)n < cap) keys[n] = (int32_t)k;
    n++;
  }
  *n_occ = n;
  return B2N_OK;
}

int b2n_pf_set_stream(b2n_pf *h, void *cuda_stream)
{
  B2N_REQUIRE(h, B2N_ERR_INVALID_ARGUMENT, "null handle");
  B2N_CUDA(cudaStreamSynchronize(h->stream));
  h->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : h->own_stream;
  return B2N_OK;
}

int b2n_pf_geometry(const b2n_pf *h, double *xmin, double *ymin, double *resolution)
{
  B2N_REQUIRE(h, B2N_ERR_INVALID_ARGUMENT, "null handle");
  if (xmin) *xmin = h->c.xmin;
  if (ymin) *ymin = h->c.ymin;
  if (resolution) *resolution = h->c.res;
  return B2N_OK;
}

int b2n_pf_write_distance_field(b2n_pf *h, float *device_out, size_t count)
{
  B2N_REQUIRE(h && device_out, B2N_ERR_INVALID_ARGUMENT, "null argument");
  B2N_REQUIRE(count == (size_t)h->c.G, B2N_ERR_INVALID_ARGUMENT, "count %zu, expected %d cells", count, h->c.G);
  B2N_REQUIRE(h->nranks == 1, B2N_ERR_UNSUPPORTED, "single-rank operation (the best particle may live on another GPU)");
  if (int rc = set_device(h)) return rc;
  if (int rc = run_best(h)) return rc;
  rbpf_export_distance_kernel<<<(h->c.G + 255) / 256, 256, 0, h->stream>>>(h->c, best_sets(h), h->d_status + 3, h->max_occ_dist, device_out);
  if (int rc = pf_barrier(h)) return rc;
  B2N_CUDA(cudaGetLastError());
  h->launches++;
  B2N_CUDA(cudaStreamSynchronize(h->stream));
  return B2N_OK;
}

int b2n_pf_launch_count(const b2n_pf *h, uint64_t *launches)
{
  B2N_REQUIRE(h && launches, B2N_ERR_INVALID_ARGUMENT, "null argument");
  *launches = h->launches;
  return B2N_OK;
}

int b2n_pf_set_kernel_timing(b2n_pf *h, int on)
{
  B2N_REQUIRE(h, B2N_ERR_INVALID_ARGUMENT, "null handle");
  if (int rc = set_device(h)) return rc;
  if (on && !h->ev[0])
    for (auto &e : h->ev) B2N_CUDA(cudaEventCreate(&e));
  h->timing = on != 0;
  return B2N_OK;
}

int b2n_pf_kernel_times(b2n_pf *h, double ms[3])
{
  B2N_REQUIRE(h && ms, B2N_ERR_INVALID_ARGUMENT, "null argument");
  for (int i = 0; i < 3; i++) ms[i] = h->last_ms[i];
  return B2N_OK;
}

int b2n_pf_distance_field_stats(b2n_pf *h, uint64_t *iterations, uint64_t *heap_max)
{
  B2N_REQUIRE(h, B2N_ERR_INVALID_ARGUMENT, "null handle");
  if (int rc = set_device(h)) return rc;
  B2N_CUDA(cudaStreamSynchronize(h->stream));
  unsigned long long s[2];
  B2N_CUDA(cudaMemcpy(s, h->d_stats, sizeof(s), cudaMemcpyDeviceToHost));
  if (iterations) *iterations = s[0];
  if (heap_max) *heap_max = s[1];
  return B2N_OK;
}

int b2n_pf_distance_field_skipped(b2n_pf *h, uint64_t *particles)
{
  B2N_REQUIRE(h && particles, B2N_ERR_INVALID_ARGUMENT, "null argument");
  if (int rc = set_device(h)) return rc;
  B2N_CUDA(cudaStreamSynchronize(h->stream));
  unsigned long long s = 0;
  B2N_CUDA(cudaMemcpy(&s, h->d_stats + 2, sizeof(s), cudaMemcpyDeviceToHost));
  *particles = s;
  return B2N_OK;
}

int b2n_pf_set_heap_capacity(b2n_pf *h, int entries)
{
  B2N_REQUIRE(h && entries >= 2 && entries <= 24576, B2N_ERR_INVALID_ARGUMENT, "heap capacity must be in [2, 24576]");
  if (int rc = set_device(h)) return rc;
  B2N_CUDA(cudaStreamSynchronize(h->stream));
  h->df_hcap_request = entries & ~1;
  return configure_df(h);
}

int b2n_pf_host_tables(const b2n_pf_params *params, double constants[4], double *beam_cs, size_t beam_count, double *pz, size_t pz_cap,
                       int *pz_n)
{
  B2N_REQUIRE(params && constants && pz_n, B2N_ERR_INVALID_ARGUMENT, "null argument");
  const b2n_pf_params &p = *params;
  B2N_REQUIRE(p.resolution > 0.0 && p.xmax > p.xmin && p.ymax > p.ymin, B2N_ERR_INVALID_ARGUMENT, "bad map geometry");
  const int xsize = (int)mapSize(p.xmin, p.xmax, p.resolution), ysize = (int)mapSize(p.ymin, p.ymax, p.resolution);
  PfConst c;
  double l_prior;
  std::vector<double> beam, table;
  host_tables(p, xsize, ysize, (long long)xsize * ysize, (int)(beam_count / 2), c, l_prior, beam, table);
  constants[0] = c.t_occ; constants[1] = c.t_free; constants[2] = c.d_free; constants[3] = c.d_occ;
  if (beam_cs) std::copy(beam.begin(), beam.end(), beam_cs);
  *pz_n = c.pz_n;
  if (pz) std::copy(table.begin(), table.begin() + std::min(pz_cap, table.size()), pz);
  return B2N_OK;
}

int b2n_pf_plan_migration(const int32_t *ancestors, int n_total, int rank, int nranks, int32_t *copy1, int32_t *copy2, int32_t *recv,
                          size_t recv_cap, int *n_recv, int32_t *send, size_t send_cap, int *n_send)
{
  B2N_REQUIRE(ancestors && copy1 && copy2 && n_recv && n_send, B2N_ERR_INVALID_ARGUMENT, "null argument");
  B2N_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks && n_total > 0 && n_total % nranks == 0, B2N_ERR_INVALID_ARGUMENT,
              "sharding must be even: n_total = %d over %d ranks", n_total, nranks);
  for (int i = 0; i < n_total; i++)
    B2N_REQUIRE(ancestors[i] >= 0 && ancestors[i] < n_total, B2N_ERR_INVALID_ARGUMENT, "ancestor %d of slot %d out of range", ancestors[i], i);
  const int n_local = n_total / nranks;
  MigrationPlan mp;
  plan_migration(ancestors, n_total, n_local, rank, nranks, mp);
  std::copy(mp.copy1.begin(), mp.copy1.end(), copy1);
  std::copy(mp.copy2.begin(), mp.copy2.end(), copy2);
  *n_recv = (int)mp.recv_slot.size();
  *n_send = (int)mp.send_idx.size();
  for (size_t i = 0; i < mp.recv_slot.size() && recv && 3 * i + 2 < recv_cap; i++) {
    recv[3 * i] = mp.recv_slot[i]; recv[3 * i + 1] = mp.recv_anc[i]; recv[3 * i + 2] = mp.recv_rank[i];
  }
  for (size_t i = 0; i < mp.send_idx.size() && send && 2 * i + 1 < send_cap; i++) {
    send[2 * i] = mp.send_idx[i]; send[2 * i + 1] = mp.send_rank[i];
  }
  return B2N_OK;
}

int b2n_pf_get_migration(const b2n_pf *h, int *received, int *sent)
{
  B2N_REQUIRE(h, B2N_ERR_INVALID_ARGUMENT, "null handle");
  if (received) *received = h->last_migrated_in;
  if (sent) *sent = h->last_migrated_out;
  return B2N_OK;
}

int b2n_pf_p2p_export(b2n_pf *h, void *handles640)
{
  B2N_REQUIRE(h && handles640, B2N_ERR_INVALID_ARGUMENT, "null argument");
  if (int rc = set_device(h)) return rc;
  B2N_CUDA(cudaStreamSynchronize(h->stream));
  char *out = static_cast<char *>(handles640);
  for (int s = 0; s < 2; s++) {
    void *planes[5] = {h->set[s].log_odds, h->set[s].d2, h->set[s].nxt, h->set[s].bkt, h->set[s].meta};
    for (int k = 0; k < 5; k++) {
      cudaIpcMemHandle_t ipc;
      B2N_CUDA(cudaIpcGetMemHandle(&ipc, planes[k]));
      std::memcpy(out + (size_t)(s * 5 + k) * sizeof(ipc), &ipc, sizeof(ipc));
    }
  }
  return B2N_OK;
}

int b2n_pf_p2p_init(b2n_pf *h, int rank, int nranks, const void *handles)
{
  B2N_REQUIRE(h && handles, B2N_ERR_INVALID_ARGUMENT, "null argument");
  B2N_REQUIRE(h->comm && rank == h->rank && nranks == h->nranks, B2N_ERR_INVALID_ARGUMENT,
              "b2n_pf_comm_init comes first (the weights still travel with ncclAllGather), with the same rank and size");
  if (int rc = set_device(h)) return rc;
  h->peer_ptrs.assign((size_t)nranks * 10, nullptr);
  std::vector<PfPlanes> sets((size_t)2 * nranks);
  for (int r = 0; r < nranks; r++) {
    for (int s = 0; s < 2; s++) {
      void *p[5];
      if (r == rank) {
        p[0] = h->set[s].log_odds; p[1] = h->set[s].d2; p[2] = h->set[s].nxt; p[3] = h->set[s].bkt; p[4] = h->set[s].meta;
      } else {
        for (int k = 0; k < 5; k++) {
          cudaIpcMemHandle_t ipc;
          std::memcpy(&ipc, static_cast<const char *>(handles) + ((size_t)r * 10 + s * 5 + k) * sizeof(ipc), sizeof(ipc));
          cudaError_t e = cudaIpcOpenMemHandle(&p[k], ipc, cudaIpcMemLazyEnablePeerAccess);
          if (e != cudaSuccess) {
            cudaGetLastError();
            set_error("cudaIpcOpenMemHandle for rank %d: %s (peer access between the GPUs is required)", r, cudaGetErrorString(e));
            return B2N_ERR_COMM;
          }
        }
      }
      for (int k = 0; k < 5; k++) h->peer_ptrs[(size_t)r * 10 + s * 5 + k] = p[k];
      PfPlanes &q = sets[(size_t)s * nranks + r];
      q.log_odds = static_cast<double *>(p[0]); q.d2 = static_cast<uint32_t *>(p[1]); q.nxt = static_cast<uint16_t *>(p[2]);
      q.bkt = static_cast<uint16_t *>(p[3]); q.meta = static_cast<PfParticle *>(p[4]);
    }
  }
  if (!h->d_peer_sets) B2N_CUDA(cudaMalloc(&h->d_peer_sets, sets.size() * sizeof(PfPlanes)));
  B2N_CUDA(cudaMemcpy(h->d_peer_sets, sets.data(), sets.size() * sizeof(PfPlanes), cudaMemcpyHostToDevice));
  h->p2p_ready = true;
  return B2N_OK;
}

int b2n_pf_comm_init(b2n_pf *h, int rank, int nranks, const void *unique_id128)
{
  B2N_REQUIRE(h && unique_id128, B2N_ERR_INVALID_ARGUMENT, "null argument");
  B2N_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, B2N_ERR_INVALID_ARGUMENT, "bad rank %d of %d", rank, nranks);
  B2N_REQUIRE(h->n_total == h->N * nranks && h->offset == rank * h->N, B2N_ERR_INVALID_ARGUMENT,
              "sharding must be even: particles_total = nranks * num_particles and particle_offset = rank * num_particles");
  if (int rc = set_device(h)) return rc;
  ncclUniqueId id;
  std::memcpy(&id, unique_id128, sizeof(id));
  ncclResult_t r = ncclCommInitRank(&h->comm, nranks, id, rank);
  B2N_REQUIRE(r == ncclSuccess, B2N_ERR_COMM, "ncclCommInitRank: %s", ncclGetErrorString(r));
  h->rank = rank; h->nranks = nranks;
  return B2N_OK;
}

} // extern "C"
