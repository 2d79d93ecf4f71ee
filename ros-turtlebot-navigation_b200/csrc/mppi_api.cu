// mppi_api.cu - extern "C" MPPI entry points of libb2nav (see include/b2nav.h).
// Host side of controller::MPPI (reference: controller/src/controller/mppi.cpp:28-69,72-140).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include <nccl.h>

#include "mppi_kernels.cuh"

using namespace b2n;

constexpr int kZBuf = 3;

struct b2n_mppi
{
  b2n_mppi_params p;
  int T = 0, K = 0, S = 4, G = 16, NW = 8, device = 0;
  double plan_abs_max = 0.0;             // bound on |u| over the plan, the tail value and the clamp
  bool force_generic = false;            // B2N_MPPI_GENERIC=1: never take the FAST rollout variant (tests)
  int n_sm = 0, grid = 0;
  size_t smem = 0;
  bool tma_store = false;

  cudaStream_t own_stream = nullptr, stream = nullptr;

  double xd[3] = {0, 0, 0}, uinit[2] = {0, 0};
  uint64_t seed = 0;
  uint32_t call = 0;

  double *d_u[2] = {nullptr, nullptr};   // ping-pong plan [2][T]
  int cur = 0;
  float *d_states = nullptr;             // [ring][K][T][3]
  int ring = 1, ring_pos = 0, last_slot = 0;
  double *d_partials = nullptr;          // [T][grid][6]
  MppiLL *d_ll_partials = nullptr;       // [T][3][grid] tagged words: the fused call's CTA partials (MppiArgs::ll_partials)
  MppiLL *d_ll_plan[2] = {nullptr, nullptr};  // [T] tagged words: the plan a fused call leaves for the next one's CTAs
  unsigned long long *d_zready = nullptr;     // CTAs of noise kernels that have finished
  unsigned long long zready_total = 0;        //   its value once every noise kernel launched so far is through
  bool full_wait = false;                     // B2N_MPPI_FULL_WAIT=1: every call waits for the grids in front of it (tuning)
  unsigned long long *d_arrive = nullptr;     // warps of rollout CTAs that have sent their partial words (a hint for the merger CTAs)
  uint32_t plan_tag = 0;                 // tag of the fused call that wrote the current plan; 0: written some other way (plain array, stream order)
  unsigned long long fused_calls = 0;    // fused calls enqueued so far
  unsigned long long *d_dbg = nullptr;   // [grid][8] stage timestamps, B2N_MPPI_DEBUG_TIMES=1 (tuning runs)
  int last_fast = 0;                     // the last call ran the FAST instantiation
  // the variates of a call, drawn ahead by mppi_noise_kernel behind the previous call: two buffers [K][T/2] float4
  // the variates are drawn TWO calls ahead (call c + 2's behind call c): three buffers in rotation
  float4 *d_z[kZBuf] = {nullptr, nullptr, nullptr};
  long long z_call[kZBuf] = {-1, -1, -1};       // the call number whose variates a buffer holds (-1: none)
  uint64_t z_seed[kZBuf] = {0, 0, 0};
  int zslot = 0;                                 // buffer of call h->call (advances with it)
  unsigned long long z_ready_at[kZBuf] = {0, 0, 0};   // MppiArgs::z_need for a call that reads the buffer
  int noise_ctas_per_sm = 8;             // B2N_MPPI_NOISE_CTAS: resident CTAs of the noise kernel per SM (it shares the SMs with a call's tail)
  bool noise_ahead = true;               // B2N_MPPI_NOISE_AHEAD=0: draw a call's variates in front of the call instead of behind the previous one
  double *d_merged = nullptr;            // [T][6]
  double *d_gathered = nullptr;          // [nranks][T][6]
  double *d_stepstats = nullptr;         // [T][2]
  double *h_out = nullptr;               // pinned, mapped [2]: the update kernel writes the controls here
  double *d_out_host = nullptr;          // device view of h_out
  unsigned long long out_seq = 0;        // sequence number of the last enqueued call (completion word in h_out[2])
  bool use_pdl = true;                   // programmatic dependent launch of call -> next call (trigger after the rollout loop, so
                                         // dependents never take residency from it); B2N_MPPI_PDL=0 turns it off
  double *d_ext = nullptr;               // [K][T][2]
  bool ext_armed = false;
  int capture = 0;
  double *d_J = nullptr, *d_du = nullptr, *d_w = nullptr;

  // obstacle field
  float *d_obs = nullptr;
  int obs_on = 0, obs_xsize = 0, obs_ysize = 0;
  double obs_xmin = 0, obs_ymin = 0, obs_res = 1, obs_weight = 0, obs_d0 = 0, obs_off = 0;

  // sharding
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  // peer-memory exchange (CUDA IPC): this rank's area and the mapped areas of all ranks
  void *xchg = nullptr;                  // [2][nranks][T] x (6 doubles) followed by [2][nranks][T] flags
  size_t xchg_bytes = 0, xchg_flag_offset = 0;
  int xchg_nranks = 0;
  std::vector<void *> peer_base;         // mapped peer areas (own entry = xchg)
  bool p2p_ready = false, p2p_local = false;   // local: the peers are handles of this process (plain device pointers, nothing to close)
  unsigned long long xchg_call = 0;

  // accounting
  uint64_t launches = 0;
  bool timing = false;
  std::vector<cudaEvent_t> ev;           // start/stop pairs
  size_t ev_used = 0;
  bool pending = false;
  bool obs_external = false;             // the obstacle field's device address is known to a producer outside this handle
  bool prof = false;                     // B2N_MPPI_HOST_BREAKDOWN=1: host time per phase of a synchronous call (tuning)
  double prof_ns[4] = {0, 0, 0, 0};      //   arguments, the call's launch, the noise kernel's launch, wait for the controls
  unsigned long long waited_seq = 0;     // sequence number of the last call whose controls b2n_mppi_wait handed out
};

namespace
{

// (S, G, NW) = steps per lane, lanes per rollout, warps per CTA; G * S >= T.  The table below is every shape that is built,
// each as three instantiations: generic (taps, external noise, partial lanes, runtime obstacle switch), FAST and FAST + obstacle term.
template <int S, int G, int NW>
cudaError_t configure_shape(b2n_mppi *h)
{
  // the generic variant decides about the obstacle term at run time: its shared memory always includes the tile
  h->smem = mppi_rollout_smem(S, G, NW, true, false);
  cudaError_t e = cudaFuncSetAttribute(mppi_rollout_kernel<S, G, false, false, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(mppi_rollout_kernel<S, G, true, false, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mppi_rollout_smem(S, G, NW, false, true));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(mppi_rollout_kernel<S, G, true, true, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mppi_rollout_smem(S, G, NW, true, true));
  if (e != cudaSuccess) return e;
  int per_sm = 0, per_sm_fast = 0, per_sm_obs = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mppi_rollout_kernel<S, G, false, false, NW>, NW * 32, h->smem);
  if (e != cudaSuccess) return e;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_fast, mppi_rollout_kernel<S, G, true, false, NW>, NW * 32, mppi_rollout_smem(S, G, NW, false, true));
  if (e != cudaSuccess) return e;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_obs, mppi_rollout_kernel<S, G, true, true, NW>, NW * 32, mppi_rollout_smem(S, G, NW, true, true));
  if (e != cudaSuccess) return e;
  per_sm = std::max(per_sm, std::max(per_sm_fast, per_sm_obs));     // the partials buffer is sized for the largest grid; every variant strides by gridDim
  if (per_sm < 1) per_sm = 1;
  // persistent grid: every SM filled to its residency limit (an evenly divided but smaller grid measured slower:
  // SMs holding one CTA more than their neighbours set the pace)
  const int per_cta = NW * (32 / G);
  const int want = (h->K + per_cta - 1) / per_cta;
  h->grid = std::max(1, std::min(want, per_sm * h->n_sm));
  return cudaSuccess;
}

#define B2N_MPPI_SHAPES(X) X(2, 8, 8) X(4, 8, 8) X(2, 16, 8) X(4, 16, 8) X(2, 32, 8) X(4, 32, 8) X(8, 32, 8) X(4, 16, 10) X(4, 16, 12) X(4, 16, 20) X(4, 32, 10) X(4, 32, 20)

cudaError_t configure(b2n_mppi *h)
{
#define X(S_, G_, W_) if (h->S == S_ && h->G == G_ && h->NW == W_) return configure_shape<S_, G_, W_>(h);
  B2N_MPPI_SHAPES(X)
#undef X
  return cudaErrorInvalidValue;
}

template <class Kernel>
void launch_rollout_kernel(b2n_mppi *h, Kernel kernel, int threads, size_t smem, const MppiArgs &a)
{
  cudaLaunchConfig_t cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(h->grid + (a.tail ? h->T : 0))); cfg.blockDim = dim3((unsigned)threads); cfg.dynamicSmemBytes = smem; cfg.stream = h->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = h->use_pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, a);
}

// variant: 0 generic, 1 FAST, 2 FAST with the obstacle term
void launch_rollout(b2n_mppi *h, const MppiArgs &a, int variant)
{
#define X(S_, G_, W_)                                                                                                           \
  if (h->S == S_ && h->G == G_ && h->NW == W_) {                                                                                \
    if (variant == 1) launch_rollout_kernel(h, mppi_rollout_kernel<S_, G_, true, false, W_>, W_ * 32, mppi_rollout_smem(S_, G_, W_, false, true), a);   \
    else if (variant == 2) launch_rollout_kernel(h, mppi_rollout_kernel<S_, G_, true, true, W_>, W_ * 32, mppi_rollout_smem(S_, G_, W_, true, true), a);  \
    else launch_rollout_kernel(h, mppi_rollout_kernel<S_, G_, false, false, W_>, W_ * 32, mppi_rollout_smem(S_, G_, W_, true, false), a);               \
    return;                                                                                                                     \
  }
  B2N_MPPI_SHAPES(X)
#undef X
}

// default shape: four steps per lane where the horizon allows it (fewest scan levels per step without running out
// of registers), 8 warps per CTA; B2N_MPPI_SHAPE="S,G[,NW]" overrides it for tuning runs
void pick_shape(int T, int &S, int &G, int &NW)
{
  NW = 8;
  if (T <= 16) { S = 2; G = 8; }
  else if (T <= 32) { S = 4; G = 8; }
  else if (T <= 64) { S = 4; G = 16; NW = 10; }      // 2 CTAs x 10 warps per SM at 96 registers: 20 warps, no spill traffic in the loop
  else if (T <= 128) { S = 4; G = 32; NW = 10; }     //   (measured at C2: 12 warps x 2 at 80 registers 23.7 us per call, 13 x 2 at 72: 26.2, 10 x 2: 20.9)
  else { S = 8; G = 32; }
  if (const char *env = std::getenv("B2N_MPPI_SHAPE")) {
    int s = 0, g = 0, w = 8;
    const int n = std::sscanf(env, "%d,%d,%d", &s, &g, &w);
    if (n >= 2 && s * g >= T) {
#define X(S_, G_, W_) if (s == S_ && g == G_ && w == W_) { S = s; G = g; NW = w; }
      B2N_MPPI_SHAPES(X)
#undef X
    }
  }
}

int set_device(const b2n_mppi *h)
{
  B2N_CUDA(cudaSetDevice(h->device));
  return B2N_OK;
}

// kernel arguments of one call (the state-tensor slot is chosen by the caller)
MppiArgs make_args(b2n_mppi *h, double x, double y, double theta)
{
  const b2n_mppi_params &p = h->p;
  MppiArgs a;
  std::memset(&a, 0, sizeof(a));
  a.c_v = (p.wheel_radius / 2.0) * (p.dt / 6.0);
  a.c_w = (p.wheel_radius / p.wheel_base) * p.dt;
  for (int i = 0; i < 3; i++) { a.Q[i] = p.Q[i]; a.P1[i] = p.P1[i]; a.xd[i] = h->xd[i]; }
  a.R[0] = p.R[0]; a.R[1] = p.R[1];
  a.inv_lambda = 1.0 / p.lambda;
  a.cut_lambda = 708.0 * p.lambda;
  a.sigL = std::sqrt(p.ul_var);       // mppi.cpp:176-177
  a.sigR = std::sqrt(p.ur_var);
  a.x0[0] = x; a.x0[1] = y; a.x0[2] = theta;   // mppi.cpp:75-76
  a.cos0 = std::cos(theta); a.sin0 = std::sin(theta);
  a.ks3 = -1.6666666666666666e-01; a.ks5 = 8.3333333333333332e-03; a.ks7 = -1.9841269841269841e-04; a.ks9 = 2.7557319223985893e-06;
  a.kc2 = -0.5; a.kc4 = 4.1666666666666664e-02; a.kc6 = -1.3888888888888889e-03; a.kc8 = 2.4801587301587302e-05; a.kc10 = -2.7557319223985888e-07;
  a.T = h->T; a.K = h->K; a.k_offset = p.rollout_offset;
  a.call = h->call;
  for (int r = 0; r < 10; r++) {
    a.key0[r] = (uint32_t)h->seed + (uint32_t)r * 0x9E3779B9u;
    a.key1[r] = (uint32_t)(h->seed >> 32) + (uint32_t)r * 0xBB67AE85u;
  }
  a.external_noise = h->ext_armed ? 1 : 0;
  a.capture = h->capture;
  a.tma_store = h->tma_store ? 1 : 0;
  a.obs_on = h->obs_on; a.obs_xsize = h->obs_xsize; a.obs_ysize = h->obs_ysize;
  a.obs_xmin = h->obs_xmin; a.obs_ymin = h->obs_ymin; a.obs_res = h->obs_res; a.obs_inv_res = 1.0 / h->obs_res;
  a.obs_xmax = h->obs_xmin + h->obs_xsize * h->obs_res;
  a.obs_ymax = h->obs_ymin + h->obs_ysize * h->obs_res;
  a.obs_weight = h->obs_weight; a.obs_d0 = h->obs_d0; a.obs_off = h->obs_off; a.obs_dist = h->d_obs;
  // the tile of the field staged through TMA: kMppiObsTile cells square around the start pose's cell, clamped into the
  // field, its first column a multiple of four cells (16-byte granules of the bulk copies)
  a.obs_ti0 = a.obs_tj0 = -1;
  if (h->obs_on && h->obs_xsize >= kMppiObsTile && h->obs_ysize >= kMppiObsTile && h->obs_ysize % 4 == 0) {
    const int ci = (int)std::floor((x - h->obs_xmin) / h->obs_res), cj = (int)std::floor((y - h->obs_ymin) / h->obs_res);
    a.obs_ti0 = std::min(std::max(ci - kMppiObsTile / 2, 0), h->obs_xsize - kMppiObsTile);
    a.obs_tj0 = std::min(std::max((cj - kMppiObsTile / 2) & ~3, 0), (h->obs_ysize - kMppiObsTile) & ~3);
  }
  // J >= 0 (sums of squares with non-negative weights): the high words of J order like J
  a.thr_on = 1;
  for (int i = 0; i < 3; i++) if (!(p.Q[i] >= 0.0) || !(p.P1[i] >= 0.0)) a.thr_on = 0;
  if (!(p.R[0] >= 0.0) || !(p.R[1] >= 0.0) || (h->obs_on && (!(h->obs_weight >= 0.0) || !(h->obs_off >= 0.0)))) a.thr_on = 0;
  a.u_plan = h->d_u[h->cur];
  a.ext = h->d_ext; a.J_out = h->d_J; a.du_out = h->d_du; a.partials = h->d_partials;
  a.tail = 0;
  a.n_roll = h->grid;
  return a;
}

// the fused tail's arguments: merge tree, update, and (sharded over peer memory) the exchange
void arm_tail(b2n_mppi *h, MppiArgs &a)
{
  const b2n_mppi_params &p = h->p;
  a.tail = 1;
  h->fused_calls++;
  a.tag = (uint32_t)(h->fused_calls % 0xFFFFFFFFull) + 1u;      // never 0, distinct from the tags a buffer can still hold
  a.plan_tag = h->plan_tag;
  a.ll_partials = h->d_ll_partials;
  a.arrive = h->d_arrive; a.arrive_need = (unsigned long long)h->grid * (unsigned long long)((h->T + 31) / 32) * h->fused_calls;
  a.ll_plan = h->d_ll_plan[(h->fused_calls & 1ull) ^ 1ull];     // fused call n writes buffer n & 1: its predecessor wrote the other one
  a.ll_plan_next = h->d_ll_plan[h->fused_calls & 1ull];
  h->plan_tag = a.tag;
  a.k_total = (double)(p.rollouts_total > 0 ? p.rollouts_total : p.rollouts);
  a.umax = p.max_wheel_vel;
  a.uinit[0] = h->uinit[0]; a.uinit[1] = h->uinit[1];
  a.u_next = h->d_u[h->cur ^ 1];
  a.out = h->d_out_host;               // mapped pinned memory: the controls land on the host without a copy operation
  a.out_seq = reinterpret_cast<unsigned long long *>(h->d_out_host + 2);
  a.dbg = h->d_dbg;
  a.seq = ++h->out_seq;
  a.stepstats = h->capture ? h->d_stepstats : nullptr;      // the weights tap needs them; a production call skips the stores
  a.merged = h->d_merged;
  a.rank = 0; a.nranks = 1;
  if (h->nranks > 1 && h->p2p_ready) {
    for (int r = 0; r < h->nranks; r++) a.peer[r] = static_cast<unsigned long long *>(h->peer_base[r]);
    a.rank = h->rank; a.nranks = h->nranks;
    h->xchg_call++;
    a.parity = (int)(h->xchg_call & 1ull);
    a.call_id = (uint32_t)(h->xchg_call % 0xFFFFFFFFull) + 1u;
  }
}

// the variates of call `call` into buffer `slot` (no-op when it already holds them)
int ensure_noise(b2n_mppi *h, uint32_t call, int slot, bool pdl)
{
  if (h->z_call[slot] == (long long)call && h->z_seed[slot] == h->seed) return B2N_OK;
  const size_t n = (size_t)h->K * (h->T / 2);
  B2N_REQUIRE(h->d_z[slot], B2N_ERR_CUDA, "no buffer for the variates (the horizon does not fit a production shape)");
  MppiNoiseArgs na;
  std::memset(&na, 0, sizeof(na));
  na.z_ready = h->d_zready; na.zbuf = h->d_z[slot]; na.K = h->K; na.half_T = h->T / 2;
  na.half_shift = -1;
  for (int s = 0; s < 30; s++) if ((1 << s) == na.half_T) na.half_shift = s;
  na.k_offset = h->p.rollout_offset; na.call = call;
  for (int r = 0; r < 10; r++) {
    na.key0[r] = (uint32_t)h->seed + (uint32_t)r * 0x9E3779B9u;
    na.key1[r] = (uint32_t)(h->seed >> 32) + (uint32_t)r * 0xBB67AE85u;
  }
  cudaLaunchConfig_t cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  const unsigned want = (unsigned)((n + 255) / 256);
  cfg.gridDim = dim3(std::max(1u, std::min(want, (unsigned)h->n_sm * (unsigned)h->noise_ctas_per_sm))); cfg.blockDim = dim3(256); cfg.stream = h->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = (pdl && h->use_pdl) ? 1 : 0;
  B2N_CUDA(cudaLaunchKernelEx(&cfg, mppi_noise_kernel, na));
  h->zready_total += cfg.gridDim.x;
  h->launches++;
  h->z_call[slot] = (long long)call; h->z_seed[slot] = h->seed; h->z_ready_at[slot] = h->zready_total;
  return B2N_OK;
}

// 0 generic, 1 FAST, 2 FAST + obstacle term.  The FAST variants need: own noise, no taps, a full last lane, TMA stores, and
// every half-step heading increment inside the short Taylor range of mppi_sincos_small (1/16): |h w / 2| <= c_w (|uL| + |uR|) / 2
// with |u| <= max|plan| + 5.78 sigma (the binary32 Box-Muller cannot exceed sqrt(48 ln 2) = 5.77 standard deviations)
int pick_variant(const b2n_mppi *h, const MppiArgs &a)
{
  const double u_bound = 2.0 * h->plan_abs_max + 5.78 * (a.sigL + a.sigR);
  const bool fast = !a.external_noise && !a.capture && a.tma_store && h->T == h->S * h->G &&
                    0.5 * std::fabs(a.c_w) * u_bound <= 0.0625 && !h->force_generic;
  return fast ? (a.obs_on ? 2 : 1) : 0;
}

// update kernel behind the rollout kernel with programmatic dependent launch: its CTAs may be scheduled while the
// rollout grid drains and block in cudaGridDependencySynchronize() until the partials are complete and visible
int launch_update(b2n_mppi *h, const MppiUpdateArgs &u)
{
  cudaLaunchConfig_t cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)h->T); cfg.blockDim = dim3(kMppiUpdateThreads); cfg.dynamicSmemBytes = 0; cfg.stream = h->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = h->use_pdl ? 1 : 0;
  B2N_CUDA(cudaLaunchKernelEx(&cfg, mppi_update_kernel, u));
  h->launches++;
  return B2N_OK;
}

int enqueue_call(b2n_mppi *h, double x, double y, double theta, bool noise_behind = true)
{
  const b2n_mppi_params &p = h->p;
  using clk = std::chrono::steady_clock;
  clk::time_point pt0, pt1, pt2;
  if (h->prof) pt0 = clk::now();
  MppiArgs a = make_args(h, x, y, theta);
  h->last_slot = h->ring_pos;
  a.states = h->d_states + (size_t)h->ring_pos * h->K * h->T * 3;
  h->ring_pos = (h->ring_pos + 1) % h->ring;
  const int variant = pick_variant(h, a);
  h->last_fast = variant != 0;
  const bool ahead = variant != 0;
  if (ahead) {
    // normally a no-op: the variates were drawn behind the previous call
    if (int rc = ensure_noise(h, h->call, h->zslot, false)) return rc;
    a.zbuf = h->d_z[h->zslot];
  }
  const bool nccl_transport = h->nranks > 1 && !h->p2p_ready;
  if (!nccl_transport) arm_tail(h, a);      // the whole call is this one launch
  // a fused call behind a fused call reads two things from the grids in front of it - the variates and the plan - and
  // waits for exactly those (MppiArgs::skip_wait).  Any other device-resident input keeps the wait for the grids themselves:
  // an obstacle field whose device address was handed out (b2n_mppi_obstacle_field_device) may be rewritten in place by
  // work on this stream; one copied from the host (b2n_mppi_set_obstacle_field) cannot change under a queued call
  a.z_ready = h->d_zready; a.z_need = h->z_ready_at[h->zslot];
  // ... and only when calls are being queued up (the previous call's controls have not been collected): behind a call the
  // host has already waited for, the grids in front are complete and the plain wait costs nothing (measured 0.4 us less
  // than the count's round trip)
  a.skip_wait = (ahead && a.tail && a.plan_tag != 0 && !(a.obs_on && h->obs_external) && h->use_pdl && !h->full_wait && h->waited_seq + 1 != a.seq) ? 1 : 0;

  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (h->timing && h->ev_used + 2 <= h->ev.size()) {
    e0 = h->ev[h->ev_used]; e1 = h->ev[h->ev_used + 1];
    h->ev_used += 2;
    B2N_CUDA(cudaEventRecord(e0, h->stream));
  }
  if (h->prof) pt1 = clk::now();
  launch_rollout(h, a, variant);
  B2N_CUDA(cudaGetLastError());
  if (e1) B2N_CUDA(cudaEventRecord(e1, h->stream));
  h->launches++;
  if (h->prof) pt2 = clk::now();

  if (nccl_transport) {
    // baseline transport of a sharded job: local merge -> one allgather of [T][6] doubles -> identical update on every
    // rank, three more launches per call
    MppiUpdateArgs u;
    std::memset(&u, 0, sizeof(u));
    u.T = h->T;
    u.inv_lambda = a.inv_lambda;
    u.k_total = (double)(p.rollouts_total > 0 ? p.rollouts_total : p.rollouts);
    u.umax = p.max_wheel_vel;
    u.uinit[0] = h->uinit[0]; u.uinit[1] = h->uinit[1];
    u.u_cur = h->d_u[h->cur];
    u.u_next = h->d_u[h->cur ^ 1];
    u.out = h->d_out_host;
    u.out_seq = reinterpret_cast<unsigned long long *>(h->d_out_host + 2);
    u.seq = ++h->out_seq;
    u.stepstats = h->d_stepstats;
    u.merged = h->d_merged;
    u.partials = h->d_partials; u.n_partials = h->grid; u.p_stride = 6; u.t_stride = 6 * h->grid; u.merge_only = 1;
    if (int rc = launch_update(h, u)) return rc;
    ncclResult_t r = ncclAllGather(h->d_merged, h->d_gathered, (size_t)h->T * 6, ncclDouble, h->comm, h->stream);
    B2N_REQUIRE(r == ncclSuccess, B2N_ERR_COMM, "ncclAllGather: %s", ncclGetErrorString(r));
    u.partials = h->d_gathered; u.n_partials = h->nranks; u.p_stride = 6 * h->T; u.t_stride = 6; u.merge_only = 0;
    if (int rc = launch_update(h, u)) return rc;
    h->plan_tag = 0;    // written by the update kernel: the plain array, ordered by the stream
  }

  h->cur ^= 1;
  h->call++;
  h->zslot = (h->zslot + 1) % kZBuf;
  h->ext_armed = false;
  h->pending = true;
  // the variates of the call after the next one, behind this call: off every call's critical path
  if (ahead && h->noise_ahead && noise_behind) {
    // (with programmatic launch behind a fused call only: that call's successor finds the plan through its tagged words, not through this grid)
    // (h->call is the NEXT call by now: its variates are normally there already, drawn behind the call before this one)
    for (uint32_t ahead_by = 0; ahead_by < 2; ahead_by++)
      if (int rc = ensure_noise(h, h->call + ahead_by, (h->zslot + (int)ahead_by) % kZBuf, !nccl_transport)) return rc;
  }
  if (h->prof) {
    h->prof_ns[0] += std::chrono::duration<double, std::nano>(pt1 - pt0).count();
    h->prof_ns[1] += std::chrono::duration<double, std::nano>(pt2 - pt1).count();
    h->prof_ns[2] += std::chrono::duration<double, std::nano>(clk::now() - pt2).count();
  }
  return B2N_OK;
}

} // namespace

extern "C" {

int b2n_mppi_create(const b2n_mppi_params *params, b2n_mppi **out)
{
  B2N_REQUIRE(params && out, B2N_ERR_INVALID_ARGUMENT, "b2n_mppi_create: null argument");
  *out = nullptr;
  const b2n_mppi_params &p = *params;
  B2N_REQUIRE(p.rollouts > 0, B2N_ERR_INVALID_ARGUMENT, "rollouts must be positive (got %d)", p.rollouts);
  B2N_REQUIRE(p.dt > 0.0 && p.horizon > 0.0, B2N_ERR_INVALID_ARGUMENT, "horizon and dt must be positive");
  B2N_REQUIRE(p.lambda > 0.0, B2N_ERR_INVALID_ARGUMENT, "lambda must be positive");
  B2N_REQUIRE(p.wheel_base != 0.0, B2N_ERR_INVALID_ARGUMENT, "wheel_base must be non-zero");
  B2N_REQUIRE(p.ul_var >= 0.0 && p.ur_var >= 0.0, B2N_ERR_INVALID_ARGUMENT, "control variances must be non-negative");
  const int T = static_cast<int>(p.horizon / p.dt);   // mppi.cpp:47, truncation included
  B2N_REQUIRE(T >= 1, B2N_ERR_INVALID_ARGUMENT, "horizon/dt gives %d steps", T);
  B2N_REQUIRE(T <= kMppiMaxT, B2N_ERR_UNSUPPORTED, "steps = %d exceeds the %d the rollout kernel is built for", T, kMppiMaxT);

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    set_error("b2n_mppi_create: no CUDA device (libb2nav has no CPU path)");
    return B2N_ERR_CUDA;
  }
  b2n_mppi *h = new (std::nothrow) b2n_mppi();
  B2N_REQUIRE(h, B2N_ERR_CUDA, "out of host memory");
  h->p = p;
  h->T = T;
  h->K = p.rollouts;
  pick_shape(T, h->S, h->G, h->NW);
  h->plan_abs_max = std::fabs(p.max_wheel_vel);   // the plan starts at 0 and every update is clamped to +-max_wheel_vel
  if (const char *env = std::getenv("B2N_MPPI_GENERIC")) h->force_generic = env[0] == '1';
  if (p.device >= 0) h->device = p.device; else cudaGetDevice(&h->device);

#define B2N_TRY(expr)                                                                          \
  do {                                                                                         \
    cudaError_t e__ = (expr);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      set_error("%s: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__);         \
      b2n_mppi_destroy(h);                                                                     \
      return B2N_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

  B2N_TRY(cudaSetDevice(h->device));
  cudaDeviceProp prop;
  B2N_TRY(cudaGetDeviceProperties(&prop, h->device));
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; libb2nav is built for sm_100a only", h->device, prop.major, prop.minor);
    b2n_mppi_destroy(h);
    return B2N_ERR_CUDA;
  }
  h->n_sm = prop.multiProcessorCount;
  B2N_TRY(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
  h->stream = h->own_stream;
  B2N_TRY(configure(h));
  h->tma_store = ((size_t)T * 3 * sizeof(float)) % 16 == 0;   // cp.async.bulk moves 16-byte granules

  const size_t KT = (size_t)h->K * T;
  B2N_TRY(cudaMalloc(&h->d_u[0], 2 * T * sizeof(double)));
  B2N_TRY(cudaMalloc(&h->d_u[1], 2 * T * sizeof(double)));
  B2N_TRY(cudaMemsetAsync(h->d_u[0], 0, 2 * T * sizeof(double), h->stream));   // mppi.cpp:157-170
  B2N_TRY(cudaMemsetAsync(h->d_u[1], 0, 2 * T * sizeof(double), h->stream));
  B2N_TRY(cudaMalloc(&h->d_states, KT * 3 * sizeof(float)));
  B2N_TRY(cudaMalloc(&h->d_partials, (size_t)h->grid * T * 6 * sizeof(double)));
  if (T == h->S * h->G) {
    // the production variants read their variates from these (allocated here, never inside a call: an allocation can
    // synchronise the device, and a sharded call in flight waits for ranks that this thread has not launched yet)
    for (int i = 0; i < kZBuf; i++) B2N_TRY(cudaMalloc(&h->d_z[i], KT / 2 * sizeof(float4)));
  }
  B2N_TRY(cudaMalloc(&h->d_ll_partials, (size_t)h->grid * T * 3 * sizeof(MppiLL)));
  B2N_TRY(cudaMemsetAsync(h->d_ll_partials, 0, (size_t)h->grid * T * 3 * sizeof(MppiLL), h->stream));     // tag 0: no call's
  B2N_TRY(cudaMalloc(&h->d_zready, sizeof(unsigned long long)));
  B2N_TRY(cudaMemsetAsync(h->d_zready, 0, sizeof(unsigned long long), h->stream));
  B2N_TRY(cudaMalloc(&h->d_arrive, sizeof(unsigned long long)));
  B2N_TRY(cudaMemsetAsync(h->d_arrive, 0, sizeof(unsigned long long), h->stream));
  for (int i = 0; i < 2; i++) {
    B2N_TRY(cudaMalloc(&h->d_ll_plan[i], (size_t)T * sizeof(MppiLL)));
    B2N_TRY(cudaMemsetAsync(h->d_ll_plan[i], 0, (size_t)T * sizeof(MppiLL), h->stream));
  }
  if (const char *env = std::getenv("B2N_MPPI_DEBUG_TIMES")) {
    if (env[0] == '1') {
      B2N_TRY(cudaMalloc(&h->d_dbg, (size_t)(h->grid + T) * kMppiDbgSlots * sizeof(unsigned long long)));
      B2N_TRY(cudaMemsetAsync(h->d_dbg, 0, (size_t)(h->grid + T) * kMppiDbgSlots * sizeof(unsigned long long), h->stream));
    }
  }
  B2N_TRY(cudaMalloc(&h->d_merged, (size_t)T * 6 * sizeof(double)));
  B2N_TRY(cudaMalloc(&h->d_stepstats, (size_t)T * 2 * sizeof(double)));
  B2N_TRY(cudaHostAlloc(&h->h_out, 4 * sizeof(double), cudaHostAllocMapped));
  std::memset(h->h_out, 0, 4 * sizeof(double));
  B2N_TRY(cudaHostGetDevicePointer(&h->d_out_host, h->h_out, 0));
  if (const char *env = std::getenv("B2N_MPPI_PDL")) h->use_pdl = env[0] != '0';
  if (const char *env = std::getenv("B2N_MPPI_NOISE_AHEAD")) h->noise_ahead = env[0] != '0';
  if (const char *env = std::getenv("B2N_MPPI_FULL_WAIT")) h->full_wait = env[0] == '1';
  if (const char *env = std::getenv("B2N_MPPI_NOISE_CTAS")) { const int n = std::atoi(env); if (n >= 1 && n <= 8) h->noise_ctas_per_sm = n; }
  B2N_TRY(cudaStreamSynchronize(h->stream));
#undef B2N_TRY
  *out = h;
  return B2N_OK;
}

void b2n_mppi_destroy(b2n_mppi *h)
{
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->comm) ncclCommDestroy(h->comm);
  for (int r = 0; r < (int)h->peer_base.size(); r++)
    if (!h->p2p_local && h->peer_base[r] && h->peer_base[r] != h->xchg) cudaIpcCloseMemHandle(h->peer_base[r]);
  cudaFree(h->xchg);
  for (auto e : h->ev) cudaEventDestroy(e);
  cudaFree(h->d_u[0]); cudaFree(h->d_u[1]); cudaFree(h->d_states); cudaFree(h->d_partials);
  cudaFree(h->d_merged); cudaFree(h->d_gathered); cudaFree(h->d_stepstats); cudaFree(h->d_ll_partials); cudaFree(h->d_arrive); cudaFree(h->d_zready); cudaFree(h->d_ll_plan[0]); cudaFree(h->d_ll_plan[1]); cudaFree(h->d_dbg);
  for (int i = 0; i < kZBuf; i++) cudaFree(h->d_z[i]);
  cudaFree(h->d_ext); cudaFree(h->d_J); cudaFree(h->d_du); cudaFree(h->d_w); cudaFree(h->d_obs);
  if (h->h_out) cudaFreeHost(h->h_out);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  cudaGetLastError();
  delete h;
}

int b2n_mppi_steps(const b2n_mppi *h) { return h ? h->T : B2N_ERR_INVALID_ARGUMENT; }

int b2n_mppi_set_initial_controls(b2n_mppi *h, double ul, double ur)
{
  B2N_REQUIRE(h, B2N_ERR_INVALID_ARGUMENT, "null handle");
  if (int rc = set_device(h)) return rc;
  h->uinit[0] = ul; h->uinit[1] = ur;
  h->plan_abs_max = std::max(h->plan_abs_max, std::max(std::fabs(ul), std::fabs(ur)));
  std::vector<double> u(2 * h->T);
  for (int t = 0; t < h->T; t++) { u[t] = ul; u[h->T + t] = ur; }
  B2N_CUDA(cudaMemcpyAsync(h->d_u[h->cur], u.data(), u.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  B2N_CUDA(cudaStreamSynchronize(h->stream));
  h->plan_tag = 0;      // the plain array is the plan now
  return B2N_OK;
}

int b2n_mppi_set_waypoint(b2n_mppi *h, double x, double y, double theta)
{
  B2N_REQUIRE(h, B2N_ERR_INVALID_ARGUMENT, "null handle");
  h->xd[0] = x; h->xd[1] = y; h->xd[2] = theta;
  return B2N_OK;
}

int b2n_mppi_enqueue(b2n_mppi *h, double x, double y, double theta)
{
  B2N_REQUIRE(h, B2N_ERR_INVALID_ARGUMENT, "null handle");
  if (int rc = set_device(h)) return rc;
  return enqueue_call(h, x, y, theta);
}

int b2n_mppi_enqueue_many(b2n_mppi *h, double x, double y, double theta, int calls)
{
  B2N_REQUIRE(h && calls >= 0, B2N_ERR_INVALID_ARGUMENT, "bad argument");
  if (int rc = set_device(h)) return rc;
  for (int i = 0; i < calls; i++)
    if (int rc = enqueue_call(h, x, y, theta)) return rc;
  return B2N_OK;
}

int b2n_mppi_wait(b2n_mppi *h, double *ul, double *ur)
{
  B2N_REQUIRE(h, B2N_ERR_INVALID_ARGUMENT, "null handle");
  B2N_REQUIRE(h->pending, B2N_ERR_INVALID_ARGUMENT, "b2n_mppi_wait: nothing enqueued");
  // the kernel publishes the controls in mapped pinned memory as four 8-byte words, each carrying a 4-byte half of a control
  // and the low half of the call's sequence number: poll until all four show this call's number (about a microsecond after
  // the stores) instead of a stream synchronisation; every so often make sure the stream has not failed
  volatile unsigned long long *w = reinterpret_cast<volatile unsigned long long *>(h->h_out);
  const unsigned long long tag = h->out_seq & 0xFFFFFFFFull;
  const auto t_start = std::chrono::steady_clock::now();
  unsigned long long v[4];
  for (unsigned spins = 0;; spins++) {
    v[0] = w[0]; v[1] = w[1]; v[2] = w[2]; v[3] = w[3];
    if ((v[0] >> 32) == tag && (v[1] >> 32) == tag && (v[2] >> 32) == tag && (v[3] >> 32) == tag) break;
    if ((spins & 0xFFFFu) == 0xFFFFu) {
      const cudaError_t e = cudaStreamQuery(h->stream);
      if (e != cudaSuccess && e != cudaErrorNotReady) { set_error("stream failed while waiting for the controls: %s", cudaGetErrorString(e)); return B2N_ERR_CUDA; }
      if (std::chrono::steady_clock::now() - t_start > std::chrono::seconds(60)) {
        set_error("no controls after 60 s (a sharded job waits for every rank: is one of them gone?)");
        return B2N_ERR_COMM;
      }
    }
  }
  const unsigned long long bl = (v[0] & 0xFFFFFFFFull) | (v[1] << 32), br = (v[2] & 0xFFFFFFFFull) | (v[3] << 32);
  double dl, dr;
  std::memcpy(&dl, &bl, 8); std::memcpy(&dr, &br, 8);
  if (ul) *ul = dl;
  if (ur) *ur = dr;
  h->waited_seq = h->out_seq;
  return B2N_OK;
}

int b2n_mppi_new_controls(b2n_mppi *h, double x, double y, double theta, double *ul, double *ur)
{
  B2N_REQUIRE(h && ul && ur, B2N_ERR_INVALID_ARGUMENT, "null argument");
  if (int rc = set_device(h)) return rc;
  // (the next call's variates are drawn right behind this call's kernel; drawing them after the controls are out, while the
  // host turns the pose around, measured 1.5 us SLOWER per call: launch + run of that kernel outlasts the turnaround)
  if (int rc = enqueue_call(h, x, y, theta)) return rc;
  return b2n_mppi_wait(h, ul, ur);
}

int b2n_mppi_seed(b2n_mppi *h, uint64_t seed, uint32_t first_call)
{
  B2N_REQUIRE(h, B2N_ERR_INVALID_ARGUMENT, "null handle");
  h->seed = seed; h->call = first_call;
  return B2N_OK;
}

int b2n_mppi_set_noise(b2n_mppi *h, const double *du, size_t count)
{
  B2N_REQUIRE(h && du, B2N_ERR_INVALID_ARGUMENT, "null argument");
  const size_t need = (size_t)h->K * h->T * 2;
  B2N_REQUIRE(count == need, B2N_ERR_INVALID_ARGUMENT, "noise count %zu, expected K*T*2 = %zu", count, need);
  if (int rc = set_device(h)) return rc;
  if (!h->d_ext) B2N_CUDA(cudaMalloc(&h->d_ext, need * sizeof(double)));
  B2N_CUDA(cudaMemcpyAsync(h->d_ext, du, need * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  B2N_CUDA(cudaStreamSynchronize(h->stream));
  h->ext_armed = true;
  return B2N_OK;
}

int b2n_mppi_set_capture(b2n_mppi *h, int on)
{
  B2N_REQUIRE(h, B2N_ERR_INVALID_ARGUMENT, "null handle");
  if (int rc = set_device(h)) return rc;
  const size_t KT = (size_t)h->K * h->T;
  if (on && !h->d_J) {
    B2N_CUDA(cudaMalloc(&h->d_J, KT * sizeof(double)));
    B2N_CUDA(cudaMalloc(&h->d_du, KT * 2 * sizeof(double)));
    B2N_CUDA(cudaMalloc(&h->d_w, KT * sizeof(double)));
  }
  h->capture = on ? 1 : 0;
  return B2N_OK;
}

static int copy_out(b2n_mppi *h, void *dst, const void *src, size_t bytes)
{
  if (int rc = set_device(h)) return rc;
  B2N_CUDA(cudaStreamSynchronize(h->stream));
  B2N_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
  return B2N_OK;
}

int b2n_mppi_get_states(b2n_mppi *h, float *out, size_t count)
{
  B2N_REQUIRE(h && out, B2N_ERR_INVALID_ARGUMENT, "null argument");
  const size_t need = (size_t)h->K * h->T * 3;
  B2N_REQUIRE(count == need, B2N_ERR_INVALID_ARGUMENT, "states count %zu, expected K*T*3 = %zu", count, need);
  return copy_out(h, out, h->d_states + (size_t)h->last_slot * need, need * sizeof(float));
}

int b2n_mppi_get_cost_to_go(b2n_mppi *h, double *out, size_t count)
{
  B2N_REQUIRE(h && out, B2N_ERR_INVALID_ARGUMENT, "null argument");
  B2N_REQUIRE(h->capture && h->d_J, B2N_ERR_INVALID_ARGUMENT, "cost-to-go needs b2n_mppi_set_capture(h, 1) before the call");
  const size_t need = (size_t)h->K * h->T;
  B2N_REQUIRE(count == need, B2N_ERR_INVALID_ARGUMENT, "count %zu, expected K*T = %zu", count, need);
  return copy_out(h, out, h->d_J, need * sizeof(double));
}

int b2n_mppi_get_noise(b2n_mppi *h, double *out, size_t count)
{
  B2N_REQUIRE(h && out, B2N_ERR_INVALID_ARGUMENT, "null argument");
  B2N_REQUIRE(h->capture && h->d_du, B2N_ERR_INVALID_ARGUMENT, "noise needs b2n_mppi_set_capture(h, 1) before the call");
  const size_t need = (size_t)h->K * h->T * 2;
  B2N_REQUIRE(count == need, B2N_ERR_INVALID_ARGUMENT, "count %zu, expected K*T*2 = %zu", count, need);
  return copy_out(h, out, h->d_du, need * sizeof(double));
}

int b2n_mppi_get_weights(b2n_mppi *h, double *out, size_t count)
{
  B2N_REQUIRE(h && out, B2N_ERR_INVALID_ARGUMENT, "null argument");
  B2N_REQUIRE(h->capture && h->d_J, B2N_ERR_INVALID_ARGUMENT, "weights need b2n_mppi_set_capture(h, 1) before the call");
  const size_t need = (size_t)h->K * h->T;
  B2N_REQUIRE(count == need, B2N_ERR_INVALID_ARGUMENT, "count %zu, expected K*T = %zu", count, need);
  if (int rc = set_device(h)) return rc;
  const int threads = 256;
  mppi_weights_kernel<<<(unsigned)((need + threads - 1) / threads), threads, 0, h->stream>>>(h->d_J, h->d_stepstats, h->d_w, h->K,
                                                                                            h->T, 1.0 / h->p.lambda);
  B2N_CUDA(cudaGetLastError());
  h->launches++;
  return copy_out(h, out, h->d_w, need * sizeof(double));
}

int b2n_mppi_get_plan(b2n_mppi *h, double *out, size_t count)
{
  B2N_REQUIRE(h && out, B2N_ERR_INVALID_ARGUMENT, "null argument");
  B2N_REQUIRE(count == (size_t)2 * h->T, B2N_ERR_INVALID_ARGUMENT, "plan count %zu, expected 2*T = %d", count, 2 * h->T);
  return copy_out(h, out, h->d_u[h->cur], count * sizeof(double));
}

int b2n_mppi_set_plan(b2n_mppi *h, const double *u, size_t count)
{
  B2N_REQUIRE(h && u, B2N_ERR_INVALID_ARGUMENT, "null argument");
  B2N_REQUIRE(count == (size_t)2 * h->T, B2N_ERR_INVALID_ARGUMENT, "plan count %zu, expected 2*T = %d", count, 2 * h->T);
  if (int rc = set_device(h)) return rc;
  for (size_t i = 0; i < count; i++) h->plan_abs_max = std::max(h->plan_abs_max, std::fabs(u[i]));
  B2N_CUDA(cudaMemcpyAsync(h->d_u[h->cur], u, count * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  B2N_CUDA(cudaStreamSynchronize(h->stream));
  h->plan_tag = 0;      // the plain array is the plan now
  return B2N_OK;
}

int b2n_mppi_get_partials(b2n_mppi *h, double *out, size_t count)
{
  B2N_REQUIRE(h && out, B2N_ERR_INVALID_ARGUMENT, "null argument");
  B2N_REQUIRE(count == (size_t)6 * h->T, B2N_ERR_INVALID_ARGUMENT, "partials count %zu, expected 6*T = %d", count, 6 * h->T);
  if (int rc = set_device(h)) return rc;
  MppiUpdateArgs u;
  std::memset(&u, 0, sizeof(u));
  u.T = h->T; u.inv_lambda = 1.0 / h->p.lambda; u.partials = h->d_partials; u.n_partials = h->grid; u.p_stride = 6; u.t_stride = 6 * h->grid;
  u.merge_only = 1; u.merged = h->d_merged;
  if (int rc = launch_update(h, u)) return rc;
  return copy_out(h, out, h->d_merged, count * sizeof(double));
}

int b2n_mppi_set_obstacle_field(b2n_mppi *h, const float *dist, int xsize, int ysize, double xmin, double ymin,
                                double resolution, double weight, double d0, double off_map)
{
  B2N_REQUIRE(h, B2N_ERR_INVALID_ARGUMENT, "null handle");
  if (int rc = set_device(h)) return rc;
  if (!dist) { h->obs_on = 0; return B2N_OK; }
  B2N_REQUIRE(xsize > 0 && ysize > 0 && resolution > 0.0, B2N_ERR_INVALID_ARGUMENT, "bad obstacle field geometry");
  // the reference's cell index is i * xsize + j with i < xsize, j < ysize (grid_mapper.cpp:890-898): square grids only, as b2n_pf_create
  B2N_REQUIRE(xsize == ysize, B2N_ERR_UNSUPPORTED, "obstacle field must be square (got %d x %d)", xsize, ysize);
  cudaFree(h->d_obs); h->d_obs = nullptr;
  const size_t n = (size_t)xsize * ysize;
  B2N_CUDA(cudaMalloc(&h->d_obs, n * sizeof(float)));
  B2N_CUDA(cudaMemcpyAsync(h->d_obs, dist, n * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  B2N_CUDA(cudaStreamSynchronize(h->stream));
  h->obs_on = 1; h->obs_xsize = xsize; h->obs_ysize = ysize; h->obs_xmin = xmin; h->obs_ymin = ymin;
  h->obs_res = resolution; h->obs_weight = weight; h->obs_d0 = d0; h->obs_off = off_map;
  h->obs_external = false;
  return B2N_OK;
}

int b2n_mppi_obstacle_field_device(b2n_mppi *h, int xsize, int ysize, double xmin, double ymin, double resolution, double weight,
                                   double d0, double off_map, float **device_field)
{
  B2N_REQUIRE(h && device_field, B2N_ERR_INVALID_ARGUMENT, "null argument");
  B2N_REQUIRE(xsize > 0 && ysize > 0 && resolution > 0.0, B2N_ERR_INVALID_ARGUMENT, "bad obstacle field geometry");
  B2N_REQUIRE(xsize == ysize, B2N_ERR_UNSUPPORTED, "obstacle field must be square (got %d x %d)", xsize, ysize);
  if (int rc = set_device(h)) return rc;
  B2N_CUDA(cudaStreamSynchronize(h->stream));
  const size_t n = (size_t)xsize * ysize;
  if (!h->d_obs || (size_t)h->obs_xsize * h->obs_ysize != n) {
    cudaFree(h->d_obs); h->d_obs = nullptr;
    B2N_CUDA(cudaMalloc(&h->d_obs, n * sizeof(float)));
    B2N_CUDA(cudaMemset(h->d_obs, 0x7f, n * sizeof(float)));     // ~3.4e38: no obstacle anywhere until the producer writes
  }
  h->obs_on = 1; h->obs_xsize = xsize; h->obs_ysize = ysize; h->obs_xmin = xmin; h->obs_ymin = ymin;
  h->obs_res = resolution; h->obs_weight = weight; h->obs_d0 = d0; h->obs_off = off_map;
  *device_field = h->d_obs;
  h->obs_external = true;
  return B2N_OK;
}

int b2n_mppi_set_stream(b2n_mppi *h, void *cuda_stream)
{
  B2N_REQUIRE(h, B2N_ERR_INVALID_ARGUMENT, "null handle");
  B2N_CUDA(cudaStreamSynchronize(h->stream));
  h->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : h->own_stream;
  return B2N_OK;
}

int b2n_mppi_set_state_ring(b2n_mppi *h, int n)
{
  B2N_REQUIRE(h && n >= 1, B2N_ERR_INVALID_ARGUMENT, "ring size must be >= 1");
  if (int rc = set_device(h)) return rc;
  B2N_CUDA(cudaStreamSynchronize(h->stream));
  float *fresh = nullptr;
  B2N_CUDA(cudaMalloc(&fresh, (size_t)n * h->K * h->T * 3 * sizeof(float)));
  cudaFree(h->d_states);
  h->d_states = fresh; h->ring = n; h->ring_pos = 0; h->last_slot = 0;
  return B2N_OK;
}

namespace
{
__global__ void box_muller_range_kernel(uint32_t first, uint32_t count, uint32_t rb, float2 *z)
{
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
    float2 o;
    box_muller_f32((first + i) << 9, rb, o.x, o.y);
    z[i] = o;
  }
}
} // namespace

int b2n_test_box_muller(uint32_t first, uint32_t count, uint32_t rb, float *z)
{
  B2N_REQUIRE(z && count > 0, B2N_ERR_INVALID_ARGUMENT, "bad argument");
  float2 *d = nullptr;
  B2N_CUDA(cudaMalloc(&d, (size_t)count * sizeof(float2)));
  box_muller_range_kernel<<<1184, 256>>>(first, count, rb, d);
  cudaError_t e = cudaMemcpy(z, d, (size_t)count * sizeof(float2), cudaMemcpyDeviceToHost);
  cudaFree(d);
  B2N_CUDA(e);
  return B2N_OK;
}

int b2n_mppi_debug_times(b2n_mppi *h, unsigned long long *out, size_t count, int *grid)
{
  B2N_REQUIRE(h && out && grid, B2N_ERR_INVALID_ARGUMENT, "null argument");
  B2N_REQUIRE(h->d_dbg, B2N_ERR_INVALID_ARGUMENT, "set B2N_MPPI_DEBUG_TIMES=1 before creating the handle");
  const size_t n = (size_t)(h->grid + h->T) * kMppiDbgSlots;
  B2N_REQUIRE(count >= n, B2N_ERR_INVALID_ARGUMENT, "need (grid + T) * %d = %zu words", kMppiDbgSlots, n);
  *grid = h->grid;
  return copy_out(h, out, h->d_dbg, n * sizeof(unsigned long long));
}

int b2n_mppi_last_variant(const b2n_mppi *h, int *fast)
{
  B2N_REQUIRE(h && fast, B2N_ERR_INVALID_ARGUMENT, "null argument");
  *fast = h->last_fast;
  return B2N_OK;
}

int b2n_mppi_launch_count(const b2n_mppi *h, uint64_t *launches)
{
  B2N_REQUIRE(h && launches, B2N_ERR_INVALID_ARGUMENT, "null argument");
  *launches = h->launches;
  return B2N_OK;
}

int b2n_mppi_set_kernel_timing(b2n_mppi *h, int on)
{
  B2N_REQUIRE(h, B2N_ERR_INVALID_ARGUMENT, "null handle");
  if (int rc = set_device(h)) return rc;
  if (on && h->ev.empty()) {
    h->ev.resize(2 * 4096);
    for (auto &e : h->ev) B2N_CUDA(cudaEventCreate(&e));
  }
  h->timing = on != 0;
  h->ev_used = 0;
  return B2N_OK;
}

int b2n_mppi_kernel_time(b2n_mppi *h, double *avg_ms, int *samples)
{
  B2N_REQUIRE(h && avg_ms && samples, B2N_ERR_INVALID_ARGUMENT, "null argument");
  if (int rc = set_device(h)) return rc;
  B2N_CUDA(cudaStreamSynchronize(h->stream));
  double total = 0.0;
  int n = 0;
  for (size_t i = 0; i + 1 < h->ev_used; i += 2) {
    float ms = 0.f;
    B2N_CUDA(cudaEventElapsedTime(&ms, h->ev[i], h->ev[i + 1]));
    total += ms; n++;
  }
  *avg_ms = n ? total / n : 0.0;
  *samples = n;
  h->ev_used = 0;
  return B2N_OK;
}

int b2n_mppi_time_rollout(b2n_mppi *h, double x, double y, double theta, int launches, double *avg_ms)
{
  B2N_REQUIRE(h && avg_ms && launches > 0, B2N_ERR_INVALID_ARGUMENT, "bad argument");
  if (int rc = set_device(h)) return rc;
  MppiArgs a = make_args(h, x, y, theta);
  const int variant = pick_variant(h, a);      // a.tail = 0: the rollout phase and the CTA partials only
  if (variant != 0) {
    if (int rc = ensure_noise(h, h->call, h->zslot, false)) return rc;
    a.zbuf = h->d_z[h->zslot];
  }
  a.plan_tag = h->plan_tag; a.ll_plan = h->d_ll_plan[h->fused_calls & 1ull];      // the plan as the last fused call left it
  a.z_ready = h->d_zready; a.z_need = h->z_ready_at[h->zslot];
  cudaEvent_t e0, e1;
  B2N_CUDA(cudaEventCreate(&e0));
  B2N_CUDA(cudaEventCreate(&e1));
  B2N_CUDA(cudaEventRecord(e0, h->stream));
  for (int i = 0; i < launches; i++) {
    a.states = h->d_states + (size_t)h->ring_pos * h->K * h->T * 3;
    h->ring_pos = (h->ring_pos + 1) % h->ring;
    a.call = h->call + (uint32_t)i;
    launch_rollout(h, a, variant);
  }
  B2N_CUDA(cudaGetLastError());
  B2N_CUDA(cudaEventRecord(e1, h->stream));
  B2N_CUDA(cudaEventSynchronize(e1));
  float ms = 0.f;
  B2N_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  h->launches += (uint64_t)launches;
  *avg_ms = (double)ms / launches;
  return B2N_OK;
}

int b2n_mppi_time_new_controls(b2n_mppi *h, double x, double y, double theta, int calls, double *avg_ms, double *ul, double *ur)
{
  B2N_REQUIRE(h && avg_ms && calls > 0, B2N_ERR_INVALID_ARGUMENT, "bad argument");
  double l = 0.0, r = 0.0;
  h->prof = std::getenv("B2N_MPPI_HOST_BREAKDOWN") != nullptr;
  for (double &v : h->prof_ns) v = 0.0;
  const auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < calls; i++)
    if (int rc = b2n_mppi_new_controls(h, x, y, theta, &l, &r)) return rc;
  const auto t1 = std::chrono::steady_clock::now();
  *avg_ms = std::chrono::duration<double, std::milli>(t1 - t0).count() / calls;
  if (h->prof) {
    const double total = std::chrono::duration<double, std::nano>(t1 - t0).count();
    std::fprintf(stderr, "host time per synchronous call: arguments %.2f us, launch of the call %.2f us, launch of the noise kernel %.2f us, rest (wait for the controls) %.2f us; total %.2f us\n",
                 h->prof_ns[0] / calls / 1e3, h->prof_ns[1] / calls / 1e3, h->prof_ns[2] / calls / 1e3,
                 (total - h->prof_ns[0] - h->prof_ns[1] - h->prof_ns[2]) / calls / 1e3, total / calls / 1e3);
    h->prof = false;
  }
  if (ul) *ul = l;
  if (ur) *ur = r;
  return B2N_OK;
}

int b2n_mppi_p2p_export(b2n_mppi *h, int nranks, void *handle64)
{
  B2N_REQUIRE(h && handle64, B2N_ERR_INVALID_ARGUMENT, "null argument");
  B2N_REQUIRE(nranks >= 2 && nranks <= kMppiMaxRanks, B2N_ERR_INVALID_ARGUMENT, "nranks must be in [2, %d]", kMppiMaxRanks);
  B2N_REQUIRE(!h->xchg, B2N_ERR_INVALID_ARGUMENT, "exchange area already exported");
  if (int rc = set_device(h)) return rc;
  h->xchg_bytes = (size_t)2 * nranks * h->T * kMppiXchgWords * sizeof(unsigned long long);
  h->xchg_flag_offset = 0;
  B2N_CUDA(cudaMalloc(&h->xchg, h->xchg_bytes));
  B2N_CUDA(cudaMemset(h->xchg, 0, h->xchg_bytes));
  h->xchg_nranks = nranks;
  cudaIpcMemHandle_t ipc;
  static_assert(sizeof(ipc) == 64, "cudaIpcMemHandle_t is 64 bytes");
  B2N_CUDA(cudaIpcGetMemHandle(&ipc, h->xchg));
  std::memcpy(handle64, &ipc, sizeof(ipc));
  return B2N_OK;
}

int b2n_mppi_p2p_init(b2n_mppi *h, int rank, int nranks, const void *handles)
{
  B2N_REQUIRE(h && handles, B2N_ERR_INVALID_ARGUMENT, "null argument");
  B2N_REQUIRE(h->xchg && nranks == h->xchg_nranks && rank >= 0 && rank < nranks, B2N_ERR_INVALID_ARGUMENT,
              "b2n_mppi_p2p_export(h, %d, ...) must come first on every rank", nranks);
  if (int rc = set_device(h)) return rc;
  h->peer_base.assign(nranks, nullptr);
  for (int r = 0; r < nranks; r++) {
    if (r == rank) { h->peer_base[r] = h->xchg; continue; }
    cudaIpcMemHandle_t ipc;
    std::memcpy(&ipc, static_cast<const char *>(handles) + (size_t)r * sizeof(ipc), sizeof(ipc));
    cudaError_t e = cudaIpcOpenMemHandle(&h->peer_base[r], ipc, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      set_error("cudaIpcOpenMemHandle for rank %d: %s (peer access between the GPUs is required)", r, cudaGetErrorString(e));
      return B2N_ERR_COMM;
    }
  }
  h->rank = rank; h->nranks = nranks; h->p2p_ready = true; h->xchg_call = 0;
  return B2N_OK;
}

int b2n_mppi_p2p_area(b2n_mppi *h, void **area)
{
  B2N_REQUIRE(h && area, B2N_ERR_INVALID_ARGUMENT, "null argument");
  B2N_REQUIRE(h->xchg, B2N_ERR_INVALID_ARGUMENT, "b2n_mppi_p2p_export must come first");
  *area = h->xchg;
  return B2N_OK;
}

int b2n_mppi_p2p_init_local(b2n_mppi *h, int rank, int nranks, void *const *areas)
{
  B2N_REQUIRE(h && areas, B2N_ERR_INVALID_ARGUMENT, "null argument");
  B2N_REQUIRE(h->xchg && nranks == h->xchg_nranks && rank >= 0 && rank < nranks, B2N_ERR_INVALID_ARGUMENT,
              "b2n_mppi_p2p_export(h, %d, ...) must come first on every rank", nranks);
  B2N_REQUIRE(areas[rank] == h->xchg, B2N_ERR_INVALID_ARGUMENT, "areas[rank] must be this handle's own area");
  h->peer_base.assign(areas, areas + nranks);
  h->p2p_local = true;
  h->rank = rank; h->nranks = nranks; h->p2p_ready = true; h->xchg_call = 0;
  return B2N_OK;
}

int b2n_mppi_comm_init(b2n_mppi *h, int rank, int nranks, const void *unique_id128)
{
  B2N_REQUIRE(h && unique_id128, B2N_ERR_INVALID_ARGUMENT, "null argument");
  B2N_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, B2N_ERR_INVALID_ARGUMENT, "bad rank %d of %d", rank, nranks);
  if (int rc = set_device(h)) return rc;
  ncclUniqueId id;
  static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
  std::memcpy(&id, unique_id128, sizeof(id));
  ncclResult_t r = ncclCommInitRank(&h->comm, nranks, id, rank);
  B2N_REQUIRE(r == ncclSuccess, B2N_ERR_COMM, "ncclCommInitRank: %s", ncclGetErrorString(r));
  h->rank = rank; h->nranks = nranks;
  cudaFree(h->d_gathered); h->d_gathered = nullptr;
  B2N_CUDA(cudaMalloc(&h->d_gathered, (size_t)nranks * h->T * 6 * sizeof(double)));
  return B2N_OK;
}

} // extern "C"
