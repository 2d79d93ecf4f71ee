// icp_api.cu - extern "C" scan-matcher entry points of libb2nav (see include/b2nav.h).
// Host side of bmapping::ScanAlignment (reference: bmapping/src/bmapping/cloud_alignment.cpp:20-72,76-157): the wrapper's
// state machine and the point clouds are built here exactly as the reference builds them; the alignment itself runs
// in icp_align_kernel (our own statement of the ICP the reference takes from PCL - see icp_kernels.cuh).
//
// Compiled with -fmad=false so that the kernel's sums round like the CPU checker's.
#include <cmath>
#include <cstring>
#include <new>
#include <vector>

#include "icp_kernels.cuh"

using namespace b2n;

struct b2n_icp
{
  b2n_icp_params p;
  int device = 0, max_beams = 0;
  cudaStream_t stream = nullptr;
  float *d_cloud[2] = {nullptr, nullptr};   // previous (target) and current (source) clouds, swapped on success
  int n_prev = 0;
  float *h_cloud = nullptr;                 // pinned staging
  double *d_out = nullptr, *h_out = nullptr;
  bool first_scan_received = false;
  int last_iterations = 0, last_pairs = 0;
  double last_mse = 0.0;
  uint64_t launches = 0;
};

extern "C" {

int b2n_icp_create(const b2n_icp_params *params, b2n_icp **out)
{
  B2N_REQUIRE(params && out, B2N_ERR_INVALID_ARGUMENT, "b2n_icp_create: null argument");
  *out = nullptr;
  B2N_REQUIRE(params->max_iter >= 1 && params->max_correspondence_dist > 0.0, B2N_ERR_INVALID_ARGUMENT, "bad ICP settings");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    set_error("b2n_icp_create: no CUDA device (libb2nav has no CPU path)");
    return B2N_ERR_CUDA;
  }
  b2n_icp *h = new (std::nothrow) b2n_icp();
  B2N_REQUIRE(h, B2N_ERR_CUDA, "out of host memory");
  h->p = *params;
  h->max_beams = params->max_beams > 0 ? params->max_beams : 1024;
  if (h->max_beams > kIcpMaxPoints) { delete h; set_error("max_beams %d exceeds %d", params->max_beams, kIcpMaxPoints); return B2N_ERR_UNSUPPORTED; }
  if (params->device >= 0) h->device = params->device; else cudaGetDevice(&h->device);
  bool ok = cudaSetDevice(h->device) == cudaSuccess && cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaMalloc(&h->d_cloud[0], sizeof(float) * 2 * h->max_beams) == cudaSuccess &&
            cudaMalloc(&h->d_cloud[1], sizeof(float) * 2 * h->max_beams) == cudaSuccess &&
            cudaMallocHost(&h->h_cloud, sizeof(float) * 2 * h->max_beams) == cudaSuccess &&
            cudaMalloc(&h->d_out, 8 * sizeof(double)) == cudaSuccess && cudaMallocHost(&h->h_out, 8 * sizeof(double)) == cudaSuccess &&
            cudaFuncSetAttribute(icp_align_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)icp_smem_bytes(kIcpMaxPoints, kIcpMaxPoints)) == cudaSuccess;
  if (!ok) {
    set_error("b2n_icp_create: %s", cudaGetErrorString(cudaGetLastError()));
    b2n_icp_destroy(h);
    return B2N_ERR_CUDA;
  }
  *out = h;
  return B2N_OK;
}

void b2n_icp_destroy(b2n_icp *h)
{
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  cudaFree(h->d_cloud[0]); cudaFree(h->d_cloud[1]); cudaFree(h->d_out);
  if (h->h_cloud) cudaFreeHost(h->h_cloud);
  if (h->h_out) cudaFreeHost(h->h_out);
  if (h->stream) cudaStreamDestroy(h->stream);
  cudaGetLastError();
  delete h;
}

int b2n_icp_align(b2n_icp *h, const float *scan, int n_beams, const double t_init[3], double t[3], int *success)
{
  B2N_REQUIRE(h && scan && t_init && t && success, B2N_ERR_INVALID_ARGUMENT, "null argument");
  B2N_REQUIRE(n_beams > 0 && n_beams <= h->max_beams, B2N_ERR_INVALID_ARGUMENT, "scan has %d beams, handle was created for %d", n_beams, h->max_beams);
  B2N_CUDA(cudaSetDevice(h->device));
  const b2n_icp_params &p = h->p;
  // ScanAlignment::createPointCloud, cloud_alignment.cpp:76-157 (Trs = identity): float beam angle with the wrap rule,
  // range gate, `range` promoted to double, std::cos / std::sin of the float angle
  int n = 0;
  float beam_angle = p.beam_min;
  for (int i = 0; i < n_beams; i++) {
    const float range = scan[i];
    if (range >= p.range_min && range < p.range_max) {
      h->h_cloud[2 * n] = (float)((double)range * (double)std::cos(beam_angle));
      h->h_cloud[2 * n + 1] = (float)((double)range * (double)std::sin(beam_angle));
      n++;
    }
    beam_angle += p.beam_delta;
    if (p.beam_max < 0.0f && beam_angle <= p.beam_max) beam_angle = p.beam_min;
    else if (p.beam_max >= 0.0f && beam_angle >= p.beam_max) beam_angle = p.beam_min;
  }
  // pclICPWrapper, cloud_alignment.cpp:37-72
  if (!h->first_scan_received) {
    B2N_CUDA(cudaMemcpyAsync(h->d_cloud[0], h->h_cloud, sizeof(float) * 2 * n, cudaMemcpyHostToDevice, h->stream));
    B2N_CUDA(cudaStreamSynchronize(h->stream));
    h->n_prev = n;
    h->first_scan_received = true;
    *success = 1;                     // the reference returns true and leaves T untouched
    return B2N_OK;
  }
  h->last_iterations = 0; h->last_pairs = 0; h->last_mse = 0.0;
  if (n < 3 || h->n_prev < 3) { *success = 0; return B2N_OK; }
  B2N_CUDA(cudaMemcpyAsync(h->d_cloud[1], h->h_cloud, sizeof(float) * 2 * n, cudaMemcpyHostToDevice, h->stream));
  IcpArgs a;
  std::memset(&a, 0, sizeof(a));
  a.tgt = h->d_cloud[0]; a.nt = h->n_prev; a.src = h->d_cloud[1]; a.ns = n; a.max_iter = p.max_iter;
  a.Tinit[0] = t_init[0]; a.Tinit[1] = t_init[1]; a.Tinit[2] = t_init[2];
  a.max_d2 = p.max_correspondence_dist * p.max_correspondence_dist;
  a.transform_eps = p.transformation_epsilon; a.fitness_eps = p.euclidean_fitness_epsilon;
  a.out = h->d_out;
  icp_align_kernel<<<1, kIcpThreads, icp_smem_bytes(a.nt, a.ns), h->stream>>>(a);
  B2N_CUDA(cudaGetLastError());
  h->launches++;
  B2N_CUDA(cudaMemcpyAsync(h->h_out, h->d_out, 8 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  B2N_CUDA(cudaStreamSynchronize(h->stream));
  h->last_iterations = (int)h->h_out[4]; h->last_pairs = (int)h->h_out[5]; h->last_mse = h->h_out[6];
  if (h->h_out[3] == 0.0) { *success = 0; return B2N_OK; }     // "ICP FAILED TO CONVERGED": the old scan is kept
  t[0] = h->h_out[0]; t[1] = h->h_out[1]; t[2] = h->h_out[2];
  std::swap(h->d_cloud[0], h->d_cloud[1]);
  h->n_prev = n;
  *success = 1;
  return B2N_OK;
}

int b2n_icp_stats(const b2n_icp *h, int *iterations, int *pairs, double *mse, uint64_t *launches)
{
  B2N_REQUIRE(h, B2N_ERR_INVALID_ARGUMENT, "null handle");
  if (iterations) *iterations = h->last_iterations;
  if (pairs) *pairs = h->last_pairs;
  if (mse) *mse = h->last_mse;
  if (launches) *launches = h->launches;
  return B2N_OK;
}

} // extern "C"
