// rbpf_api.cu - TEMPORARY: RBPF entry points not built yet in this commit.
#include "common.cuh"
using namespace b2n;
struct b2n_pf {};
extern "C" {
int b2n_pf_create(const b2n_pf_params *params, b2n_pf **out) { set_error("RBPF path not built yet"); return B2N_ERR_UNSUPPORTED; }
void b2n_pf_destroy(b2n_pf *h) {}
int b2n_pf_slam(b2n_pf *h, const float *scan, int n_beams, const double twist[3], const double cur_odom[3], const double prev_odom[3], int icp_ok, const double icp_pose[3]) { set_error("RBPF path not built yet"); return B2N_ERR_UNSUPPORTED; }
int b2n_pf_get_robot_state(b2n_pf *h, double pose[3]) { set_error("RBPF path not built yet"); return B2N_ERR_UNSUPPORTED; }
int b2n_pf_new_map(b2n_pf *h, int8_t *out, size_t count) { set_error("RBPF path not built yet"); return B2N_ERR_UNSUPPORTED; }
int b2n_pf_seed(b2n_pf *h, uint64_t seed, uint32_t first_call) { set_error("RBPF path not built yet"); return B2N_ERR_UNSUPPORTED; }
int b2n_pf_set_noise(b2n_pf *h, const double *z, size_t count) { set_error("RBPF path not built yet"); return B2N_ERR_UNSUPPORTED; }
int b2n_pf_grid_size(const b2n_pf *h, int *xsize, int *ysize) { set_error("RBPF path not built yet"); return B2N_ERR_UNSUPPORTED; }
int b2n_pf_get_weights(b2n_pf *h, double *out, size_t count) { set_error("RBPF path not built yet"); return B2N_ERR_UNSUPPORTED; }
int b2n_pf_set_weights(b2n_pf *h, const double *w, size_t count) { set_error("RBPF path not built yet"); return B2N_ERR_UNSUPPORTED; }
int b2n_pf_get_poses(b2n_pf *h, double *poses, double *prev_poses, size_t count) { set_error("RBPF path not built yet"); return B2N_ERR_UNSUPPORTED; }
int b2n_pf_set_poses(b2n_pf *h, const double *poses, size_t count) { set_error("RBPF path not built yet"); return B2N_ERR_UNSUPPORTED; }
int b2n_pf_get_resample(b2n_pf *h, int *neff, int *resampled, int32_t *ancestors, size_t count) { set_error("RBPF path not built yet"); return B2N_ERR_UNSUPPORTED; }
int b2n_pf_get_grid(b2n_pf *h, int particle, double *log_odds, float *occ_dist, int8_t *state, size_t count) { set_error("RBPF path not built yet"); return B2N_ERR_UNSUPPORTED; }
int b2n_pf_set_grid(b2n_pf *h, int particle, const double *log_odds, const float *occ_dist, const int8_t *state, const int32_t *occ_order, int n_occ, size_t count) { set_error("RBPF path not built yet"); return B2N_ERR_UNSUPPORTED; }
int b2n_pf_get_occ_order(b2n_pf *h, int particle, int32_t *keys, size_t cap, int *n_occ) { set_error("RBPF path not built yet"); return B2N_ERR_UNSUPPORTED; }
int b2n_pf_likelihoods(b2n_pf *h, const float *scan, int n_beams, double *out, size_t count) { set_error("RBPF path not built yet"); return B2N_ERR_UNSUPPORTED; }
int b2n_pf_set_stream(b2n_pf *h, void *cuda_stream) { set_error("RBPF path not built yet"); return B2N_ERR_UNSUPPORTED; }
int b2n_pf_launch_count(const b2n_pf *h, uint64_t *launches) { set_error("RBPF path not built yet"); return B2N_ERR_UNSUPPORTED; }
int b2n_pf_comm_init(b2n_pf *h, int rank, int nranks, const void *unique_id128) { set_error("RBPF path not built yet"); return B2N_ERR_UNSUPPORTED; }
}
