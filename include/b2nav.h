/* b2nav.h - C ABI of libb2nav.so: the B200 (sm_100a) implementation of the two data-parallel hot
 * loops of bostoncleek/ROS-Turtlebot-Navigation.
 *
 *   controller::MPPI::newControls()       reference: controller/src/controller/mppi.cpp:72-140
 *   bmapping::ParticleFilter::SLAM()      reference: bmapping/src/bmapping/particle_filter.cpp:141-251
 *
 * The reference has no FFI: its boundary is two C++ class surfaces linked statically into ROS nodes
 * (callers: nuturtle_robot/src/mppi_waypoints_node.cpp:186-199,216,257,265 and
 * bmapping/src/turtle_mapping_node.cpp:392-410,474,479,494).  The drop-in C++ classes with the same
 * names and signatures live in include/controller/mppi.hpp and include/bmapping/particle_filter.hpp
 * and are thin pimpl wrappers over the functions below.  Each entry point cites the reference
 * interface it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ or torch types;
 *   - every function returns B2N_OK (0) or a negative b2n_status; b2n_last_error() gives the text
 *     of the most recent failure on the calling thread;
 *   - handles own their device and host memory; in/out arrays are owned by the caller and only
 *     borrowed for the duration of the call;
 *   - handles are not thread-safe (the reference's callers are single-threaded spin loops);
 *   - there is NO CPU fallback: without a usable CUDA device create() fails with B2N_ERR_CUDA.
 */
#ifndef B2NAV_H
#define B2NAV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum b2n_status {
  B2N_OK = 0,
  B2N_ERR_INVALID_ARGUMENT = -1,   /* reference: std::invalid_argument / std::out_of_range */
  B2N_ERR_CUDA = -2,               /* no device, launch or allocation failure */
  B2N_ERR_OFF_MAP = -3,            /* reference: world2Grid/world2RowMajor throw, grid_mapper.cpp:817-825,854-862 */
  B2N_ERR_UNSUPPORTED = -4,        /* size outside what the kernels are built for */
  B2N_ERR_COMM = -5,               /* NCCL failure */
  B2N_ERR_NUMERIC = -6             /* reference: "eta is 0", particle_filter.cpp:577-580 */
} b2n_status;

const char *b2n_last_error(void);
/* number of CUDA devices visible to the library (0 when there is no usable driver) */
int b2n_device_count(void);
/* writes "libb2nav <version> sm_100a"; returns the number of bytes needed */
int b2n_version(char *buf, size_t cap);

/* ======================================================================================== MPPI */

typedef struct b2n_mppi b2n_mppi;

/* constructor arguments of controller::CartModel (mppi.hpp:33), controller::LossFunc (mppi.hpp:63)
 * and controller::MPPI (mppi.hpp:133-141), plus where this handle sits in a sharded job */
typedef struct b2n_mppi_params {
  double wheel_radius, wheel_base;
  double Q[3], R[2], P1[3];          /* diagonals */
  double lambda, max_wheel_vel, ul_var, ur_var, horizon, dt;
  int32_t rollouts;                  /* rollouts simulated by THIS handle */
  int32_t rollout_offset;            /* global index of this handle's rollout 0 (0 when unsharded) */
  int32_t rollouts_total;            /* K of the whole job (0 or == rollouts when unsharded) */
  int32_t device;                    /* CUDA device ordinal, -1 = current device */
} b2n_mppi_params;

/* controller::MPPI::MPPI, mppi.cpp:28-51; steps = (int)(horizon/dt) exactly as mppi.cpp:47 */
int b2n_mppi_create(const b2n_mppi_params *params, b2n_mppi **out);
void b2n_mppi_destroy(b2n_mppi *h);
int b2n_mppi_steps(const b2n_mppi *h);

/* controller::MPPI::setInitialControls, mppi.cpp:54-61 */
int b2n_mppi_set_initial_controls(b2n_mppi *h, double ul, double ur);
/* controller::MPPI::setWaypoint, mppi.cpp:64-69 */
int b2n_mppi_set_waypoint(b2n_mppi *h, double x, double y, double theta);
/* controller::MPPI::newControls, mppi.cpp:72-140: synchronous; the controls are on the host at return */
int b2n_mppi_new_controls(b2n_mppi *h, double x, double y, double theta, double *ul, double *ur);

/* The same call split in two so that many steps can be queued without a host round trip:
 * enqueue() launches the kernels for one newControls() on the handle's stream, wait() blocks until
 * the most recently enqueued call is done and returns its controls. */
int b2n_mppi_enqueue(b2n_mppi *h, double x, double y, double theta);
int b2n_mppi_wait(b2n_mppi *h, double *ul, double *ur);
/* `calls` times b2n_mppi_enqueue() with the same pose from one C loop (what a C++ caller's loop does; the bench's
 * device-resident leg uses it so that no interpreter sits between two launches) */
int b2n_mppi_enqueue_many(b2n_mppi *h, double x, double y, double theta, int calls);

/* Noise.  The reference draws from one process-global std::mt19937_64 (rigid2d/src/rigid2d/
 * utilities.cpp:12-24); a serial engine cannot feed K*T lanes, so the kernels use a counter-based
 * Philox4x32-10 stream keyed by (seed; call, rollout, step) - results do not depend on how the
 * rollouts are sharded.  For runs that must consume the reference's own variates, pass them in. */
int b2n_mppi_seed(b2n_mppi *h, uint64_t seed, uint32_t first_call);
/* perturbations for the NEXT call only, du[k][t][0..1] = (duL, duR), k local to this handle */
int b2n_mppi_set_noise(b2n_mppi *h, const double *du, size_t count);

/* Test / bench taps (the reference keeps these as private Eigen members, mppi.hpp:169-183) */
int b2n_mppi_set_capture(b2n_mppi *h, int on);                    /* keep J and du of each call */
int b2n_mppi_get_states(b2n_mppi *h, float *out, size_t count);   /* [K][T][3] x,y,theta of the last call */
int b2n_mppi_get_cost_to_go(b2n_mppi *h, double *out, size_t count);  /* [K][T], needs capture */
int b2n_mppi_get_noise(b2n_mppi *h, double *out, size_t count);   /* [K][T][2], needs capture */
int b2n_mppi_get_weights(b2n_mppi *h, double *out, size_t count); /* [K][T] normalised, needs capture */
int b2n_mppi_get_plan(b2n_mppi *h, double *out, size_t count);    /* [2][T] the control plan u */
int b2n_mppi_set_plan(b2n_mppi *h, const double *u, size_t count);
/* [T][6] per-step (min J, sum e, sum e*duL, sum e*duR, sum duL, sum duR) of this handle's rollouts */
int b2n_mppi_get_partials(b2n_mppi *h, double *out, size_t count);

/* Extension absent from the reference: running loss +=
 * weight * max(0, d0 - dist(x,y))^2 with dist looked up as grid_mapper.cpp:852-898; off-map states
 * pay off_map.  dist == NULL switches the term off. */
int b2n_mppi_set_obstacle_field(b2n_mppi *h, const float *dist, int xsize, int ysize, double xmin, double ymin,
                                double resolution, double weight, double d0, double off_map);

/* The same term fed on the device: allocates the field inside the handle and returns its DEVICE
 * address, for a producer that fills it without a host round trip (b2n_pf_write_distance_field below). */
int b2n_mppi_obstacle_field_device(b2n_mppi *h, int xsize, int ysize, double xmin, double ymin, double resolution, double weight,
                                   double d0, double off_map, float **device_field);

/* Launch on a caller-owned cudaStream_t (pass the pointer value); NULL restores the handle's own */
int b2n_mppi_set_stream(b2n_mppi *h, void *cuda_stream);
/* Write the K*T*3 state tensor round-robin into n buffers (n*K*T*12 bytes) so that back-to-back
 * calls stream through HBM instead of re-dirtying the same L2 lines; n = 1 is the default */
int b2n_mppi_set_state_ring(b2n_mppi *h, int n);
/* kernels launched by this handle since creation, and average device time (ms) of the rollout
 * kernel measured with CUDA events on the launching stream since the last call to this function */
int b2n_mppi_launch_count(const b2n_mppi *h, uint64_t *launches);
/* *fast = 1 when the last call ran the production (FAST) instantiation of the kernel, 0 for the generic one (capture taps,
 * caller-supplied noise, horizons that leave lanes partially filled): lets a parity test prove which one it checked */
int b2n_mppi_last_variant(const b2n_mppi *h, int *fast);
/* tuning hook (B2N_MPPI_DEBUG_TIMES=1 at create): globaltimer stamps [grid + T][24] of thread 0 of every CTA of the last call
 * (rollout CTAs, then the T merger CTAs); the slots are listed in tools/mppi_stages.py */
int b2n_mppi_debug_times(b2n_mppi *h, unsigned long long *out, size_t count, int *grid);
/* test hook: the Box-Muller stage of the perturbation generator alone, z[2 i], z[2 i + 1] for the first words
 * (first + i) << 9, i < count (the 2^23 values cover every radius the generator can produce) and one second word rb */
int b2n_test_box_muller(uint32_t first, uint32_t count, uint32_t rb, float *z);
int b2n_mppi_set_kernel_timing(b2n_mppi *h, int on);
int b2n_mppi_kernel_time(b2n_mppi *h, double *avg_ms, int *samples);
/* bench hook: the rollout kernel alone, `launches` times back to back on the handle's stream between two CUDA events
 * (same arguments as the last call would use, state tensor rotating through the ring; the plan is not updated);
 * avg_ms = elapsed / launches */
int b2n_mppi_time_rollout(b2n_mppi *h, double x, double y, double theta, int launches, double *avg_ms);

/* bench hook: `calls` synchronous b2n_mppi_new_controls() in a row from a C loop (host pose in, host controls out, every call
 * complete before the next starts: what a C++ node's control loop pays), wall-clock mean in ms; the last controls in ul, ur */
int b2n_mppi_time_new_controls(b2n_mppi *h, double x, double y, double theta, int calls, double *avg_ms, double *ul, double *ur);

/* Sharded operation: every rank simulates its slice of the rollouts, the [T][6]
 * partials are exchanged with ONE ncclAllGather, every rank applies the identical update.
 * unique_id is the 128-byte ncclUniqueId made by b2n_comm_unique_id() on one rank and distributed
 * by the caller (e.g. a torch.distributed broadcast). */
int b2n_comm_unique_id(void *out128);
int b2n_mppi_comm_init(b2n_mppi *h, int rank, int nranks, const void *unique_id128);
/* Same sharding, exchange over NVLink peer memory instead of NCCL (ranks = processes on ONE node, peer access between
 * their GPUs): merge of the CTA partials, push of the [T][6] result into every rank's exchange area, wait for the
 * others and the control update are ONE kernel (mppi_exchange_update_kernel).  Every rank calls
 * b2n_mppi_p2p_export (allocates its area, returns a 64-byte cudaIpcMemHandle_t), the caller gathers the handles of
 * all ranks in rank order (nranks x 64 bytes) and passes them to b2n_mppi_p2p_init.  Takes precedence over the NCCL
 * path once initialised. */
int b2n_mppi_p2p_export(b2n_mppi *h, int nranks, void *handle64);
int b2n_mppi_p2p_init(b2n_mppi *h, int rank, int nranks, const void *handles);
/* The same wiring for ranks that are handles of ONE process (several GPUs driven by one process, or - in the tests - two
 * ranks sharing one GPU on separate streams): b2n_mppi_p2p_export on every handle, then the device addresses of all
 * areas (b2n_mppi_p2p_area) in rank order.  The handles must outlive each other's use.  Ranks that SHARE a GPU must
 * collect every call (b2n_mppi_wait) before enqueuing the next one: the merger CTAs of a call spin while resident, waiting
 * for the other ranks' words, and a queued next call of one rank could take the SM slots another rank's mergers still need.
 * With one GPU per rank (the supported deployment) calls may be queued freely. */
int b2n_mppi_p2p_area(b2n_mppi *h, void **area);
int b2n_mppi_p2p_init_local(b2n_mppi *h, int rank, int nranks, void *const *areas);

/* ======================================================================================== RBPF */

typedef struct b2n_pf b2n_pf;

/* bmapping::LaserProperties (sensor_model.hpp:63-76), bmapping::GridMapper ctor (grid_mapper.hpp:121-122,
 * square maps as the reference assumes, grid_mapper.cpp:352-353) and bmapping::ParticleFilter ctor
 * (particle_filter.hpp:112-130).  Trs (robot->sensor) is the identity, as in turtle_mapping_node.cpp:397. */
typedef struct b2n_pf_params {
  /* lidar */
  float beam_min, beam_max, beam_delta, range_min, range_max;
  double z_hit, z_short, z_max, z_rand, sigma_hit;
  /* map */
  double resolution, xmin, xmax, ymin, ymax;
  /* filter */
  int32_t num_particles;             /* particles held by THIS handle */
  int32_t k;                         /* mode samples of the improved proposal */
  double srr, srt, str, stt;
  double motion_noise_theta, motion_noise_x, motion_noise_y;
  double sample_range_theta, sample_range_x, sample_range_y;
  double scan_likelihood_min, scan_likelihood_max, pose_likelihood_min, pose_likelihood_max;
  double init_pose[3];               /* theta, x, y  (particle_filter.cpp:133) */
  int32_t particle_offset;           /* global index of this handle's particle 0 */
  int32_t particles_total;           /* N of the whole job (0 or == num_particles when unsharded) */
  int32_t device;
  int32_t max_beams;                 /* capacity for scan length; 0 -> 1024 */
} b2n_pf_params;

int b2n_pf_create(const b2n_pf_params *params, b2n_pf **out);
void b2n_pf_destroy(b2n_pf *h);

/* bmapping::ParticleFilter::SLAM, particle_filter.cpp:141-251.
 * twist = (w, vx, vy) body twist of the scan interval, odometry poses = (theta, x, y).
 * The scan matcher (PCL ICP, cloud_alignment.cpp:37-72) stays on the caller's side of the boundary:
 * icp_ok == 0 selects the motion-model branch (particle_filter.cpp:161-176), icp_ok != 0 the
 * improved-proposal branch (:178-233) with icp_pose = (theta, x, y) of Ticp.
 * Returns B2N_ERR_OFF_MAP where the reference would throw because an end point left the map. */
int b2n_pf_slam(b2n_pf *h, const float *scan, int n_beams, const double twist[3], const double cur_odom[3],
                const double prev_odom[3], int icp_ok, const double icp_pose[3]);
/* bmapping::ParticleFilter::getRobotState, particle_filter.cpp:255-274 -> (theta, x, y).
 * On a SHARDED filter (b2n_pf_comm_init + b2n_pf_p2p_init) this, b2n_pf_new_map and b2n_pf_write_distance_field return the
 * argmax over the particles of ALL ranks, read from the owning GPU over peer memory; they are then COLLECTIVE: every rank
 * calls the same function between the same two SLAM() calls (two small all-reduces keep a rank from moving its particles
 * while another still reads them).  Without peer memory they return B2N_ERR_UNSUPPORTED on a sharded filter.
 * In the same way a sharded SLAM() fails on EVERY rank when any rank's particles or end points left the map. */
int b2n_pf_get_robot_state(b2n_pf *h, double pose[3]);
/* bmapping::ParticleFilter::newMap, particle_filter.cpp:277-291 + grid_mapper.cpp:185-226 */
int b2n_pf_new_map(b2n_pf *h, int8_t *out, size_t count);

int b2n_pf_seed(b2n_pf *h, uint64_t seed, uint32_t first_call);
/* standard normals for the NEXT call only, in the order the reference draws them per particle
 * (bmapping::sampleStandardNormal, particle_filter.cpp:25-36): motion-model branch z[n][3] (theta, x, y);
 * proposal branch z[n][3*(k+1)] (k mode samples, then the new pose); then ONE more value, the resampling
 * draw (used only if the filter resamples): count = 3*N + 1 or 3*(k+1)*N + 1 */
int b2n_pf_set_noise(b2n_pf *h, const double *z, size_t count);

/* Taps */
int b2n_pf_grid_size(const b2n_pf *h, int *xsize, int *ysize);
int b2n_pf_get_weights(b2n_pf *h, double *out, size_t count);
int b2n_pf_set_weights(b2n_pf *h, const double *w, size_t count);
int b2n_pf_get_poses(b2n_pf *h, double *poses, double *prev_poses, size_t count);   /* [N][3] theta,x,y */
int b2n_pf_set_poses(b2n_pf *h, const double *poses, size_t count);
/* outcome of the last SLAM(): N_eff as the reference prints it, whether it resampled, and the
 * ancestor of every slot (identity when it did not) */
int b2n_pf_get_resample(b2n_pf *h, int *neff, int *resampled, int32_t *ancestors, size_t count);
/* one particle's map: log-odds, distance to the obstacle cell that claimed the cell (grid_mapper.hpp Cell::occ_dist,
 * reconstructed exactly as sqrt(di^2 + dj^2) * resolution), class (-1 unknown / 0 free / 1 occupied) */
int b2n_pf_get_grid(b2n_pf *h, int particle, double *log_odds, double *occ_dist, int8_t *state, size_t count);
/* iteration order of the particle's occupied-cell set (std::unordered_set<int> occ_cells_, the seeds of the
 * distance transform, grid_mapper.cpp:348-361); writes up to cap keys, *n_occ = size of the set */
int b2n_pf_get_occ_order(b2n_pf *h, int particle, int32_t *keys, size_t cap, int *n_occ);
/* run only parts of SLAM(), for parity tests and kernel timing */
int b2n_pf_likelihoods(b2n_pf *h, const float *scan, int n_beams, double *out, size_t count);
/* normalizeWeights + effectiveParticles + lowVarianceResampling alone (particle_filter.cpp:244-249) */
int b2n_pf_normalize_resample(b2n_pf *h);
int b2n_pf_set_stream(b2n_pf *h, void *cuda_stream);
/* map geometry of the handle: origin and resolution (cells: b2n_pf_grid_size) */
int b2n_pf_geometry(const b2n_pf *h, double *xmin, double *ymin, double *resolution);
/* the best particle's distance field (GridMapper Cell::occ_dist, row-major i * xsize + j as grid_mapper.cpp:890-898)
 * as fp32 written to a DEVICE buffer of `count` floats on the same GPU - e.g. the one returned by
 * b2n_mppi_obstacle_field_device: the filter's map becomes the controller's obstacle term without touching the host.
 * Synchronous: the field is complete at return. */
int b2n_pf_write_distance_field(b2n_pf *h, float *device_out, size_t count);
int b2n_pf_launch_count(const b2n_pf *h, uint64_t *launches);
/* CUDA-event durations (ms) of the last SLAM(): [0] sample + weight + ray integration, [1] distance field,
 * [2] normalise + resample + particle copies */
int b2n_pf_set_kernel_timing(b2n_pf *h, int on);
int b2n_pf_kernel_times(b2n_pf *h, double ms[3]);
/* brushfire iterations summed over all particles and calls, largest heap seen */
int b2n_pf_distance_field_stats(b2n_pf *h, uint64_t *iterations, uint64_t *heap_max);
/* particle-scans (summed over calls) whose distance field was NOT regrown because their occupied set had not changed
 * since it was last grown - the field is a deterministic function of that set's iteration order */
int b2n_pf_distance_field_skipped(b2n_pf *h, uint64_t *particles);
/* heap entries kept in shared memory per particle in flight (the rest spills to global memory); default 2048 */
int b2n_pf_set_heap_capacity(b2n_pf *h, int entries);
/* what create() derives on the HOST from the constructor arguments, without needing a device (CPU tests):
 * constants = {t_occ, t_free, d_free, d_occ} - the log-odds thresholds equivalent to GridMapper::updateCellState's
 * prob >= 0.9 / prob <= 0.35 under this host's exp(), and the two log-odds steps (grid_mapper.cpp:42-47,438-477);
 * beam_cs[beam_count] = cos, sin of LaserScanner's accumulated beam angles (sensor_model.cpp:66-108);
 * pz = likelihood term z_hit * pdfNormal(sqrt(d2) * res, sigma_hit^2) + z_rand / z_max for every squared cell distance
 * d2 the distance field can hold, last entry = never-reached cell (grid_mapper.cpp:101-128) */
int b2n_pf_host_tables(const b2n_pf_params *params, double constants[4], double *beam_cs, size_t beam_count, double *pz,
                       size_t pz_cap, int *pz_n);
int b2n_pf_comm_init(b2n_pf *h, int rank, int nranks, const void *unique_id128);
/* Resampling across ranks.  Every rank runs the identical walk on the allgathered weights and so holds
 * the same ancestor vector; b2n_pf_slam then moves particles whose ancestor lives on another GPU with one group of
 * ncclSend/ncclRecv (each migrating particle once per destination rank) and copies the rest on the device.
 * b2n_pf_plan_migration is the pure host part, exposed for CPU tests: for `rank`, copy1[n_local] = local source index
 * in the old set or -1; copy2[n_local] = local slot of the new set to copy from after the exchange or -1;
 * recv = (local slot, global ancestor, source rank) triples; send = (local particle, destination rank) pairs, both in
 * the order the exchange posts them (ascending ancestor per peer). */
int b2n_pf_plan_migration(const int32_t *ancestors, int n_total, int rank, int nranks, int32_t *copy1, int32_t *copy2, int32_t *recv,
                          size_t recv_cap, int *n_recv, int32_t *send, size_t send_cap, int *n_send);
/* The same migration over NVLink peer memory instead of ncclSend/ncclRecv (ranks = processes on ONE node): every rank
 * exports the CUDA IPC handles of its plane allocations (b2n_pf_p2p_export: 10 x 64 bytes), the caller gathers them in
 * rank order (nranks x 640 bytes) and hands them to b2n_pf_p2p_init after b2n_pf_comm_init.  From then on a resampling
 * step is ONE copy kernel in which every slot reads its ancestor where it lives - local HBM or a peer's. */
int b2n_pf_p2p_export(b2n_pf *h, void *handles640);
int b2n_pf_p2p_init(b2n_pf *h, int rank, int nranks, const void *handles);
/* particles received from / sent to other ranks by the last SLAM() */
int b2n_pf_get_migration(const b2n_pf *h, int *received, int *sent);

/* ================================================================================ scan matcher */

/* SURVEY.md 8f row 2.  The reference takes its scan matcher from PCL (pcl::IterativeClosestPoint behind
 * bmapping::ScanAlignment, cloud_alignment.cpp:28-80,160-223), which is neither vendored nor pinned; this is libb2nav's
 * own point-to-point ICP with the reference's settings and wrapper semantics, parity-checked against
 * oracle/icp_oracle.cpp only.  Its (success, T) pair is what b2n_pf_slam takes as (icp_ok, icp_pose). */
typedef struct b2n_icp b2n_icp;

typedef struct b2n_icp_params {
  float beam_min, beam_max, beam_delta, range_min, range_max;   /* LaserProperties, sensor_model.hpp:63-76 */
  int32_t max_iter;                                             /* cloud_alignment.cpp:21: 100 */
  double max_correspondence_dist;                               /* :22: 0.5 */
  double transformation_epsilon;                                /* :23: 1e-8 */
  double euclidean_fitness_epsilon;                             /* :24: 1e-6 */
  int32_t device;
  int32_t max_beams;                                            /* capacity for scan length; 0 -> 1024 */
} b2n_icp_params;

int b2n_icp_create(const b2n_icp_params *params, b2n_icp **out);
void b2n_icp_destroy(b2n_icp *h);
/* bmapping::ScanAlignment::pclICPWrapper, cloud_alignment.cpp:37-72: the first call stores the scan and reports success
 * with t untouched; later calls align the new scan (source) onto the stored one (target) from the initial guess
 * t_init = (theta, x, y); on success t = (theta, x, y) of the transform and the new scan becomes the stored one. */
int b2n_icp_align(b2n_icp *h, const float *scan, int n_beams, const double t_init[3], double t[3], int *success);
/* iterations, correspondences and mean squared pair distance of the last alignment; kernel launches so far */
int b2n_icp_stats(const b2n_icp *h, int *iterations, int *pairs, double *mse, uint64_t *launches);

#ifdef __cplusplus
}
#endif
#endif /* B2NAV_H */
