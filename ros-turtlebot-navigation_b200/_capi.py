"""ctypes binding of include/b2nav.h.  Loading fails loudly when libb2nav.so is missing: there is
no Python or CPU implementation behind these classes."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libb2nav.so")

D = C.c_double
OK = 0
ERR_INVALID_ARGUMENT, ERR_CUDA, ERR_OFF_MAP, ERR_UNSUPPORTED, ERR_COMM, ERR_NUMERIC = -1, -2, -3, -4, -5, -6


class B2NError(RuntimeError):
    def __init__(self, code, text):
        super().__init__("libb2nav error %d: %s" % (code, text))
        self.code = code


class MppiParams(C.Structure):
    _fields_ = [
        ("wheel_radius", D), ("wheel_base", D),
        ("Q", D * 3), ("R", D * 2), ("P1", D * 3),
        ("lambda_", D), ("max_wheel_vel", D), ("ul_var", D), ("ur_var", D), ("horizon", D), ("dt", D),
        ("rollouts", C.c_int32), ("rollout_offset", C.c_int32), ("rollouts_total", C.c_int32), ("device", C.c_int32),
    ]


class PfParams(C.Structure):
    _fields_ = [
        ("beam_min", C.c_float), ("beam_max", C.c_float), ("beam_delta", C.c_float),
        ("range_min", C.c_float), ("range_max", C.c_float),
        ("z_hit", D), ("z_short", D), ("z_max", D), ("z_rand", D), ("sigma_hit", D),
        ("resolution", D), ("xmin", D), ("xmax", D), ("ymin", D), ("ymax", D),
        ("num_particles", C.c_int32), ("k", C.c_int32),
        ("srr", D), ("srt", D), ("str_", D), ("stt", D),
        ("motion_noise_theta", D), ("motion_noise_x", D), ("motion_noise_y", D),
        ("sample_range_theta", D), ("sample_range_x", D), ("sample_range_y", D),
        ("scan_likelihood_min", D), ("scan_likelihood_max", D),
        ("pose_likelihood_min", D), ("pose_likelihood_max", D),
        ("init_pose", D * 3),
        ("particle_offset", C.c_int32), ("particles_total", C.c_int32), ("device", C.c_int32), ("max_beams", C.c_int32),
    ]


_P = C.POINTER
_vp = C.c_void_p
_sz = C.c_size_t

# name -> (restype, argtypes); every symbol include/b2nav.h declares
PROTOTYPES = {
    "b2n_last_error": (C.c_char_p, []),
    "b2n_device_count": (C.c_int, []),
    "b2n_version": (C.c_int, [C.c_char_p, _sz]),
    "b2n_mppi_create": (C.c_int, [_P(MppiParams), _P(_vp)]),
    "b2n_mppi_destroy": (None, [_vp]),
    "b2n_mppi_steps": (C.c_int, [_vp]),
    "b2n_mppi_set_initial_controls": (C.c_int, [_vp, D, D]),
    "b2n_mppi_set_waypoint": (C.c_int, [_vp, D, D, D]),
    "b2n_mppi_new_controls": (C.c_int, [_vp, D, D, D, _P(D), _P(D)]),
    "b2n_mppi_enqueue": (C.c_int, [_vp, D, D, D]),
    "b2n_mppi_enqueue_many": (C.c_int, [_vp, D, D, D, C.c_int]),
    "b2n_mppi_wait": (C.c_int, [_vp, _P(D), _P(D)]),
    "b2n_mppi_seed": (C.c_int, [_vp, C.c_uint64, C.c_uint32]),
    "b2n_mppi_set_noise": (C.c_int, [_vp, _vp, _sz]),
    "b2n_mppi_set_capture": (C.c_int, [_vp, C.c_int]),
    "b2n_mppi_get_states": (C.c_int, [_vp, _vp, _sz]),
    "b2n_mppi_get_cost_to_go": (C.c_int, [_vp, _vp, _sz]),
    "b2n_mppi_get_noise": (C.c_int, [_vp, _vp, _sz]),
    "b2n_mppi_get_weights": (C.c_int, [_vp, _vp, _sz]),
    "b2n_mppi_get_plan": (C.c_int, [_vp, _vp, _sz]),
    "b2n_mppi_set_plan": (C.c_int, [_vp, _vp, _sz]),
    "b2n_mppi_get_partials": (C.c_int, [_vp, _vp, _sz]),
    "b2n_mppi_set_obstacle_field": (C.c_int, [_vp, _vp, C.c_int, C.c_int, D, D, D, D, D, D]),
    "b2n_mppi_obstacle_field_device": (C.c_int, [_vp, C.c_int, C.c_int, D, D, D, D, D, D, _P(_vp)]),
    "b2n_mppi_set_stream": (C.c_int, [_vp, _vp]),
    "b2n_mppi_set_state_ring": (C.c_int, [_vp, C.c_int]),
    "b2n_mppi_launch_count": (C.c_int, [_vp, _P(C.c_uint64)]),
    "b2n_mppi_last_variant": (C.c_int, [_vp, _P(C.c_int)]),
    "b2n_mppi_debug_times": (C.c_int, [_vp, _vp, _sz, _P(C.c_int)]),
    "b2n_test_box_muller": (C.c_int, [C.c_uint32, C.c_uint32, C.c_uint32, _vp]),
    "b2n_mppi_set_kernel_timing": (C.c_int, [_vp, C.c_int]),
    "b2n_mppi_kernel_time": (C.c_int, [_vp, _P(D), _P(C.c_int)]),
    "b2n_mppi_time_rollout": (C.c_int, [_vp, D, D, D, C.c_int, _P(D)]),
    "b2n_mppi_time_new_controls": (C.c_int, [_vp, D, D, D, C.c_int, _P(D), _P(D), _P(D)]),
    "b2n_comm_unique_id": (C.c_int, [_vp]),
    "b2n_mppi_comm_init": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "b2n_mppi_p2p_export": (C.c_int, [_vp, C.c_int, _vp]),
    "b2n_mppi_p2p_init": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "b2n_mppi_p2p_area": (C.c_int, [_vp, _P(_vp)]),
    "b2n_mppi_p2p_init_local": (C.c_int, [_vp, C.c_int, C.c_int, _P(_vp)]),
    "b2n_pf_create": (C.c_int, [_P(PfParams), _P(_vp)]),
    "b2n_pf_destroy": (None, [_vp]),
    "b2n_pf_slam": (C.c_int, [_vp, _vp, C.c_int, _P(D), _P(D), _P(D), C.c_int, _P(D)]),
    "b2n_pf_get_robot_state": (C.c_int, [_vp, _P(D)]),
    "b2n_pf_new_map": (C.c_int, [_vp, _vp, _sz]),
    "b2n_pf_seed": (C.c_int, [_vp, C.c_uint64, C.c_uint32]),
    "b2n_pf_set_noise": (C.c_int, [_vp, _vp, _sz]),
    "b2n_pf_grid_size": (C.c_int, [_vp, _P(C.c_int), _P(C.c_int)]),
    "b2n_pf_get_weights": (C.c_int, [_vp, _vp, _sz]),
    "b2n_pf_set_weights": (C.c_int, [_vp, _vp, _sz]),
    "b2n_pf_get_poses": (C.c_int, [_vp, _vp, _vp, _sz]),
    "b2n_pf_set_poses": (C.c_int, [_vp, _vp, _sz]),
    "b2n_pf_get_resample": (C.c_int, [_vp, _P(C.c_int), _P(C.c_int), _vp, _sz]),
    "b2n_pf_get_grid": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _sz]),
    "b2n_pf_get_occ_order": (C.c_int, [_vp, C.c_int, _vp, _sz, _P(C.c_int)]),
    "b2n_pf_likelihoods": (C.c_int, [_vp, _vp, C.c_int, _vp, _sz]),
    "b2n_pf_normalize_resample": (C.c_int, [_vp]),
    "b2n_pf_set_stream": (C.c_int, [_vp, _vp]),
    "b2n_pf_geometry": (C.c_int, [_vp, _P(D), _P(D), _P(D)]),
    "b2n_pf_write_distance_field": (C.c_int, [_vp, _vp, _sz]),
    "b2n_pf_launch_count": (C.c_int, [_vp, _P(C.c_uint64)]),
    "b2n_pf_set_kernel_timing": (C.c_int, [_vp, C.c_int]),
    "b2n_pf_kernel_times": (C.c_int, [_vp, _P(D)]),
    "b2n_pf_distance_field_stats": (C.c_int, [_vp, _P(C.c_uint64), _P(C.c_uint64)]),
    "b2n_pf_distance_field_skipped": (C.c_int, [_vp, _P(C.c_uint64)]),
    "b2n_pf_set_heap_capacity": (C.c_int, [_vp, C.c_int]),
    "b2n_pf_host_tables": (C.c_int, [_P(PfParams), _P(D), _vp, _sz, _vp, _sz, _P(C.c_int)]),
    "b2n_pf_comm_init": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "b2n_icp_create": (C.c_int, [_vp, _P(_vp)]),
    "b2n_icp_destroy": (None, [_vp]),
    "b2n_icp_align": (C.c_int, [_vp, _vp, C.c_int, _P(D), _P(D), _P(C.c_int)]),
    "b2n_icp_stats": (C.c_int, [_vp, _P(C.c_int), _P(C.c_int), _P(D), _P(C.c_uint64)]),
    "b2n_pf_p2p_export": (C.c_int, [_vp, _vp]),
    "b2n_pf_p2p_init": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "b2n_pf_plan_migration": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _sz, _P(C.c_int), _vp, _sz, _P(C.c_int)]),
    "b2n_pf_get_migration": (C.c_int, [_vp, _P(C.c_int), _P(C.c_int)]),
}

_lib = None


class IcpParams(C.Structure):
    _fields_ = [("beam_min", C.c_float), ("beam_max", C.c_float), ("beam_delta", C.c_float), ("range_min", C.c_float),
                ("range_max", C.c_float), ("max_iter", C.c_int32), ("max_correspondence_dist", D), ("transformation_epsilon", D),
                ("euclidean_fitness_epsilon", D), ("device", C.c_int32), ("max_beams", C.c_int32)]


def load_library():
    """dlopen libb2nav.so (built by build.py / __graft_entry__.build()) and bind every prototype."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: run `python __graft_entry__.py build` (there is no fallback path)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != OK:
        raise B2NError(rc, load_library().b2n_last_error().decode("utf-8", "replace"))


def as_ptr(arr):
    """void* of a C-contiguous numpy array."""
    return arr.ctypes.data_as(C.c_void_p)
