"""ctypes access to the two CPU checkers.  TEST INFRASTRUCTURE ONLY.

  OracleMppi  -> oracle/liboracle_nav.so   (our Eigen-free restatement)
  RefMppi     -> oracle/_ref/libref_nav.so (the unmodified reference sources, compiled in place)
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle_nav.so")
REF_SO = os.environ.get("B2N_REF_SO", os.path.join(ROOT, "oracle", "_ref", "libref_nav.so"))
D = C.c_double
nd = np.ctypeslib.ndpointer

# synthetic inputs and shipped parameters live with the package (plain numpy, shared with bench.py and tools/)
import importlib.util as _ilu  # noqa: E402
_spec = _ilu.spec_from_file_location("b2n_synthetic", os.path.join(ROOT, "ros-turtlebot-navigation_b200", "synthetic.py"))
_syn = _ilu.module_from_spec(_spec)
_spec.loader.exec_module(_syn)
SHIPPED, MILD, PF_SHIPPED = _syn.SHIPPED, _syn.MILD, _syn.PF_SHIPPED
pf_params, room_scan, circle_path, unicycle_step = _syn.pf_params, _syn.room_scan, _syn.circle_path, _syn.unicycle_step

class _P(C.Structure):
    _fields_ = [("wheel_radius", D), ("wheel_base", D), ("Q", D * 3), ("R", D * 2), ("P1", D * 3),
                ("lambda_", D), ("max_wheel_vel", D), ("ul_var", D), ("ur_var", D), ("horizon", D), ("dt", D),
                ("rollouts", C.c_int)]


_olib = None
_rlib = None


def have_ref():
    return os.path.exists(REF_SO)


def _preload_cxx_runtime():
    """libref_nav.so prints through std::cout inside the reference's hot path; when libstdc++ first entered the
    process RTLD_LOCAL (as a dependency of numpy) those iostream objects resolve to a second, uninitialised copy and
    the first `std::cout <<` crashes.  Making the runtime global before loading the checkers avoids that."""
    try:
        C.CDLL("libstdc++.so.6", mode=C.RTLD_GLOBAL)
    except OSError:
        pass


def oracle_lib():
    global _olib
    if _olib is None:
        _preload_cxx_runtime()
        L = C.CDLL(ORACLE_SO)
        L.orc_mppi_create.restype = C.c_void_p
        L.orc_mppi_create.argtypes = [C.POINTER(_P)]
        L.orc_mppi_destroy.argtypes = [C.c_void_p]
        L.orc_mppi_steps.argtypes = [C.c_void_p]
        L.orc_mppi_set_initial_controls.argtypes = [C.c_void_p, D, D]
        L.orc_mppi_set_waypoint.argtypes = [C.c_void_p, D, D, D]
        L.orc_mppi_noise_mt19937.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_mppi_noise_philox.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32]
        L.orc_mppi_noise_external.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_mppi_set_shard.argtypes = [C.c_void_p, C.c_int]
        L.orc_mppi_set_plan.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_mppi_set_obstacles.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, D, D, D, D, D, D]
        L.orc_mppi_new_controls.argtypes = [C.c_void_p, D, D, D, C.POINTER(D), C.POINTER(D)]
        L.orc_mppi_get.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.orc_mt_normals.argtypes = [C.c_uint64, C.c_int, D, D, nd(np.float64)]
        L.orc_philox_raw.argtypes = [nd(np.uint32), nd(np.uint32), nd(np.uint32)]
        L.orc_philox_normal_pair.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, nd(np.float64)]
        _olib = L
    return _olib


def ref_lib():
    global _rlib
    if _rlib is None:
        _preload_cxx_runtime()
        L = C.CDLL(REF_SO)
        L.ref_rigid2d_seed.argtypes = [C.c_uint64]
        L.ref_bmapping_seed.argtypes = [C.c_uint64]
        L.ref_rigid2d_normals.argtypes = [C.c_int, D, D, nd(np.float64)]
        L.ref_bmapping_std_normals.argtypes = [C.c_int, nd(np.float64)]
        L.ref_normalize_angle_pi.restype = D
        L.ref_normalize_angle_pi.argtypes = [D]
        L.ref_wheels_to_twist.argtypes = [D, D, D, D, nd(np.float64)]
        L.ref_twist_to_wheels.argtypes = [D, D, D, D, nd(np.float64)]
        L.ref_integrate_twist.argtypes = [nd(np.float64)] * 3
        L.ref_feedforward.argtypes = [D, D, nd(np.float64), nd(np.float64), nd(np.float64)]
        L.ref_dd_create.restype = C.c_void_p
        L.ref_dd_create.argtypes = [nd(np.float64), D, D]
        L.ref_dd_destroy.argtypes = [C.c_void_p]
        L.ref_dd_feedforward.argtypes = [C.c_void_p, D, D]
        L.ref_dd_update_odometry.argtypes = [C.c_void_p, D, D, nd(np.float64)]
        L.ref_dd_state.argtypes = [C.c_void_p, nd(np.float64)]
        L.ref_mppi_create.restype = C.c_void_p
        L.ref_mppi_create.argtypes = [D, D, C.POINTER(D), C.POINTER(D), C.POINTER(D), D, D, D, D, D, D, C.c_int]
        L.ref_mppi_destroy.argtypes = [C.c_void_p]
        L.ref_mppi_steps.argtypes = [C.c_void_p]
        L.ref_mppi_set_initial_controls.argtypes = [C.c_void_p, D, D]
        L.ref_mppi_set_waypoint.argtypes = [C.c_void_p, D, D, D]
        L.ref_mppi_new_controls.argtypes = [C.c_void_p, D, D, D, C.POINTER(D), C.POINTER(D)]
        L.ref_mppi_get.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.ref_mppi_rollout.argtypes = [C.c_void_p, nd(np.float64), nd(np.float64), nd(np.float64), nd(np.float64)]
        _rlib = L
    return _rlib


class RefDiffDrive:
    """The compiled reference rigid2d::DiffDrive (oracle/_ref), kept alive across calls: the plant and the odometer of the
    closed-loop tests.  Same method names and (theta, x, y) / tuple conventions as synthetic.DiffDrive."""

    def __init__(self, pose=(0.0, 0.0, 0.0), wheel_base=0.16, wheel_radius=0.033):
        self.L = ref_lib()
        self.base, self.radius = float(wheel_base), float(wheel_radius)
        self.h = self.L.ref_dd_create(np.asarray(pose, dtype=np.float64), self.base, self.radius)

    def _state(self):
        out = np.zeros(7)
        self.L.ref_dd_state(self.h, out)
        return out

    def feedforward(self, w, vx, vy=0.0):
        self.L.ref_dd_feedforward(self.h, float(w), float(vx))

    def updateOdometry(self, left, right):
        v = np.zeros(2)
        self.L.ref_dd_update_odometry(self.h, float(left), float(right), v)
        return float(v[0]), float(v[1])

    def wheelsToTwist(self, ul, ur):
        out = np.zeros(3)
        self.L.ref_wheels_to_twist(self.base, self.radius, float(ul), float(ur), out)
        return float(out[0]), float(out[1]), float(out[2])

    def pose(self):
        return tuple(float(v) for v in self._state()[0:3])

    def getEncoders(self):
        return tuple(float(v) for v in self._state()[3:5])

    def wheelVelocities(self):
        return tuple(float(v) for v in self._state()[5:7])

    def __del__(self):
        try:
            self.L.ref_dd_destroy(self.h)
        except Exception:
            pass


def _vp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleMppi:
    def __init__(self, horizon, dt, rollouts, **kw):
        self.L = oracle_lib()
        q = dict(SHIPPED)
        q.update(kw)
        p = _P(q["wheel_radius"], q["wheel_base"], (D * 3)(*q["Q"]), (D * 2)(*q["R"]), (D * 3)(*q["P1"]), q["lambda_"],
               q["max_wheel_vel"], q["ul_var"], q["ur_var"], horizon, dt, rollouts)
        self.h = C.c_void_p(self.L.orc_mppi_create(C.byref(p)))
        self.K = rollouts
        self.T = self.L.orc_mppi_steps(self.h)
        self._ext = None

    def setInitialControls(self, ul, ur):
        self.L.orc_mppi_set_initial_controls(self.h, ul, ur)

    def setWaypoint(self, x, y, theta):
        self.L.orc_mppi_set_waypoint(self.h, x, y, theta)

    def noise_mt19937(self, seed):
        self.L.orc_mppi_noise_mt19937(self.h, seed)

    def noise_philox(self, seed, first_call=0):
        self.L.orc_mppi_noise_philox(self.h, seed, first_call)

    def noise_external(self, du):
        self._ext = np.ascontiguousarray(du, dtype=np.float64)
        assert self._ext.shape == (self.K, self.T, 2)
        self.L.orc_mppi_noise_external(self.h, _vp(self._ext))

    def set_shard(self, k_offset):
        self.L.orc_mppi_set_shard(self.h, k_offset)

    def set_plan(self, u):
        u = np.ascontiguousarray(u, dtype=np.float64)
        self.L.orc_mppi_set_plan(self.h, _vp(u))

    def set_obstacles(self, dist, xmin, ymin, res, weight, d0, off_map):
        dist = np.ascontiguousarray(dist, dtype=np.float32)
        self.L.orc_mppi_set_obstacles(self.h, _vp(dist), dist.shape[0], dist.shape[1], xmin, ymin, res, weight, d0, off_map)

    def newControls(self, x, y, theta):
        ul, ur = D(), D()
        self.L.orc_mppi_new_controls(self.h, x, y, theta, C.byref(ul), C.byref(ur))
        return ul.value, ur.value

    def get(self):
        K, T = self.K, self.T
        out = dict(plan=np.zeros((2, T)), du=np.zeros((K, T, 2)), states=np.zeros((K, T, 3)), J=np.zeros((T, K)),
                   w=np.zeros((T, K)))
        self.L.orc_mppi_get(self.h, _vp(out["plan"]), _vp(out["du"]), _vp(out["states"]), _vp(out["J"]), _vp(out["w"]))
        return out

    def __del__(self):
        try:
            self.L.orc_mppi_destroy(self.h)
        except Exception:
            pass


class RefMppi:
    """The compiled reference controller::MPPI.  Its RNG is the process-global rigid2d engine."""

    def __init__(self, horizon, dt, rollouts, **kw):
        self.L = ref_lib()
        q = dict(SHIPPED)
        q.update(kw)
        self.h = C.c_void_p(self.L.ref_mppi_create(q["wheel_radius"], q["wheel_base"], (D * 3)(*q["Q"]), (D * 2)(*q["R"]),
                                                   (D * 3)(*q["P1"]), q["lambda_"], q["max_wheel_vel"], q["ul_var"],
                                                   q["ur_var"], horizon, dt, rollouts))
        self.K = rollouts
        self.T = self.L.ref_mppi_steps(self.h)

    def seed(self, s):
        self.L.ref_rigid2d_seed(s)

    def setInitialControls(self, ul, ur):
        self.L.ref_mppi_set_initial_controls(self.h, ul, ur)

    def setWaypoint(self, x, y, theta):
        self.L.ref_mppi_set_waypoint(self.h, x, y, theta)

    def newControls(self, x, y, theta):
        ul, ur = D(), D()
        self.L.ref_mppi_new_controls(self.h, x, y, theta, C.byref(ul), C.byref(ur))
        return ul.value, ur.value

    def get(self):
        K, T = self.K, self.T
        out = dict(plan=np.zeros((2, T)), Jsub=np.zeros((T, K)), duL=np.zeros((T, K)), duR=np.zeros((T, K)))
        self.L.ref_mppi_get(self.h, _vp(out["plan"]), _vp(out["Jsub"]), _vp(out["duL"]), _vp(out["duR"]))
        return out

    def rollout(self, x0, u_pert):
        T = self.T
        traj, loss = np.zeros((T, 3)), np.zeros(T)
        self.L.ref_mppi_rollout(self.h, np.ascontiguousarray(x0, dtype=np.float64),
                                np.ascontiguousarray(u_pert, dtype=np.float64), traj, loss)
        return traj, loss

    def __del__(self):
        try:
            self.L.ref_mppi_destroy(self.h)
        except Exception:
            pass


# =========================================================================================== RBPF
class _OPF(C.Structure):
    _fields_ = [("beam_min", C.c_float), ("beam_max", C.c_float), ("beam_delta", C.c_float), ("range_min", C.c_float),
                ("range_max", C.c_float), ("z_hit", D), ("z_short", D), ("z_max", D), ("z_rand", D), ("sigma_hit", D),
                ("resolution", D), ("xmin", D), ("xmax", D), ("ymin", D), ("ymax", D),
                ("num_particles", C.c_int32), ("k", C.c_int32), ("srr", D), ("srt", D), ("str_", D), ("stt", D),
                ("motion_noise", D * 3), ("sample_range", D * 3), ("scan_min", D), ("scan_max", D), ("pose_min", D),
                ("pose_max", D), ("init_pose", D * 3)]


_opf_bound = False
_rpf_bound = False
F32P = nd(np.float32)
F64P = nd(np.float64)
I32P = nd(np.int32)


def _bind_oracle_pf():
    global _opf_bound
    L = oracle_lib()
    if _opf_bound:
        return L
    L.orc_pf_create.restype = C.c_void_p
    L.orc_pf_create.argtypes = [C.POINTER(_OPF)]
    L.orc_pf_destroy.argtypes = [C.c_void_p]
    L.orc_pf_noise_mt19937.argtypes = [C.c_void_p, C.c_uint64]
    L.orc_pf_noise_philox.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32]
    L.orc_pf_noise_external.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.orc_pf_set_shard.argtypes = [C.c_void_p, C.c_int]
    L.orc_pf_grid_size.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.orc_pf_slam.argtypes = [C.c_void_p, F32P, C.c_int, F64P, F64P, F64P, C.c_int, F64P]
    L.orc_pf_get.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_pf_set_weights.argtypes = [C.c_void_p, F64P]
    L.orc_pf_set_poses.argtypes = [C.c_void_p, F64P]
    L.orc_pf_get_resample.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p]
    L.orc_pf_normalize_resample.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_void_p]
    L.orc_pf_robot_state.argtypes = [C.c_void_p, F64P]
    L.orc_pf_new_map.argtypes = [C.c_void_p, nd(np.int8)]
    L.orc_pf_grid_dump.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4
    L.orc_pf_grid_occ_order.argtypes = [C.c_void_p, C.c_int, I32P, C.c_int]
    L.orc_pf_grid_bucket_count.argtypes = [C.c_void_p, C.c_int]
    L.orc_pf_grid_likelihood.argtypes = [C.c_void_p, C.c_int, F32P, C.c_int, F64P, C.POINTER(D)]
    L.orc_pf_grid_integrate.argtypes = [C.c_void_p, C.c_int, F32P, C.c_int, F64P]
    L.orc_pf_grid_end_points.argtypes = [C.c_void_p, C.c_int, F32P, C.c_int, F64P, F64P]
    L.orc_pf_grid_free_cells.argtypes = [C.c_void_p, C.c_int, F64P, F64P, I32P, C.c_int]
    L.orc_pf_grid_map.argtypes = [C.c_void_p, C.c_int, nd(np.int8)]
    L.orc_pf_grid_stats.argtypes = [C.c_void_p, C.c_int, nd(np.uint64)]
    L.orc_selftest_occset.argtypes = [C.c_uint64, C.c_int, C.c_int]
    L.orc_selftest_heap.argtypes = [C.c_uint64, C.c_int, C.c_int]
    _opf_bound = True
    return L


def _bind_ref_pf():
    global _rpf_bound
    L = ref_lib()
    if _rpf_bound:
        return L
    F5, D5 = C.c_float * 5, D * 5
    L.ref_grid_create.restype = C.c_void_p
    L.ref_grid_create.argtypes = [F5, D5, D, D, D, D, D]
    L.ref_grid_destroy.argtypes = [C.c_void_p]
    L.ref_grid_clone.restype = C.c_void_p
    L.ref_grid_clone.argtypes = [C.c_void_p]
    L.ref_grid_size.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.ref_grid_likelihood.argtypes = [C.c_void_p, F32P, C.c_int, F64P, C.POINTER(D)]
    L.ref_grid_integrate.argtypes = [C.c_void_p, F32P, C.c_int, F64P]
    L.ref_grid_dump.argtypes = [C.c_void_p] + [C.c_void_p] * 4
    L.ref_grid_occ_order.argtypes = [C.c_void_p, I32P, C.c_int]
    L.ref_grid_bucket_count.argtypes = [C.c_void_p]
    L.ref_grid_map.argtypes = [C.c_void_p, nd(np.int8)]
    L.ref_grid_end_points.argtypes = [C.c_void_p, F32P, C.c_int, F64P, F64P]
    L.ref_grid_free_cells.argtypes = [C.c_void_p, F64P, F64P, I32P, C.c_int]
    L.ref_pf_create.restype = C.c_void_p
    L.ref_pf_create.argtypes = [C.c_int, C.c_int, D * 14, F5, D5, D, D, D, D, D, D * 3]
    L.ref_pf_destroy.argtypes = [C.c_void_p]
    L.ref_pf_set_icp.argtypes = [C.c_int, F64P]
    L.ref_pf_slam.argtypes = [C.c_void_p, F32P, C.c_int, F64P, F64P, F64P, C.POINTER(C.c_int)]
    L.ref_pf_num.argtypes = [C.c_void_p]
    L.ref_pf_get.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ref_pf_set_weights.argtypes = [C.c_void_p, F64P]
    L.ref_pf_grid_dump.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4
    L.ref_pf_grid_occ_order.argtypes = [C.c_void_p, C.c_int, I32P, C.c_int]
    L.ref_pf_robot_state.argtypes = [C.c_void_p, F64P]
    L.ref_pf_new_map.argtypes = [C.c_void_p, nd(np.int8)]
    L.ref_pf_normalize_resample.argtypes = [C.c_void_p, C.POINTER(C.c_int), I32P]
    _rpf_bound = True
    return L


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class _PfCommon:
    """Shared convenience over either CPU particle filter (same method names on both)."""

    def grid(self, particle=0):
        G = self.G
        out = dict(log_odds=np.zeros(G), prob=np.zeros(G), occ_dist=np.zeros(G), state=np.zeros(G, dtype=np.int32))
        self._grid_dump(particle, out)
        return out

    def state(self):
        N = self.N
        w, p, pp = np.zeros(N), np.zeros((N, 3)), np.zeros((N, 3))
        self._get(w, p, pp)
        return dict(weights=w, poses=p, prev_poses=pp)


class OraclePf(_PfCommon):
    """oracle/liboracle_nav.so restatement of bmapping::ParticleFilter (+ per-particle GridMapper)."""

    def __init__(self, **kw):
        self.L = _bind_oracle_pf()
        q = pf_params(**kw)
        self.q = q
        p = _OPF(q["beam_min"], q["beam_max"], q["beam_delta"], q["range_min"], q["range_max"], q["z_hit"], q["z_short"],
                 q["z_max"], q["z_rand"], q["sigma_hit"], q["resolution"], q["xmin"], q["xmax"], q["ymin"], q["ymax"],
                 q["num_particles"], q["k"], q["srr"], q["srt"], q["str_"], q["stt"], (D * 3)(*q["motion_noise"]),
                 (D * 3)(*q["sample_range"]), q["scan_min"], q["scan_max"], q["pose_min"], q["pose_max"],
                 (D * 3)(*q["init_pose"]))
        self.h = C.c_void_p(self.L.orc_pf_create(C.byref(p)))
        self.N = q["num_particles"]
        xs, ys = C.c_int(), C.c_int()
        self.G = self.L.orc_pf_grid_size(self.h, C.byref(xs), C.byref(ys))
        self.xsize, self.ysize = xs.value, ys.value
        self._ext = None

    def noise_mt19937(self, seed):
        self.L.orc_pf_noise_mt19937(self.h, seed)

    def noise_philox(self, seed, first_call=0):
        self.L.orc_pf_noise_philox(self.h, seed, first_call)

    def noise_external(self, z, per_particle):
        self._ext = _f64(z)
        assert self._ext.size == self.N * per_particle + 1
        self.L.orc_pf_noise_external(self.h, _vp(self._ext), per_particle)

    def set_shard(self, offset):
        self.L.orc_pf_set_shard(self.h, offset)

    def slam(self, scan, twist, cur_odom, prev_odom, icp_ok=0, icp_pose=(0.0, 0.0, 0.0)):
        scan = _f32(scan)
        return self.L.orc_pf_slam(self.h, scan, scan.size, _f64(twist), _f64(cur_odom), _f64(prev_odom), int(icp_ok), _f64(icp_pose))

    def _get(self, w, p, pp):
        self.L.orc_pf_get(self.h, _vp(w), _vp(p), _vp(pp))

    def set_weights(self, w):
        self.L.orc_pf_set_weights(self.h, _f64(w))

    def set_poses(self, p):
        self.L.orc_pf_set_poses(self.h, _f64(p))

    def resample_info(self):
        neff, rs = C.c_int(), C.c_int()
        anc = np.zeros(self.N, dtype=np.int32)
        self.L.orc_pf_get_resample(self.h, C.byref(neff), C.byref(rs), _vp(anc))
        return neff.value, rs.value, anc

    def normalize_resample(self):
        rs = C.c_int()
        anc = np.zeros(self.N, dtype=np.int32)
        self.L.orc_pf_normalize_resample(self.h, C.byref(rs), _vp(anc))
        return rs.value, anc

    def robot_state(self):
        out = np.zeros(3)
        self.L.orc_pf_robot_state(self.h, out)
        return out

    def new_map(self):
        out = np.zeros(self.G, dtype=np.int8)
        self.L.orc_pf_new_map(self.h, out)
        return out

    def _grid_dump(self, particle, out):
        self.L.orc_pf_grid_dump(self.h, particle, _vp(out["log_odds"]), _vp(out["prob"]), _vp(out["occ_dist"]), _vp(out["state"]))

    def occ_order(self, particle=0):
        keys = np.zeros(self.G, dtype=np.int32)
        n = self.L.orc_pf_grid_occ_order(self.h, particle, keys, self.G)
        return keys[:n].copy()

    def bucket_count(self, particle=0):
        return self.L.orc_pf_grid_bucket_count(self.h, particle)

    def grid_likelihood(self, scan, pose, particle=0):
        scan = _f32(scan)
        p = D()
        rc = self.L.orc_pf_grid_likelihood(self.h, particle, scan, scan.size, _f64(pose), C.byref(p))
        return rc, p.value

    def grid_integrate(self, scan, pose, particle=0):
        scan = _f32(scan)
        return self.L.orc_pf_grid_integrate(self.h, particle, scan, scan.size, _f64(pose))

    def grid_end_points(self, scan, pose, particle=0):
        scan = _f32(scan)
        xy = np.zeros((scan.size, 2))
        n = self.L.orc_pf_grid_end_points(self.h, particle, scan, scan.size, _f64(pose), xy)
        return xy[:n].copy()

    def grid_free_cells(self, pt, pose, particle=0):
        cells = np.zeros(4 * (self.xsize + self.ysize), dtype=np.int32)
        n = self.L.orc_pf_grid_free_cells(self.h, particle, _f64(pt), _f64(pose), cells, cells.size)
        return None if n < 0 else cells[:n].copy()

    def grid_map(self, particle=0):
        out = np.zeros(self.G, dtype=np.int8)
        self.L.orc_pf_grid_map(self.h, particle, out)
        return out

    def grid_stats(self, particle=0):
        out = np.zeros(4, dtype=np.uint64)
        self.L.orc_pf_grid_stats(self.h, particle, out)
        return dict(ray_cells=int(out[0]), esdf_iterations=int(out[1]), esdf_pushes=int(out[2]), heap_max=int(out[3]))

    def __del__(self):
        try:
            self.L.orc_pf_destroy(self.h)
        except Exception:
            pass


def _laser_arrays(q):
    lf = (C.c_float * 5)(q["beam_min"], q["beam_max"], q["beam_delta"], q["range_min"], q["range_max"])
    ld = (D * 5)(q["z_hit"], q["z_short"], q["z_max"], q["z_rand"], q["sigma_hit"])
    return lf, ld


class RefPf(_PfCommon):
    """The compiled reference bmapping::ParticleFilter.  RNG = the process-global bmapping engine; the ICP
    outcome is injected (process-global too), see oracle/ref_capi.cpp."""

    def __init__(self, **kw):
        self.L = _bind_ref_pf()
        q = pf_params(**kw)
        self.q = q
        lf, ld = _laser_arrays(q)
        pfp = (D * 14)(q["srr"], q["srt"], q["str_"], q["stt"], *q["motion_noise"], *q["sample_range"], q["scan_min"],
                       q["scan_max"], q["pose_min"], q["pose_max"])
        self.h = C.c_void_p(self.L.ref_pf_create(q["num_particles"], q["k"], pfp, lf, ld, q["resolution"], q["xmin"],
                                                 q["xmax"], q["ymin"], q["ymax"], (D * 3)(*q["init_pose"])))
        self.N = q["num_particles"]
        self.xsize = self.ysize = int(np.ceil((q["xmax"] - q["xmin"]) / q["resolution"]))
        self.G = self.xsize * self.ysize

    def seed(self, s):
        self.L.ref_bmapping_seed(s)

    def slam(self, scan, twist, cur_odom, prev_odom, icp_ok=0, icp_pose=(0.0, 0.0, 0.0)):
        scan = _f32(scan)
        self.L.ref_pf_set_icp(int(icp_ok), _f64(icp_pose))
        rs = C.c_int()
        rc = self.L.ref_pf_slam(self.h, scan, scan.size, _f64(twist), _f64(cur_odom), _f64(prev_odom), C.byref(rs))
        self.last_resampled = rs.value
        return rc

    def _get(self, w, p, pp):
        self.L.ref_pf_get(self.h, _vp(w), _vp(p), _vp(pp))

    def set_weights(self, w):
        self.L.ref_pf_set_weights(self.h, _f64(w))

    def normalize_resample(self):
        rs = C.c_int()
        anc = np.zeros(self.N, dtype=np.int32)
        self.L.ref_pf_normalize_resample(self.h, C.byref(rs), anc)
        return rs.value, anc

    def robot_state(self):
        out = np.zeros(3)
        self.L.ref_pf_robot_state(self.h, out)
        return out

    def new_map(self):
        out = np.zeros(self.G, dtype=np.int8)
        self.L.ref_pf_new_map(self.h, out)
        return out

    def _grid_dump(self, particle, out):
        self.L.ref_pf_grid_dump(self.h, particle, _vp(out["log_odds"]), _vp(out["prob"]), _vp(out["occ_dist"]), _vp(out["state"]))

    def occ_order(self, particle=0):
        keys = np.zeros(self.G, dtype=np.int32)
        n = self.L.ref_pf_grid_occ_order(self.h, particle, keys, self.G)
        return keys[:n].copy()

    def __del__(self):
        try:
            self.L.ref_pf_destroy(self.h)
        except Exception:
            pass


class RefGrid:
    """A stand-alone compiled reference bmapping::GridMapper."""

    def __init__(self, **kw):
        self.L = _bind_ref_pf()
        q = pf_params(**kw)
        lf, ld = _laser_arrays(q)
        self.h = C.c_void_p(self.L.ref_grid_create(lf, ld, q["resolution"], q["xmin"], q["xmax"], q["ymin"], q["ymax"]))
        xs, ys = C.c_int(), C.c_int()
        self.G = self.L.ref_grid_size(self.h, C.byref(xs), C.byref(ys))
        self.xsize, self.ysize = xs.value, ys.value

    def likelihood(self, scan, pose):
        scan = _f32(scan)
        p = D()
        rc = self.L.ref_grid_likelihood(self.h, scan, scan.size, _f64(pose), C.byref(p))
        return rc, p.value

    def integrate(self, scan, pose):
        scan = _f32(scan)
        return self.L.ref_grid_integrate(self.h, scan, scan.size, _f64(pose))

    def grid(self):
        G = self.G
        out = dict(log_odds=np.zeros(G), prob=np.zeros(G), occ_dist=np.zeros(G), state=np.zeros(G, dtype=np.int32))
        self.L.ref_grid_dump(self.h, _vp(out["log_odds"]), _vp(out["prob"]), _vp(out["occ_dist"]), _vp(out["state"]))
        return out

    def occ_order(self):
        keys = np.zeros(self.G, dtype=np.int32)
        n = self.L.ref_grid_occ_order(self.h, keys, self.G)
        return keys[:n].copy()

    def bucket_count(self):
        return self.L.ref_grid_bucket_count(self.h)

    def grid_map(self):
        out = np.zeros(self.G, dtype=np.int8)
        self.L.ref_grid_map(self.h, out)
        return out

    def end_points(self, scan, pose):
        scan = _f32(scan)
        xy = np.zeros((scan.size, 2))
        n = self.L.ref_grid_end_points(self.h, scan, scan.size, _f64(pose), xy)
        return xy[:n].copy()

    def free_cells(self, pt, pose):
        cells = np.zeros(4 * (self.xsize + self.ysize), dtype=np.int32)
        n = self.L.ref_grid_free_cells(self.h, _f64(pt), _f64(pose), cells, cells.size)
        return None if n < 0 else cells[:n].copy()

    def __del__(self):
        try:
            self.L.ref_grid_destroy(self.h)
        except Exception:
            pass




# ---- scan matcher -------------------------------------------------------------------------------------
class _OICP(C.Structure):
    _fields_ = [("beam_min", C.c_float), ("beam_max", C.c_float), ("beam_delta", C.c_float), ("range_min", C.c_float),
                ("range_max", C.c_float), ("max_iter", C.c_int32), ("max_corr_dist", D), ("transform_eps", D), ("fitness_eps", D)]


class OracleIcp:
    """oracle/icp_oracle.cpp: CPU statement of the scan matcher (ScanAlignment::pclICPWrapper semantics)."""

    def __init__(self, max_iter=100, max_corr_dist=0.5, transform_eps=1e-8, fitness_eps=1e-6, **kw):
        self.L = oracle_lib()
        self.L.orc_icp_create.restype = C.c_void_p
        self.L.orc_icp_create.argtypes = [C.POINTER(_OICP)]
        self.L.orc_icp_destroy.argtypes = [C.c_void_p]
        self.L.orc_icp_align.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(D), C.POINTER(D)]
        self.L.orc_icp_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(D)]
        q = pf_params(**kw)
        p = _OICP(q["beam_min"], q["beam_max"], q["beam_delta"], q["range_min"], q["range_max"], max_iter, max_corr_dist,
                  transform_eps, fitness_eps)
        self.h = C.c_void_p(self.L.orc_icp_create(C.byref(p)))
        self._T = (0.0, 0.0, 0.0)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_icp_destroy(self.h)
            self.h = None

    def pclICPWrapper(self, T_init, scan):
        scan = np.ascontiguousarray(scan, dtype=np.float32)
        Ti = (0.0, 0.0, 0.0) if T_init is None else tuple(T_init)
        T = (D * 3)(*self._T)
        ok = self.L.orc_icp_align(self.h, scan.ctypes.data, scan.size, (D * 3)(*Ti), T)
        if ok:
            self._T = (T[0], T[1], T[2])
        return bool(ok), self._T

    def stats(self):
        it, pairs, mse = C.c_int(0), C.c_int(0), D(0)
        self.L.orc_icp_stats(self.h, C.byref(it), C.byref(pairs), C.byref(mse))
        return it.value, pairs.value, mse.value
