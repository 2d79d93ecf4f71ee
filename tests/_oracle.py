"""ctypes access to the two CPU checkers.  TEST INFRASTRUCTURE ONLY.

  OracleMppi  -> oracle/liboracle_nav.so   (our Eigen-free restatement)
  RefMppi     -> oracle/_ref/libref_nav.so (the unmodified reference sources, compiled in place)
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle_nav.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_nav.so")
D = C.c_double
nd = np.ctypeslib.ndpointer

# shipped parameters: controller/config/mppi_params.yaml, nuturtle_description/config/diff_params.yaml
SHIPPED = dict(wheel_radius=0.033, wheel_base=0.16, Q=(1e4, 1e4, 1.0), R=(0.1, 0.1), P1=(1e3, 1e3, 1e3),
               lambda_=0.01, max_wheel_vel=6.35495, ul_var=0.9, ur_var=0.9)
# well-conditioned set
MILD = dict(wheel_radius=0.033, wheel_base=0.16, Q=(1.0, 1.0, 0.1), R=(0.1, 0.1), P1=(10.0, 10.0, 1.0),
            lambda_=1.0, max_wheel_vel=6.35495, ul_var=0.9, ur_var=0.9)


class _P(C.Structure):
    _fields_ = [("wheel_radius", D), ("wheel_base", D), ("Q", D * 3), ("R", D * 2), ("P1", D * 3),
                ("lambda_", D), ("max_wheel_vel", D), ("ul_var", D), ("ur_var", D), ("horizon", D), ("dt", D),
                ("rollouts", C.c_int)]


_olib = None
_rlib = None


def have_ref():
    return os.path.exists(REF_SO)


def oracle_lib():
    global _olib
    if _olib is None:
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


def unicycle_step(pose, ul, ur, dt, r=0.033, L=0.16):
    """Plant used by closed-loop tests: exact arc integration of the wheel command over dt."""
    x, y, th = pose
    v = r / 2.0 * (ul + ur)
    w = r / L * (ur - ul)
    if abs(w) < 1e-12:
        return (x + v * dt * np.cos(th), y + v * dt * np.sin(th), th)
    return (x + v / w * (np.sin(th + w * dt) - np.sin(th)), y - v / w * (np.cos(th + w * dt) - np.cos(th)), th + w * dt)
