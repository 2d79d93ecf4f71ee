"""Host-side mirror of the reference's controller package surface
(controller/include/controller/mppi.hpp:31-53,58-112,121-155) over the C ABI of libb2nav.so.

Same class names, constructor argument order, method names and error behaviour, so tests read like
tests of the reference.  Every numeric operation happens in the CUDA kernels; this file only
marshals arguments.
"""
import ctypes as C

import numpy as np

from . import _capi
from .rigid2d import Pose, WheelVelocities


class CartModel:
    """controller::CartModel(wheel_radius, wheel_base), mppi.hpp:33."""

    def __init__(self, wheel_radius, wheel_base):
        self.wheel_radius = float(wheel_radius)
        self.wheel_base = float(wheel_base)


class LossFunc:
    """controller::LossFunc(Qdiag, Rdiag, P1diag), mppi.hpp:63-80.  The reference indexes with .at(),
    so short vectors raise (std::out_of_range there, IndexError here)."""

    def __init__(self, Qdiag, Rdiag, P1diag):
        self.Q = [float(Qdiag[0]), float(Qdiag[1]), float(Qdiag[2])]
        self.R = [float(Rdiag[0]), float(Rdiag[1])]
        self.P1 = [float(P1diag[0]), float(P1diag[1]), float(P1diag[2])]


class MPPI:
    """controller::MPPI, mppi.hpp:121-183.

    MPPI(cart_model, loss_func, lambda_, max_wheel_vel, ul_var, ur_var, horizon, dt, rollouts)
    Keyword-only extras place the handle in a sharded job: rollout_offset, rollouts_total, device.
    """

    def __init__(self, cart_model, loss_func, lambda_, max_wheel_vel, ul_var, ur_var, horizon, dt, rollouts,
                 *, rollout_offset=0, rollouts_total=0, device=-1):
        self._lib = _capi.load_library()
        p = _capi.MppiParams()
        p.wheel_radius, p.wheel_base = cart_model.wheel_radius, cart_model.wheel_base
        p.Q[:] = loss_func.Q
        p.R[:] = loss_func.R
        p.P1[:] = loss_func.P1
        p.lambda_, p.max_wheel_vel, p.ul_var, p.ur_var = lambda_, max_wheel_vel, ul_var, ur_var
        p.horizon, p.dt, p.rollouts = horizon, dt, int(rollouts)
        p.rollout_offset, p.rollouts_total, p.device = int(rollout_offset), int(rollouts_total), int(device)
        self._h = C.c_void_p()
        _capi.check(self._lib.b2n_mppi_create(C.byref(p), C.byref(self._h)))
        self.rollouts = int(rollouts)
        self.steps = self._lib.b2n_mppi_steps(self._h)
        # out-parameters of newControls(), made once: the call sits in 50 Hz loops and in the bench's end-to-end leg
        self._ul, self._ur = C.c_double(), C.c_double()
        self._pul, self._pur = C.byref(self._ul), C.byref(self._ur)
        self._new_controls = self._lib.b2n_mppi_new_controls

    # ---- the reference's public methods ---------------------------------------------------
    def setInitialControls(self, uL, uR):
        _capi.check(self._lib.b2n_mppi_set_initial_controls(self._h, uL, uR))

    def setWaypoint(self, wpt):
        _capi.check(self._lib.b2n_mppi_set_waypoint(self._h, wpt.x, wpt.y, wpt.theta))

    def newControls(self, ps):
        rc = self._new_controls(self._h, ps.x, ps.y, ps.theta, self._pul, self._pur)
        if rc:
            _capi.check(rc)
        return WheelVelocities(self._ul.value, self._ur.value)

    # ---- noise seam, taps, bench hooks --------------------------------------------------------
    def seed(self, seed, first_call=0):
        _capi.check(self._lib.b2n_mppi_seed(self._h, seed, first_call))

    def setNoise(self, du):
        du = np.ascontiguousarray(du, dtype=np.float64)
        _capi.check(self._lib.b2n_mppi_set_noise(self._h, _capi.as_ptr(du), du.size))

    def setCapture(self, on=True):
        _capi.check(self._lib.b2n_mppi_set_capture(self._h, int(bool(on))))

    def enqueue(self, ps):
        _capi.check(self._lib.b2n_mppi_enqueue(self._h, ps.x, ps.y, ps.theta))

    def enqueueMany(self, ps, calls):
        """`calls` queued calls with the same pose from one C loop"""
        _capi.check(self._lib.b2n_mppi_enqueue_many(self._h, ps.x, ps.y, ps.theta, int(calls)))

    def wait(self):
        ul, ur = C.c_double(), C.c_double()
        _capi.check(self._lib.b2n_mppi_wait(self._h, C.byref(ul), C.byref(ur)))
        return WheelVelocities(ul.value, ur.value)

    def _get(self, fn, shape, dtype):
        out = np.empty(shape, dtype=dtype)
        _capi.check(fn(self._h, _capi.as_ptr(out), out.size))
        return out

    def states(self):
        return self._get(self._lib.b2n_mppi_get_states, (self.rollouts, self.steps, 3), np.float32)

    def costToGo(self):
        return self._get(self._lib.b2n_mppi_get_cost_to_go, (self.rollouts, self.steps), np.float64)

    def noise(self):
        return self._get(self._lib.b2n_mppi_get_noise, (self.rollouts, self.steps, 2), np.float64)

    def weights(self):
        return self._get(self._lib.b2n_mppi_get_weights, (self.rollouts, self.steps), np.float64)

    def plan(self):
        return self._get(self._lib.b2n_mppi_get_plan, (2, self.steps), np.float64)

    def setPlan(self, u):
        u = np.ascontiguousarray(u, dtype=np.float64)
        _capi.check(self._lib.b2n_mppi_set_plan(self._h, _capi.as_ptr(u), u.size))

    def partials(self):
        return self._get(self._lib.b2n_mppi_get_partials, (self.steps, 6), np.float64)

    def setObstacleField(self, dist, xmin, ymin, resolution, weight, d0, off_map):
        if dist is None:
            _capi.check(self._lib.b2n_mppi_set_obstacle_field(self._h, None, 0, 0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0))
            return
        dist = np.ascontiguousarray(dist, dtype=np.float32)
        xs, ys = dist.shape
        _capi.check(self._lib.b2n_mppi_set_obstacle_field(self._h, _capi.as_ptr(dist), xs, ys, xmin, ymin, resolution,
                                                          weight, d0, off_map))

    def obstacleFieldFrom(self, pf, weight, d0, off_map):
        """the filter's best map becomes the obstacle term, device to device (pf: bmapping.ParticleFilter)"""
        xmin, ymin, res = C.c_double(), C.c_double(), C.c_double()
        _capi.check(self._lib.b2n_pf_geometry(pf._h, C.byref(xmin), C.byref(ymin), C.byref(res)))
        dev = C.c_void_p()
        _capi.check(self._lib.b2n_mppi_obstacle_field_device(self._h, pf.xsize, pf.ysize, xmin.value, ymin.value, res.value, weight, d0,
                                                             off_map, C.byref(dev)))
        _capi.check(self._lib.b2n_pf_write_distance_field(pf._h, dev, pf.cells))

    def setStream(self, cuda_stream):
        _capi.check(self._lib.b2n_mppi_set_stream(self._h, C.c_void_p(cuda_stream)))

    def setStateRing(self, n):
        _capi.check(self._lib.b2n_mppi_set_state_ring(self._h, int(n)))

    def lastVariant(self):
        """which instantiation of the kernel the last call ran: "fast" (production) or "generic" (taps / external noise / ragged horizon)"""
        f = C.c_int()
        _capi.check(self._lib.b2n_mppi_last_variant(self._h, C.byref(f)))
        return "fast" if f.value else "generic"

    def debugTimes(self):
        """tuning: [grid][8] globaltimer stamps of the last call (needs B2N_MPPI_DEBUG_TIMES=1 when the handle was made)"""
        out = np.zeros(24 * 4096, dtype=np.uint64)
        g = C.c_int()
        _capi.check(self._lib.b2n_mppi_debug_times(self._h, _capi.as_ptr(out), out.size, C.byref(g)))
        n = g.value + self.steps
        return out[:24 * n].reshape(n, 24)

    def launchCount(self):
        n = C.c_uint64()
        _capi.check(self._lib.b2n_mppi_launch_count(self._h, C.byref(n)))
        return n.value

    def setKernelTiming(self, on=True):
        _capi.check(self._lib.b2n_mppi_set_kernel_timing(self._h, int(bool(on))))

    def kernelTime(self):
        ms, n = C.c_double(), C.c_int()
        _capi.check(self._lib.b2n_mppi_kernel_time(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def timeRollout(self, ps, launches):
        """bench hook: average duration (ms) of the rollout kernel over `launches` back-to-back launches"""
        ms = C.c_double()
        _capi.check(self._lib.b2n_mppi_time_rollout(self._h, ps.x, ps.y, ps.theta, int(launches), C.byref(ms)))
        return ms.value

    def timeNewControls(self, ps, calls):
        """bench hook: mean wall-clock time (ms) of `calls` synchronous newControls() issued from a C loop, and the last controls"""
        ms, ul, ur = C.c_double(), C.c_double(), C.c_double()
        _capi.check(self._lib.b2n_mppi_time_new_controls(self._h, ps.x, ps.y, ps.theta, int(calls), C.byref(ms), C.byref(ul), C.byref(ur)))
        return ms.value, WheelVelocities(ul.value, ur.value)

    def p2pExport(self, nranks):
        """allocate this rank's exchange area; returns the 64-byte CUDA IPC handle to hand to the other ranks"""
        buf = C.create_string_buffer(64)
        _capi.check(self._lib.b2n_mppi_p2p_export(self._h, int(nranks), buf))
        return buf.raw

    def p2pInit(self, rank, nranks, handles):
        """handles: the nranks x 64 bytes of every rank's p2pExport(), in rank order"""
        buf = C.create_string_buffer(bytes(handles), 64 * int(nranks))
        _capi.check(self._lib.b2n_mppi_p2p_init(self._h, int(rank), int(nranks), buf))

    def p2pArea(self):
        a = C.c_void_p()
        _capi.check(self._lib.b2n_mppi_p2p_area(self._h, C.byref(a)))
        return a.value

    def p2pInitLocal(self, rank, nranks, areas):
        """ranks that live in this process: areas = p2pArea() of every rank's handle, in rank order"""
        arr = (C.c_void_p * int(nranks))(*areas)
        _capi.check(self._lib.b2n_mppi_p2p_init_local(self._h, int(rank), int(nranks), arr))

    def commInit(self, rank, nranks, unique_id):
        buf = C.create_string_buffer(bytes(unique_id), 128)
        _capi.check(self._lib.b2n_mppi_comm_init(self._h, rank, nranks, buf))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.b2n_mppi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def comm_unique_id():
    """128-byte NCCL bootstrap id; make it on one rank and hand it to the others."""
    buf = C.create_string_buffer(128)
    _capi.check(_capi.load_library().b2n_comm_unique_id(buf))
    return buf.raw
