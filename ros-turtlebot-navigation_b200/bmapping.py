"""Host-side mirror of the reference's bmapping package surface over the C ABI of libb2nav.so:
LaserProperties (bmapping/include/bmapping/sensor_model.hpp:20-79), GridMapper (grid_mapper.hpp:117-122),
ScanAlignment (cloud_alignment.hpp:28-46) and ParticleFilter (particle_filter.hpp:89-144).

Same class names, constructor argument order, method names and error behaviour, so tests read like tests of the
reference.  GridMapper and LaserProperties only carry parameters here: the per-particle maps live on the GPU.
Every numeric operation of SLAM() happens in the CUDA kernels; this file only marshals arguments.
"""
import ctypes as C
import math

import numpy as np

from . import _capi
from .rigid2d import Pose, Transform2D, Twist2D, Vector2D  # noqa: F401


class LaserProperties:
    """bmapping::LaserProperties(beam_min, beam_max, beam_delta, range_min, range_max, z_hit, z_short, z_max, z_rand,
    sigma_hit); the first five are float in the reference and are rounded to float32 here as well."""

    def __init__(self, beam_min, beam_max, beam_delta, range_min, range_max, z_hit, z_short, z_max, z_rand, sigma_hit):
        f = lambda v: float(np.float32(v))  # noqa: E731
        self.beam_min, self.beam_max, self.beam_delta = f(beam_min), f(beam_max), f(beam_delta)
        self.range_min, self.range_max = f(range_min), f(range_max)
        self.z_hit, self.z_short, self.z_max, self.z_rand, self.sigma_hit = (float(z_hit), float(z_short), float(z_max),
                                                                              float(z_rand), float(sigma_hit))


class GridMapper:
    """bmapping::GridMapper(resolution, xmin, xmax, ymin, ymax, props, Trs): the prototype map every particle starts from
    (particle_filter.cpp:125-138).  Trs must be the identity, as in turtle_mapping_node.cpp:397."""

    def __init__(self, resolution, xmin, xmax, ymin, ymax, props, Trs=None):
        if Trs is not None and Trs.displacement() != (0.0, 0.0, 0.0):
            raise ValueError("only Trs = identity is supported (turtle_mapping_node.cpp:397)")
        self.resolution, self.xmin, self.xmax, self.ymin, self.ymax = (float(resolution), float(xmin), float(xmax),
                                                                      float(ymin), float(ymax))
        self.props = props


class ScanAlignment:
    """bmapping::ScanAlignment(props, Trs).  The reference's implementation wraps PCL's ICP (cloud_alignment.cpp:37-72),
    which stays on the caller's side of the boundary.  This stand-in reports 'no match' unless an outcome is injected
    with setResult(); a caller with PCL overrides pclICPWrapper."""

    def __init__(self, props, Trs=None):
        self.props = props
        self._ok, self._T = False, (0.0, 0.0, 0.0)

    def setResult(self, ok, T=None):
        self._ok = bool(ok)
        if T is not None:
            self._T = T.displacement() if hasattr(T, "displacement") else tuple(T)

    def pclICPWrapper(self, T_init, beam_length):
        """-> (success, (theta, x, y) of Ticp)"""
        return self._ok, self._T


class GpuScanAlignment(ScanAlignment):
    """ScanAlignment whose matcher runs on the GPU (b2n_icp_*): libb2nav's own point-to-point ICP with the reference's
    settings (cloud_alignment.cpp:20-25) and wrapper semantics (:37-72).  The reference's matcher is PCL's, which is
    not pinned here - results are checked against oracle/icp_oracle.cpp, not against PCL."""

    def __init__(self, props, Trs=None, *, max_iter=100, max_correspondence_dist=0.5, transformation_epsilon=1e-8,
                 euclidean_fitness_epsilon=1e-6, device=-1, max_beams=0):
        super().__init__(props, Trs)
        self._lib = _capi.load_library()
        p = _capi.IcpParams()
        p.beam_min, p.beam_max, p.beam_delta, p.range_min, p.range_max = props.beam_min, props.beam_max, props.beam_delta, props.range_min, props.range_max
        p.max_iter, p.max_correspondence_dist = int(max_iter), max_correspondence_dist
        p.transformation_epsilon, p.euclidean_fitness_epsilon = transformation_epsilon, euclidean_fitness_epsilon
        p.device, p.max_beams = int(device), int(max_beams)
        h = C.c_void_p()
        _capi.check(self._lib.b2n_icp_create(C.byref(p), C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.b2n_icp_destroy(self._h)
            self._h = None

    __del__ = close

    def pclICPWrapper(self, T_init, beam_length):
        scan = np.ascontiguousarray(beam_length, dtype=np.float32)
        D3 = C.c_double * 3
        Ti = (0.0, 0.0, 0.0) if T_init is None else tuple(T_init)
        T, ok = D3(*self._T), C.c_int(0)
        _capi.check(self._lib.b2n_icp_align(self._h, _capi.as_ptr(scan), scan.size, D3(*Ti), T, C.byref(ok)))
        if ok.value:
            self._T = (T[0], T[1], T[2])
        return bool(ok.value), self._T

    def stats(self):
        """-> (iterations, correspondences, mean squared pair distance, kernel launches)"""
        it, pairs, mse, launches = C.c_int(0), C.c_int(0), C.c_double(0), C.c_uint64(0)
        _capi.check(self._lib.b2n_icp_stats(self._h, C.byref(it), C.byref(pairs), C.byref(mse), C.byref(launches)))
        return it.value, pairs.value, mse.value, launches.value


def icpInitGuess(cur_odom, prev_odom):
    """particle_filter.cpp:602-612 -> (theta, x, y): the odometry difference (world-frame dx, dy, like the reference)"""
    def npi(rad):                      # rigid2d::normalize_angle_PI, rigid2d.hpp:52-64
        q = math.floor((rad + math.pi) / (2.0 * math.pi))
        rad = (rad + math.pi) - q * 2.0 * math.pi
        if rad < 0:
            rad += 2.0 * math.pi
        return rad - math.pi
    return (npi(npi(cur_odom.theta) - npi(prev_odom.theta)), cur_odom.x - prev_odom.x, cur_odom.y - prev_odom.y)


class ParticleFilter:
    """bmapping::ParticleFilter, particle_filter.hpp:112-144.

    ParticleFilter(num_particles, k, srr, srt, str, stt, motion_noise_theta, motion_noise_x, motion_noise_y,
                   sample_range_theta, sample_range_x, sample_range_y, scan_likelihood_min, scan_likelihood_max,
                   pose_likelihood_min, pose_likelihood_max, scan_matcher, pose, mapper)
    Keyword-only extras place the handle in a sharded job: particle_offset, particles_total, device, max_beams.
    """

    def __init__(self, num_particles, k, srr, srt, str_, stt, motion_noise_theta, motion_noise_x, motion_noise_y,
                 sample_range_theta, sample_range_x, sample_range_y, scan_likelihood_min, scan_likelihood_max,
                 pose_likelihood_min, pose_likelihood_max, scan_matcher, pose, mapper,
                 *, particle_offset=0, particles_total=0, device=-1, max_beams=0):
        self._lib = _capi.load_library()
        L = mapper.props
        p = _capi.PfParams()
        p.beam_min, p.beam_max, p.beam_delta, p.range_min, p.range_max = L.beam_min, L.beam_max, L.beam_delta, L.range_min, L.range_max
        p.z_hit, p.z_short, p.z_max, p.z_rand, p.sigma_hit = L.z_hit, L.z_short, L.z_max, L.z_rand, L.sigma_hit
        p.resolution, p.xmin, p.xmax, p.ymin, p.ymax = mapper.resolution, mapper.xmin, mapper.xmax, mapper.ymin, mapper.ymax
        p.num_particles, p.k = int(num_particles), int(k)
        p.srr, p.srt, p.str_, p.stt = srr, srt, str_, stt
        p.motion_noise_theta, p.motion_noise_x, p.motion_noise_y = motion_noise_theta, motion_noise_x, motion_noise_y
        p.sample_range_theta, p.sample_range_x, p.sample_range_y = sample_range_theta, sample_range_x, sample_range_y
        p.scan_likelihood_min, p.scan_likelihood_max = scan_likelihood_min, scan_likelihood_max
        p.pose_likelihood_min, p.pose_likelihood_max = pose_likelihood_min, pose_likelihood_max
        th, x, y = pose.displacement()
        p.init_pose[:] = [th, x, y]
        p.particle_offset, p.particles_total, p.device, p.max_beams = int(particle_offset), int(particles_total), int(device), int(max_beams)
        self._h = C.c_void_p()
        _capi.check(self._lib.b2n_pf_create(C.byref(p), C.byref(self._h)))
        self.num_particles = int(num_particles)
        self.particles_total = int(particles_total) if particles_total else int(num_particles)
        self.k = int(k)
        self.scan_matcher = scan_matcher
        xs, ys = C.c_int(), C.c_int()
        _capi.check(self._lib.b2n_pf_grid_size(self._h, C.byref(xs), C.byref(ys)))
        self.xsize, self.ysize = xs.value, ys.value
        self.cells = xs.value * ys.value

    # ---- the reference's public methods -------------------------------------------------------------
    def SLAM(self, scan, u, cur_odom, prev_odom):
        """particle_filter.cpp:141-251.  scan: float ranges; u: Twist2D; cur_odom / prev_odom: Pose."""
        scan = np.ascontiguousarray(scan, dtype=np.float32)
        # icpInitGuess, then the matcher, exactly where the reference calls them (particle_filter.cpp:150-153)
        ok, T = self.scan_matcher.pclICPWrapper(icpInitGuess(cur_odom, prev_odom), scan)
        D3 = C.c_double * 3
        _capi.check(self._lib.b2n_pf_slam(self._h, _capi.as_ptr(scan), scan.size, D3(u.w, u.vx, u.vy),
                                          D3(cur_odom.theta, cur_odom.x, cur_odom.y), D3(prev_odom.theta, prev_odom.x, prev_odom.y),
                                          int(bool(ok)), D3(*T)))

    def getRobotState(self):
        """particle_filter.cpp:255-274 -> Transform2D of the best particle's pose"""
        out = (C.c_double * 3)()
        _capi.check(self._lib.b2n_pf_get_robot_state(self._h, out))
        return Transform2D(Vector2D(out[1], out[2]), out[0])

    def newMap(self):
        """particle_filter.cpp:277-291: the best particle's map as int8 occupancy values (transposed, like the reference)"""
        out = np.empty(self.cells, dtype=np.int8)
        _capi.check(self._lib.b2n_pf_new_map(self._h, _capi.as_ptr(out), out.size))
        return out

    # ---- noise seam, taps, bench hooks ----------------------------------------------------------------
    def seed(self, seed, first_call=0):
        _capi.check(self._lib.b2n_pf_seed(self._h, seed, first_call))

    def setNoise(self, z):
        z = np.ascontiguousarray(z, dtype=np.float64)
        _capi.check(self._lib.b2n_pf_set_noise(self._h, _capi.as_ptr(z), z.size))

    def weights(self):
        out = np.empty(self.num_particles)
        _capi.check(self._lib.b2n_pf_get_weights(self._h, _capi.as_ptr(out), out.size))
        return out

    def setWeights(self, w):
        w = np.ascontiguousarray(w, dtype=np.float64)
        _capi.check(self._lib.b2n_pf_set_weights(self._h, _capi.as_ptr(w), w.size))

    def poses(self):
        p, pp = np.empty((self.num_particles, 3)), np.empty((self.num_particles, 3))
        _capi.check(self._lib.b2n_pf_get_poses(self._h, _capi.as_ptr(p), _capi.as_ptr(pp), p.size))
        return p, pp

    def setPoses(self, poses):
        poses = np.ascontiguousarray(poses, dtype=np.float64)
        _capi.check(self._lib.b2n_pf_set_poses(self._h, _capi.as_ptr(poses), poses.size))

    def resampleInfo(self):
        neff, rs = C.c_int(), C.c_int()
        anc = np.empty(self.particles_total, dtype=np.int32)
        _capi.check(self._lib.b2n_pf_get_resample(self._h, C.byref(neff), C.byref(rs), _capi.as_ptr(anc), anc.size))
        return neff.value, rs.value, anc

    def normalizeResample(self):
        _capi.check(self._lib.b2n_pf_normalize_resample(self._h))
        return self.resampleInfo()

    def grid(self, particle=0):
        G = self.cells
        out = dict(log_odds=np.empty(G), occ_dist=np.empty(G), state=np.empty(G, dtype=np.int8))
        _capi.check(self._lib.b2n_pf_get_grid(self._h, particle, _capi.as_ptr(out["log_odds"]), _capi.as_ptr(out["occ_dist"]),
                                              _capi.as_ptr(out["state"]), G))
        return out

    def occOrder(self, particle=0):
        keys = np.empty(self.cells, dtype=np.int32)
        n = C.c_int()
        _capi.check(self._lib.b2n_pf_get_occ_order(self._h, particle, _capi.as_ptr(keys), keys.size, C.byref(n)))
        return keys[:n.value].copy()

    def likelihoods(self, scan):
        scan = np.ascontiguousarray(scan, dtype=np.float32)
        out = np.empty(self.num_particles)
        _capi.check(self._lib.b2n_pf_likelihoods(self._h, _capi.as_ptr(scan), scan.size, _capi.as_ptr(out), out.size))
        return out

    def setStream(self, cuda_stream):
        _capi.check(self._lib.b2n_pf_set_stream(self._h, C.c_void_p(cuda_stream)))

    def launchCount(self):
        n = C.c_uint64()
        _capi.check(self._lib.b2n_pf_launch_count(self._h, C.byref(n)))
        return n.value

    def setKernelTiming(self, on=True):
        _capi.check(self._lib.b2n_pf_set_kernel_timing(self._h, int(bool(on))))

    def kernelTimes(self):
        ms = (C.c_double * 3)()
        _capi.check(self._lib.b2n_pf_kernel_times(self._h, ms))
        return tuple(ms)

    def distanceFieldStats(self):
        it, hm = C.c_uint64(), C.c_uint64()
        _capi.check(self._lib.b2n_pf_distance_field_stats(self._h, C.byref(it), C.byref(hm)))
        return it.value, hm.value

    def distanceFieldSkipped(self):
        n = C.c_uint64()
        _capi.check(self._lib.b2n_pf_distance_field_skipped(self._h, C.byref(n)))
        return n.value

    def setHeapCapacity(self, entries):
        _capi.check(self._lib.b2n_pf_set_heap_capacity(self._h, int(entries)))

    def p2pExport(self):
        """the 640 bytes of CUDA IPC handles of this rank's planes, to hand to the other ranks"""
        buf = C.create_string_buffer(640)
        _capi.check(self._lib.b2n_pf_p2p_export(self._h, buf))
        return buf.raw

    def p2pInit(self, rank, nranks, handles):
        """handles: nranks x 640 bytes, every rank's p2pExport() in rank order (after commInit)"""
        buf = C.create_string_buffer(bytes(handles), 640 * int(nranks))
        _capi.check(self._lib.b2n_pf_p2p_init(self._h, int(rank), int(nranks), buf))

    def migration(self):
        """(particles received from, sent to) other ranks by the last SLAM()"""
        a, b = C.c_int(), C.c_int()
        _capi.check(self._lib.b2n_pf_get_migration(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def commInit(self, rank, nranks, unique_id):
        buf = C.create_string_buffer(bytes(unique_id), 128)
        _capi.check(self._lib.b2n_pf_comm_init(self._h, rank, nranks, buf))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.b2n_pf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def make_filter(q, **kw):
    """ParticleFilter from a flat parameter dict (the keys of tests/_oracle.PF_SHIPPED)."""
    props = LaserProperties(q["beam_min"], q["beam_max"], q["beam_delta"], q["range_min"], q["range_max"], q["z_hit"],
                            q["z_short"], q["z_max"], q["z_rand"], q["sigma_hit"])
    mapper = GridMapper(q["resolution"], q["xmin"], q["xmax"], q["ymin"], q["ymax"], props, Transform2D())
    matcher = ScanAlignment(props, Transform2D())
    th, x, y = q["init_pose"]
    return ParticleFilter(q["num_particles"], q["k"], q["srr"], q["srt"], q["str_"], q["stt"], *q["motion_noise"],
                          *q["sample_range"], q["scan_min"], q["scan_max"], q["pose_min"], q["pose_max"], matcher,
                          Transform2D(Vector2D(x, y), th), mapper, **kw)
