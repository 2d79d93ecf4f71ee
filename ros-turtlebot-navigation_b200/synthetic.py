"""Synthetic inputs and the reference's shipped parameters - plain numpy, no checker, no GPU.

Shared by the bench, the tools and the tests: robot / cost constants from the reference's config files, the synthetic
room with its analytic lidar, the circular test path and the unicycle plant of the closed-loop runs.
Nothing here touches oracle/ or libb2nav.
"""
import numpy as np

# shipped parameters: controller/config/mppi_params.yaml, nuturtle_description/config/diff_params.yaml
SHIPPED = dict(wheel_radius=0.033, wheel_base=0.16, Q=(1e4, 1e4, 1.0), R=(0.1, 0.1), P1=(1e3, 1e3, 1e3),
               lambda_=0.01, max_wheel_vel=6.35495, ul_var=0.9, ur_var=0.9)
# well-conditioned set
MILD = dict(wheel_radius=0.033, wheel_base=0.16, Q=(1.0, 1.0, 0.1), R=(0.1, 0.1), P1=(10.0, 10.0, 1.0),
            lambda_=1.0, max_wheel_vel=6.35495, ul_var=0.9, ur_var=0.9)


def unicycle_step(pose, ul, ur, dt, r=0.033, L=0.16):
    """Plant used by closed-loop tests: exact arc integration of the wheel command over dt."""
    x, y, th = pose
    v = r / 2.0 * (ul + ur)
    w = r / L * (ur - ul)
    if abs(w) < 1e-12:
        return (x + v * dt * np.cos(th), y + v * dt * np.sin(th), th)
    return (x + v / w * (np.sin(th + w * dt) - np.sin(th)), y - v / w * (np.cos(th + w * dt) - np.cos(th)), th + w * dt)


# ------------------------------------------------------------------------------ the reference's plant ---
# rigid2d::DiffDrive (rigid2d/src/rigid2d/diff_drive.cpp:34-239) with the Transform2D pieces it uses
# (rigid2d/src/rigid2d/rigid2d.cpp:118-238,239-300), statement by statement in Python floats (C doubles, the same libm):
# the simulated robot of fake_diff_encoders (rigid2d/src/fake_diff_encoders_node.cpp:100-135: feedforward of the commanded
# twist scaled by 1 / frequency, encoder angles out) and the odometer of the nodes (updateOdometry on the encoder angles).
# tests/test_host_logic.py pins it against the compiled reference (oracle/_ref), call by call.
import math

_PI = 3.14159265358979323846


def almost_equal(d1, d2, epsilon=1.0e-12):
    """rigid2d.hpp:24-27"""
    return math.fabs(d1 - d2) < epsilon


def normalize_angle_PI(rad):
    """rigid2d.hpp:52-64"""
    q = math.floor((rad + _PI) / (2.0 * _PI))
    rad = (rad + _PI) - q * 2.0 * _PI
    if rad < 0:
        rad += 2.0 * _PI
    return rad - _PI


def _integrate_twist_from(theta, x, y, w, vx, vy):
    """Transform2D(theta, (x, y)).integrateTwist(twist).displacement(), rigid2d.cpp:239-300 (with its quirks: the new
    transform carries cos / sin of the OLD heading, and the product adds the headings, :218-227)."""
    ctheta, stheta = math.cos(theta), math.sin(theta)
    beta, Sw, Svx, Svy = 0.0, 0.0, 0.0, 0.0
    if not almost_equal(w, 0.0):
        beta = abs(w)
        Sw, Svx, Svy = w / beta, vx / beta, vy / beta
    elif almost_equal(w, 0.0) and almost_equal(vx, 0.0) and almost_equal(vy, 0.0):
        return theta, x, y
    else:
        beta = math.sqrt(math.pow(vx, 2) + math.pow(vy, 2))
        Svx, Svy = vx / beta, vy / beta
    cbeta, sbeta = math.cos(beta), math.sin(beta)
    theta_new = math.atan2(sbeta * Sw, 1 + (1 - cbeta) * (-1.0 * math.pow(Sw, 2)))
    x_new = Svx * (beta + (beta - sbeta) * (-1.0 * math.pow(Sw, 2))) + Svy * ((1 - cbeta) * (-1.0 * Sw))
    y_new = Svx * ((1 - cbeta) * Sw) + Svy * (beta + (beta - sbeta) * (-1.0 * math.pow(Sw, 2)))
    # T * T_new (operator*=, :218-227)
    xo = ctheta * x_new - stheta * y_new + x
    yo = stheta * x_new + ctheta * y_new + y
    return theta + theta_new, xo, yo


class DiffDrive:
    """rigid2d::DiffDrive(pose, wheel_base, wheel_radius), diff_drive.cpp:34-53; pose = (theta, x, y)"""

    def __init__(self, pose=(0.0, 0.0, 0.0), wheel_base=0.16, wheel_radius=0.033):
        self.theta, self.x, self.y = float(pose[0]), float(pose[1]), float(pose[2])
        self.wheel_base, self.wheel_radius = float(wheel_base), float(wheel_radius)
        self.left_curr = self.right_curr = 0.0
        self.ul = self.ur = 0.0

    def twistToWheels(self, w, vx, vy=0.0):
        """diff_drive.cpp:56-76 -> (ul, ur)"""
        d = self.wheel_base / 2
        ul = (1 / self.wheel_radius) * (-d * w + vx)
        ur = (1 / self.wheel_radius) * (d * w + vx)
        if vy != 0:
            raise ValueError("Twist cannot have y velocity component")
        return ul, ur

    def wheelsToTwist(self, ul, ur):
        """diff_drive.cpp:79-94 -> (w, vx, vy)"""
        d = 1 / self.wheel_base
        return self.wheel_radius * d * (ur - ul), self.wheel_radius * 0.5 * (ul + ur), 0.0

    def _advance(self, w, vx, vy):
        # Tb_bprime = identity.integrateTwist(twist); Tw_bprime = Twb * Tb_bprime (diff_drive.cpp:124-147, 171-194)
        tb, xb, yb = _integrate_twist_from(0.0, 0.0, 0.0, w, vx, vy)
        c, s = math.cos(self.theta), math.sin(self.theta)
        x = c * xb - s * yb + self.x
        y = s * xb + c * yb + self.y
        self.theta = normalize_angle_PI(self.theta + tb)
        self.x, self.y = x, y

    def updateOdometry(self, left, right):
        """diff_drive.cpp:97-152 -> wheel velocities (ul, ur)"""
        ul = normalize_angle_PI(left - self.left_curr)
        ur = normalize_angle_PI(right - self.right_curr)
        self.ul, self.ur = ul, ur
        self.left_curr, self.right_curr = normalize_angle_PI(left), normalize_angle_PI(right)
        self._advance(*self.wheelsToTwist(ul, ur))
        return ul, ur

    def feedforward(self, w, vx, vy=0.0):
        """diff_drive.cpp:155-195"""
        ul, ur = self.twistToWheels(w, vx, vy)
        self.ul, self.ur = normalize_angle_PI(ul), normalize_angle_PI(ur)
        self.left_curr = normalize_angle_PI(self.left_curr + ul)
        self.right_curr = normalize_angle_PI(self.right_curr + ur)
        self._advance(w, vx, vy)

    def pose(self):
        """diff_drive.cpp:198-206 -> (theta, x, y)"""
        return normalize_angle_PI(self.theta), self.x, self.y

    def wheelVelocities(self):
        return self.ul, self.ur

    def getEncoders(self):
        return self.left_curr, self.right_curr


class WaypointSwitch:
    """The waypoint bookkeeping of the reference's control loop (nuturtle_robot/src/mppi_waypoints_node.cpp:231-258):
    when the pose comes within goal_thresh of the current waypoint, advance to the next one (wrapping) and tell the
    controller; after len + 1 switches one cycle is complete.  Waypoints: lists of x, y, theta."""

    def __init__(self, xs, ys, thetas, goal_thresh):
        self.xs, self.ys, self.thetas, self.goal_thresh = list(xs), list(ys), list(thetas), float(goal_thresh)
        self.wpt_id, self.cnt, self.cycle_complete = 0, 0, False

    def current(self):
        return self.xs[self.wpt_id], self.ys[self.wpt_id], self.thetas[self.wpt_id]

    def update(self, x, y):
        """-> the new waypoint (x, y, theta) when the controller must be told, else None"""
        wx, wy, _ = self.current()
        dx, dy = wx - x, wy - y
        d2g = math.sqrt(dx * dx + dy * dy)                               # rigid2d::euclideanDistance, utilities.cpp:59-64
        if not d2g < self.goal_thresh:
            return None
        self.wpt_id += 1
        self.cnt += 1
        if self.wpt_id % len(self.xs) == 0:
            self.wpt_id = 0
        if self.cnt == len(self.xs) + 1:
            self.cycle_complete = True
        return self.current()


# bmapping/launch/slam.launch:19-42 + bmapping/config/LDS_01_lidar.yaml, with the synthetic-bench
# changes of SURVEY.md 8d (10 m map -> 200x200 cells, motion noise raised so that weights diverge)
PF_SHIPPED = dict(
    beam_min=0.0, beam_max=float(np.float32(np.deg2rad(360.0))), beam_delta=float(np.float32(np.deg2rad(1.0))),
    range_min=0.12, range_max=3.5, z_hit=0.95, z_short=0.0, z_max=0.04, z_rand=0.01, sigma_hit=0.5,
    resolution=0.05, xmin=-5.0, xmax=5.0, ymin=-5.0, ymax=5.0,
    num_particles=40, k=50, srr=0.01, srt=0.02, str_=0.01, stt=0.02,
    motion_noise=(1e-4, 1e-4, 1e-4), sample_range=(1e-4, 1e-4, 1e-4),
    scan_min=1.0, scan_max=20.0, pose_min=1.0, pose_max=10.0, init_pose=(0.0, 0.0, 0.0))


def pf_params(**kw):
    q = dict(PF_SHIPPED)
    q.update(kw)
    q["k"], q["num_particles"] = int(q["k"]), int(q["num_particles"])
    return q


# ------------------------------------------------------------------------------ synthetic world ---
def room_scan(pose, half=2.5, boxes=((0.8, 1.4, -0.3, 0.4),), n_beams=360, beam_delta=np.deg2rad(1.0), sigma=0.01, rng=None,
              range_max=3.5):
    """Analytic ray cast from pose = (theta, x, y) in an axis-aligned square room of half-width `half` with
    axis-aligned boxes (xlo, xhi, ylo, yhi); Gaussian range noise sigma (Gazebo lidar,
    nuturtle_gazebo/urdf/diff_drive.gazebo.xacro:101-105); float32 ranges; beams past range_max read range_max + 1
    (filtered by the range gate like an LDS-01 'inf')."""
    th, x, y = pose
    out = np.zeros(n_beams, dtype=np.float32)
    for b in range(n_beams):
        a = th + b * beam_delta
        dx, dy = np.cos(a), np.sin(a)
        best = np.inf
        rects = [(-half, half, -half, half)] + list(boxes)
        for (xl, xh, yl, yh) in rects:
            for (px, horiz) in ((xl, False), (xh, False), (yl, True), (yh, True)):
                if not horiz:
                    if abs(dx) < 1e-12:
                        continue
                    t = (px - x) / dx
                    q = y + t * dy
                    ok = yl - 1e-12 <= q <= yh + 1e-12
                else:
                    if abs(dy) < 1e-12:
                        continue
                    t = (px - y) / dy
                    q = x + t * dx
                    ok = xl - 1e-12 <= q <= xh + 1e-12
                if ok and t > 1e-9 and t < best:
                    best = t
        r = best + (rng.normal(0.0, sigma) if (rng is not None and sigma > 0) else 0.0)
        out[b] = np.float32(r if r < range_max else range_max + 1.0)
    return out


def circle_path(n_scans, radius=0.5, step=0.05):
    """Robot poses (theta, x, y) on a circle of `radius`, arc length `step` per scan, heading tangent; with the
    per-scan body twist (w, vx, vy) that takes one pose to the next."""
    poses, twists = [], []
    dphi = step / radius
    for i in range(n_scans + 1):
        phi = i * dphi
        poses.append((np.pi / 2 + phi, radius * np.cos(phi), radius * np.sin(phi)))
    for i in range(n_scans):
        twists.append((dphi, step, 0.0))
    return np.array(poses), np.array(twists)
