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
