"""BASELINE configs[4] on one GPU's share: RBPF (4096 particles) + MPPI (K = 16384, T = 64) in closed loop on a synthetic
trajectory - control at 50 Hz, lidar at 5 Hz (LDS-01), plant = exact unicycle arcs, waypoints = the reference's pentagon
(nuturtle_robot/config/real_waypoints.yaml:3-7) scaled into the synthetic room.  Prints one JSON line with the per-tick
latencies against the 20 ms control budget and the 200 ms scan budget.  Not the bench; a measured demo of 8(f) row 1.

    python tools/closed_loop.py [ticks] [particles] [rollouts]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _pkg  # noqa: E402

ticks = int(sys.argv[1]) if len(sys.argv) > 1 else 300
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
K = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
pkg = _pkg.load()
orc = pkg.synthetic          # shipped parameters and synthetic inputs (plain numpy)
prm = orc.SHIPPED
dt, hor, scan_every = 0.02, 0.64 * 2, 10          # 50 Hz control, T = 64 at dt = 0.02, 5 Hz lidar
rng = np.random.default_rng(0)
start = (0.0, 0.0, 0.0)
f = pkg.bmapping.make_filter(orc.pf_params(num_particles=N, init_pose=start, motion_noise=(2e-3, 1e-3, 1e-3)))
f.seed(1)
m = pkg.MPPI(pkg.CartModel(prm["wheel_radius"], prm["wheel_base"]), pkg.LossFunc(prm["Q"], prm["R"], prm["P1"]), prm["lambda_"],
             prm["max_wheel_vel"], prm["ul_var"], prm["ur_var"], hor, dt, K)
m.seed(2)
wpts = [(0.8, 0.0), (1.05, 0.76), (0.4, 1.23), (-0.25, 0.76), (0.0, 0.0)]
wi = 0
m.setWaypoint(pkg.Pose(theta=0.0, x=wpts[0][0], y=wpts[0][1]))
true = (0.0, 0.0, 0.0)                             # x, y, theta
est = start
odom_prev = start
t_mppi, t_slam, reached = [], [], 0
# first scan builds the map before the loop starts
f.SLAM(orc.room_scan(start, rng=rng), pkg.Twist2D(0.0, 0.0, 0.0), pkg.Pose(*start), pkg.Pose(*start))
for k in range(ticks):
    t0 = time.perf_counter()
    v = m.newControls(pkg.Pose(theta=est[0], x=est[1], y=est[2]))
    t_mppi.append(time.perf_counter() - t0)
    true = orc.unicycle_step(true, v.ul, v.ur, dt)
    est = (true[2], true[0], true[1]) if (k + 1) % scan_every else est     # dead-reckon between scans (perfect odometry)
    if (k + 1) % scan_every == 0:
        odom_cur = (true[2], true[0], true[1])
        twist = (odom_cur[0] - odom_prev[0], float(np.hypot(odom_cur[1] - odom_prev[1], odom_cur[2] - odom_prev[2])), 0.0)
        scan = orc.room_scan(odom_cur, rng=rng)
        t0 = time.perf_counter()
        f.SLAM(scan, pkg.Twist2D(*twist), pkg.Pose(*odom_cur), pkg.Pose(*odom_prev))
        T = f.getRobotState().displacement()
        t_slam.append(time.perf_counter() - t0)
        odom_prev = odom_cur
        est = tuple(T)
    if np.hypot(true[0] - wpts[wi][0], true[1] - wpts[wi][1]) < 0.08:
        reached += 1
        wi = (wi + 1) % len(wpts)
        m.setWaypoint(pkg.Pose(theta=0.0, x=wpts[wi][0], y=wpts[wi][1]))
err = float(np.hypot(est[1] - true[0], est[2] - true[1]))
print(json.dumps({
    "what": "closed loop RBPF + MPPI, one GPU", "particles": N, "rollouts": K, "horizon_steps": m.steps, "ticks": ticks,
    "control_hz": 1.0 / dt, "scan_hz": 1.0 / (dt * scan_every),
    "mppi_ms_per_tick": {"mean": 1e3 * float(np.mean(t_mppi)), "p99": 1e3 * float(np.percentile(t_mppi, 99)), "budget": 1e3 * dt},
    "slam_ms_per_scan": {"mean": 1e3 * float(np.mean(t_slam)), "max": 1e3 * float(np.max(t_slam)), "budget": 1e3 * dt * scan_every},
    "waypoints_reached": reached, "final_position_error_m": err,
}))
