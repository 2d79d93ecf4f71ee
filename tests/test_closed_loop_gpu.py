"""BASELINE configs[4] in miniature: RBPF pose estimate -> MPPI wheel command -> plant -> odometry + lidar scan, tick by
tick, the GPU stack and the CPU oracle stack fed the same scans and odometry (mppi_waypoints_node.cpp:231-282 and
turtle_mapping_node.cpp:451-494 fused into one loop; the plant stands in for fake_diff_encoders + Gazebo).

Every tick: the filter's best pose within 1e-9 of the oracle's, the controller's command within 1e-5, ancestors equal.
"""
import numpy as np
import pytest

import _oracle as orc

pytestmark = pytest.mark.gpu


def test_rbpf_feeds_mppi_closed_loop(gpu_pkg):
    pkg = gpu_pkg
    N, K, hor, dt, ticks = 48, 512, 0.5, 0.02, 8
    scan_every = 2                                     # a lidar scan every other control tick
    rng = np.random.default_rng(21)
    start = (0.0, 0.6, 0.0)                            # theta, x, y
    q = dict(num_particles=N, init_pose=start, motion_noise=(2e-3, 1e-3, 1e-3))
    f = pkg.bmapping.make_filter(orc.pf_params(**q))
    of = orc.OraclePf(**q)
    f.seed(5)
    of.noise_philox(5)
    prm = orc.SHIPPED
    m = pkg.MPPI(pkg.CartModel(prm["wheel_radius"], prm["wheel_base"]), pkg.LossFunc(prm["Q"], prm["R"], prm["P1"]), prm["lambda_"],
                 prm["max_wheel_vel"], prm["ul_var"], prm["ur_var"], hor, dt, K)
    om = orc.OracleMppi(hor, dt, K)
    m.seed(6)
    om.noise_philox(6)
    m.setWaypoint(pkg.Pose(theta=0.0, x=1.6, y=0.3))
    om.setWaypoint(1.6, 0.3, 0.0)

    true = (start[1], start[2], start[0])              # plant state x, y, theta
    odom_prev = start
    est = start
    resampled = 0
    for k in range(ticks):
        # controller on the current estimate
        v = m.newControls(pkg.Pose(theta=est[0], x=est[1], y=est[2]))
        c = om.newControls(est[1], est[2], est[0])
        assert max(abs(v.ul - c[0]), abs(v.ur - c[1])) / max(1e-3, abs(c[0]), abs(c[1])) < 1e-5, (k, v, c)
        # plant + perfect wheel odometry
        true = orc.unicycle_step(true, c[0], c[1], dt)
        if (k + 1) % scan_every:
            continue
        odom_cur = (true[2], true[0], true[1])
        dth = odom_cur[0] - odom_prev[0]
        dist = np.hypot(odom_cur[1] - odom_prev[1], odom_cur[2] - odom_prev[2])
        twist = (dth, dist, 0.0)                        # body twist integrated over the scan interval
        scan = orc.room_scan(odom_cur, rng=rng)
        f.SLAM(scan, pkg.Twist2D(*twist), pkg.Pose(*odom_cur), pkg.Pose(*odom_prev))
        of.slam(scan, twist, odom_cur, odom_prev)
        odom_prev = odom_cur
        T = f.getRobotState().displacement()
        want = of.robot_state()
        assert np.max(np.abs(np.array(T) - want)) < 1e-9, (k, T, want)
        neff, rs, anc = f.resampleInfo()
        oneff, ors, oanc = of.resample_info()
        assert (neff, rs) == (oneff, ors) and np.array_equal(anc, oanc)
        resampled += rs
        assert np.array_equal(f.newMap(), of.new_map())
        est = tuple(want)
    assert resampled >= 1


def test_filter_map_feeds_the_obstacle_term_on_the_device(gpu_pkg):
    """the best particle's distance field goes device-to-device into the controller's obstacle term;
    the oracle gets the same field through the host."""
    pkg = gpu_pkg
    N, K, hor, dt = 24, 1024, 0.64, 0.01
    poses, twists = orc.circle_path(3)
    rng = np.random.default_rng(8)
    q = dict(num_particles=N, init_pose=tuple(poses[0]), motion_noise=(2e-3, 1e-3, 1e-3))
    f = pkg.bmapping.make_filter(orc.pf_params(**q))
    f.seed(2)
    for i in range(3):
        f.SLAM(orc.room_scan(poses[i + 1], rng=rng), pkg.Twist2D(*twists[i]), pkg.Pose(*poses[i + 1]), pkg.Pose(*poses[i]))
    prm = orc.SHIPPED
    m = pkg.MPPI(pkg.CartModel(prm["wheel_radius"], prm["wheel_base"]), pkg.LossFunc(prm["Q"], prm["R"], prm["P1"]), prm["lambda_"],
                 prm["max_wheel_vel"], prm["ul_var"], prm["ur_var"], hor, dt, K)
    m.obstacleFieldFrom(f, 5e4, 0.4, 1e6)
    # the same field for the oracle: best particle = first maximum of the weights (particle_filter.cpp:255-274)
    best = int(np.argmax(f.weights()))
    dist = f.grid(best)["occ_dist"].astype(np.float32).reshape(f.xsize, f.ysize)
    o = orc.OracleMppi(hor, dt, K)
    o.set_obstacles(dist, -5.0, -5.0, 0.05, 5e4, 0.4, 1e6)
    m.setCapture(True)
    m.seed(4)
    o.noise_philox(4)
    th, x, y = poses[3]
    m.setWaypoint(pkg.Pose(theta=0.0, x=2.2, y=0.2))        # towards the box and the wall: the term is active
    o.setWaypoint(2.2, 0.2, 0.0)
    v = m.newControls(pkg.Pose(theta=th, x=x, y=y))
    c = o.newControls(x, y, th)
    g = o.get()
    J = m.costToGo()
    assert np.max(np.abs(J - g["J"].T) / np.maximum(np.abs(g["J"].T), 1e-6)) < 1e-11
    assert max(abs(v.ul - c[0]), abs(v.ur - c[1])) / max(1e-3, abs(c[0]), abs(c[1])) < 1e-5
    # and the term really bites: the same call without it costs less somewhere
    o2 = orc.OracleMppi(hor, dt, K)
    o2.noise_philox(4)
    o2.setWaypoint(2.2, 0.2, 0.0)
    o2.newControls(x, y, th)
    assert np.max(g["J"] - o2.get()["J"]) > 1.0
