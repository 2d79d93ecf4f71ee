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
