"""BASELINE configs[4] in miniature: RBPF pose estimate -> MPPI wheel command -> plant -> odometry + lidar scan, tick by
tick, the GPU stack and the CPU oracle stack fed the same scans and odometry (mppi_waypoints_node.cpp:231-282 and
turtle_mapping_node.cpp:451-494 fused into one loop; the plant is the reference's rigid2d::DiffDrive, compiled in
oracle/_ref, driven the way fake_diff_encoders drives it; the lidar is the analytic ray cast of the synthetic room).

Every tick: the filter's best pose within 1e-9 of the oracle's, the controller's command within 1e-5, ancestors equal.
"""
import numpy as np
import pytest

import _oracle as orc

pytestmark = pytest.mark.gpu


def test_rbpf_feeds_mppi_closed_loop(gpu_pkg):
    """The loop of the reference's nodes with the reference's OWN plant: the simulated robot is rigid2d::DiffDrive::feedforward
    on the commanded twist scaled by 1 / frequency (fake_diff_encoders_node.cpp:100-135), its encoder angles feed two odometers
    (DiffDrive::updateOdometry: the controller's pose source and the mapper's pf_drive, turtle_mapping_node.cpp:456-472), the
    wheel command becomes a twist through DiffDrive::wheelsToTwist (mppi_waypoints_node.cpp:276), waypoints switch as in
    mppi_waypoints_node.cpp:231-258 - all through the compiled reference (oracle/_ref)."""
    pkg = gpu_pkg
    syn = pkg.synthetic
    N, K, hor, dt, ticks = 48, 512, 0.5, 0.02, 12
    scan_every = 2                                     # a lidar scan every other control tick
    freq = 1.0 / dt
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
    # waypoints: a short leg first so that the switch fires inside the test
    sw = syn.WaypointSwitch([0.62, 1.6, 1.6], [0.0, 0.3, 1.0], [0.0, 0.0, 1.5707], 0.03)
    wx, wy, wth = sw.current()
    m.setWaypoint(pkg.Pose(theta=wth, x=wx, y=wy))
    om.setWaypoint(wx, wy, wth)

    Plant = orc.RefDiffDrive if orc.have_ref() else syn.DiffDrive
    robot = Plant(start, prm["wheel_base"], prm["wheel_radius"])       # the simulated robot (fake_diff_encoders)
    odo = Plant(start, prm["wheel_base"], prm["wheel_radius"])         # the controller's odometer
    pf_drive = Plant(start, prm["wheel_base"], prm["wheel_radius"])    # the mapper's odometer
    prev_odom = start
    est = start
    resampled = switched = 0
    for k in range(ticks):
        # waypoint bookkeeping, then the controller on the current pose estimate (mppi_waypoints_node.cpp:231-265)
        nw = sw.update(est[1], est[2])
        if nw is not None:
            switched += 1
            m.setWaypoint(pkg.Pose(theta=nw[2], x=nw[0], y=nw[1]))
            om.setWaypoint(nw[0], nw[1], nw[2])
        v = m.newControls(pkg.Pose(theta=est[0], x=est[1], y=est[2]))
        c = om.newControls(est[1], est[2], est[0])
        assert max(abs(v.ul - c[0]), abs(v.ur - c[1])) / max(1e-3, abs(c[0]), abs(c[1])) < 1e-5, (k, v, c)
        # the wheel command as a body twist (:276), the simulated robot one period further, the odometers on its encoders
        w, vx, _ = robot.wheelsToTwist(c[0], c[1])
        robot.feedforward(w / freq, vx / freq)
        left, right = robot.getEncoders()
        odo.updateOdometry(left, right)
        est = odo.pose()
        if (k + 1) % scan_every:
            continue
        # the mapper: odometry since the last scan, the twist of the wheel velocities, SLAM (turtle_mapping_node.cpp:466-494)
        pf_drive.updateOdometry(left, right)
        cur_odom = pf_drive.pose()
        twist = pf_drive.wheelsToTwist(*pf_drive.wheelVelocities())
        scan = orc.room_scan(robot.pose(), rng=rng)
        f.SLAM(scan, pkg.Twist2D(*twist), pkg.Pose(*cur_odom), pkg.Pose(*prev_odom))
        of.slam(scan, twist, cur_odom, prev_odom)
        prev_odom = cur_odom
        T = f.getRobotState().displacement()
        want = of.robot_state()
        assert np.max(np.abs(np.array(T) - want)) < 1e-9, (k, T, want)
        neff, rs, anc = f.resampleInfo()
        oneff, ors, oanc = of.resample_info()
        assert (neff, rs) == (oneff, ors) and np.array_equal(anc, oanc)
        resampled += rs
        assert np.array_equal(f.newMap(), of.new_map())
        est = tuple(want)                               # the filter's pose replaces the odometer's until the next tick
    assert resampled >= 1 and switched >= 1


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
