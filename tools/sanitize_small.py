"""Small invocations of every kernel of libb2nav, meant to be run under compute-sanitizer
(memcheck / racecheck / synccheck / initcheck):  compute-sanitizer --tool memcheck python tools/sanitize_small.py
Sizes are tiny so that the sanitizer's slow-down stays within a minute or two."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _pkg  # noqa: E402

pkg = _pkg.load()
orc = pkg.synthetic
which = sys.argv[1] if len(sys.argv) > 1 else "all"

if which in ("all", "mppi"):
    prm = orc.SHIPPED
    for K, horizon in ((300, 0.16), (1024, 0.64)):
        m = pkg.MPPI(pkg.CartModel(prm["wheel_radius"], prm["wheel_base"]), pkg.LossFunc(prm["Q"], prm["R"], prm["P1"]),
                     prm["lambda_"], prm["max_wheel_vel"], prm["ul_var"], prm["ur_var"], horizon, 0.01, K)
        m.seed(7)
        m.setWaypoint(pkg.Pose(theta=1.5707, x=1.0, y=0.0))
        pose = pkg.Pose(theta=0.0, x=0.0, y=0.0)
        for _ in range(3):
            v = m.newControls(pose)
        print("mppi K=%d ok" % K, v.ul, v.ur, flush=True)

if which in ("all", "rbpf"):
    poses, twists = orc.circle_path(4)
    rng = np.random.default_rng(2)
    q = orc.pf_params(num_particles=8, init_pose=tuple(poses[0]), motion_noise=(2e-3, 1e-3, 1e-3), xmin=-2.0, xmax=2.0, ymin=-2.0, ymax=2.0)
    f = pkg.bmapping.make_filter(q)
    f.seed(3)
    for i in range(4):
        scan = np.minimum(orc.room_scan(poses[i + 1], rng=rng), 1.6).astype(np.float32)
        if i >= 2:
            w, d = twists[i][0], twists[i][1]
            f.scan_matcher.setResult(True, (w, d * np.cos(w / 2), d * np.sin(w / 2)))
        f.SLAM(scan, pkg.Twist2D(*twists[i]), pkg.Pose(*poses[i + 1]), pkg.Pose(*poses[i]))
    print("rbpf ok", f.getRobotState().displacement(), int(np.sum(f.newMap() > 50)), flush=True)

if which in ("all", "icp"):
    poses, _ = orc.circle_path(3)
    g = pkg.bmapping.GpuScanAlignment(pkg.bmapping.make_filter(orc.pf_params(num_particles=2)).scan_matcher.props, None)
    for i in range(3):
        r = g.pclICPWrapper((0.0, 0.0, 0.0), orc.room_scan(poses[i]))
    print("icp ok", r, g.stats(), flush=True)
