"""Tuning: host-clock time of a burst of n queued MPPI calls (enqueue x n, then wait) against n: the pipeline's fill and
drain next to its steady-state period.  python tools/mppi_burst.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkg  # noqa: E402

pkg = _pkg.load()
prm = pkg.synthetic.SHIPPED
m = pkg.MPPI(pkg.CartModel(prm["wheel_radius"], prm["wheel_base"]), pkg.LossFunc(prm["Q"], prm["R"], prm["P1"]),
             prm["lambda_"], prm["max_wheel_vel"], prm["ul_var"], prm["ur_var"], 0.64, 0.01, 16384)
m.setStateRing(16)
m.seed(42)
m.setWaypoint(pkg.Pose(theta=1.5707, x=1.0, y=0.0))
pose = pkg.Pose(theta=0.0, x=0.0, y=0.0)
for _ in range(3000):
    m.enqueue(pose)
m.wait()
prev = None
for n in (1, 2, 3, 4, 5, 6, 8, 10, 15, 20, 40, 80, 200):
    ts = []
    for rep in range(200):
        time.sleep(0.0002)
        t0 = time.perf_counter()
        for _ in range(n):
            m.enqueue(pose)
        m.wait()
        ts.append(time.perf_counter() - t0)
    t = float(np.median(ts)) * 1e6
    print("burst of %3d calls: %7.2f us (median of 200), %6.2f us per call%s" % (n, t, t / n, "" if prev is None else ", marginal %.2f us per call" % ((t - prev[1]) / (n - prev[0]))), flush=True)
    prev = (n, t)
