"""Tuning/debug: the in-process sharded MPPI path (ranks = handles sharing one GPU), prints per-call agreement."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkg  # noqa: E402

pkg = _pkg.load()
prm = pkg.synthetic.SHIPPED
nranks, K, hor = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
Kr = K // nranks
ranks = [pkg.MPPI(pkg.CartModel(prm["wheel_radius"], prm["wheel_base"]), pkg.LossFunc(prm["Q"], prm["R"], prm["P1"]),
                  prm["lambda_"], prm["max_wheel_vel"], prm["ul_var"], prm["ur_var"], hor, 0.01, Kr,
                  rollout_offset=r * Kr, rollouts_total=K) for r in range(nranks)]
for m in ranks:
    m.p2pExport(nranks)
areas = [m.p2pArea() for m in ranks]
for r, m in enumerate(ranks):
    m.p2pInitLocal(r, nranks, areas)
    m.seed(42)
    m.setWaypoint(pkg.Pose(theta=1.5707, x=1.0, y=0.0))
P = pkg.Pose(theta=0.0, x=0.0, y=0.0)
for c in range(4):
    t0 = time.time()
    for m in ranks:
        m.enqueue(P)
    vs = [m.wait() for m in ranks]
    print("call", c, "%.3f s" % (time.time() - t0), [(v.ul, v.ur) for v in vs][:2], flush=True)
print("ok")
