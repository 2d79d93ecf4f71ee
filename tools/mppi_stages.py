"""Tuning: where one MPPI call spends its time (globaltimer stamps of every CTA's thread 0, B2N_MPPI_DEBUG_TIMES=1).
  python tools/mppi_stages.py [K] [horizon]
Stamps are microseconds from the first CTA's first instruction.  Rollout CTAs: entry, loop start, loop end (the stamp sits
behind a barrier that defers blocking: it is when warp 0 arrived), sets merged, partial words sent, landed (read back from
L2).  Merger CTAs: resident, partials in, merged, (sharded: received), updated."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["B2N_MPPI_DEBUG_TIMES"] = "1"
import _pkg  # noqa: E402

pkg = _pkg.load()
prm = pkg.synthetic.SHIPPED
K = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
hor = float(sys.argv[2]) if len(sys.argv) > 2 else 0.64
m = pkg.MPPI(pkg.CartModel(prm["wheel_radius"], prm["wheel_base"]), pkg.LossFunc(prm["Q"], prm["R"], prm["P1"]),
             prm["lambda_"], prm["max_wheel_vel"], prm["ul_var"], prm["ur_var"], hor, 0.01, K)
m.setStateRing(16)
m.seed(42)
m.setWaypoint(pkg.Pose(theta=1.5707, x=1.0, y=0.0))
pose = pkg.Pose(theta=0.0, x=0.0, y=0.0)
ROLL = [(13, "entry"), (0, "loop start"), (1, "loop end"), (3, "sets merged"), (2, "partial sent"), (7, "landed + read back")]
MERG = [(4, "resident"), (3, "partials in"), (5, "merged"), (6, "updated")]


def show(tag):
    d = m.debugTimes().astype(np.int64)
    T = m.steps
    r, g = d[:-T], d[-T:]
    e0 = r[:, 13].min()
    f = lambda c: "%6.2f .. %6.2f .. %6.2f" % ((c.min() - e0) / 1e3, (np.median(c) - e0) / 1e3, (c.max() - e0) / 1e3)
    print("%s: %d rollout CTAs + %d merger CTAs (min .. median .. max, us)" % (tag, r.shape[0], T))
    for j, n in ROLL:
        print("  rollout CTAs  %-20s %s" % (n, f(r[:, j])))
    for j, n in MERG:
        print("  merger CTAs   %-20s %s" % (n, f(g[:, j])))
    print("  previous call's last update -> this call's first loop start %.2f us; period (loop start to loop start, median) %.2f us"
          % ((r[:, 0].min() - g[:, 12].max()) / 1e3, np.median(r[:, 0] - r[:, 11]) / 1e3))


for rep in range(3):
    for _ in range(10):
        m.newControls(pose)
    show("synchronous calls, run %d" % rep)
for rep in range(3):
    for _ in range(100):
        m.enqueue(pose)
    m.wait()
    show("queued calls, run %d" % rep)
