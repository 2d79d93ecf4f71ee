"""Tuning: where one MPPI call spends its time (stage stamps of every CTA, B2N_MPPI_DEBUG_TIMES=1)."""
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
names = ["loop start", "loop end", "partial written", "merger: partials in", "merger: resident", "merger: merged", "merger: updated", "-"]
for rep in range(6):
    for _ in range(10):
        m.newControls(pose)
    d = m.debugTimes().astype(np.int64)
    t0 = d[:-m.steps, 0].min()
    print("call %d: grid %d" % (rep, d.shape[0]))
    for j, n in enumerate(names):
        col = d[:, j]
        ok = col >= t0
        col = col[ok] - t0
        if len(col):
            print("  %-16s n=%4d  min %7.2f  median %7.2f  max %7.2f us" % (n, len(col), col.min() / 1e3, np.median(col) / 1e3, col.max() / 1e3))
    # per-CTA stage durations
    for i, j in ((0, 1), (1, 2), (4, 3), (3, 5), (5, 6)):
        ok = (d[:, j] >= t0) & (d[:, i] >= t0)
        if ok.any():
            dd = (d[ok, j] - d[ok, i]) / 1e3
            print("  %-16s -> %-16s n=%4d  min %6.2f median %6.2f max %6.2f us" % (names[i], names[j], ok.sum(), dd.min(), np.median(dd), dd.max()))
    ph = ["z load", "D sums", "(decl)", "local integration", "rot scan", "pos scan", "stage wait", "loss + stores", "cost scan", "softmax", "bulk store"]
    c = d[:-m.steps, 8:19]
    dc = np.diff(c, axis=1)
    print("  first pass of warp 0, SM cycles per phase (median over CTAs | CTA 0):")
    for j, n in enumerate(ph[1:] + ["?"]):
        if j < dc.shape[1]:
            print("    %-20s %8.0f | %8d" % (ph[j] if j == 0 else ph[j], np.median(dc[:, j]), dc[0, j]))
    print("    total pass           %8.0f" % np.median(c[:, -1] - c[:, 0]))
