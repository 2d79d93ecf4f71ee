import os, sys
import numpy as np
ROOT = "/root/repo"
sys.path.insert(0, ROOT)
os.environ["B2N_MPPI_DEBUG_TIMES"] = "1"
import _pkg
pkg = _pkg.load()
prm = pkg.synthetic.SHIPPED
m = pkg.MPPI(pkg.CartModel(prm["wheel_radius"], prm["wheel_base"]), pkg.LossFunc(prm["Q"], prm["R"], prm["P1"]),
             prm["lambda_"], prm["max_wheel_vel"], prm["ul_var"], prm["ur_var"], 0.64, 0.01, 16384)
m.setStateRing(16); m.seed(42); m.setWaypoint(pkg.Pose(theta=1.5707, x=1.0, y=0.0))
pose = pkg.Pose(theta=0.0, x=0.0, y=0.0)
out = {}
for rep in range(3):
    for _ in range(10): m.newControls(pose)
    out["sync%d" % rep] = m.debugTimes().astype(np.int64)
for rep in range(3):
    for _ in range(100): m.enqueue(pose)
    m.wait()
    out["pipe%d" % rep] = m.debugTimes().astype(np.int64)
np.savez("gpurun_out/r04c_stages.npz", **out)
