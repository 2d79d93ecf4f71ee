"""A handful of MPPI calls at BASELINE configs[1] (for ncu captures)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _pkg  # noqa: E402
pkg = _pkg.load()
orc = pkg.synthetic          # shipped parameters and synthetic inputs (plain numpy)
prm = orc.SHIPPED
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
K = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
m = pkg.MPPI(pkg.CartModel(prm["wheel_radius"], prm["wheel_base"]), pkg.LossFunc(prm["Q"], prm["R"], prm["P1"]),
             prm["lambda_"], prm["max_wheel_vel"], prm["ul_var"], prm["ur_var"], 0.64, 0.01, K)
m.setStateRing(16)
m.seed(42)
m.setWaypoint(pkg.Pose(theta=1.5707, x=1.0, y=0.0))
for _ in range(n):
    m.newControls(pkg.Pose(theta=0.0, x=0.0, y=0.0))
