"""Per-call time of the MPPI path as the receding-horizon plan evolves (chunks of 100 back-to-back calls)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _pkg  # noqa: E402
pkg = _pkg.load()
orc = pkg.synthetic          # shipped parameters and synthetic inputs (plain numpy)
prm = orc.SHIPPED
m = pkg.MPPI(pkg.CartModel(prm["wheel_radius"], prm["wheel_base"]), pkg.LossFunc(prm["Q"], prm["R"], prm["P1"]), prm["lambda_"],
             prm["max_wheel_vel"], prm["ul_var"], prm["ur_var"], 0.64, 0.01, 16384)
m.setStateRing(16); m.seed(42)
m.setWaypoint(pkg.Pose(theta=1.5707, x=1.0, y=0.0))
pose = pkg.Pose(theta=0.0, x=0.0, y=0.0)
for chunk in range(30):
    t0 = time.perf_counter()
    for _ in range(100):
        m.enqueue(pose)
    v = m.wait()
    el = (time.perf_counter() - t0) / 100
    k = m.timeRollout(pose, 50)
    p = m.plan()
    print("calls %4d-%4d: %.2f us/call, rollout kernel %.2f us, controls (%.3f, %.3f), plan[0][:3] %s" % (chunk * 100, chunk * 100 + 99, el * 1e6, k * 1e3, v.ul, v.ur, p[0][:3]), flush=True)
