"""Quick timing probe of the MPPI path on one GPU (not the bench): per-call and rollout-kernel times for the
rollout-kernel shapes (S steps per lane, G lanes per rollout) that fit the horizon."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _pkg  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
hor = float(sys.argv[2]) if len(sys.argv) > 2 else 0.64
dt = float(sys.argv[3]) if len(sys.argv) > 3 else 0.01
params = sys.argv[4] if len(sys.argv) > 4 else "shipped"
pkg = _pkg.load()
orc = pkg.synthetic          # shipped parameters and synthetic inputs (plain numpy)
prm = orc.SHIPPED if params == "shipped" else orc.MILD
shapes = [(2, 8, 8), (4, 8, 8), (2, 16, 8), (4, 16, 8), (2, 32, 8), (4, 32, 8), (8, 32, 8), (4, 16, 10), (4, 16, 12), (4, 16, 20), (4, 32, 10), (4, 32, 20)]
if os.environ.get("PROBE_SHAPES"):
    shapes = [tuple(int(v) for v in x.split(",")) for x in os.environ["PROBE_SHAPES"].split(";")]
for shape in shapes:
    S, G = shape[0], shape[1]
    NW = shape[2] if len(shape) > 2 else 8
    os.environ["B2N_MPPI_SHAPE"] = "%d,%d,%d" % (S, G, NW)
    m = pkg.MPPI(pkg.CartModel(prm["wheel_radius"], prm["wheel_base"]), pkg.LossFunc(prm["Q"], prm["R"], prm["P1"]),
                 prm["lambda_"], prm["max_wheel_vel"], prm["ul_var"], prm["ur_var"], hor, dt, K)
    T = m.steps
    if S * G < T:
        continue
    m.setStateRing(16)
    m.seed(42)
    m.setWaypoint(pkg.Pose(theta=1.5707, x=1.0, y=0.0))
    pose = pkg.Pose(theta=0.0, x=0.0, y=0.0)
    for _ in range(20):
        m.newControls(pose)
    n = 1000
    t0 = time.perf_counter()
    for _ in range(n):
        m.enqueue(pose)
    m.wait()
    el = (time.perf_counter() - t0) / n
    t0 = time.perf_counter()
    for _ in range(n):
        m.newControls(pose)
    sync = (time.perf_counter() - t0) / n
    kms = m.timeRollout(pose, n)
    print("K=%d T=%d %s shape S=%d G=%d NW=%d: %.2f us/call pipelined (%.3e traj-steps/s), %.2f us/call synchronous, rollout phase %.2f us (back to back), variant %s"
          % (K, T, params, S, G, NW, el * 1e6, K * T / el, sync * 1e6, kms * 1e3, m.lastVariant()), flush=True)
    del m
