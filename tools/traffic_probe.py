"""Steady-state DRAM traffic of the MPPI rollout kernel: 16 consecutive launches writing round-robin through the 16-buffer
state ring, bracketed by cudaProfilerStart / cudaProfilerStop so that `ncu --replay-mode range` measures the WHOLE range as
one unit (a single profiled launch shows no DRAM writes: its 12.6 MB state tensor sits in the write-back L2 until later
launches evict it).  Driven by tools/measure_traffic.sh; not the bench."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["B2N_MPPI_PDL"] = "0"          # plain stream order inside the replayed range
import torch  # noqa: E402
import _pkg  # noqa: E402

pkg = _pkg.load()
prm = pkg.synthetic.SHIPPED
LAUNCHES = 16
m = pkg.MPPI(pkg.CartModel(prm["wheel_radius"], prm["wheel_base"]), pkg.LossFunc(prm["Q"], prm["R"], prm["P1"]),
             prm["lambda_"], prm["max_wheel_vel"], prm["ul_var"], prm["ur_var"], 0.64, 0.01, 16384)
m.setStateRing(16)
m.seed(42)
m.setWaypoint(pkg.Pose(theta=1.5707, x=1.0, y=0.0))
pose = pkg.Pose(theta=0.0, x=0.0, y=0.0)
for _ in range(8):
    m.newControls(pose)
m.timeRollout(pose, 64)                   # the ring is dirty in L2 / HBM as in the bench's steady state
torch.cuda.synchronize()
torch.cuda.profiler.start()
m.timeRollout(pose, LAUNCHES)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("launches in range:", LAUNCHES)
