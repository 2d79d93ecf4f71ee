"""Quick timing probe of the RBPF path on one GPU (not the bench): per-kernel CUDA-event times at a given size."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _pkg  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
scans = int(sys.argv[2]) if len(sys.argv) > 2 else 6
hcap = int(sys.argv[3]) if len(sys.argv) > 3 else 0
icp = int(sys.argv[4]) if len(sys.argv) > 4 else 0     # 1: improved-proposal branch from the second scan on (Ticp = odometry increment)
pkg = _pkg.load()
orc = pkg.synthetic          # shipped parameters and synthetic inputs (plain numpy)
poses, twists = orc.circle_path(scans)
rng = np.random.default_rng(4)
f = pkg.bmapping.make_filter(orc.pf_params(num_particles=N, init_pose=tuple(poses[0]), motion_noise=(2e-3, 1e-3, 1e-3)))
if hcap:
    f.setHeapCapacity(hcap)
f.seed(1)
f.setKernelTiming(True)
if icp == 2:                                   # 2: the GPU scan matcher in the scan_matcher slot
    f.scan_matcher = pkg.bmapping.GpuScanAlignment(f.scan_matcher.props, None)
for i in range(scans):
    scan = orc.room_scan(poses[i + 1], rng=rng)
    if icp == 1 and i > 0:
        w, d = twists[i][0], twists[i][1]
        f.scan_matcher.setResult(True, (w, d * np.cos(w / 2), d * np.sin(w / 2)))
    if icp == 2:                               # time a second matcher on the same scan sequence
        m = pkg.bmapping.GpuScanAlignment(f.scan_matcher.props, None) if i == 0 else m
        t1 = time.perf_counter()
        m.pclICPWrapper((0.0, 0.0, 0.0), scan)
        print("   matcher: %.1f us for pclICPWrapper (host wall), (iterations, pairs, mse, launches) = %s" % ((time.perf_counter() - t1) * 1e6, m.stats()))
    t0 = time.perf_counter()
    f.SLAM(scan, pkg.Twist2D(*twists[i]), pkg.Pose(*poses[i + 1]), pkg.Pose(*poses[i]))
    wall = (time.perf_counter() - t0) * 1e3
    ms = f.kernelTimes()
    neff, rs, _ = f.resampleInfo()
    print("scan %d: wall %.2f ms | update %.3f ms, distance field %.3f ms, normalise+resample %.3f ms | N_eff %d resampled %d | %.0f particle-updates/s"
          % (i, wall, ms[0], ms[1], ms[2], neff, rs, N / (wall * 1e-3)))
print("distance-field stats (iterations, heap max):", f.distanceFieldStats(), "particle-scans skipped (occupied set unchanged):", f.distanceFieldSkipped())
