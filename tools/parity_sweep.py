"""Extended randomized parity sweep on the GPU: the random-problem tests of tests/test_mppi_gpu.py and
tests/test_rbpf_gpu.py with many more seeds than the suite runs (50 MPPI problems, 30 RBPF scan sequences by default).

    python tools/parity_sweep.py [first_seed] [n_mppi] [n_rbpf]
"""
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _pkg  # noqa: E402
import test_mppi_gpu as tm  # noqa: E402
import test_rbpf_gpu as tr  # noqa: E402

first = int(sys.argv[1]) if len(sys.argv) > 1 else 10
n_mppi = int(sys.argv[2]) if len(sys.argv) > 2 else 50
n_rbpf = int(sys.argv[3]) if len(sys.argv) > 3 else 30
pkg = _pkg.load()
bad = 0
for seed in range(first, first + n_mppi):
    try:
        tm.test_random_problems_match_oracle(pkg, seed)
    except Exception:
        bad += 1
        print("MPPI seed", seed, "FAILED")
        traceback.print_exc(limit=2)
for seed in range(first, first + n_rbpf):
    try:
        tr.test_random_scans_exercise_every_ray_direction(pkg, seed)
    except Exception:
        bad += 1
        print("RBPF seed", seed, "FAILED")
        traceback.print_exc(limit=2)
print("sweep done: %d MPPI problems, %d RBPF sequences, failures: %d" % (n_mppi, n_rbpf, bad))
sys.exit(1 if bad else 0)
