"""Sharded MPPI / RBPF on two B200s (NCCL) against the unsharded CPU oracle.  Skipped on a one-GPU box."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_two_gpus_match_the_unsharded_oracle(gpu_pkg):
    if gpu_pkg.load_library().b2n_device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(HERE, "_mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("MGPU_RESULT ")][-1]
    for res in json.loads(line[len("MGPU_RESULT "):]):
        for mode in ("mppi_nccl", "mppi_p2p"):
            assert res[mode]["controls_rel_err"] < 1e-5 and res[mode]["plan_rel_err"] < 1e-5, (mode, res[mode])
            assert res[mode]["plan_replicated_bitwise"], mode       # identical update on every rank, no broadcast
        for mode in ("rbpf_nccl", "rbpf_p2p"):
            rb = res[mode]
            assert rb["ancestors_equal"] and rb["map_equal"], mode      # resampling indices and maps bit-exact
            assert rb["weights"] < 1e-9 and rb["poses"] < 1e-9, mode
            assert rb["resampled"] >= 1
        rb = res["rbpf_p2p"]
        assert rb["best_pose"] < 1e-9 and rb["best_map_equal"]            # global argmax, read from the owning GPU
        assert res["rbpf_nccl"]["best_unsupported_without_peer_memory"]
    for mode in ("rbpf_nccl", "rbpf_p2p"):
        assert sum(res[mode]["migrated"] for res in json.loads(line[len("MGPU_RESULT "):])) >= 1   # particles really changed GPU
