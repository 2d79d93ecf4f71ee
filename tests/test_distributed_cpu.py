"""The N > 1 paths on CPU: two processes, torch.distributed gloo on 127.0.0.1.

MPPI: rollouts sharded over ranks, ONE all_gather of the [T][6] partial per call, identical update on every rank - must
give the controls and plan of the unsharded oracle.  RBPF: weights all_gather, identical walk, particle migration along
the plan computed by the product's host code (b2n_pf_plan_migration) - every slot must end up with its ancestor.
"""
import json
import os
import socket
import subprocess
import sys

import numpy as np

import _oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return str(p)


def _run(case, tmp_path, world=2):
    port = _free_port()
    outs = [str(tmp_path / ("%s_%d.json" % (case, r))) for r in range(world)]
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "_dist_worker.py"), case, str(r), str(world), port, outs[r]],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    for p in procs:
        try:
            log, _ = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        assert p.returncode == 0, log
    return [json.load(open(o)) for o in outs]


def test_mppi_sharded_rollouts_one_allgather(tmp_path, pkg):
    res = _run("mppi", tmp_path)
    assert res[0] == res[1]                                   # every rank holds the same controls and plan, no broadcast
    # the unsharded oracle on the same global noise stream
    K, hor, dt = 512, 0.5, 0.02
    o = orc.OracleMppi(hor, dt, K, **orc.MILD)
    o.noise_philox(7)
    o.setWaypoint(1.0, 0.0, 1.5707)
    pose = (0.0, 0.0, 0.0)
    for c, got in enumerate(res[0]["controls"]):
        want = o.newControls(*pose)
        assert np.allclose(got, want, rtol=1e-9, atol=1e-12), (c, got, want)
        pose = orc.unicycle_step(pose, want[0], want[1], dt)
    assert np.allclose(res[0]["plan"], o.get()["plan"], rtol=1e-9, atol=1e-12)


def test_rbpf_sharded_resampling_migrates_particles(tmp_path, pkg):
    res = _run("rbpf", tmp_path)
    assert res[0]["ancestors"] == res[1]["ancestors"]         # identical walk on every rank
    for r in res:
        assert r["ok"]                                        # every slot holds its ancestor's payload
        assert r["n_recv"] == r["unique_remote"]              # a migrating particle travels once per destination rank
    assert res[0]["n_send"] == res[1]["n_recv"] and res[1]["n_send"] == res[0]["n_recv"]
    assert res[0]["n_recv"] + res[1]["n_recv"] > 0            # the case really crosses ranks


def test_migration_plan_single_rank_is_a_plain_copy(pkg):
    import ctypes as C
    lib = pkg.load_library()
    anc = np.array([0, 0, 2, 5, 5, 5, 6, 7], dtype=np.int32)
    c1, c2 = np.zeros(8, np.int32), np.zeros(8, np.int32)
    nr, ns = C.c_int(), C.c_int()
    pkg._capi.check(lib.b2n_pf_plan_migration(pkg._capi.as_ptr(anc), 8, 0, 1, pkg._capi.as_ptr(c1), pkg._capi.as_ptr(c2), None, 0,
                                              C.byref(nr), None, 0, C.byref(ns)))
    assert np.array_equal(c1, anc) and np.all(c2 == -1) and nr.value == 0 and ns.value == 0
