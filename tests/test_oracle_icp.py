"""The scan-matcher checker (oracle/icp_oracle.cpp).  The reference's matcher is PCL's IterativeClosestPoint
(cloud_alignment.cpp:160-223), which is not vendored: parity with PCL is UNPINNED.  What is checked here is that the
restated algorithm does its job - it recovers the robot's motion between two synthetic scans of the same room - and the
wrapper semantics of pclICPWrapper (cloud_alignment.cpp:37-72)."""
import numpy as np

import _oracle as O


def _relative(p0, p1):
    """pose of frame p1 seen from frame p0: (theta, x, y)"""
    dth = p1[0] - p0[0]
    dx, dy = p1[1] - p0[1], p1[2] - p0[2]
    c, s = np.cos(p0[0]), np.sin(p0[0])
    return dth, c * dx + s * dy, -s * dx + c * dy


def test_first_call_stores_the_scan_and_leaves_T():
    icp = O.OracleIcp()
    poses, _ = O.circle_path(2)
    ok, T = icp.pclICPWrapper((0.3, 0.2, 0.1), O.room_scan(poses[0]))
    assert ok and T == (0.0, 0.0, 0.0)
    assert icp.stats()[0] == 0


def test_recovers_motion_between_scans():
    icp = O.OracleIcp()
    poses, _ = O.circle_path(12)
    icp.pclICPWrapper(None, O.room_scan(poses[0]))
    for i in range(1, 12):
        ok, T = icp.pclICPWrapper((0.0, 0.0, 0.0), O.room_scan(poses[i]))
        rel = _relative(poses[i - 1], poses[i])
        it, pairs, mse = icp.stats()
        assert ok and 1 <= it < 100 and pairs > 300
        assert abs(T[0] - rel[0]) < 0.02 and abs(T[1] - rel[1]) < 0.02 and abs(T[2] - rel[2]) < 0.02, (i, T, rel)


def test_identical_scans_give_identity():
    icp = O.OracleIcp()
    poses, _ = O.circle_path(1)
    s = O.room_scan(poses[0])
    icp.pclICPWrapper(None, s)
    ok, T = icp.pclICPWrapper((0.0, 0.0, 0.0), s)
    assert ok and np.allclose(T, 0.0, atol=1e-12)


def test_too_few_returns_is_a_failure_and_keeps_the_old_scan():
    icp = O.OracleIcp()
    poses, _ = O.circle_path(2)
    icp.pclICPWrapper(None, O.room_scan(poses[0]))
    empty = np.zeros(360, np.float32)                      # every beam outside the range gate
    ok, T = icp.pclICPWrapper((0.0, 0.0, 0.0), empty)
    assert not ok and T == (0.0, 0.0, 0.0)
    ok, T = icp.pclICPWrapper((0.0, 0.0, 0.0), O.room_scan(poses[1]))     # still aligned against scan 0
    rel = _relative(poses[0], poses[1])
    assert ok and abs(T[1] - rel[1]) < 0.02 and abs(T[2] - rel[2]) < 0.02
