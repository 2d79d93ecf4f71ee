"""GPU scan matcher (b2n_icp_*) against oracle/icp_oracle.cpp on the same scans.  The kernel forms the same ordered
fp64 sums as the checker, so the only differences are the device's atan2/sin/cos (<= 2 ulp): transforms agree to 1e-9,
iteration and correspondence counts exactly.  Parity with PCL itself is unpinned (see the oracle's header)."""
import numpy as np
import pytest

import _oracle as O
import _pkg

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    return _pkg.load()


def _matchers(pkg, **kw):
    q = O.pf_params()
    props = pkg.bmapping.LaserProperties(q["beam_min"], q["beam_max"], q["beam_delta"], q["range_min"], q["range_max"],
                                         q["z_hit"], q["z_short"], q["z_max"], q["z_rand"], q["sigma_hit"])
    return pkg.bmapping.GpuScanAlignment(props, None, **kw), O.OracleIcp(**{{"max_correspondence_dist": "max_corr_dist"}.get(k, k): v for k, v in kw.items()})


def test_matches_the_checker_along_a_path(pkg):
    g, o = _matchers(pkg)
    poses, _ = O.circle_path(20)
    rng = np.random.default_rng(3)
    for i in range(20):
        scan = O.room_scan(poses[i], rng=rng)
        guess = (0.0, 0.0, 0.0) if i % 2 else (0.004, 0.08, 0.03)
        okg, Tg = g.pclICPWrapper(guess, scan)
        oko, To = o.pclICPWrapper(guess, scan)
        assert okg == oko
        np.testing.assert_allclose(Tg, To, rtol=0, atol=1e-9)
        assert g.stats()[:2] == o.stats()[:2]
        assert abs(g.stats()[2] - o.stats()[2]) <= 1e-12
    assert g.stats()[3] == 19                                 # one kernel launch per aligned scan, none for the first


def test_failure_keeps_previous_scan(pkg):
    g, o = _matchers(pkg)
    poses, _ = O.circle_path(3)
    for m in (g, o):
        m.pclICPWrapper(None, O.room_scan(poses[0]))
    empty = np.zeros(360, np.float32)
    assert g.pclICPWrapper((0, 0, 0), empty)[0] is False and o.pclICPWrapper((0, 0, 0), empty)[0] is False
    far = O.room_scan(poses[0]) * 0 + 3.4                       # a circle of returns far from every stored point
    rg, ro = g.pclICPWrapper((0, 0, 0), far), o.pclICPWrapper((0, 0, 0), far)
    assert rg[0] == ro[0]
    rg, ro = g.pclICPWrapper((0, 0, 0), O.room_scan(poses[1])), o.pclICPWrapper((0, 0, 0), O.room_scan(poses[1]))
    assert rg[0] and ro[0]
    np.testing.assert_allclose(rg[1], ro[1], rtol=0, atol=1e-9)


def test_iteration_cap_and_tight_gate(pkg):
    g, o = _matchers(pkg, max_iter=3, max_correspondence_dist=0.05)
    poses, _ = O.circle_path(4)
    for i in range(4):
        s = O.room_scan(poses[i])
        rg, ro = g.pclICPWrapper((0, 0, 0), s), o.pclICPWrapper((0, 0, 0), s)
        assert rg[0] == ro[0]
        np.testing.assert_allclose(rg[1], ro[1], rtol=0, atol=1e-9)
        assert g.stats()[:2] == o.stats()[:2]


def test_filter_with_gpu_matcher_tracks_the_path(pkg):
    """ParticleFilter.SLAM with the GPU matcher in the scan_matcher slot: the improved-proposal branch runs from the
    second scan on and the estimate follows the true path."""
    poses, twists = O.circle_path(8)
    f = pkg.bmapping.make_filter(O.pf_params(num_particles=64, init_pose=tuple(poses[0])))
    f.scan_matcher = pkg.bmapping.GpuScanAlignment(f.scan_matcher.props, None)
    f.seed(5)
    for i in range(8):
        f.SLAM(O.room_scan(poses[i + 1]), pkg.Twist2D(*twists[i]), pkg.Pose(*poses[i + 1]), pkg.Pose(*poses[i]))
    th, x, y = f.getRobotState().displacement()
    assert abs(x - poses[8][1]) < 0.15 and abs(y - poses[8][2]) < 0.15
    assert f.scan_matcher.stats()[3] == 7
