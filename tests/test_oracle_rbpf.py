"""CPU tests (no GPU): pin the RBPF oracle (oracle/rbpf_oracle.cpp) bit-exactly against (1) outputs of the
UNMODIFIED reference committed under tests/golden/rbpf_*.npz, (2) the compiled reference itself when
oracle/_ref exists, (3) the real libstdc++ containers whose behaviour it restates, (4) the survey's
hand-derived known answers and the reference's rigid2d gtest vectors."""
import os

import numpy as np
import pytest

import _oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
needs_ref = pytest.mark.skipif(not orc.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
SMALL_ROOM = dict(half=1.5, boxes=((0.5, 0.9, -0.2, 0.3),))


def _load(name):
    return np.load(os.path.join(GOLD, name))


def _params(g):
    return {k[2:]: (tuple(g[k]) if g[k].ndim else float(g[k])) for k in g.files if k.startswith("p_")}


# ---------------------------------------------------------------- standard-library restatements ---
@pytest.mark.parametrize("seed,keyspace,steps", [(1, 64, 4000), (2, 6400, 30000), (3, 40000, 150000)])
def test_occupied_set_iteration_order_equals_std_unordered_set(seed, keyspace, steps):
    assert orc._bind_oracle_pf().orc_selftest_occset(seed, keyspace, steps) == 0


@pytest.mark.parametrize("seed,steps,distinct", [(1, 60000, 2), (2, 60000, 40), (3, 60000, 5000)])
def test_heap_order_equals_std_priority_queue(seed, steps, distinct):
    assert orc._bind_oracle_pf().orc_selftest_heap(seed, steps, distinct) == 0


# ------------------------------------------------------------------------------ known answers ---
def test_kat4_single_hit_and_single_miss_sit_on_the_thresholds():
    """one hit gives prob == 0.9 (occupied), one miss prob == 0.35 (free)."""
    o = orc.OraclePf(num_particles=1, xmin=-2.0, xmax=2.0, ymin=-2.0, ymax=2.0)
    scan = np.full(360, 10.0, dtype=np.float32)     # everything out of range ...
    scan[0] = 1.0                                   # ... except one beam along +x
    assert o.grid_integrate(scan, (0.0, 0.0, 0.0)) == 0
    g = o.grid()
    xs = o.xsize
    hit = (40 + 20) * xs + 40                        # cell of (1.0, 0.0): i = floor(3/0.05) = 60, j = 40
    assert g["state"][hit] == 1 and g["prob"][hit] == 1.0
    assert g["log_odds"][hit] == np.log(0.9 / (1 - 0.9)) == 2.1972245773362196
    ray = [(40 + d) * xs + 40 for d in range(20)]
    assert all(g["state"][c] == 0 for c in ray)
    assert all(g["log_odds"][c] == np.log(0.35 / (1 - 0.35)) for c in ray)
    assert list(o.occ_order()) == [hit]
    # distance field: brushfire from the single seed, 4-connected propagation inheriting the source
    assert g["occ_dist"][hit] == 0.0
    assert g["occ_dist"][hit + 1] == 0.05 and g["occ_dist"][hit + xs] == 0.05
    assert g["occ_dist"][hit + xs + 1] == np.sqrt(2.0) * 0.05
    # occupied then free -> unknown band, removed from the occupied set
    scan2 = np.full(360, 10.0, dtype=np.float32)
    scan2[0] = 1.5
    assert o.grid_integrate(scan2, (0.0, 0.0, 0.0)) == 0
    g2 = o.grid()
    assert g2["state"][hit] == -1 and abs(g2["prob"][hit] - 0.8289473684210527) < 1e-15
    assert hit not in set(o.occ_order())


def test_pdf_normal_and_floor_term():
    """KAT4: pz = 0.95 * N(0.1; 0, 0.25) + 0.01/0.04 for one beam whose end point is 2 cells from the only obstacle."""
    o = orc.OraclePf(num_particles=1, xmin=-2.0, xmax=2.0, ymin=-2.0, ymax=2.0)
    scan = np.full(360, 10.0, dtype=np.float32)
    scan[0] = 1.0
    o.grid_integrate(scan, (0.0, 0.0, 0.0))
    probe = np.full(360, 10.0, dtype=np.float32)
    probe[0] = 0.92                                  # ends 2 cells short of the obstacle: occ_dist = 0.1
    rc, p = o.grid_likelihood(probe, (0.0, 0.0, 0.0))
    assert rc == 0 and abs(p - 0.9929811185533661) < 1e-15


def test_empty_map_likelihood_is_one_and_off_map_is_an_error():
    o = orc.OraclePf(num_particles=1, xmin=-2.0, xmax=2.0, ymin=-2.0, ymax=2.0)
    scan = np.full(360, 3.0, dtype=np.float32)
    assert o.grid_likelihood(scan, (0.0, 0.0, 0.0)) == (0, 1.0)         # grid_mapper.cpp:94-98, even with off-map beams
    assert o.grid_integrate(scan, (0.0, 0.0, 0.0)) == 1                  # world2Grid throws, :817-825
    ok = np.full(360, 1.0, dtype=np.float32)
    assert o.grid_integrate(ok, (0.0, 0.0, 0.0)) == 0
    assert o.grid_likelihood(scan, (0.0, 0.0, 0.0))[0] == 1              # now the map has obstacles -> throws


# -------------------------------------------------------------------------- golden fixtures ---
def test_grid_fixture_bit_exact():
    g = _load("rbpf_grid_ref.npz")
    o = orc.OraclePf(num_particles=1, **_params(g))
    f = 0
    for i in range(g["scans"].shape[0]):
        scan, pose = g["scans"][i], g["poses"][i]
        ep = o.grid_end_points(scan, pose)
        assert len(ep) == g["n_valid"][i] and np.array_equal(ep, g["end_points"][i][:len(ep)])
        for _ in range(4):
            cells = o.grid_free_cells(g["free_pt"][f], pose)
            assert np.array_equal(cells, g["free_cells"][f][:g["free_n"][f]])
            f += 1
        assert o.grid_likelihood(scan, g["lik_pose"][i]) == (0, g["lik"][i])
        assert o.grid_integrate(scan, pose) == 0
        m = o.grid()
        for k in ("log_odds", "prob", "occ_dist", "state"):
            assert np.array_equal(m[k], g[k][i]), (i, k)
        assert np.array_equal(o.occ_order(), g["occ_order"][i][:g["n_occ"][i]])
        assert o.bucket_count() == g["bucket_count"][i]
        assert np.array_equal(o.grid_map(), g["grid_map"][i])


@pytest.mark.parametrize("name", ["rbpf_slam_motion_ref.npz", "rbpf_slam_icp_ref.npz"])
@pytest.mark.parametrize("noise", ["mt19937", "external"])
def test_slam_fixture_bit_exact(name, noise):
    g = _load(name)
    N = int(g["N"])
    o = orc.OraclePf(num_particles=N, init_pose=tuple(g["odom"][0]), **_params(g))
    if noise == "mt19937":
        o.noise_mt19937(int(g["seed"]))
    for i in range(g["scans"].shape[0]):
        per = int(g["per_particle"][i])
        if noise == "external":
            o.noise_external(g["z"][i][:N * per + 1], per)
        rc = o.slam(g["scans"][i], g["twists"][i], g["odom"][i + 1], g["odom"][i], int(g["icp_ok"][i]), g["icp_pose"][i])
        assert rc == 0
        st = o.state()
        assert np.array_equal(st["weights"], g["weights"][i]), i
        assert np.array_equal(st["poses"], g["poses"][i]) and np.array_equal(st["prev_poses"], g["prev_poses"][i])
        assert o.resample_info()[1] == g["resampled"][i]
        assert np.array_equal(o.robot_state(), g["robot_state"][i])
        g0 = o.grid(0)
        assert np.array_equal(g0["occ_dist"], g["occ_dist0"][i]) and np.array_equal(g0["log_odds"], g["log_odds0"][i])
        assert np.array_equal(o.occ_order(0), g["occ_order0"][i][:g["n_occ0"][i]])
        assert np.array_equal(o.new_map(), g["new_map"][i])
    assert g["resampled"].sum() >= (1 if "motion" in name else 0)


def test_resample_fixture_bit_exact():
    g = _load("rbpf_resample_ref.npz")
    for c in range(int(g["n_cases"])):
        N = int(g["c%d_N" % c])
        o = orc.OraclePf(num_particles=N, xmin=-2.0, xmax=2.0, ymin=-2.0, ymax=2.0)
        o.noise_external(np.concatenate([np.zeros(N * 3), [g["c%d_z" % c]]]), 3)
        o.set_weights(g["c%d_w" % c])
        rs, anc = o.normalize_resample()
        assert rs == g["c%d_resampled" % c]
        assert np.array_equal(anc, g["c%d_anc" % c])                      # ancestors bit-exact
        assert np.array_equal(o.state()["weights"], g["c%d_w_after" % c])


# ----------------------------------------------------------------------- live against oracle/_ref ---
@needs_ref
@pytest.mark.parametrize("icp", [False, True])
def test_oracle_matches_live_reference_on_the_200x200_map(icp):
    rng = np.random.default_rng(21)
    scans = 5
    N = 6 if icp else 16
    poses, twists = orc.circle_path(scans)
    kw = dict(motion_noise=(2e-3, 1e-3, 1e-3), k=8)
    o = orc.OraclePf(num_particles=N, init_pose=tuple(poses[0]), **kw)
    r = orc.RefPf(num_particles=N, init_pose=tuple(poses[0]), **kw)
    o.noise_mt19937(77)
    r.seed(77)
    for i in range(scans):
        scan = orc.room_scan(poses[i + 1], rng=rng)
        icp_ok = int(icp and i > 0)
        icp_pose = (twists[i][0], twists[i][1] * np.cos(twists[i][0] / 2), twists[i][1] * np.sin(twists[i][0] / 2))
        assert o.slam(scan, twists[i], poses[i + 1], poses[i], icp_ok, icp_pose) == 0
        assert r.slam(scan, twists[i], poses[i + 1], poses[i], icp_ok, icp_pose) == 0
        so, sr = o.state(), r.state()
        for k in so:
            assert np.array_equal(so[k], sr[k]), (i, k)
        assert o.resample_info()[1] == r.last_resampled
        for p in (0, N - 1):
            go, gr = o.grid(p), r.grid(p)
            for k in go:
                assert np.array_equal(go[k], gr[k]), (i, p, k)
            assert np.array_equal(o.occ_order(p), r.occ_order(p))
        assert np.array_equal(o.new_map(), r.new_map())


@needs_ref
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_oracle_matches_live_reference_on_random_scans(seed):
    """Independent random ranges per beam (gated-out beams included) from random poses: rays in every direction class,
    scattered obstacle cells, distance fields grown from scattered seeds.  Same inputs as the GPU test of the same name."""
    rng = np.random.default_rng(1000 + seed)
    N, scans = 6, 5
    start = (float(rng.uniform(-3, 3)), float(rng.uniform(-0.8, 0.8)), float(rng.uniform(-0.8, 0.8)))
    kw = dict(num_particles=N, init_pose=start, motion_noise=(4e-3, 2e-3, 2e-3))
    o = orc.OraclePf(**kw)
    r = orc.RefPf(**kw)
    o.noise_mt19937(seed)
    r.seed(seed)
    prev = start
    for i in range(scans):
        cur = (prev[0] + float(rng.uniform(-0.3, 0.3)), prev[1] + float(rng.uniform(-0.05, 0.05)), prev[2] + float(rng.uniform(-0.05, 0.05)))
        twist = (cur[0] - prev[0], float(np.hypot(cur[1] - prev[1], cur[2] - prev[2])), 0.0)
        scan = rng.uniform(0.05, 3.45, 360).astype(np.float32)
        scan[rng.integers(0, 360, 25)] = np.float32(4.5)
        scan[rng.integers(0, 360, 10)] = np.float32(0.1)
        assert o.slam(scan, twist, cur, prev) == 0
        assert r.slam(scan, twist, cur, prev) == 0
        prev = cur
        so, sr = o.state(), r.state()
        for k in so:
            assert np.array_equal(so[k], sr[k]), (i, k)
        for p in range(N):
            go, gr = o.grid(p), r.grid(p)
            for k in go:
                assert np.array_equal(go[k], gr[k]), (i, p, k)
            assert np.array_equal(o.occ_order(p), r.occ_order(p))
        assert np.array_equal(o.new_map(), r.new_map())


@needs_ref
def test_bresenham_all_octants_match_reference():
    """Every direction class of GridMapper::freeGridIndex (grid_mapper.cpp:549-704), including the reversed-order
    and start-cell quirks, on random end points around random poses."""
    rng = np.random.default_rng(5)
    o = orc.OraclePf(num_particles=1)
    r = orc.RefGrid()
    for _ in range(400):
        pose = np.array([rng.uniform(-3, 3), rng.uniform(-1, 1), rng.uniform(-1, 1)])
        kind = rng.integers(0, 4)
        d = rng.uniform(0.05, 3.0)
        a = [0.0, np.pi / 2, np.pi / 4, rng.uniform(0, 2 * np.pi)][kind] + (np.pi if rng.integers(0, 2) else 0.0)
        pt = np.array([pose[1] + d * np.cos(a), pose[2] + d * np.sin(a)])
        if kind == 2:    # exact diagonal in cell space
            n = rng.integers(1, 40) * (1 if rng.integers(0, 2) else -1)
            pt = np.array([pose[1] + 0.05 * n, pose[2] + 0.05 * n * (1 if rng.integers(0, 2) else -1)])
        assert np.array_equal(o.grid_free_cells(pt, pose), r.free_cells(pt, pose))


def test_philox_noise_is_shard_invariant():
    """Mode B keys the motion noise by the GLOBAL particle id: two half-size filters reproduce the poses of one."""
    poses, twists = orc.circle_path(1)
    scan = orc.room_scan(poses[1])
    kw = dict(xmin=-3.0, xmax=3.0, ymin=-3.0, ymax=3.0, init_pose=tuple(poses[0]), motion_noise=(1e-3, 1e-3, 1e-3))
    full = orc.OraclePf(num_particles=8, **kw)
    full.noise_philox(42)
    full.slam(scan, twists[0], poses[1], poses[0])
    halves = []
    for off in (0, 4):
        h = orc.OraclePf(num_particles=4, **kw)
        h.noise_philox(42)
        h.set_shard(off)
        h.slam(scan, twists[0], poses[1], poses[0])
        halves.append(h.state()["poses"])
    assert np.array_equal(np.concatenate(halves), full.state()["poses"])
