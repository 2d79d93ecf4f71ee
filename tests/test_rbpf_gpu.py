"""GPU parity tests for the RBPF path, all through the C ABI of libb2nav.so (ctypes mirror in bmapping.py).

Bar (BASELINE.json north_star): resampling indices bit-exact; weights and poses within 1e-5 relative.  Everything that is
integer work - log-odds sums in beam order, cell classes, the occupied set's iteration order, the brushfire distance
field (as squared cell distances) - is compared BIT-EXACTLY; weights and poses are compared at 1e-12 (the only
differences are CUDA-vs-glibc sin/cos in the last ulp).
"""
import os

import numpy as np
import pytest

import _oracle as orc

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-5
TIGHT = 1e-12


def _params(g):
    q = {k[2:]: (tuple(g[k]) if g[k].ndim else float(g[k])) for k in g.files if k.startswith("p_")}
    return q


def make_gpu(pkg, **kw):
    extra = {k: kw.pop(k) for k in ("particle_offset", "particles_total", "device", "max_beams") if k in kw}
    return pkg.bmapping.make_filter(orc.pf_params(**kw), **extra)


def slam_gpu(pkg, f, scan, twist, cur, prev, icp_ok=0, icp_pose=(0.0, 0.0, 0.0)):
    f.scan_matcher.setResult(icp_ok, icp_pose)
    f.SLAM(scan, pkg.Twist2D(*twist), pkg.Pose(*cur), pkg.Pose(*prev))


def rel(a, b, floor=1e-300):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0


def assert_grid_equal(gg, go):
    assert np.array_equal(gg["log_odds"], go["log_odds"])
    assert np.array_equal(gg["occ_dist"], go["occ_dist"])
    assert np.array_equal(gg["state"].astype(np.int32), go["state"])


@pytest.mark.parametrize("name", ["rbpf_slam_motion_ref.npz", "rbpf_slam_icp_ref.npz"])
def test_reference_fixture_with_reference_variates(gpu_pkg, name):
    """The reference's own mt19937_64 normals (recorded from oracle/_ref) go in; its weights, poses, ancestors-by-effect,
    best pose, map of particle 0 and exported map must come out, scan after scan."""
    g = np.load(os.path.join(GOLD, name))
    N = int(g["N"])
    f = make_gpu(gpu_pkg, num_particles=N, init_pose=tuple(g["odom"][0]), **_params(g))
    for i in range(g["scans"].shape[0]):
        per = int(g["per_particle"][i])
        f.setNoise(g["z"][i][:N * per + 1])
        slam_gpu(gpu_pkg, f, g["scans"][i], g["twists"][i], g["odom"][i + 1], g["odom"][i], int(g["icp_ok"][i]), g["icp_pose"][i])
        assert f.resampleInfo()[1] == g["resampled"][i]
        assert rel(f.weights(), g["weights"][i]) < TIGHT
        p, pp = f.poses()
        assert np.max(np.abs(p - g["poses"][i])) < TIGHT and np.max(np.abs(pp - g["prev_poses"][i])) < TIGHT
        T = f.getRobotState()
        assert np.max(np.abs(np.array(T.displacement()) - g["robot_state"][i])) < TIGHT
        g0 = f.grid(0)
        assert np.array_equal(g0["log_odds"], g["log_odds0"][i])
        assert np.array_equal(g0["occ_dist"], g["occ_dist0"][i])
        assert np.array_equal(f.occOrder(0), g["occ_order0"][i][:g["n_occ0"][i]])
        assert np.array_equal(f.newMap(), g["new_map"][i])


def test_grid_fixture_single_particle(gpu_pkg):
    """One particle with zero motion noise IS a GridMapper: rbpf_grid_ref.npz scan by scan (poses forced)."""
    g = np.load(os.path.join(GOLD, "rbpf_grid_ref.npz"))
    q = _params(g)
    f = make_gpu(gpu_pkg, num_particles=1, motion_noise=(0.0, 0.0, 0.0), **q)
    for i in range(g["scans"].shape[0]):
        pose = g["poses"][i]
        f.setPoses(np.array([g["lik_pose"][i]]))
        assert rel(f.likelihoods(g["scans"][i]), [g["lik"][i]]) < TIGHT
        f.setPoses(np.array([pose]))
        slam_gpu(gpu_pkg, f, g["scans"][i], (0.0, 0.0, 0.0), pose, pose)
        m = f.grid(0)
        assert np.array_equal(m["log_odds"], g["log_odds"][i])
        assert np.array_equal(m["state"].astype(np.int32), g["state"][i])
        assert np.array_equal(m["occ_dist"], g["occ_dist"][i])
        assert np.array_equal(f.occOrder(0), g["occ_order"][i][:g["n_occ"][i]])
        assert np.array_equal(f.newMap(), g["grid_map"][i])


def test_resample_fixture_bit_exact(gpu_pkg):
    g = np.load(os.path.join(GOLD, "rbpf_resample_ref.npz"))
    for c in range(int(g["n_cases"])):
        N = int(g["c%d_N" % c])
        f = make_gpu(gpu_pkg, num_particles=N, xmin=-2.0, xmax=2.0, ymin=-2.0, ymax=2.0)
        f.setWeights(g["c%d_w" % c])
        f.setNoise(np.concatenate([np.zeros(3 * N), [g["c%d_z" % c]]]))
        neff, rs, anc = f.normalizeResample()
        assert rs == g["c%d_resampled" % c]
        assert np.array_equal(anc, g["c%d_anc" % c])
        assert np.array_equal(f.weights(), g["c%d_w_after" % c])          # sequential-order sums: bit-exact


@pytest.mark.parametrize("icp,N,scans", [(False, 64, 8), (True, 16, 5)])
def test_philox_closed_loop_matches_oracle_on_the_200x200_map(gpu_pkg, icp, N, scans):
    """Counter-based noise on both sides, config-C3 geometry (200x200 cells, 360 beams), several scans with resampling."""
    rng = np.random.default_rng(31)
    poses, twists = orc.circle_path(scans)
    kw = dict(num_particles=N, init_pose=tuple(poses[0]), motion_noise=(2e-3, 1e-3, 1e-3), k=12)
    f = make_gpu(gpu_pkg, **kw)
    o = orc.OraclePf(**kw)
    f.seed(42)
    o.noise_philox(42)
    n_res = 0
    for i in range(scans):
        scan = orc.room_scan(poses[i + 1], rng=rng)
        icp_ok = int(icp and i > 0)
        icp_pose = (twists[i][0], twists[i][1] * np.cos(twists[i][0] / 2), twists[i][1] * np.sin(twists[i][0] / 2))
        assert o.slam(scan, twists[i], poses[i + 1], poses[i], icp_ok, icp_pose) == 0
        slam_gpu(gpu_pkg, f, scan, twists[i], poses[i + 1], poses[i], icp_ok, icp_pose)
        so = o.state()
        neff_o, rs_o, anc_o = o.resample_info()
        neff_g, rs_g, anc_g = f.resampleInfo()
        assert (neff_g, rs_g) == (neff_o, rs_o) and np.array_equal(anc_g, anc_o)      # resampling indices bit-exact
        n_res += rs_o
        assert rel(f.weights(), so["weights"]) < 1e-9
        p, pp = f.poses()
        assert np.max(np.abs(p - so["poses"])) < 1e-11 and np.max(np.abs(pp - so["prev_poses"])) < 1e-11
        for k in (0, N // 2, N - 1):
            assert_grid_equal(f.grid(k), o.grid(k))
            assert np.array_equal(f.occOrder(k), o.occ_order(k))
        assert np.array_equal(f.newMap(), o.new_map())
        assert np.max(np.abs(np.array(f.getRobotState().displacement()) - o.robot_state())) < 1e-11
    if not icp:
        assert n_res >= 1


def test_heap_spill_path_gives_the_same_field(gpu_pkg):
    """A tiny shared-memory heap forces most heap entries through the global spill area; results must not change."""
    rng = np.random.default_rng(8)
    poses, twists = orc.circle_path(2)
    kw = dict(num_particles=4, init_pose=tuple(poses[0]), motion_noise=(1e-3, 1e-3, 1e-3))
    a, b = make_gpu(gpu_pkg, **kw), make_gpu(gpu_pkg, **kw)
    b.setHeapCapacity(16)
    for f in (a, b):
        f.seed(5)
    for i in range(2):
        scan = orc.room_scan(poses[i + 1], rng=rng)
        for f in (a, b):
            slam_gpu(gpu_pkg, f, scan, twists[i], poses[i + 1], poses[i])
    for k in range(4):
        ga, gb = a.grid(k), b.grid(k)
        assert np.array_equal(ga["occ_dist"], gb["occ_dist"])
    it, hm = b.distanceFieldStats()
    assert it > 0 and hm > 16


def test_shards_reproduce_the_full_filter_before_resampling(gpu_pkg):
    """Particles are independent until normalisation: two handles holding halves (global ids through particle_offset)
    produce the same poses, un-normalised likelihood ratios and maps as one handle holding all."""
    poses, twists = orc.circle_path(2)
    rng = np.random.default_rng(2)
    kw = dict(init_pose=tuple(poses[0]), motion_noise=(1e-3, 1e-3, 1e-3))
    full = make_gpu(gpu_pkg, num_particles=8, **kw)
    halves = [make_gpu(gpu_pkg, num_particles=4, particle_offset=o, **kw) for o in (0, 4)]
    for f in [full] + halves:
        f.seed(9)
    for i in range(2):
        scan = orc.room_scan(poses[i + 1], rng=rng)
        for f in [full] + halves:
            slam_gpu(gpu_pkg, f, scan, twists[i], poses[i + 1], poses[i])
    pf, _ = full.poses()
    ph = np.concatenate([h.poses()[0] for h in halves])
    assert np.array_equal(pf, ph)
    for k in range(8):
        assert np.array_equal(full.grid(k)["occ_dist"], halves[k // 4].grid(k % 4)["occ_dist"])


def test_config_c3_properties_at_full_size(gpu_pkg):
    """BASELINE configs[2]: 4096 particles, 360 beams, 200x200 map.  Size-independent properties: weights sum to 1 (until a resample),
    ancestors sorted and in range, every particle's occupied cells have distance 0, maps of resampled twins are
    identical, a second identical filter gives identical results (determinism)."""
    N, scans = 4096, 3
    poses, twists = orc.circle_path(scans)
    rng = np.random.default_rng(4)
    kw = dict(num_particles=N, init_pose=tuple(poses[0]), motion_noise=(2e-3, 1e-3, 1e-3))
    a, b = make_gpu(gpu_pkg, **kw), make_gpu(gpu_pkg, **kw)
    for f in (a, b):
        f.seed(1234)
    resampled = 0
    for i in range(scans):
        scan = orc.room_scan(poses[i + 1], rng=rng)
        for f in (a, b):
            slam_gpu(gpu_pkg, f, scan, twists[i], poses[i + 1], poses[i])
        w = a.weights()
        neff, rs, anc = a.resampleInfo()
        # the reference does not reset weights after resampling: copies carry their normalised weight (particle_filter.cpp:497-499)
        assert np.all(w >= 0) and (rs or abs(w.sum() - 1.0) < 1e-12)
        assert np.all(np.diff(anc) >= 0) and anc.min() >= 0 and anc.max() < N
        if rs:
            resampled += 1
            twins = np.where(np.diff(anc) == 0)[0]
            if len(twins):
                m = int(twins[0])
                assert np.array_equal(a.grid(m)["log_odds"], a.grid(m + 1)["log_odds"])
        assert np.array_equal(w, b.weights())
    for k in (0, 1777, N - 1):
        g = a.grid(k)
        assert np.all(g["occ_dist"][g["state"] == 1] == 0.0)
        assert np.array_equal(g["occ_dist"], b.grid(k)["occ_dist"])
        assert len(a.occOrder(k)) == int(np.sum(g["state"] == 1))


def test_config_c3_full_size_against_the_oracle(gpu_pkg):
    """BASELINE configs[2] at FULL size against the CPU oracle (not against itself): 4096 particles, 360 beams, 200x200 map,
    two scans on the same Philox streams.  Weights and poses at 1e-9, N_eff / resample flag / ancestors bit-exact, three full
    maps (log-odds, distance field, cell classes, occupied-set iteration order) bit-exact, exported map equal.  The oracle
    needs about 15 s per scan at this size."""
    N, scans = 4096, 2
    poses, twists = orc.circle_path(scans)
    rng = np.random.default_rng(4)
    kw = dict(num_particles=N, init_pose=tuple(poses[0]), motion_noise=(2e-3, 1e-3, 1e-3))
    f = make_gpu(gpu_pkg, **kw)
    o = orc.OraclePf(**kw)
    f.seed(1234)
    o.noise_philox(1234)
    for i in range(scans):
        scan = orc.room_scan(poses[i + 1], rng=rng)
        slam_gpu(gpu_pkg, f, scan, twists[i], poses[i + 1], poses[i])
        o.slam(scan, twists[i], poses[i + 1], poses[i])
        st = o.state()
        assert rel(f.weights(), st["weights"]) < 1e-9, i
        assert np.max(np.abs(f.poses()[0] - st["poses"])) < 1e-9, i
        ng, rg, ag = f.resampleInfo()
        no, ro, ao = o.resample_info()
        assert (ng, rg) == (no, ro) and np.array_equal(ag, ao), i
    for k in (0, 2049, N - 1):
        assert_grid_equal(f.grid(k), o.grid(k))
        assert np.array_equal(f.occOrder(k), o.occ_order(k))
    assert np.array_equal(f.newMap(), o.new_map())
    assert np.max(np.abs(np.subtract(f.getRobotState().displacement(), o.robot_state()))) < 1e-9


def test_error_behaviour(gpu_pkg):
    poses, twists = orc.circle_path(1)
    f = make_gpu(gpu_pkg, num_particles=4, xmin=-2.0, xmax=2.0, ymin=-2.0, ymax=2.0)
    scan = np.full(360, 3.0, dtype=np.float32)          # end points 3 m away on a 2 m map
    with pytest.raises(gpu_pkg.B2NError) as e:
        slam_gpu(gpu_pkg, f, scan, (0.0, 0.0, 0.0), (0.0, 0.0, 0.0), (0.0, 0.0, 0.0))
    assert e.value.code == -3                            # B2N_ERR_OFF_MAP (reference: world2Grid throws)
    with pytest.raises(gpu_pkg.B2NError):
        make_gpu(gpu_pkg, num_particles=0)
    with pytest.raises(gpu_pkg.B2NError) as e:
        make_gpu(gpu_pkg, num_particles=2, xmin=-5.0, xmax=5.0, ymin=-2.0, ymax=2.0)
    assert e.value.code == -4                            # non-square map
    g = make_gpu(gpu_pkg, num_particles=2)
    with pytest.raises(gpu_pkg.B2NError):
        g.setNoise(np.zeros(5))
    with pytest.raises(gpu_pkg.B2NError):
        slam_gpu(gpu_pkg, g, np.zeros(4000, dtype=np.float32), (0, 0, 0), (0, 0, 0), (0, 0, 0))


@pytest.mark.parametrize("env", [{"B2N_PF_DF_LANES": "8"}, {"B2N_PF_DF_LANES": "16"}, {"B2N_PF_DF_LANES": "32"},
                                 {"B2N_PF_DF_LANES": "32", "B2N_PF_DF_SMEM_MARKS": "1"}])
def test_every_distance_field_kernel_variant_is_bit_exact(gpu_pkg, env, monkeypatch):
    """The brushfire exists in four forms (lane groups of 8 / 16 with the visited bitmap in tensor memory staged through
    shared memory, one particle per warp with the bitmap in tensor memory or in shared memory); the handle reads the choice
    from the environment at creation.  All must reproduce the oracle's field, heap order quirks included."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    rng = np.random.default_rng(77)
    N, scans = 37, 4                                   # not a multiple of any group size
    poses, twists = orc.circle_path(scans)
    kw = dict(num_particles=N, init_pose=tuple(poses[0]), motion_noise=(2e-3, 1e-3, 1e-3))
    f = make_gpu(gpu_pkg, **kw)
    o = orc.OraclePf(**kw)
    f.seed(3)
    o.noise_philox(3)
    for i in range(scans):
        scan = orc.room_scan(poses[i + 1], rng=rng)
        assert o.slam(scan, twists[i], poses[i + 1], poses[i]) == 0
        slam_gpu(gpu_pkg, f, scan, twists[i], poses[i + 1], poses[i])
        for k in (0, 17, N - 1):
            assert_grid_equal(f.grid(k), o.grid(k))
    assert rel(f.weights(), o.state()["weights"]) < 1e-9


def test_shipped_launch_map_80x80(gpu_pkg):
    """bmapping/launch/slam.launch:40-42: a 4 m x 4 m map (80 x 80 cells) - another grid size, another bitmap width."""
    rng = np.random.default_rng(5)
    N, scans = 24, 4
    poses, twists = orc.circle_path(scans, radius=0.3, step=0.04)
    kw = dict(num_particles=N, init_pose=tuple(poses[0]), motion_noise=(1e-3, 5e-4, 5e-4), xmin=-2.0, xmax=2.0, ymin=-2.0, ymax=2.0)
    f = make_gpu(gpu_pkg, **kw)
    o = orc.OraclePf(**kw)
    f.seed(8)
    o.noise_philox(8)
    for i in range(scans):
        scan = orc.room_scan(poses[i + 1], half=1.6, boxes=((0.6, 0.9, -0.2, 0.3),), rng=rng)
        assert o.slam(scan, twists[i], poses[i + 1], poses[i]) == 0
        slam_gpu(gpu_pkg, f, scan, twists[i], poses[i + 1], poses[i])
        assert np.array_equal(f.resampleInfo()[2], o.resample_info()[2])
        for k in (0, N - 1):
            assert_grid_equal(f.grid(k), o.grid(k))
        assert np.array_equal(f.newMap(), o.new_map())
    assert rel(f.weights(), o.state()["weights"]) < 1e-9


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_scans_exercise_every_ray_direction(gpu_pkg, seed):
    """Scans with independent random ranges per beam (gated-out beams included) from random poses: rays in every octant,
    axis-aligned and diagonal cases, isolated obstacle cells, distance fields grown from scattered seeds - log-odds,
    occupied-set order and distance field bit-exact after every scan, weights and ancestors as the oracle's."""
    rng = np.random.default_rng(1000 + seed)
    N, scans = 6, 5
    start = (float(rng.uniform(-3, 3)), float(rng.uniform(-0.8, 0.8)), float(rng.uniform(-0.8, 0.8)))
    kw = dict(num_particles=N, init_pose=start, motion_noise=(4e-3, 2e-3, 2e-3))
    f = make_gpu(gpu_pkg, **kw)
    o = orc.OraclePf(**kw)
    f.seed(seed)
    o.noise_philox(seed)
    prev = start
    for i in range(scans):
        cur = (prev[0] + float(rng.uniform(-0.3, 0.3)), prev[1] + float(rng.uniform(-0.05, 0.05)), prev[2] + float(rng.uniform(-0.05, 0.05)))
        twist = (cur[0] - prev[0], float(np.hypot(cur[1] - prev[1], cur[2] - prev[2])), 0.0)
        scan = rng.uniform(0.05, 3.45, 360).astype(np.float32)
        scan[rng.integers(0, 360, 25)] = np.float32(4.5)                  # beyond range_max: filtered by the gate
        scan[rng.integers(0, 360, 10)] = np.float32(0.1)                  # below range_min
        assert o.slam(scan, twist, cur, prev) == 0
        slam_gpu(gpu_pkg, f, scan, twist, cur, prev)
        prev = cur
        for k in range(N):
            assert_grid_equal(f.grid(k), o.grid(k))
            assert np.array_equal(f.occOrder(k), o.occ_order(k))
        assert np.array_equal(f.resampleInfo()[2], o.resample_info()[2])
        assert rel(f.weights(), o.state()["weights"]) < 1e-9
        assert np.array_equal(f.newMap(), o.new_map())


def test_unchanged_occupied_set_skips_the_distance_field_and_changes_nothing(gpu_pkg):
    """A robot standing still with four isolated returns: after the first scan no cell enters or leaves the occupied set,
    so the brushfire would reproduce the field that is already there - the kernel skips it; the oracle (which regrows it
    every scan, like the reference) must still agree bit for bit.  (With a full room scan the set is touched every scan:
    grazing rays erase and re-insert wall cells, which also reorders the set.)"""
    N, scans = 5, 4
    pose = (0.3, 0.4, -0.2)
    kw = dict(num_particles=N, init_pose=pose, motion_noise=(0.0, 0.0, 0.0))
    f = make_gpu(gpu_pkg, **kw)
    o = orc.OraclePf(**kw)
    f.seed(1)
    o.noise_philox(1)
    scan = np.full(360, 4.5, dtype=np.float32)          # everything out of range ...
    scan[[10, 100, 190, 280]] = np.float32(2.0)         # ... except four well separated hits: no ray crosses another's end cell
    for i in range(scans):
        assert o.slam(scan, (0.0, 0.0, 0.0), pose, pose) == 0
        slam_gpu(gpu_pkg, f, scan, (0.0, 0.0, 0.0), pose, pose)
        for k in range(N):
            assert_grid_equal(f.grid(k), o.grid(k))
            assert np.array_equal(f.occOrder(k), o.occ_order(k))
        assert rel(f.weights(), o.state()["weights"]) < 1e-9
    assert f.distanceFieldSkipped() == (scans - 1) * N


def test_fifty_scans_match_oracle(gpu_pkg):
    """>= 50 consecutive scans with the filter state carried (most of a lap of the circular path), every
    scan compared: resampling decisions and ancestors bit-exact, weights / poses at 1e-9, maps of two particles and the
    exported map equal."""
    rng = np.random.default_rng(50)
    N, scans = 16, 50
    poses, twists = orc.circle_path(scans)
    kw = dict(num_particles=N, init_pose=tuple(poses[0]), motion_noise=(2e-3, 1e-3, 1e-3))
    f = make_gpu(gpu_pkg, **kw)
    o = orc.OraclePf(**kw)
    f.seed(50)
    o.noise_philox(50)
    n_res = 0
    for i in range(scans):
        scan = orc.room_scan(poses[i + 1], rng=rng)
        assert o.slam(scan, twists[i], poses[i + 1], poses[i]) == 0
        slam_gpu(gpu_pkg, f, scan, twists[i], poses[i + 1], poses[i])
        neff_o, rs_o, anc_o = o.resample_info()
        neff_g, rs_g, anc_g = f.resampleInfo()
        assert (neff_g, rs_g) == (neff_o, rs_o) and np.array_equal(anc_g, anc_o), i
        n_res += rs_o
        so = o.state()
        assert rel(f.weights(), so["weights"]) < 1e-9, i
        assert np.max(np.abs(f.poses()[0] - so["poses"])) < 1e-9, i
        if i % 7 == 0 or i == scans - 1:
            for k in (0, N - 1):
                assert_grid_equal(f.grid(k), o.grid(k))
            assert np.array_equal(f.newMap(), o.new_map())
    assert n_res >= 5


def test_largest_maps_and_other_scan_geometries(gpu_pkg):
    """A 250 x 250 map (62 500 cells: 16-bit cell ids nearly exhausted, 62 bitmap columns per particle in tensor memory)
    and a 180-beam, 2-degree scanner."""
    rng = np.random.default_rng(9)
    N, scans = 6, 3
    poses, twists = orc.circle_path(scans)
    kw = dict(num_particles=N, init_pose=tuple(poses[0]), motion_noise=(2e-3, 1e-3, 1e-3), xmin=-6.25, xmax=6.25, ymin=-6.25, ymax=6.25,
              beam_delta=float(np.float32(np.deg2rad(2.0))))
    f = make_gpu(gpu_pkg, **kw)
    assert f.xsize == 250
    o = orc.OraclePf(**kw)
    f.seed(4)
    o.noise_philox(4)
    for i in range(scans):
        scan = orc.room_scan(poses[i + 1], n_beams=180, beam_delta=np.deg2rad(2.0), rng=rng)
        assert o.slam(scan, twists[i], poses[i + 1], poses[i]) == 0
        slam_gpu(gpu_pkg, f, scan, twists[i], poses[i + 1], poses[i])
        for k in (0, N - 1):
            assert_grid_equal(f.grid(k), o.grid(k))
            assert np.array_equal(f.occOrder(k), o.occ_order(k))
        assert np.array_equal(f.resampleInfo()[2], o.resample_info()[2])
        assert np.array_equal(f.newMap(), o.new_map())
    assert rel(f.weights(), o.state()["weights"]) < 1e-9
