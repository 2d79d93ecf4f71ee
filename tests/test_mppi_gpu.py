"""GPU parity tests for the MPPI path, all through the C ABI of libb2nav.so (ctypes).

Tolerances (BASELINE.json north_star): 1e-5 relative on trajectory states, weights and chosen
controls.  The fp64 kernels land orders of magnitude inside that; the asserts below use the
contract tolerance for the headline quantities and tighter ones where a regression would hide.
"""
import ctypes as C
import os

import numpy as np
import pytest

import _oracle as orc

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-5


def _params(g):
    return {k[2:]: (tuple(g[k]) if g[k].ndim else float(g[k])) for k in g.files if k.startswith("p_")}


def make_gpu(pkg, horizon, dt, K, prm=orc.SHIPPED, **kw):
    return pkg.MPPI(pkg.CartModel(prm["wheel_radius"], prm["wheel_base"]), pkg.LossFunc(prm["Q"], prm["R"], prm["P1"]),
                    prm["lambda_"], prm["max_wheel_vel"], prm["ul_var"], prm["ur_var"], horizon, dt, K, **kw)


def rel_err(a, b, floor=0.0):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)) if a.size else 0.0


def check_call(gpu, o, ctrl_gpu, ctrl_o, states=True):
    """Compare everything one newControls() produced on the GPU with the oracle's."""
    g = o.get()
    K, T = o.K, o.T
    assert rel_err(ctrl_gpu, ctrl_o, 1e-3) < RTOL
    assert rel_err(gpu.plan(), g["plan"], 1e-3) < RTOL
    J = gpu.costToGo()                                  # [K][T]
    assert rel_err(J, g["J"].T, 1e-6) < 1e-11          # what the 1e-5 on weights needs at lambda = 0.01
    w = gpu.weights()
    wo = g["w"].T
    assert rel_err(w, wo, 1e-300) < RTOL
    assert np.allclose(w.sum(axis=0), 1.0, rtol=1e-12, atol=0)
    if states:
        s = gpu.states()                                # fp32 tensor
        so = g["states"]
        assert np.max(np.abs(s - so) / np.maximum(np.abs(so), 1e-2)) < RTOL
        assert np.array_equal(s, so.astype(np.float32)) or np.max(np.abs(s - so.astype(np.float32))) <= 2.4e-7 * np.max(np.abs(so))


@pytest.mark.parametrize("name", ["mppi_c1_shipped_ref.npz", "mppi_c1_mild_ref.npz", "mppi_t64_shipped_ref.npz",
                                  "mppi_t100_shipped_ref.npz"])
def test_reference_fixture_with_reference_variates(gpu_pkg, name):
    """The reference's own mt19937_64 variates (recorded from oracle/_ref) go in; its controls, plan and
    min-subtracted cost-to-go must come out, call after call with the plan carried on the device."""
    g = np.load(os.path.join(GOLD, name))
    K, T = int(g["K"]), int(g["T"])
    gpu = make_gpu(gpu_pkg, float(g["horizon"]), float(g["dt"]), K, _params(g))
    assert gpu.steps == T
    gpu.setCapture(True)
    gpu.setInitialControls(0.0, 0.0)
    gpu.setWaypoint(gpu_pkg.Pose(theta=g["wpt"][2], x=g["wpt"][0], y=g["wpt"][1]))
    for c in range(g["poses"].shape[0]):
        gpu.setNoise(g["du"][c])
        x, y, th = g["poses"][c]
        v = gpu.newControls(gpu_pkg.Pose(theta=th, x=x, y=y))
        assert rel_err([v.ul, v.ur], g["controls"][c], 1e-3) < RTOL, (c, v, g["controls"][c])
        assert rel_err(gpu.plan(), g["plans"][c], 1e-3) < RTOL
        J = gpu.costToGo().T
        Jsub = J - J.min(axis=1, keepdims=True)
        assert np.max(np.abs(Jsub - g["Jsub"][c])) < 1e-11 * np.max(J)
        assert np.array_equal(gpu.noise(), g["du"][c])


@pytest.mark.parametrize("hor,dt,K,prm", [
    (0.5, 0.02, 128, orc.SHIPPED),      # config C1 (T = 25: generic store path, one step per lane)
    (0.64, 0.01, 1000, orc.SHIPPED),    # T = 64, K not a multiple of the warps per CTA
    (1.0, 0.01, 77, orc.SHIPPED),       # shipped horizon, T = 100: four steps per lane, last lanes idle
    (1.28, 0.01, 256, orc.MILD),        # T = 128
    (2.56, 0.01, 40, orc.MILD),         # T = 256: eight steps per lane
    (0.01, 0.01, 5, orc.SHIPPED),       # T = 1: terminal loss only
    (0.32, 0.01, 1, orc.SHIPPED),       # K = 1: weight is 1, update is the perturbation itself
])
def test_philox_closed_loop_matches_oracle(gpu_pkg, hor, dt, K, prm):
    o = orc.OracleMppi(hor, dt, K, **prm)
    gpu = make_gpu(gpu_pkg, hor, dt, K, prm)
    assert gpu.steps == o.T
    gpu.setCapture(True)
    gpu.seed(42)
    o.noise_philox(42)
    for m in (gpu, o):
        m.setInitialControls(0.2, 0.1)
    o.setWaypoint(1.0, 0.0, 1.5707)
    gpu.setWaypoint(gpu_pkg.Pose(theta=1.5707, x=1.0, y=0.0))
    pose = (0.0, 0.0, 0.0)
    for c in range(8):
        v = gpu.newControls(gpu_pkg.Pose(theta=pose[2], x=pose[0], y=pose[1]))
        co = o.newControls(*pose)
        assert np.array_equal(gpu.noise(), o.get()["du"])               # same Philox stream, exact-op Box-Muller: bit for bit
        check_call(gpu, o, (v.ul, v.ur), co)
        pose = orc.unicycle_step(pose, co[0], co[1], dt)


def test_config_c1_hundred_calls(gpu_pkg):
    """>= 100 consecutive newControls with the receding-horizon state carried, every call compared."""
    hor, dt, K = 0.5, 0.02, 128
    o = orc.OracleMppi(hor, dt, K)
    gpu = make_gpu(gpu_pkg, hor, dt, K)
    gpu.seed(7)
    o.noise_philox(7)
    o.setWaypoint(1.0, 0.0, 1.5707)
    gpu.setWaypoint(gpu_pkg.Pose(theta=1.5707, x=1.0, y=0.0))
    pose = (0.0, 0.0, 0.0)
    worst = 0.0
    for c in range(100):
        v = gpu.newControls(gpu_pkg.Pose(theta=pose[2], x=pose[0], y=pose[1]))
        co = o.newControls(*pose)
        worst = max(worst, rel_err([v.ul, v.ur], co, 1e-3))
        pose = orc.unicycle_step(pose, co[0], co[1], dt)
    assert worst < RTOL
    assert rel_err(gpu.plan(), o.get()["plan"], 1e-3) < RTOL


def test_config_c2_full_size(gpu_pkg):
    """BASELINE config 2: K = 16384, T = 64, shipped cost, against the oracle at full size."""
    hor, dt, K = 0.64, 0.01, 16384
    o = orc.OracleMppi(hor, dt, K)
    gpu = make_gpu(gpu_pkg, hor, dt, K)
    gpu.setCapture(True)
    gpu.seed(42)
    o.noise_philox(42)
    o.setWaypoint(1.0, 0.0, 1.5707)
    gpu.setWaypoint(gpu_pkg.Pose(theta=1.5707, x=1.0, y=0.0))
    pose = (0.0, 0.0, 0.0)
    for c in range(3):
        v = gpu.newControls(gpu_pkg.Pose(theta=pose[2], x=pose[0], y=pose[1]))
        co = o.newControls(*pose)
        check_call(gpu, o, (v.ul, v.ur), co)
        pose = orc.unicycle_step(pose, co[0], co[1], dt)


def test_box_muller_every_radius_bit_exact(gpu_pkg):
    """The perturbation generator's Box-Muller stage for ALL 2^23 possible first words (every radius: ln by polynomial, square
    root by SFU seed + one fused Newton step) and four second words (one per quadrant): bit for bit the oracle's fmaf / sqrtf."""
    lib = gpu_pkg.load_library()
    L = orc.oracle_lib()
    L.orc_box_muller_range.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
    L.orc_box_muller_range.restype = None
    n = 1 << 23
    for rb in (0x12345678, 0x6badf00d, 0x9e3779b9, 0xfedcba98):
        zg, zo = np.empty(2 * n, dtype=np.float32), np.empty(2 * n, dtype=np.float32)
        assert lib.b2n_test_box_muller(0, n, rb, zg.ctypes.data_as(C.c_void_p)) == 0
        L.orc_box_muller_range(0, n, rb, zo.ctypes.data_as(C.c_void_p))
        assert np.array_equal(zg.view(np.uint32), zo.view(np.uint32)), rb
    # second word: 2^16 angles spread over the range, a few radii
    rbs = np.arange(0, 1 << 32, 65537, dtype=np.uint64).astype(np.uint32)
    zg, zo = np.empty(2, dtype=np.float32), np.empty(2, dtype=np.float32)
    for rb in rbs[::257]:
        assert lib.b2n_test_box_muller(4242, 1, int(rb), zg.ctypes.data_as(C.c_void_p)) == 0
        L.orc_box_muller_range(4242, 1, int(rb), zo.ctypes.data_as(C.c_void_p))
        assert np.array_equal(zg.view(np.uint32), zo.view(np.uint32)), rb


@pytest.mark.parametrize("K,ring,calls", [(16384, 1, 4), (2048, 3, 5), (16384, 3, 3)])
def test_production_variant_matches_oracle_at_c2(gpu_pkg, K, ring, calls):
    """BASELINE configs[1] on the instantiation the bench times: capture OFF, so newControls() runs the FAST
    (S = 4, G = 16) rollout variant with the fused tail.  Controls, plan and the fp32 state tensor against the
    oracle over receding-horizon calls; K = 2048 is the 8-GPU share of C2, ring = 3 the bench's state ring."""
    hor, dt = 0.64, 0.01
    o = orc.OracleMppi(hor, dt, K)
    gpu = make_gpu(gpu_pkg, hor, dt, K)
    if ring > 1:
        gpu.setStateRing(ring)
    gpu.seed(42)
    o.noise_philox(42)
    o.setWaypoint(1.0, 0.0, 1.5707)
    gpu.setWaypoint(gpu_pkg.Pose(theta=1.5707, x=1.0, y=0.0))
    pose = (0.0, 0.0, 0.0)
    for c in range(calls):
        v = gpu.newControls(gpu_pkg.Pose(theta=pose[2], x=pose[0], y=pose[1]))
        assert gpu.lastVariant() == "fast", "the production variant did not run"
        co = o.newControls(*pose)
        g = o.get()
        assert rel_err([v.ul, v.ur], co, 1e-3) < RTOL, (c, v, co)
        assert rel_err(gpu.plan(), g["plan"], 1e-3) < RTOL
        s, so = gpu.states(), g["states"]
        assert np.max(np.abs(s - so) / np.maximum(np.abs(so), 1e-2)) < RTOL
        assert np.max(np.abs(s - so.astype(np.float32))) <= 2.4e-7 * np.max(np.abs(so))
        # the merged per-step sums the update consumed (min J, sum e, sum e*du, sum du) against the oracle's J and du
        p = gpu.partials()
        Jo = g["J"]
        m = Jo.min(axis=1)
        e = np.exp(-(Jo - m[:, None]) / orc.SHIPPED["lambda_"])
        assert rel_err(p[:, 0], m, 1e-6) < 1e-11
        assert rel_err(p[:, 1], e.sum(axis=1), 1e-300) < 1e-6
        assert np.allclose(p[:, 2], (e * g["du"][:, :, 0].T).sum(axis=1), rtol=1e-5, atol=1e-9)
        assert np.allclose(p[:, 3], (e * g["du"][:, :, 1].T).sum(axis=1), rtol=1e-5, atol=1e-9)
        assert np.allclose(p[:, 4], g["du"][:, :, 0].sum(axis=0), rtol=1e-5, atol=1e-3)
        pose = orc.unicycle_step(pose, co[0], co[1], dt)


def test_config_c4_shard_size_with_obstacles(gpu_pkg):
    """BASELINE config 4 per-GPU shard (8192 of 65536 rollouts, T = 128) with the obstacle term on."""
    hor, dt, K = 1.28, 0.01, 8192
    xs = np.arange(200)
    dist = np.hypot((xs[:, None] - 110) * 0.05, (xs[None, :] - 104) * 0.05).astype(np.float32)   # one obstacle cell
    o = orc.OracleMppi(hor, dt, K)
    gpu = make_gpu(gpu_pkg, hor, dt, K, rollout_offset=3 * 8192, rollouts_total=65536)
    o.set_shard(3 * 8192)
    o.set_obstacles(dist, -5.0, -5.0, 0.05, 5e4, 0.4, 1e6)
    gpu.setObstacleField(dist, -5.0, -5.0, 0.05, 5e4, 0.4, 1e6)
    gpu.setCapture(True)
    gpu.seed(42)
    o.noise_philox(42)
    o.setWaypoint(1.0, 0.0, 0.0)
    gpu.setWaypoint(gpu_pkg.Pose(theta=0.0, x=1.0, y=0.0))
    gpu.newControls(gpu_pkg.Pose(theta=0.0, x=0.3, y=0.1))
    o.newControls(0.3, 0.1, 0.0)
    g = o.get()
    assert rel_err(gpu.costToGo(), g["J"].T, 1e-6) < 1e-11
    assert np.max(np.abs(gpu.states() - g["states"])) < 1e-6
    # the shard's control differs from the oracle's (the +1e-8 floor is normalised by the JOB's K),
    # so compare the per-step partial sums instead: min J, sum e, sum e*du
    p = gpu.partials()
    Jo = g["J"]
    m = Jo.min(axis=1)
    e = np.exp(-(Jo - m[:, None]) / 0.01)
    assert rel_err(p[:, 0], m, 1e-6) < 1e-11
    assert rel_err(p[:, 1], e.sum(axis=1), 1e-300) < 1e-7
    assert np.allclose(p[:, 2], (e * g["du"][:, :, 0].T).sum(axis=1), rtol=1e-6, atol=1e-9)
    assert np.allclose(p[:, 4], g["du"][:, :, 0].sum(axis=0), rtol=1e-9, atol=1e-9)


def test_shards_compose_to_the_full_job(gpu_pkg):
    """Two half-size handles (as two ranks would hold) see the same noise and produce partials whose merge
    equals the full job's partials."""
    hor, dt, K = 0.64, 0.01, 512
    full = make_gpu(gpu_pkg, hor, dt, K)
    halves = [make_gpu(gpu_pkg, hor, dt, K // 2, rollout_offset=i * K // 2, rollouts_total=K) for i in range(2)]
    P = gpu_pkg.Pose(theta=0.1, x=0.0, y=0.0)
    for m in [full] + halves:
        m.seed(5)
        m.setCapture(True)
        m.setWaypoint(gpu_pkg.Pose(theta=0.0, x=1.0, y=0.2))
        m.newControls(P)
    assert np.array_equal(np.concatenate([h.noise() for h in halves]), full.noise())
    assert np.array_equal(np.concatenate([h.states() for h in halves]), full.states())
    pf = full.partials()
    ph = [h.partials() for h in halves]
    m = np.minimum(ph[0][:, 0], ph[1][:, 0])
    f = [np.exp((m - p[:, 0]) / 0.01) for p in ph]
    assert np.allclose(m, pf[:, 0], rtol=1e-15)
    for j in (1, 2, 3):
        assert np.allclose(ph[0][:, j] * f[0] + ph[1][:, j] * f[1], pf[:, j], rtol=1e-9, atol=1e-12)
    for j in (4, 5):
        assert np.allclose(ph[0][:, j] + ph[1][:, j], pf[:, j], rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("nranks,K,T_cfg,ahead", [(2, 16384, (0.64, 0.01), False), (2, 2048, (0.64, 0.01), True), (4, 2048, (0.64, 0.01), False),
                                                  (2, 4096, (1.28, 0.01), False), (8, 1024, (0.32, 0.01), False)])
def test_sharded_exchange_on_one_gpu_matches_the_unsharded_oracle(gpu_pkg, monkeypatch, nranks, K, T_cfg, ahead):
    """The sharded path - every rank's merger CTAs exchange their step's [6] sums over peer memory inside the call's one
    kernel and apply the identical update - with the ranks as handles of THIS process sharing ONE GPU (separate streams):
    the production (FAST) instantiation against the UNSHARDED oracle on the whole job, and the plan replicated bit for bit
    on every rank.  Merger CTAs wait for the other ranks' words while resident, so on a shared GPU every rank's grid must
    be able to start: the sizes keep the merger CTAs of all ranks well inside the GPU, and the kernels that draw the next
    call's variates behind a call (they queue more CTAs than the GPU holds) are switched off except in the smallest case.
    With one GPU per rank - the way the path is deployed - there is no such coupling."""
    monkeypatch.setenv("B2N_MPPI_NOISE_AHEAD", "1" if ahead else "0")
    hor, dt = T_cfg
    Kr = K // nranks
    o = orc.OracleMppi(hor, dt, K)
    ranks = [make_gpu(gpu_pkg, hor, dt, Kr, rollout_offset=r * Kr, rollouts_total=K) for r in range(nranks)]
    for m in ranks:
        m.p2pExport(nranks)
    areas = [m.p2pArea() for m in ranks]
    for r, m in enumerate(ranks):
        m.p2pInitLocal(r, nranks, areas)
        m.seed(42)
        m.setWaypoint(gpu_pkg.Pose(theta=1.5707, x=1.0, y=0.0))
    o.noise_philox(42)
    o.setWaypoint(1.0, 0.0, 1.5707)
    pose = (0.0, 0.0, 0.0)
    for c in range(4):
        P = gpu_pkg.Pose(theta=pose[2], x=pose[0], y=pose[1])
        for m in ranks:
            m.enqueue(P)
        vs = [m.wait() for m in ranks]
        assert all(m.lastVariant() == "fast" for m in ranks)
        co = o.newControls(*pose)
        plans = [m.plan() for m in ranks]
        for v, p in zip(vs, plans):
            assert rel_err([v.ul, v.ur], co, 1e-3) < RTOL, (c, v, co)
            assert rel_err(p, o.get()["plan"], 1e-3) < RTOL
            assert (v.ul, v.ur) == (vs[0].ul, vs[0].ur) and np.array_equal(p, plans[0])
        so = o.get()["states"]
        for r, m in enumerate(ranks):
            s = m.states()
            assert np.max(np.abs(s - so[r * Kr:(r + 1) * Kr]) / np.maximum(np.abs(so[r * Kr:(r + 1) * Kr]), 1e-2)) < RTOL
        pose = orc.unicycle_step(pose, co[0], co[1], dt)


def test_zero_variance_leaves_the_plan_alone(gpu_pkg):
    prm = dict(orc.SHIPPED, ul_var=0.0, ur_var=0.0)
    gpu = make_gpu(gpu_pkg, 0.64, 0.01, 64, prm)
    gpu.setInitialControls(0.5, -0.25)
    gpu.setWaypoint(gpu_pkg.Pose(theta=0.0, x=1.0, y=0.0))
    v = gpu.newControls(gpu_pkg.Pose())
    assert (v.ul, v.ur) == (0.5, -0.25)
    assert np.array_equal(gpu.plan(), np.repeat([[0.5], [-0.25]], 64, axis=1))


def test_controls_saturate(gpu_pkg):
    prm = dict(orc.SHIPPED, max_wheel_vel=0.05)
    gpu = make_gpu(gpu_pkg, 0.64, 0.01, 256, prm)
    gpu.seed(1)
    gpu.setWaypoint(gpu_pkg.Pose(theta=0.0, x=5.0, y=0.0))
    for _ in range(3):
        v = gpu.newControls(gpu_pkg.Pose())
        assert abs(v.ul) <= 0.05 and abs(v.ur) <= 0.05
    assert np.max(np.abs(gpu.plan())) <= 0.05


def test_set_initial_controls_fills_plan_and_tail(gpu_pkg):
    gpu = make_gpu(gpu_pkg, 0.5, 0.02, 32)
    gpu.setInitialControls(1.5, -0.5)                    # mppi.cpp:54-61
    assert np.array_equal(gpu.plan(), np.repeat([[1.5], [-0.5]], 25, axis=1))
    gpu.seed(3)
    gpu.newControls(gpu_pkg.Pose())
    p = gpu.plan()
    assert p[0, -1] == 1.5 and p[1, -1] == -0.5          # mppi.cpp:136-137


@pytest.mark.parametrize("K,obstacles,calls", [(256, False, 5), (4096, False, 40), (2048, True, 24)])
def test_async_queue_equals_synchronous_calls(gpu_pkg, K, obstacles, calls):
    """Queued calls (which wait for the previous call's plan words and for the variates' count instead of for the grids in
    front of them, with the variates drawn two calls ahead into three rotating buffers) against synchronous calls: the same
    controls and plan bit for bit, and both equal to the oracle's - also with the obstacle term (field copied from the host)."""
    a = make_gpu(gpu_pkg, 0.64, 0.01, K)
    b = make_gpu(gpu_pkg, 0.64, 0.01, K)
    o = orc.OracleMppi(0.64, 0.01, K)
    P = gpu_pkg.Pose(theta=0.2, x=0.1, y=-0.1)
    xs = np.arange(200)
    dist = np.hypot((xs[:, None] - 108) * 0.05, (xs[None, :] - 99) * 0.05).astype(np.float32)   # one obstacle cell near the path
    for m in (a, b):
        m.seed(11)
        m.setWaypoint(gpu_pkg.Pose(theta=0.0, x=1.0, y=0.0))
        if obstacles:
            m.setObstacleField(dist, -5.0, -5.0, 0.05, 5e4, 0.4, 1e6)
    o.noise_philox(11)
    o.setWaypoint(1.0, 0.0, 0.0)
    if obstacles:
        o.set_obstacles(dist, -5.0, -5.0, 0.05, 5e4, 0.4, 1e6)
    for _ in range(calls):
        va = a.newControls(P)
    for _ in range(calls):
        b.enqueue(P)
    vb = b.wait()
    assert a.lastVariant() == "fast" and b.lastVariant() == "fast"
    assert (va.ul, va.ur) == (vb.ul, vb.ur)
    assert np.array_equal(a.plan(), b.plan())
    assert np.array_equal(a.states(), b.states())
    for _ in range(calls):
        co = o.newControls(0.1, -0.1, 0.2)
    assert rel_err([vb.ul, vb.ur], co, 1e-3) < RTOL
    # (entries of the plan that happen to lie near zero are held to 1e-5 of 0.1 rad/s, not of themselves: after two dozen
    # receding-horizon calls the binary32 weights of the production variant have moved them by some 1e-7 of the control scale)
    assert rel_err(b.plan(), o.get()["plan"], 0.1) < RTOL


def test_state_ring_and_launch_count(gpu_pkg):
    gpu = make_gpu(gpu_pkg, 0.64, 0.01, 128)
    gpu.setStateRing(3)
    gpu.seed(2)
    gpu.setWaypoint(gpu_pkg.Pose(theta=0.0, x=1.0, y=0.0))
    n0 = gpu.launchCount()
    first = None
    for i in range(4):
        gpu.newControls(gpu_pkg.Pose())
        if i == 0:
            first = gpu.states().copy()
    # ONE kernel per call (rollouts, merge tree, update) + the kernel that draws a later call's variates behind it (two calls
    # ahead); the first call also draws its own and the next call's
    assert gpu.launchCount() - n0 == 10
    assert not np.array_equal(first, gpu.states())


def test_error_behaviour(gpu_pkg):
    B2NError = gpu_pkg.B2NError
    with pytest.raises(IndexError):                      # reference: std::out_of_range from .at(), mppi.hpp:66-79
        gpu_pkg.LossFunc([1.0, 1.0], [1.0, 1.0], [1.0, 1.0, 1.0])
    with pytest.raises(B2NError) as e:
        make_gpu(gpu_pkg, 0.64, 0.01, 0)
    assert e.value.code == -1
    with pytest.raises(B2NError) as e:
        make_gpu(gpu_pkg, 2.57, 0.01, 8)                 # 257 steps
    assert e.value.code == -4
    gpu = make_gpu(gpu_pkg, 0.64, 0.01, 8)
    with pytest.raises(B2NError):
        gpu.costToGo()                                   # capture was not switched on
    with pytest.raises(B2NError):
        gpu.setNoise(np.zeros((8, 63, 2)))


@pytest.mark.parametrize("seed", range(6))
def test_random_problems_match_oracle(gpu_pkg, seed):
    """Random cost weights, temperatures, robot geometry, horizons (odd step counts: partially filled lanes), rollout counts,
    waypoints and poses; capture taps on (generic kernel variant) for the first calls and off (FAST variant where the
    horizon allows it) for the rest."""
    rng = np.random.default_rng(500 + seed)
    T = int(rng.choice([3, 17, 25, 37, 64, 90, 128, 200]))
    dt = float(rng.choice([0.01, 0.02, 0.05]))
    hor = (T + 0.5) * dt
    K = int(rng.integers(2, 700))
    prm = dict(wheel_radius=float(rng.uniform(0.02, 0.06)), wheel_base=float(rng.uniform(0.1, 0.3)),
               Q=tuple(10.0 ** rng.uniform(-1, 4, 3)), R=tuple(10.0 ** rng.uniform(-2, 0, 2)), P1=tuple(10.0 ** rng.uniform(0, 3, 3)),
               lambda_=float(10.0 ** rng.uniform(-2, 1)), max_wheel_vel=float(rng.uniform(2.0, 8.0)),
               ul_var=float(rng.uniform(0.1, 1.5)), ur_var=float(rng.uniform(0.1, 1.5)))
    o = orc.OracleMppi(hor, dt, K, **prm)
    gpu = make_gpu(gpu_pkg, hor, dt, K, prm)
    assert gpu.steps == o.T == T
    gpu.seed(seed)
    o.noise_philox(seed)
    u0 = (float(rng.uniform(-1, 1)), float(rng.uniform(-1, 1)))
    gpu.setInitialControls(*u0)
    o.setInitialControls(*u0)
    wp = (float(rng.uniform(-2, 2)), float(rng.uniform(-2, 2)), float(rng.uniform(-3, 3)))
    gpu.setWaypoint(gpu_pkg.Pose(theta=wp[2], x=wp[0], y=wp[1]))
    o.setWaypoint(*wp)
    pose = (float(rng.uniform(-1, 1)), float(rng.uniform(-1, 1)), float(rng.uniform(-3, 3)))
    for c in range(6):
        gpu.setCapture(c < 2)
        v = gpu.newControls(gpu_pkg.Pose(theta=pose[2], x=pose[0], y=pose[1]))
        co = o.newControls(*pose)
        if c < 2:
            check_call(gpu, o, (v.ul, v.ur), co)
        else:
            assert rel_err([v.ul, v.ur], co, 1e-3) < RTOL
            assert rel_err(gpu.plan(), o.get()["plan"], 1e-3) < RTOL
            s, so = gpu.states(), o.get()["states"]
            assert np.max(np.abs(s - so) / np.maximum(np.abs(so), 1e-2)) < RTOL
        pose = orc.unicycle_step(pose, co[0], co[1], dt)
