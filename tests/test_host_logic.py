"""CPU tests of the product's HOST logic (no GPU, no compute calls): the C-ABI library loads, exports every symbol
include/b2nav.h declares, refuses to create handles without a device, and derives the same constants and tables on the
host that the reference's constructors and the oracle do."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import _oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.load_library()
    text = open(os.path.join(ROOT, "include", "b2nav.h")).read()
    declared = set(re.findall(r"\b(b2n_[a-z0-9_]+)\s*\(", text))
    assert len(declared) > 50
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(pkg._capi.PROTOTYPES), declared ^ set(pkg._capi.PROTOTYPES)


def test_no_cpu_path(pkg):
    lib = pkg.load_library()
    if lib.b2n_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.B2NError) as e:
        pkg.MPPI(pkg.CartModel(0.033, 0.16), pkg.LossFunc([1, 1, 1], [1, 1], [1, 1, 1]), 1.0, 6.0, 0.9, 0.9, 0.5, 0.02, 16)
    assert e.value.code == -2 and "no CPU path" in str(e.value)
    with pytest.raises(pkg.B2NError) as e:
        pkg.bmapping.make_filter(orc.pf_params(num_particles=4))
    assert e.value.code == -2
    q = orc.pf_params()
    props = pkg.bmapping.LaserProperties(q["beam_min"], q["beam_max"], q["beam_delta"], q["range_min"], q["range_max"],
                                         q["z_hit"], q["z_short"], q["z_max"], q["z_rand"], q["sigma_hit"])
    with pytest.raises(pkg.B2NError) as e:
        pkg.bmapping.GpuScanAlignment(props, None)
    assert e.value.code == -2 and "no CPU path" in str(e.value)


def _host_tables(pkg, q, n_beams=360):
    lib = pkg.load_library()
    p = pkg._capi.PfParams()
    for k in ("beam_min", "beam_max", "beam_delta", "range_min", "range_max", "z_hit", "z_short", "z_max", "z_rand", "sigma_hit",
              "resolution", "xmin", "xmax", "ymin", "ymax", "num_particles", "k"):
        setattr(p, k, q[k])
    const = (C.c_double * 4)()
    beam = np.zeros(2 * n_beams)
    pz = np.zeros(50000)
    n = C.c_int()
    pkg._capi.check(lib.b2n_pf_host_tables(C.byref(p), const, pkg._capi.as_ptr(beam), beam.size, pkg._capi.as_ptr(pz), pz.size, C.byref(n)))
    return list(const), beam.reshape(n_beams, 2), pz[:n.value]


def test_log_odds_thresholds_reproduce_the_reference_classification(pkg):
    """updateCellState classifies on prob = 1 - 1/(1 + exp(l)); the product classifies on l.  Single hits and single
    misses sit EXACTLY on the reference's thresholds and must fall on the same side."""
    (t_occ, t_free, d_free, d_occ), _, _ = _host_tables(pkg, orc.pf_params())
    assert d_occ == np.log(0.9 / (1 - 0.9)) and d_free == np.log(0.35 / (1 - 0.35)) - np.log(0.5 / 0.5)
    prob = lambda l: 1 - (1 / (1 + np.exp(l)))  # noqa: E731  (numpy's exp is the same libm exp on this host)
    assert prob(t_occ) >= 0.9 and prob(np.nextafter(t_occ, -np.inf)) < 0.9
    assert prob(t_free) <= 0.35 and prob(np.nextafter(t_free, np.inf)) > 0.35
    assert d_occ >= t_occ and d_free <= t_free               # one hit -> occupied, one miss -> free
    # every log-odds value reachable by up to 6 hits/misses in any order classifies like the oracle's updateCellState
    vals = {0.0}
    for _ in range(6):
        vals |= {v + d_occ for v in vals} | {v + d_free for v in vals}
    for l in vals:
        p = prob(l)
        want = 1 if (p != 0.5 and p >= 0.9) else (0 if (p != 0.5 and p <= 0.35) else -1)
        got = 1 if l >= t_occ else (0 if l <= t_free else -1)
        assert got == want, l


def test_beam_table_and_likelihood_table_match_the_oracle(pkg):
    q = orc.pf_params(xmin=-2.0, xmax=2.0, ymin=-2.0, ymax=2.0)
    _, beam, pz = _host_tables(pkg, q)
    o = orc.OraclePf(num_particles=1, **{k: q[k] for k in ("xmin", "xmax", "ymin", "ymax")})
    # beam table: end points of a unit-range scan from the origin ARE (cos, sin) of the accumulated beam angles
    ep = o.grid_end_points(np.ones(360, dtype=np.float32), (0.0, 0.0, 0.0))
    assert np.array_equal(ep, beam)
    # likelihood table: one obstacle, probe beams ending d cells short of it along +x
    scan = np.full(360, 10.0, dtype=np.float32)
    scan[0] = 1.5
    assert o.grid_integrate(scan, (0.0, 0.0, 0.0)) == 0
    occ = o.grid()["occ_dist"]
    xs = o.xsize
    for d in (0, 1, 2, 5, 17):
        probe = np.full(360, 10.0, dtype=np.float32)
        probe[0] = np.float32(1.5 - 0.05 * d + 0.02)
        rc, lik = o.grid_likelihood(probe, (0.0, 0.0, 0.0))
        cell = int(np.floor((float(probe[0]) + 2.0) / 0.05)) * xs + 40
        d2 = int(round((occ[cell] / 0.05) ** 2))
        assert rc == 0 and lik == pz[d2], (d, d2)
    assert len(pz) == 200 * 200 + 2
    assert pz[-1] == 0.95 * (1.0 / np.sqrt(2.0 * np.pi * 0.25)) * np.exp(-0.5 * (10.0 * 10.0) / 0.25) + 0.01 / 0.04


@pytest.mark.skipif(not orc.have_ref(), reason="needs oracle/_ref (the compiled reference)")
def test_synthetic_plant_is_the_reference_diff_drive(pkg):
    """synthetic.DiffDrive / WaypointSwitch (the plant and node bookkeeping bench.py's closed-loop leg runs on) against the
    compiled reference rigid2d::DiffDrive: a simulated robot driven by feedforward (fake_diff_encoders_node.cpp:100-135) and an
    odometer fed its encoder angles (updateOdometry), 400 random commands incl. pure rotations, straight lines and stand-still:
    poses, encoders and wheel velocities agree to the last bit or two (Python floats are C doubles on the same libm)."""
    syn = pkg.synthetic
    rng = np.random.default_rng(3)
    start = (0.3, -0.2, 0.7)
    robots = (syn.DiffDrive(start, 0.16, 0.033), orc.RefDiffDrive(start, 0.16, 0.033))
    odos = (syn.DiffDrive(start, 0.16, 0.033), orc.RefDiffDrive(start, 0.16, 0.033))
    for k in range(400):
        kind = k % 8
        ul, ur = rng.uniform(-6.3, 6.3, 2)
        if kind == 5:
            ul = ur
        if kind == 6:
            ul = -ur
        if kind == 7:
            ul = ur = 0.0
        states = []
        for rb, od in zip(robots, odos):
            w, vx, vy = rb.wheelsToTwist(ul, ur)
            rb.feedforward(w / 50.0, vx / 50.0)
            el, er = rb.getEncoders()
            vel = od.updateOdometry(el, er)
            states.append(np.array(rb.pose() + rb.getEncoders() + rb.wheelVelocities() + od.pose() + tuple(vel) + od.wheelsToTwist(*vel)))
        assert np.max(np.abs(states[0] - states[1])) < 1e-13, (k, states)
    for a in np.concatenate([rng.uniform(-20, 20, 200), [0.0, np.pi, -np.pi, 3 * np.pi]]):
        assert syn.normalize_angle_PI(float(a)) == orc.ref_lib().ref_normalize_angle_pi(float(a))
    # the node's waypoint bookkeeping (mppi_waypoints_node.cpp:231-258) on the shipped pentagon
    sw = syn.WaypointSwitch([0, 1, 1, 0.5, 0], [0, 0, 1, 2, 1], [0, 1.5707, 2.3562, -2.3562, -1.5707], 0.1)
    assert sw.update(0.5, 0.5) is None and sw.wpt_id == 0
    seen = []
    for _ in range(6):
        x, y, _th = sw.current()
        seen.append(sw.update(x + 0.05, y))
    assert [s[:2] for s in seen] == [(1, 0), (1, 1), (0.5, 2), (0, 1), (0, 0), (1, 0)] and sw.cycle_complete and sw.cnt == 6
