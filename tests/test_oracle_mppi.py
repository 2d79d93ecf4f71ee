"""CPU tests (no GPU): pin the MPPI oracle against (1) the reference's own outputs committed under
tests/golden/, (2) the compiled reference when oracle/_ref exists, (3) hand-derived known answers
(SURVEY.md section 4, KAT1-3)."""
import os

import numpy as np
import pytest

import _oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
needs_ref = pytest.mark.skipif(not orc.have_ref(), reason="oracle/_ref not built (needs /root/reference)")


def _load(name):
    return np.load(os.path.join(GOLD, name))


def _params(g):
    return {k[2:]: (tuple(g[k]) if g[k].ndim else float(g[k])) for k in g.files if k.startswith("p_")}


def test_mt19937_normal_restatement_matches_reference_stream():
    g = _load("rng_ref.npz")
    out = np.zeros(256)
    orc.oracle_lib().orc_mt_normals(42, 256, 0.0, np.sqrt(0.9), out)
    assert np.array_equal(out, g["rigid2d_seed42_sigma_sqrt0p9"])      # bit-exact
    orc.oracle_lib().orc_mt_normals(7, 256, 0.0, 1.0, out)
    assert np.array_equal(out, g["bmapping_seed7_std"])


def test_philox_known_answer():
    # Random123 kat_vectors: philox4x32-10, ctr = key = 0 and the all-ones / pi-digits vectors
    L = orc.oracle_lib()
    out = np.zeros(4, dtype=np.uint32)
    L.orc_philox_raw(np.zeros(4, dtype=np.uint32), np.zeros(2, dtype=np.uint32), out)
    assert [hex(v) for v in out] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    L.orc_philox_raw(np.full(4, 0xFFFFFFFF, dtype=np.uint32), np.full(2, 0xFFFFFFFF, dtype=np.uint32), out)
    assert [hex(v) for v in out] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    L.orc_philox_raw(np.array([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], dtype=np.uint32),
                     np.array([0xa4093822, 0x299f31d0], dtype=np.uint32), out)
    assert [hex(v) for v in out] == ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_philox_normals_are_standard_normal():
    L = orc.oracle_lib()
    z = np.zeros(2)
    zs = []
    for i in range(20000):
        L.orc_philox_normal_pair(42, 0x4D505049, 0, i % 97, i, z)
        zs.append(z.copy())
    zs = np.array(zs).ravel()
    assert abs(zs.mean()) < 0.02 and abs(zs.std() - 1.0) < 0.02
    assert abs(np.mean(zs ** 4) - 3.0) < 0.15


def test_kat_rk4_and_loss():
    # SURVEY.md section 4 KAT1/KAT2 (numpy restatement of mppi.hpp:41-48,87-105, rk4.cpp:95-115)
    o = orc.OracleMppi(0.02, 0.01, 1)
    o.setWaypoint(1.0, 0.0, 0.0)
    du = np.array([[[1.0, 2.0], [2.0, -1.0]]])
    o.noise_external(du)
    o.newControls(0.0, 0.0, 0.0)
    g = o.get()
    assert np.allclose(g["states"][0, 0], [0.0004949996490528122, 5.104685690428451e-07, 0.0020625], rtol=1e-14, atol=0)
    assert np.allclose(g["states"][0, 1], [0.0006599992981059043, 3.4031262063788525e-07, -0.004125], rtol=1e-14, atol=0)
    # J[1] = terminal loss, J[0] = running loss + terminal
    assert g["J"][1, 0] == pytest.approx(998.6974526279776, rel=1e-14)
    assert g["J"][0, 0] - g["J"][1, 0] == pytest.approx(9990.602461521983, rel=1e-12)


def test_kat_mini_mppi():
    # K=2, T=2, u=0, lambda=0.01
    o = orc.OracleMppi(0.02, 0.01, 2)
    o.setWaypoint(1.0, 0.0, 0.0)
    du = np.array([[[0.5, 1.0], [-0.25, 0.75]], [[-0.5, 0.1], [0.25, -0.2]]])
    o.noise_external(du)
    ul, ur = o.newControls(0.0, 0.0, 0.0)
    g = o.get()
    J = np.array([[10994.525295160216, 11001.461643774279], [999.3496806569262, 1000.1155990197891]])
    assert np.allclose(g["J"], J, rtol=1e-13, atol=0)
    assert np.allclose(g["w"], [[0.9999999900000003, 9.999999800000005e-09]] * 2, rtol=1e-9, atol=0)
    # plan before the shift was [[.4999..., -.2499...],[.9999..., .7499...]]; after: column 1 moved left
    assert ul == pytest.approx(0.4999999900000002, rel=1e-12) and ur == pytest.approx(0.9999999910000003, rel=1e-12)
    assert g["plan"][0, 0] == pytest.approx(-0.2499999950000001, rel=1e-12)
    assert g["plan"][1, 0] == pytest.approx(0.7499999905000003, rel=1e-12)
    assert g["plan"][0, 1] == 0.0 and g["plan"][1, 1] == 0.0


def test_single_rollouts_match_reference_fixture():
    g = _load("mppi_rollouts_ref.npz")
    T = g["u"].shape[2]
    o = orc.OracleMppi(0.64, 0.01, 1)
    for i in range(g["u"].shape[0]):
        o.setWaypoint(*g["wpt"])
        o.set_plan(np.zeros((2, T)))
        o.noise_external(np.ascontiguousarray(g["u"][i].T[None]))
        o.newControls(*g["x0"][i])
        got = o.get()
        assert np.array_equal(got["states"][0], g["traj"][i])          # bit-exact vs the reference RK4
        loss = got["J"][:, 0] - np.append(got["J"][1:, 0], 0.0)
        assert np.allclose(loss, g["loss"][i], rtol=1e-12, atol=1e-9)   # J differencing loses a few ulps


@pytest.mark.parametrize("name", ["mppi_c1_shipped_ref.npz", "mppi_c1_mild_ref.npz", "mppi_t64_shipped_ref.npz",
                                  "mppi_t100_shipped_ref.npz"])
def test_closed_loop_matches_reference_fixture_bit_exact(name):
    """Mode A (mt19937_64 + libstdc++ normal): the oracle reproduces the reference's controls, plan,
    perturbations and min-subtracted cost-to-go exactly, call after call."""
    g = _load(name)
    o = orc.OracleMppi(float(g["horizon"]), float(g["dt"]), int(g["K"]), **_params(g))
    assert o.T == int(g["T"])
    o.noise_mt19937(int(g["seed"]))
    o.setInitialControls(0.0, 0.0)
    o.setWaypoint(*g["wpt"])
    for c in range(g["poses"].shape[0]):
        ul, ur = o.newControls(*g["poses"][c])
        got = o.get()
        assert (ul, ur) == tuple(g["controls"][c])
        assert np.array_equal(got["plan"], g["plans"][c])
        assert np.array_equal(got["du"], g["du"][c])
        assert np.array_equal(got["J"] - got["J"].min(axis=1, keepdims=True), g["Jsub"][c])


@needs_ref
def test_oracle_matches_live_reference_long_run():
    """Same check against the compiled reference itself, 30 receding-horizon calls, other sizes."""
    for (hor, dt, K, prm) in [(0.5, 0.02, 64, orc.SHIPPED), (0.3, 0.01, 33, orc.MILD)]:
        ref = orc.RefMppi(hor, dt, K, **prm)
        o = orc.OracleMppi(hor, dt, K, **prm)
        ref.seed(1234)
        o.noise_mt19937(1234)
        for m in (ref, o):
            m.setInitialControls(0.1, -0.1)
            m.setWaypoint(0.5, 0.5, 0.0)
        pose = (0.0, 0.0, 0.0)
        for c in range(30):
            a = ref.newControls(*pose)
            b = o.newControls(*pose)
            assert a == b, (c, a, b)
            pose = orc.unicycle_step(pose, a[0], a[1], dt)
        assert np.array_equal(ref.get()["plan"], o.get()["plan"])


def test_steps_truncation_hazard():
    # mppi.cpp:47: steps = (int)(horizon/dt); 0.29/0.01 -> 28, not 29
    assert orc.OracleMppi(0.29, 0.01, 1).T == 28
    assert orc.OracleMppi(0.64, 0.01, 1).T == 64
    assert orc.OracleMppi(1.28, 0.01, 1).T == 128


def test_philox_mode_is_shard_invariant():
    """Mode B noise is keyed by the GLOBAL rollout index: two half-size shards see exactly the
    perturbations the full job sees (what makes multi-GPU results independent of the split)."""
    full = orc.OracleMppi(0.5, 0.02, 64)
    full.noise_philox(42)
    full.setWaypoint(1, 0, 0)
    full.newControls(0, 0, 0)
    du = full.get()["du"]
    for off in (0, 32):
        sh = orc.OracleMppi(0.5, 0.02, 32)
        sh.noise_philox(42)
        sh.set_shard(off)
        sh.setWaypoint(1, 0, 0)
        sh.newControls(0, 0, 0)
        assert np.array_equal(sh.get()["du"], du[off:off + 32])


def test_mppi_noise_quad_known_answers_and_statistics():
    """The MPPI perturbation generator (Philox4x32-10 -> four binary32 normals by an exact-operation Box-Muller) is a
    CONVENTION of this repo shared by oracle/noise.hpp and csrc/common.cuh: pin it with known answers (bit patterns), so
    that neither side can drift unnoticed, and check that it is a standard normal."""
    import ctypes as C
    L = orc.oracle_lib()
    f = L.orc_philox_normal_quad_f32
    f.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_float)]
    f.restype = None
    out = (C.c_float * 4)()
    kat = {(42, 0x4D505049, 0, 0, 0): [1074173011, 1072889872, 3197487796, 3213492764],
           (42, 0x4D505049, 7, 16383, 31): [3216721748, 1057485327, 3185289316, 1068932674],
           (0xDEADBEEFCAFEF00D, 0x4D505049, 123456, 99, 5): [3210637932, 1076113122, 3192928633, 1047146036]}
    for args, want in kat.items():
        f(*args, out)
        assert [int(np.float32(v).view(np.uint32)) for v in out] == want, args
    z = []
    for i in range(20000):
        f(7, 0x4D505049, i % 50, i // 50, i % 13, out)
        z.extend(out[:])
    z = np.array(z, dtype=np.float64)
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1.0) < 0.02 and np.max(np.abs(z)) < 5.78
    assert abs(np.corrcoef(z[0::4], z[1::4])[0, 1]) < 0.03 and abs(np.corrcoef(z[0::4], z[2::4])[0, 1]) < 0.03
    from scipy import stats
    assert stats.kstest(z, "norm").pvalue > 1e-3
